#!/bin/bash
# ncu launch list (per-kernel durations) of one bench step; bench first without ncu
mkdir -p gpurun_out
B=${1:-256}
timeout 300 python bench.py --no-cpu-baseline --batch $B --steps 3 --warmup 3 > gpurun_out/bench_pre_ncu.json 2> gpurun_out/bench_pre_ncu.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_b$B.csv \
  python bench.py --no-cpu-baseline --batch $B --steps 1 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_b$B.csv 45
