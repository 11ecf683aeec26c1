#!/bin/bash
# ncu --set full of the bench's dominant kernel (ss2d_core_fwd, stage-1 shape) -- only after the bench itself exits 0
mkdir -p gpurun_out
B=${1:-1024}
timeout 600 python bench.py --no-cpu-baseline --batch $B --steps 3 --warmup 3 > gpurun_out/bench_pre_ncu.json 2> gpurun_out/bench_pre_ncu.err || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ss2d_core_fwd_kernel -s 30 -c 1 -f \
  -o gpurun_out/core_bench_stage1_b$B python bench.py --no-cpu-baseline --batch $B --steps 1 --warmup 3 > gpurun_out/ncu_core.log 2>&1
ls -la gpurun_out/*.ncu-rep
