#!/bin/bash
# end-of-session evidence run on one GPU: tests, default bench (+ CPU baseline), reference arm, the other BASELINE configs.
# `ncu` (launch list + --set full of the core kernel) only with --ncu, after the plain runs have exited 0.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.txt 2>&1; tail -2 gpurun_out/final_pytest.txt
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err || exit 1
python tools/kern.py gpurun_out/final_bench.json xxxx
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err
timeout 600 python bench.py --no-cpu-baseline --batch 256 > gpurun_out/final_bench_b256.json 2>> gpurun_out/final_bench.err
timeout 600 python bench.py --no-cpu-baseline --batch 32 --res 512 > gpurun_out/final_bench_res512_b32.json 2>> gpurun_out/final_bench.err
timeout 600 python bench.py --no-cpu-baseline --dtype f32 --batch 256 > gpurun_out/final_bench_f32_b256.json 2>> gpurun_out/final_bench.err
timeout 600 python bench.py --workload train --batch 128 > gpurun_out/final_bench_train_b128.json 2>> gpurun_out/final_bench.err
for f in reference b256 res512_b32 f32_b256 train_b128; do python -c "
import json; d=json.load(open('gpurun_out/final_bench_$f.json')); print('$f', d['value'], d['ms_per_step'], d.get('e2e'))"; done
if [ "$1" == "--ncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_b1024.csv \
  python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/final_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/final_launches_b1024.csv 30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ss2d_core_fwd_kernel -s 30 -c 1 -f \
  -o gpurun_out/final_core_stage1_b1024 python bench.py --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/final_ncu_core.log 2>&1
ls -la gpurun_out/final_core_stage1_b1024.ncu-rep
fi
