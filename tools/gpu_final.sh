#!/bin/bash
# end-of-session evidence run on one GPU: tests, default bench (+ CPU baseline), reference arm, the other BASELINE configs.
# `ncu` (launch list + --set full of the core kernels) only with --ncu, after the plain runs have exited 0.
# Every step has its own timeout and writes to a file: a hung step costs its timeout, not the call.
P=${PREFIX:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${P}_pytest.txt 2>&1; tail -2 gpurun_out/${P}_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err || { tail -5 gpurun_out/${P}_bench.err; exit 1; }
python tools/show_bench.py gpurun_out/${P}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${P}_bench_reference.json 2>> gpurun_out/${P}_bench.err
timeout 300 python bench.py --no-cpu-baseline --no-extras --batch 256 > gpurun_out/${P}_bench_b256.json 2>> gpurun_out/${P}_bench.err
timeout 300 python bench.py --no-cpu-baseline --no-extras --dtype f32 --batch 256 > gpurun_out/${P}_bench_f32_b256.json 2>> gpurun_out/${P}_bench.err
timeout 300 python bench.py --workload train --batch 128 --steps 10 > gpurun_out/${P}_bench_train_b128.json 2>> gpurun_out/${P}_bench.err
for f in reference b256 f32_b256 train_b128; do python -c "
import json; d=json.loads([l for l in open('gpurun_out/${P}_bench_$f.json') if l.startswith('{')][-1]); print('$f', d['value'], d['ms_per_step'], d.get('e2e'))"; done
if [ "$1" == "--ncu" ]; then
BENCH="python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_launches_b1024.csv $BENCH > gpurun_out/${P}_ncu_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${P}_launches_b1024.csv 60 > gpurun_out/${P}_launches_b1024_summary.txt; head -12 gpurun_out/${P}_launches_b1024_summary.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:ss2d_core_fwd_kernel -s 30 -c 1 -o gpurun_out/${P}_core_fwd_stage1_b1024 $BENCH > gpurun_out/${P}_ncu_core.log 2>&1
BW="python tools/core_bwd_bench.py --batch 128 --stage 0 --iters 1"
timeout 200 $BW > gpurun_out/${P}_bwd_plain.log 2>&1 && timeout 600 $NCU -k regex:ss2d_core_bwd_kernel -s 2 -c 1 -o gpurun_out/${P}_core_bwd_stage1_b128 $BW > gpurun_out/${P}_ncu_bwd.log 2>&1
for f in gpurun_out/${P}_*.ncu-rep; do
  b=${f%.ncu-rep}
  ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null && python tools/ncu_summary.py ${b}_raw.csv > ${b}_metrics.txt
  ncu -i $f --page source --csv --print-source sass > ${b}_source.csv 2>/dev/null && python tools/ncu_opcodes.py ${b}_source.csv 30 > ${b}_opcodes.txt
  rm -f ${b}_raw.csv ${b}_source.csv
done
ls -la gpurun_out/${P}_*.ncu-rep; rm -f gpurun_out/${P}_core_bwd_stage1_b128.ncu-rep
du -sh gpurun_out
fi
