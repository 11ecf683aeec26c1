#!/bin/bash
# Round-2 evidence: launch list of one bench step and ncu --set full captures of the kernels DESIGN.md discusses.
# Every ncu run follows a plain run of the same command that exited 0 (B200_PROFILING.md).
mkdir -p gpurun_out
BENCH="python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 3"
timeout 600 python bench.py --no-cpu-baseline --no-extras --steps 3 --warmup 3 > gpurun_out/r2p_bench_pre.json 2> gpurun_out/r2p_bench_pre.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_launches_b1024.csv $BENCH > gpurun_out/r2p_ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/r2p_launches_b1024.csv 60 > gpurun_out/r2p_launches_b1024_summary.txt
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:ss2d_core_fwd_kernel -s 30 -c 1 -o gpurun_out/r2p_core_fwd_stage1_b1024 $BENCH > gpurun_out/r2p_ncu_core.log 2>&1
timeout 900 $NCU -k regex:outnorm_gate_bf16x8 -s 30 -c 1 -o gpurun_out/r2p_outnorm_bf16x8_stage1_b1024 $BENCH > gpurun_out/r2p_ncu_outnorm.log 2>&1
timeout 900 $NCU -k regex:patch_embed_ln_mma -s 3 -c 1 -o gpurun_out/r2p_patch_embed_mma_b1024 $BENCH > gpurun_out/r2p_ncu_pe.log 2>&1
C5="python tools/core_bench.py --bf16 --batch 32 --res 512 --stage 0 --iters 1"
timeout 300 $C5 > gpurun_out/r2p_c5_plain.log 2>&1 && timeout 900 $NCU -k regex:ss2d_core_fwd_kernel -s 6 -c 2 -o gpurun_out/r2p_core_fwd_lparallel_c5 $C5 > gpurun_out/r2p_ncu_c5.log 2>&1
BW="python tools/core_bwd_bench.py --batch 128 --stage 0 --iters 1"
timeout 300 $BW > gpurun_out/r2p_bwd_plain.log 2>&1 && timeout 900 $NCU -k regex:ss2d_core_bwd_kernel -s 2 -c 1 -o gpurun_out/r2p_core_bwd_stage1_b128 $BW > gpurun_out/r2p_ncu_bwd.log 2>&1
# keep the box's gpurun_out small (64 MiB cap): text summaries of every capture, the .ncu-rep of the dominant kernel only
for f in gpurun_out/r2p_*.ncu-rep; do
  b=${f%.ncu-rep}
  ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null && python tools/ncu_summary.py ${b}_raw.csv > ${b}_metrics.txt
  ncu -i $f --page source --csv --print-source sass > ${b}_source.csv 2>/dev/null && python tools/ncu_opcodes.py ${b}_source.csv 30 > ${b}_opcodes.txt
  rm -f ${b}_raw.csv ${b}_source.csv
done
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/r2p_*.ncu-rep; do case $f in *core_fwd_stage1_b1024*) ;; *) rm -f $f;; esac; done
du -sh gpurun_out
