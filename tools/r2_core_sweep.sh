#!/bin/bash
out=gpurun_out/r2_core_sweep.jsonl
: > $out
run() { echo "# $*" >> $out; timeout 300 python tools/core_bench.py "$@" >> $out 2>&1; }
run --bf16 --batch 1024 --iters 5 --variants "default,MMB_CORE_YF32=1"
run --bf16 --batch 256 --iters 7
run --batch 256 --iters 7
run --bf16 --batch 32 --res 512 --iters 10 --variants "default,MMB_CORE_SEGS=1"
run --bf16 --batch 32 --res 512 --iters 10 --stage 0 --variants "MMB_CORE_SEGS=3,MMB_CORE_SEGS=5,MMB_CORE_SEGS=8,MMB_CORE_SEGS=12,MMB_CORE_SEGS=24"
run --bf16 --batch 32 --res 512 --iters 10 --stage 1 --variants "MMB_CORE_SEGS=2,MMB_CORE_SEGS=3,MMB_CORE_SEGS=4,MMB_CORE_SEGS=6"
run --batch 8 --iters 20 --variants "default,MMB_CORE_SEGS=1"
run --bf16 --batch 64 --iters 10 --variants "default,MMB_CORE_SEGS=2,MMB_CORE_SEGS=3,MMB_CORE_SEGS=4"
run --bf16 --batch 1 --iters 20 --variants "default,MMB_CORE_SEGS=1,MMB_CORE_S=4"
