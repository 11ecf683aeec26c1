#!/bin/bash
# sweep (S, cap) for the fused core kernel; prints stage, S, cap, ms
for b in "$@"; do
for s in 1 2 4; do for c in 4 6 8 12 16 24 31; do
  MMB_CORE_S=$s MMB_CORE_CAP=$c python tools/core_bench.py --batch $b --iters 5 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    d = json.loads(line); print('b=$b S=$s cap=$c stage', d['stage'], 'ms', d['ms'], 'mufu', d['mufu_frac'])
"
done; done; done
