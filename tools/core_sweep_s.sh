#!/bin/bash
for b in "$@"; do for s in 1 2 4; do
  MMB_CORE_S=$s python tools/core_bench.py --batch $b --iters 5 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    d = json.loads(line); print('b=$b S=$s stage', d['stage'], 'ms', d['ms'])
"
done; done
