"""Prints value / ms_per_step / e2e and the per-kernel averages matching a substring from a bench.py JSON line.
python tools/kern.py bench.json [substring]"""
import json, sys
d = json.load(open(sys.argv[1]))
pat = sys.argv[2] if len(sys.argv) > 2 else ""
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["alu"]["frac"] if d.get("roofline") else None)
for k, v in d.get("kernels", {}).items():
    if pat in k:
        print(k, v)
