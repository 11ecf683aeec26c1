"""Flattens `ncu -i X.ncu-rep --page raw --csv` into `metric [unit] = value` lines, one block per captured kernel.
python tools/ncu_summary.py raw.csv > profiles/NAME_metrics.txt"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for n, vals in enumerate(rows[2:]):
    if len(vals) != len(hdr):
        continue
    if n:
        print("\n" + "=" * 100)
    for h, u, v in zip(hdr, units, vals):
        print(f"{h} [{u}] = {v}")
