"""Flattens `ncu -i X.ncu-rep --page raw --csv` into `metric [unit] = value` lines (one kernel).
python tools/ncu_summary.py raw.csv > profiles/NAME_metrics.txt"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
for h, u, v in zip(hdr, units, vals):
    print(f"{h} [{u}] = {v}")
