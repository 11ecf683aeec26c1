"""Executed warp instructions per opcode (and stall samples) from `ncu -i X.ncu-rep --page source --csv --print-source sass`.
python tools/ncu_opcodes.py source.csv [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
si, ei, st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
agg, stall = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    try:
        n, s = int(r[ei]), int(r[st])
    except ValueError:
        continue
    toks = r[si].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    agg[op] += n
    stall[op] += s
tot, stot = sum(agg.values()), sum(stall.values()) or 1
print(f"total warp instructions {tot}, stall samples {stot}")
for k, v in agg.most_common(top):
    print(f"{k:10s} {v:12d} {v / tot * 100:5.1f}%   stall samples {stall[k] / stot * 100:5.1f}%")
