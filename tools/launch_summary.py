"""Summarises an ncu launch list (gpu__time_duration.sum per launch, --csv) by kernel: share of the LAST step.
python tools/launch_summary.py launches.csv [top]"""
import collections, csv, sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr, recs = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        recs.append((d["Kernel Name"], float(d["Metric Value"].replace(",", "")), d["Grid Size"], d["Block Size"]))
    except ValueError:
        pass
# the last forward pass: from the last patch-embed kernel onwards (bench runs warm-up steps first)
names = [n for n, *_ in recs]
starts = [i for i, n in enumerate(names) if "patch_embed_ln" in n]
lo = starts[-1] if starts else 0
sel = recs[max(lo, 0):]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v, g, b in sel:
    agg[n[:100]][0] += 1
    agg[n[:100]][1] += v
tot = sum(v for _, v in agg.values())
print(f"launches in the last step: {len(sel)}, total {tot / 1e6:.3f} ms (serialised, cold)")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{t / tot * 100:5.1f}% {c:4d} {t / 1e3:9.1f} us  {n}")
