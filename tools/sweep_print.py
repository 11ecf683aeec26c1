"""Pretty-prints the JSON lines of tools/core_bench.py / core_bwd_bench.py sweeps."""
import json, sys
# a file name as the first argument, else standard input (never block on a terminal-less stdin by accident)
for l in (open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin):
    try:
        d = json.loads(l)
    except Exception:
        print(l.strip()[:200]); continue
    env = ",".join(f"{k[4:]}={v}" for k, v in d.get("env", {}).items())
    print(f"stage {d.get('stage')} B={d.get('batch')} {env:24s} ms={d.get('ms')} min={d.get('min_ms')} mufu={d.get('mufu_frac')} {d.get('error') or ''}")
