"""Prints the interesting parts of a bench.py JSON line: python tools/show_bench.py gpurun_out/x.json"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d.get('gpu_launches'), 'clocks', d.get('clocks'))
for s in d.get('roofline_stages') or []:
    print('  stage', s['stage'], 'ms', s['avg_ms'], 'hbm', s['frac'], 'alu', s['alu']['frac'])
ks = d.get('kernels', {})
tot = {}
for k, v in ks.items():
    n = k.split('[')[0]
    tot[n] = tot.get(n, 0) + v['avg_ms'] * v['count'] / d['steps']
for n, v in sorted(tot.items(), key=lambda t: -t[1]):
    print(f'  {n:28s} {v:8.3f} ms/step')
t = d.get('train')
if t:
    print('train', t['value'], 'img/s', t['ms_per_step'], 'ms', t.get('allreduce'))
for k, v in (d.get('configs') or {}).items():
    if isinstance(v, dict):
        print(k, v['ms_per_step'], 'ms', v['value'], 'img/s', [(s['stage'], s['avg_ms'], s['alu']['frac']) for s in v['roofline_stages']])
    else:
        for r in v:
            print('  scan', r['dtype'], r.get('bc_layout', ''), 'b%d' % r.get('batch', 64), r['KD'], r['L'], 'fwd', r['fwd']['ms'], r['fwd']['hbm_frac'], 'fwd+bwd', r['fwd_bwd']['ms'], r['fwd_bwd']['hbm_frac'])
print('cpu', d.get('cpu_baseline'))
