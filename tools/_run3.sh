timeout 600 python -m pytest tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --workload train --batch 128 --steps 10 --no-cpu-baseline --no-extras > gpurun_out/r2s3_train_a.json 2>gpurun_out/r2s3_train_a.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2s3_train_a.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'])
for k,v in sorted(d['kernels'].items()):
    if 'dwconv' in k or 'outnorm_gate_bwd' in k or 'layernorm_bwd' in k: print(k, v)
P
