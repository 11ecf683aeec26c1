timeout 800 python -m pytest tests/test_train_gpu.py tests/test_trainer.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/core_bwd_bench.py --batch 128 > gpurun_out/r2s3_bwd_b.txt 2>&1; cat gpurun_out/r2s3_bwd_b.txt
timeout 300 python bench.py --workload train --batch 128 --steps 10 --no-cpu-baseline --no-extras > gpurun_out/r2s3_train_b.json 2>gpurun_out/r2s3_train_b.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/r2s3_train_b.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'])
for k,v in sorted(d['kernels'].items()):
    if 'core' in k: print(k, v)
P
