mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2s4_pytest.txt 2>&1; tail -3 gpurun_out/r2s4_pytest.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2s4_bench_try1.json 2> gpurun_out/r2s4_bench_try1.err
python tools/show_bench.py gpurun_out/r2s4_bench_try1.json
