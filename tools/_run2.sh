NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:dwconv3x3_silu_bwd_ds -c 1 -o gpurun_out/r2s3_dwconv_bwd_ds python bench.py --workload train --batch 128 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2s3_ncu_dw.log 2>&1
tail -3 gpurun_out/r2s3_ncu_dw.log
f=gpurun_out/r2s3_dwconv_bwd_ds
ncu -i $f.ncu-rep --page raw --csv > ${f}_raw.csv 2>/dev/null && python tools/ncu_summary.py ${f}_raw.csv > ${f}_metrics.txt
ncu -i $f.ncu-rep --page source --csv --print-source sass > ${f}_source.csv 2>/dev/null && python tools/ncu_opcodes.py ${f}_source.csv 30 > ${f}_opcodes.txt
rm -f ${f}_raw.csv ${f}_source.csv
