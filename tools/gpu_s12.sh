#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "int8x2|passed|failed|Error" | head
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --precision int8x2 > gpurun_out/s12_bench_c3_x2.log 2>&1
grep '^{' gpurun_out/s12_bench_c3_x2.log > gpurun_out/s12_x2.json; python -c "
import json
d=json.load(open('gpurun_out/s12_x2.json')); r=d['roofline']
print({k:d[k] for k in ['value','time_to_pcs_s','dtype']}, d['e2e']['time_to_pcs_s'], {k:r[k] for k in ['tc_g_ms_per_pca','tc_h_ms_per_pca','orth_ms_per_pca','small_stage_ms_per_pca','frac']}, d['config']['top_eigenvalues'])"
PCAONE_ORTH_PROF=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/s12_bench_c3.log 2>&1
grep "orth_fused rows" gpurun_out/s12_bench_c3.log | head -1
grep '^{' gpurun_out/s12_bench_c3.log > gpurun_out/s12_x3.json; python -c "
import json
d=json.load(open('gpurun_out/s12_x3.json')); r=d['roofline']
print({k:d[k] for k in ['value','time_to_pcs_s','dtype']}, d['e2e']['time_to_pcs_s'], {k:r[k] for k in ['tc_g_ms_per_pca','tc_h_ms_per_pca','orth_ms_per_pca','small_stage_ms_per_pca','frac']}, d['config']['top_eigenvalues'])"
