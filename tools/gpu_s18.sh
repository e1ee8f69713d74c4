#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "int8x2|passed|failed|Error" | head
for sk in 1 0; do
PCAONE_OMEGA_SKIP2=$sk timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/s18_bench_skip$sk.log 2>&1
grep '^{' gpurun_out/s18_bench_skip$sk.log > gpurun_out/s18_skip$sk.json; python -c "
import json
d=json.load(open('gpurun_out/s18_skip$sk.json')); r=d['roofline']
print('skip2=$sk', {k:d[k] for k in ['value','time_to_pcs_s']}, {k:r[k] for k in ['tc_g_ms_per_pca','tc_h_ms_per_pca','orth_ms_per_pca','small_stage_ms_per_pca']}, ['%.12f' % x for x in d['config']['top_eigenvalues']], d['config']['U_orthonormality_err'])"
done
