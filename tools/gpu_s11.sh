#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s11_pytest.log
tail -15 gpurun_out/s11_pytest.log
PCAONE_ORTH_PROF=2 timeout 300 python bench.py --scale 0.125 --steps 2 --warmup 1 --no-cpu --no-e2e 2>&1 | grep -v "^{" | head -4
PCAONE_ORTH_PROF=2 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/s11_bench_c3.log 2>&1; grep "orth_fused rows" gpurun_out/s11_bench_c3.log | head -2
grep '^{' gpurun_out/s11_bench_c3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print({k:d[k] for k in ['value','time_to_pcs_s']}, d['e2e']['time_to_pcs_s'], {k:r[k] for k in ['tc_g_ms_per_pca','tc_h_ms_per_pca','orth_ms_per_pca','small_stage_ms_per_pca','gemm_g_ms_per_pca','gemm_h_ms_per_pca']}, d['config']['top_eigenvalues'])"
