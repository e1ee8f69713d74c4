#!/bin/bash
# A/B of the H-pass finish fused into the GEMM epilogue (PCAONE_FUSE_FINISH_H), then the GPU suite on the fused path
mkdir -p gpurun_out
for f in 0 1; do
  PCAONE_FUSE_FINISH_H=$f timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/s21_bench_fuse$f.json 2> gpurun_out/s21_bench_fuse$f.err; echo "bench fuse=$f rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/s21_bench_fuse$f.json')); r=d['roofline']
print({k:d[k] for k in ['value','time_to_pcs_s','gpu_launches']}, {k:r[k] for k in ['frac','tc_g_ms_per_pca','tc_h_ms_per_pca','gemm_g_ms_per_pca','gemm_h_ms_per_pca','orth_ms_per_pca']}, r['late_pass']['ms'])"
done
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s21_pytest.log
tail -6 gpurun_out/s21_pytest.log
