#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/s19_bench_n8.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/s19_bench_n8.log > gpurun_out/s19_n8.json; python -c "
import json
d=json.load(open('gpurun_out/s19_n8.json')); r=d['roofline']
print({k:d[k] for k in ['value','time_to_pcs_s']}, d['e2e']['time_to_pcs_s'], {k:r[k] for k in ['tc_g_ms_per_pca','tc_h_ms_per_pca','orth_ms_per_pca','small_stage_ms_per_pca','allreduce_ms_per_pca','gemm_g_ms_per_pca','gemm_h_ms_per_pca']}, r['late_pass']['ms'], d['config']['top_eigenvalues'][0])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 5 --warmup 3 --no-e2e > gpurun_out/s19_bench_n4.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/s19_bench_n4.log > gpurun_out/s19_n4.json; python -c "
import json
d=json.load(open('gpurun_out/s19_n4.json')); r=d['roofline']
print('n4', {k:d[k] for k in ['value','time_to_pcs_s']}, r['orth_ms_per_pca'], r['late_pass']['ms'])"
