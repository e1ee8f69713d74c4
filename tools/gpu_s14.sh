#!/bin/bash
mkdir -p gpurun_out
for mode in 1 0; do
PCAONE_PEER_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$mode bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e > gpurun_out/s14_bench_n8_peer$mode.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/s14_bench_n8_peer$mode.log > gpurun_out/s14_n8_peer$mode.json; python -c "
import json
d=json.load(open('gpurun_out/s14_n8_peer$mode.json')); r=d['roofline']
print('peer=$mode', {k:d[k] for k in ['value','time_to_pcs_s']}, {k:r[k] for k in ['tc_g_ms_per_pca','tc_h_ms_per_pca','orth_ms_per_pca','small_stage_ms_per_pca','allreduce_ms_per_pca','gemm_g_ms_per_pca','gemm_h_ms_per_pca']}, r['late_pass']['ms'], d['config']['top_eigenvalues'][0])"
done
