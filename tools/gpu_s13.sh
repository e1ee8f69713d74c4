#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/s13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s13_pytest.log
tail -12 gpurun_out/s13_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/s13_bench_n2.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/s13_bench_n2.log > gpurun_out/s13_n2.json; python -c "
import json
d=json.load(open('gpurun_out/s13_n2.json')); r=d['roofline']
print({k:d[k] for k in ['value','time_to_pcs_s']}, {k:r[k] for k in ['tc_g_ms_per_pca','tc_h_ms_per_pca','orth_ms_per_pca','small_stage_ms_per_pca','allreduce_ms_per_pca','gemm_g_ms_per_pca','gemm_h_ms_per_pca']}, d['config']['top_eigenvalues'])"
PCAONE_PEER_EXCHANGE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/s13_bench_n2_nccl.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/s13_bench_n2_nccl.log > gpurun_out/s13_n2_nccl.json; python -c "
import json
d=json.load(open('gpurun_out/s13_n2_nccl.json')); r=d['roofline']
print('nccl-phases:', {k:d[k] for k in ['value','time_to_pcs_s']}, {k:r[k] for k in ['orth_ms_per_pca','allreduce_ms_per_pca']})"
