#!/bin/bash
# final code state on 2 GPUs: NCCL inside the library + in-kernel peer exchange (multi-rank parity), then the bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/s33_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s33_multi.log
tail -4 gpurun_out/s33_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s33_bench_n2.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/s33_bench_n2.log > gpurun_out/s33_n2.json; python -c "
import json
d=json.load(open('gpurun_out/s33_n2.json')); r=d['roofline']
print({k:d[k] for k in ['value','time_to_pcs_s','n_gpus','scaling']}, d['e2e']['time_to_pcs_s'], {k:r[k] for k in ['tc_g_ms_per_pca','tc_h_ms_per_pca','orth_ms_per_pca','small_stage_ms_per_pca','allreduce_ms_per_pca']}, r['late_pass']['ms'], d['config']['top_eigenvalues'][0], d['config'].get('parallelism'))"
