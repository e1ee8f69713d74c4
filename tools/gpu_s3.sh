#!/bin/bash
# round-2 session 3: strong-scaling bench at N=8 (+ the 8-rank parity script)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/s3_bench_n8.log 2>&1; echo "rc=$?" >> gpurun_out/s3_bench_n8.log
tail -c 4500 gpurun_out/s3_bench_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 tests/mgpu_check.py --transport nccl > gpurun_out/s3_mgpu8.log 2>&1; echo "rc=$?" >> gpurun_out/s3_mgpu8.log
grep -v Warning gpurun_out/s3_mgpu8.log | tail -12
