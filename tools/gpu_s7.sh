#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/s7_bench_n8.log 2>&1; echo "rc=$?" >> gpurun_out/s7_bench_n8.log
tail -c 3500 gpurun_out/s7_bench_n8.log
