#!/bin/bash
# round-2 session 2: 2-GPU NCCL parity (in-library communicator) + strong-scaling bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
tail -15 gpurun_out/s2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/s2_bench_n2.log 2>&1; echo "rc=$?" >> gpurun_out/s2_bench_n2.log
tail -c 5000 gpurun_out/s2_bench_n2.log
