#!/bin/bash
# round-2 session 1: regression of the refactor + first configs[2] bench on one B200
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/s1_smi.txt 2>&1
free -g >> gpurun_out/s1_smi.txt; nproc >> gpurun_out/s1_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
timeout 300 python bench.py --scale 0.1 --steps 2 --warmup 1 --no-cpu > gpurun_out/s1_bench_small.log 2>&1; echo "rc=$?" >> gpurun_out/s1_bench_small.log
tail -c 1500 gpurun_out/s1_bench_small.log
PCAONE_ORTH_PROF=6 PCAONE_SMALL_PROF=2 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/s1_bench_prof.log 2>&1; echo "rc=$?" >> gpurun_out/s1_bench_prof.log
tail -c 3000 gpurun_out/s1_bench_prof.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/s1_bench_c3.log 2>&1; echo "rc=$?" >> gpurun_out/s1_bench_c3.log
tail -c 6000 gpurun_out/s1_bench_c3.log
