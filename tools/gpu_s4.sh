#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cache.py tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/s4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest.log
tail -25 gpurun_out/s4_pytest.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/s4_bench_c3.log 2>&1; echo "rc=$?" >> gpurun_out/s4_bench_c3.log
tail -c 4000 gpurun_out/s4_bench_c3.log
