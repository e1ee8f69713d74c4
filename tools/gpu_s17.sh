#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py tests/test_gpu_ld.py -m gpu -x -q > gpurun_out/s17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s17_pytest.log
tail -25 gpurun_out/s17_pytest.log
