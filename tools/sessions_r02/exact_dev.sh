#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact.py tests/test_cli.py -q -x -k "exact or sym_svd or covariance or svd3 or usv_projection or clump" > gpurun_out/s25_exact.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s25_exact.log
tail -40 gpurun_out/s25_exact.log
