#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bgen.py -q -x > gpurun_out/s31_bgen.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s31_bgen.log
tail -30 gpurun_out/s31_bgen.log
