#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_downstream.py -q -x > gpurun_out/s37_project.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s37_project.log
tail -30 gpurun_out/s37_project.log
