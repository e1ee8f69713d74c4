#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_beagle.py tests/test_cli.py -q -x -k "beagle or pcangsd or grm" > gpurun_out/s26_grm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s26_grm.log
tail -40 gpurun_out/s26_grm.log
