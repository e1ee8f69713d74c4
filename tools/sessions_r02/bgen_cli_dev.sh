#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py -q -x -k "bgen or rejects or help" > gpurun_out/s34_bgen_cli.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s34_bgen_cli.log
tail -30 gpurun_out/s34_bgen_cli.log
