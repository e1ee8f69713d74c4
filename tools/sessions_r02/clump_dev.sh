#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py -q -x -k "clump or print_r2 or prune" > gpurun_out/s24_clump.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s24_clump.log
tail -40 gpurun_out/s24_clump.log
