#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cli.py -q -x -k "csv or bgen" > gpurun_out/s35_csv_cli.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s35_csv_cli.log
tail -40 gpurun_out/s35_csv_cli.log
