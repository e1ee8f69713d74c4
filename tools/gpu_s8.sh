#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s8_pytest.log
tail -30 gpurun_out/s8_pytest.log
PCAONE_ORTH_PROF=3 PCAONE_SMALL_PROF=2 timeout 300 python bench.py --scale 0.125 --steps 2 --warmup 1 --no-cpu --no-e2e 2>&1 | grep -v "^{" | head -8
