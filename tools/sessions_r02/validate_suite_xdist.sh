#!/bin/bash
# the whole -m gpu suite once more on the final code state, three workers on the one GPU (budget: < 3 GPU-minutes)
mkdir -p gpurun_out
timeout 190 python -m pytest tests -m gpu -q -n 3 -p no:cacheprovider > gpurun_out/s38_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s38_pytest.log
tail -15 gpurun_out/s38_pytest.log
