#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_emu.py tests/test_gpu_scale.py::test_configs3_shape_emu_vs_reference -q -x > gpurun_out/s30_emu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s30_emu.log
tail -3 gpurun_out/s30_emu.log
bash tools/sessions_r02/emu_sb_try.sh
