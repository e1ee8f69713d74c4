#!/bin/bash
# EMU update passes on the int8 route: new tests + the EMU cases of the existing suites
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_emu.py tests/test_gpu_scale.py::test_configs3_shape_emu_vs_reference "tests/test_gpu_parity.py::test_emu_vs_golden" tests/test_cli.py::test_cli_emu_vs_reference -q -x > gpurun_out/s22_emu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s22_emu.log
tail -40 gpurun_out/s22_emu.log
