#!/bin/bash
# ncu --set full of the EMU correction kernels at 1/10 of configs[3]'s SNP axis (50k x 50k, 10 % missing): one late-pass launch each
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_emu_fix -s 40 -c 2 -f -o gpurun_out/s27_prof_emu \
  python tools/run_configs.py c4 --scale 0.1 --out gpurun_out/s27_c4_small.jsonl > gpurun_out/s27_ncu_emu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/s27*.ncu-rep
tail -3 gpurun_out/s27_ncu_emu.log
