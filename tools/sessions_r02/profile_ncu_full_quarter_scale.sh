#!/bin/bash
# ncu --set full at 1/4 linear scale (125k x 125k: the 500k x 500k run holds 130 GB, which kernel replay cannot save / restore)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 258 -c 4 -f -o gpurun_out/s10_prof_tc \
  python bench.py --scale 0.25 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/s10_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_orth_fused -s 128 -c 1 -f -o gpurun_out/s10_prof_orth \
  python bench.py --scale 0.25 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/s10_ncu_orth.log 2>&1; echo "ncu orth rc=$?"
ls -la gpurun_out/s10*.ncu-rep
timeout 600 python tools/run_configs.py c4 c5 --out gpurun_out/s10_configs.jsonl > gpurun_out/s10_configs.log 2>&1; cat gpurun_out/s10_configs.jsonl
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | head -c 900
