#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu > gpurun_out/s20_bench_c2.log 2>&1; echo "rc=$?"
grep '^{' gpurun_out/s20_bench_c2.log > gpurun_out/s20_c2.json; python -c "
import json
d=json.load(open('gpurun_out/s20_c2.json')); r=d['roofline']
print({k:d[k] for k in ['value','time_to_pcs_s','epochs_per_pca']}, d['e2e']['time_to_pcs_s'], {k:r[k] for k in ['frac','tc_g_ms_per_pca','tc_h_ms_per_pca','orth_ms_per_pca','small_stage_ms_per_pca']}, r['late_pass']['ms'], r['late_pass']['gbs'])" || tail -5 gpurun_out/s20_bench_c2.log
