#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s16_pytest.log
tail -4 gpurun_out/s16_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s16_bench_n1.json 2> gpurun_out/s16_bench_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/s16_bench_n1.json')); r=d['roofline']
print({k:d[k] for k in ['value','time_to_pcs_s','gpu_launches']}, d['e2e']['time_to_pcs_s'], d['cpu_baseline']['value'], {k:r[k] for k in ['frac','frac_of_int8_peak','traffic','tc_g_ms_per_pca','tc_h_ms_per_pca','orth_ms_per_pca','small_stage_ms_per_pca']}, r['late_pass']['ms'], d['clocks'])"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
