#!/bin/bash
# final state of round 2 on one B200: the whole -m gpu suite, the default bench line, smoke(), configs[3] legs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s32_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s32_pytest.log
tail -5 gpurun_out/s32_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s32_bench_n1.json 2> gpurun_out/s32_bench_n1.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/s32_bench_n1.json')); r=d['roofline']
print({k:d[k] for k in ['value','time_to_pcs_s','gpu_launches']}, d['e2e']['time_to_pcs_s'], d['cpu_baseline']['value'], {k:r[k] for k in ['frac','frac_of_int8_peak','traffic','tc_g_ms_per_pca','tc_h_ms_per_pca','orth_ms_per_pca','small_stage_ms_per_pca']}, r['late_pass']['ms'], d['clocks'])"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
rm -f gpurun_out/s32_c4.jsonl
timeout 900 python tools/run_configs.py c4 --out gpurun_out/s32_c4.jsonl > gpurun_out/s32_c4.log 2>&1; echo "c4 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/s32_c4.jsonl'):
    d=json.loads(l); print({k:d[k] for k in d if k in ('workload','time_to_pcs_s','genotypes_per_s','em_run_s','late_pass_s','cores')}, d.get('late_update_pass',{}).get('ms'), d.get('first_update_pass',{}).get('ms'))
PY
