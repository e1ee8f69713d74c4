#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/s29_c4.jsonl
timeout 600 python tools/run_configs.py c4 --c4-legs 1 --out gpurun_out/s29_c4.jsonl > gpurun_out/s29_c4.log 2>&1; echo "rc=$?"
python - <<PY
import json
for l in open('gpurun_out/s29_c4.jsonl'):
    d=json.loads(l); print(d['time_to_pcs_s'], 'late', d['late_update_pass']['ms'], d['late_update_pass']['emu_fix_ms'], 'first', d['first_update_pass']['ms'], d['first_update_pass']['emu_fix_ms'])
PY
