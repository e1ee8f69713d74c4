#!/bin/bash
mkdir -p gpurun_out
bash tools/sessions_r02/emu_tc_dev.sh
if grep -q "pytest rc=0" gpurun_out/s22_emu.log; then
  rm -f gpurun_out/s23_c4.jsonl
  timeout 900 python tools/run_configs.py c4 --out gpurun_out/s23_c4.jsonl > gpurun_out/s23_c4.log 2>&1; echo "rc=$?"
  python - <<'PY'
import json
for l in open('gpurun_out/s23_c4.jsonl'):
    d=json.loads(l); print(d['workload'][-50:], d['time_to_pcs_s'], d['late_update_pass'], d['late_plain_pass']['ms'])
PY
fi
