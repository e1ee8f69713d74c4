#!/bin/bash
# configs[3] (EMU, 50k x 500k, 10 % missing): int8 route + FP64 correction vs update passes on the FP64 DMMA kernels
mkdir -p gpurun_out
rm -f gpurun_out/s23_c4.jsonl
timeout 900 python tools/run_configs.py c4 --out gpurun_out/s23_c4.jsonl > gpurun_out/s23_c4.log 2>&1; echo "rc=$?"
tail -5 gpurun_out/s23_c4.log
python - <<'PY'
import json
for l in open('gpurun_out/s23_c4.jsonl'):
    d=json.loads(l); print(json.dumps({k:d[k] for k in d if k not in ('eigvals_top5',)}, indent=None)[:1500])
PY
