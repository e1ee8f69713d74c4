#!/bin/bash
# A/B of the sliced sample axis in k_emu_fix_g (PCAONE_EMU_SPLIT) on configs[3], early (64 windows) and late epochs
mkdir -p gpurun_out
for sp in 1 0; do
  rm -f gpurun_out/s28_c4_split$sp.jsonl
  PCAONE_EMU_SPLIT=$sp timeout 600 python tools/run_configs.py c4 --c4-legs 1 --out gpurun_out/s28_c4_split$sp.jsonl > gpurun_out/s28_c4_split$sp.log 2>&1; echo "split=$sp rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/s28_c4_split$sp.jsonl'):
    d=json.loads(l); print('split=$sp', d['time_to_pcs_s'], 'late', d['late_update_pass']['ms'], d['late_update_pass']['emu_fix_ms'], 'first', d['first_update_pass']['ms'], d['first_update_pass']['emu_fix_ms'], d['first_update_pass']['gemm_g_ms'], d['first_update_pass']['gemm_h_ms'], d['first_update_pass']['tc_ranges'], 'plain first', d['first_plain_pass']['ms'])
PY
done
