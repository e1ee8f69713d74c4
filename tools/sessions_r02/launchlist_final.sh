#!/bin/bash
# ncu launch list of the bench command on the final code state (every launch of our kernels with its device time)
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 20000 --csv --log-file gpurun_out/s36_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/s36_ncu_list.log 2>&1; echo "ncu list rc=$?"
wc -l gpurun_out/s36_launches.csv
gzip -f gpurun_out/s36_launches.csv
