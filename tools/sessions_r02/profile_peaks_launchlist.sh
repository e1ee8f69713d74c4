#!/bin/bash
# round-2 profiling session (one B200): peaks, ncu launch list, ncu --set full of the two top kernels, bench + reference arm
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/probe_umma tools/probe_umma.cu 2>/dev/null
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/probe_dmma tools/probe_dmma.cu 2>/dev/null
gpurun_out/probe_umma > gpurun_out/s9_probe_umma.txt 2>&1; tail -3 gpurun_out/s9_probe_umma.txt
gpurun_out/probe_dmma > gpurun_out/s9_probe_dmma.txt 2>&1; tail -1 gpurun_out/s9_probe_dmma.txt
rm -f gpurun_out/probe_umma gpurun_out/probe_dmma
# the bench itself (value / e2e / cpu baseline) and the reference arm
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s9_bench_n1.json 2> gpurun_out/s9_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/s9_bench_n1.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/s9_bench_ref.json 2> gpurun_out/s9_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/s9_bench_ref.json | head -c 1500
# every launch of our kernels with its device time
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 20000 --csv --log-file gpurun_out/s9_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/s9_ncu_list.log 2>&1; echo "ncu list rc=$?"
wc -l gpurun_out/s9_launches.csv
# full captures: the dominant kernel on the merged late-epoch ranges, and one Omega update at 500k rows
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 258 -c 4 -f -o gpurun_out/s9_prof_tc \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/s9_ncu_tc.log 2>&1; echo "ncu tc rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_orth_fused -s 128 -c 1 -f -o gpurun_out/s9_prof_orth \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/s9_ncu_orth.log 2>&1; echo "ncu orth rc=$?"
ls -la gpurun_out/*.ncu-rep
gzip -f gpurun_out/s9_launches.csv
