"""Why is bench.py's timed region slower than tools/step_times.py? Toggle the suspects."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pcaone_b200 import halko, synth
import bench

n, m, K, BANDS = 10000, 1000000, 20, 64
packed = synth.torch_packed(n, m, k_pop=K + 4, seed=1, device="cuda:0", chunk=16384)
p = halko.Param(k=K, svd=2, bands=BANDS, maxp=20, tol=1e-4, no_shuffle=True, precision=3)
d = halko.FileBed(p, packed=packed, nsamples=n)
op = halko.FancyRsvdOpData(d, p.k, p.oversamples)
op.setFlags(False, True)
stream = torch.cuda.ExternalStream(op.L.pcaone_stream(op.h))

def region(timing, sampler_on, steps=7):
    op.enable_timing(timing); op.timers(reset=True)
    s = bench.ClockSampler(0)
    if sampler_on: s.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(steps):
        op._chk(op.L.pcaone_compute_gandh(op.h, i)); op._chk(op.L.pcaone_small_stage(op.h))
    t1 = time.perf_counter()
    e1.record(stream); op.sync(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    if sampler_on: s.stop()
    op.timers(reset=True)
    return e0.elapsed_time(e1), 1e3 * (t1 - t0), 1e3 * (t2 - t0)

for w in range(3):
    op._chk(op.L.pcaone_compute_gandh(op.h, w)); op._chk(op.L.pcaone_small_stage(op.h))
op.sync()
for rep in range(3):
    for timing in (True, False):
        for samp in (True, False):
            dev, issue, wall = region(timing, samp)
            print(f"rep {rep} timing={timing} sampler={samp}: device {dev:7.2f} ms  host-issue {issue:7.2f} ms  wall {wall:7.2f} ms  ({dev/7:.2f} ms/step)")
