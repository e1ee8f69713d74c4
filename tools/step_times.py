"""Per-epoch wall/device breakdown of the bench workload (debug aid, run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pcaone_b200 import halko, synth

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n, m, K, BANDS = 10000, int(sys.argv[2]) if len(sys.argv) > 2 else 1000000, 20, 64
packed = synth.torch_packed(n, m, k_pop=K + 4, seed=1, device="cuda:0", chunk=16384)
p = halko.Param(k=K, svd=2, bands=BANDS, maxp=20, tol=1e-4, no_shuffle=True, precision=prec)
d = halko.FileBed(p, packed=packed, nsamples=n)
op = halko.FancyRsvdOpData(d, p.k, p.oversamples)
op.setFlags(False, True)
op.enable_timing(True)
for rep in range(2):
    for i in range(7):
        op.sync(); op.timers(reset=True)
        t0 = time.perf_counter()
        op._chk(op.L.pcaone_compute_gandh(op.h, i)); op.sync()
        t1 = time.perf_counter()
        op._chk(op.L.pcaone_small_stage(op.h)); op.sync()
        t2 = time.perf_counter()
        tm = op.timers(reset=True)
        print(f"rep {rep} epoch {i}: gandh {1e3*(t1-t0):7.2f} ms  small {1e3*(t2-t1):6.2f} ms | g {tm.gemm_g_ms:6.2f} (tc {tm.tc_g_ms:6.2f}) h {tm.gemm_h_ms:6.2f} (tc {tm.tc_h_ms:6.2f}) orth {tm.orth_ms:6.2f} small {tm.small_ms:6.2f} launches {tm.kernel_launches} omega_updates {tm.omega_updates}")

# cost of the per-scope CUDA events themselves: the same 7 epochs with the library's timers on / off
for timing in (True, False, True, False):
    op.enable_timing(timing)
    op.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(op.L.pcaone_stream(op.h))
    e0.record(st)
    for i in range(7):
        op._chk(op.L.pcaone_compute_gandh(op.h, i))
        op._chk(op.L.pcaone_small_stage(op.h))
    e1.record(st)
    op.sync()
    torch.cuda.synchronize()
    print(f"7 epochs, library timers {'on ' if timing else 'off'}: {e0.elapsed_time(e1):7.2f} ms")
