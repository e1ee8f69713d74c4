"""SM clock / power under a sustained loop of late-epoch passes (the big k_tc_gemm launches)."""
import sys, time, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pynvml
from pcaone_b200 import halko, synth

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
n, m, K, BANDS = 10000, 1000000, 20, 64
packed = synth.torch_packed(n, m, k_pop=K + 4, seed=1, device="cuda:0", chunk=16384)
p = halko.Param(k=K, svd=2, bands=BANDS, maxp=20, tol=1e-4, no_shuffle=True, precision=prec)
d = halko.FileBed(p, packed=packed, nsamples=n)
op = halko.FancyRsvdOpData(d, p.k, p.oversamples)
op.setFlags(False, True)
for i in range(7):
    op._chk(op.L.pcaone_compute_gandh(op.h, i))
op.sync()
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []
stop = threading.Event()
def loop():
    while not stop.is_set():
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        stop.wait(0.02)
th = threading.Thread(target=loop, daemon=True); th.start()
op.enable_timing(True); op.timers(reset=True)
t0 = time.perf_counter(); reps = 0
while time.perf_counter() - t0 < secs:
    for _ in range(10):
        op._chk(op.L.pcaone_compute_gandh(op.h, 7))
    op.sync(); reps += 10
el = time.perf_counter() - t0
stop.set(); th.join()
tm = op.timers(reset=True)
sm = [s[0] for s in samples]; pw = [s[1] for s in samples]
rs = 0
for s in samples: rs |= s[2]
print(f"prec={prec} dbg={os.environ.get('PCAONE_TC_DBG','0')}: {reps} passes, {1e3*el/reps:.3f} ms/pass | tc g {tm.tc_g_ms/reps:.3f} h {tm.tc_h_ms/reps:.3f} | "
      f"gemm g {tm.gemm_g_ms/reps:.3f} h {tm.gemm_h_ms/reps:.3f} orth {tm.orth_ms/reps:.3f} | SM MHz median {np.median(sm):.0f} min {min(sm)} max {max(sm)} | power W median {np.median(pw):.0f} max {max(pw):.0f} | reasons 0x{rs:x}")
