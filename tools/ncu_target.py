"""One winSVD run (epochs 0..6) of the bench workload, for ncu captures of specific launches:
the last k_tc_gemm launches (epoch 6) are the big merged half-matrix ranges."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcaone_b200 import halko, synth

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n, m, K, BANDS = 10000, 1000000, 20, 64
packed = synth.torch_packed(n, m, k_pop=K + 4, seed=1, device="cuda:0", chunk=16384)
p = halko.Param(k=K, svd=2, bands=BANDS, maxp=20, tol=1e-4, no_shuffle=True, precision=prec)
d = halko.FileBed(p, packed=packed, nsamples=n)
op = halko.FancyRsvdOpData(d, p.k, p.oversamples)
op.setFlags(False, True)
for i in range(7):
    op._chk(op.L.pcaone_compute_gandh(op.h, i))
    op._chk(op.L.pcaone_small_stage(op.h))
op.sync()
print("done")
