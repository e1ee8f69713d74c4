"""Effect of the single-pass QR(G) gate (k_orth_fused skip2) on configs[1]: run with and without
PCAONE_QR2_ALWAYS=1 and compare (debug aid, run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pcaone_b200 import halko, synth

n, m, K = 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, 20
packed = synth.torch_packed(n, m, k_pop=K + 4, seed=1, device="cuda:0", chunk=16384)
p = halko.Param(k=K, svd=2, bands=64, maxp=20, tol=1e-4, no_shuffle=True, precision=3)
d = halko.FileBed(p, packed=packed, nsamples=n)
op = halko.FancyRsvdOpData(d, p.k, p.oversamples)
op.setFlags(False, True)
op.computeUSV(p.maxp, p.tol)
V, U, S = op.V, op.U, op.S
print("QR2_ALWAYS", os.environ.get("PCAONE_QR2_ALWAYS", "0"), "epochs", op.epochs,
      "|V'V-I|", np.abs(V.T @ V - np.eye(K)).max(), "|U'U-I|", np.abs(U.T @ U - np.eye(K)).max())
np.save(f"/tmp/S_{os.environ.get('PCAONE_QR2_ALWAYS', '0')}.npy", S)
if os.path.exists("/tmp/S_0.npy") and os.path.exists("/tmp/S_1.npy"):
    a, b = np.load("/tmp/S_0.npy"), np.load("/tmp/S_1.npy")
    print("eigenvalue rel diff single vs two-pass:", np.max(np.abs(a ** 2 - b ** 2) / b ** 2))
