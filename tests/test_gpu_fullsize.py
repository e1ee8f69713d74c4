"""BASELINE.json configs at FULL size on one B200, checked through size-independent properties
(the numpy oracle cannot finish these sizes in seconds): bit-exact allele frequencies on a
sample of SNPs, orthonormal factors, the two precision routes against each other at the
north_star tolerance (eigenvalues <= 1e-6 relative, PCs |corr| >= 0.9999), the streamed path
against the resident one, and configs[0] against the compiled reference when it travelled."""
import numpy as np
import pytest

from conftest import assert_usv_close, col_cos
from oracle import pcaone_oracle as orc
from pcaone_b200 import halko, synth

pytestmark = pytest.mark.gpu


def _run(packed, N, cls, **kw):
    p = halko.Param(**kw)
    d = halko.FileBed(p, packed=packed, nsamples=N)
    d.prepare()
    op = cls(d, p.k, p.oversamples)
    op.setFlags(False, True)
    op.computeUSV(p.maxp, p.tol)
    out = dict(U=op.U.copy(), S=op.S.copy(), V=op.V.copy(), F=op.F().copy(), epochs=op.epochs, perm=d.perm)
    op.close()
    return out


def _orthonormal(Q, tol=1e-10):
    return np.abs(Q.T @ Q - np.eye(Q.shape[1])).max() <= tol


def test_configs1_winsvd_full_size_properties():
    """configs[1]: winSVD in memory, N = 10k x M = 1M, k = 20 (Halko.cpp:155-269)."""
    N, M, k = 10_000, 1_000_000, 20
    packed = synth.torch_packed(N, M, k_pop=k + 4, seed=1, device="cuda:0", chunk=16384)
    kw = dict(k=k, svd=2, bands=64, maxp=20, tol=1e-4, no_shuffle=True)
    r3 = _run(packed, N, halko.FancyRsvdOpData, precision=3, **kw)
    # allele frequencies: bit-exact against the oracle on every 997th SNP (FilePlink.cpp:37-60)
    idx = np.arange(0, M, 997)
    sub = packed[idx.tolist()].cpu().numpy()
    assert np.array_equal(r3["F"][idx], orc.allele_freq(orc.unpack_codes(sub, N)))
    assert r3["U"].shape == (N, k) and r3["V"].shape == (M, k)
    assert _orthonormal(r3["U"]) and _orthonormal(r3["V"])
    assert np.all(np.diff(r3["S"]) <= 0) and r3["S"][-1] > 0
    assert r3["epochs"] >= 7  # winSVD is forced until 2^pi >= bands (Halko.cpp:86-89)
    # the exact-integer tensor-core route against the FP64 DMMA route on the same input and Omega
    r0 = _run(packed, N, halko.FancyRsvdOpData, precision=0, **kw)
    assert np.array_equal(r0["F"], r3["F"])
    assert r0["epochs"] == r3["epochs"]
    assert_usv_close(r3["U"], r3["S"], r3["V"], r0["U"], r0["S"], r0["V"])
    # the planted structure is found: k_pop - 1 = 23 > k population axes, so the top-k eigenvalues
    # sit well above the bulk edge (1 + sqrt(N / M))^2 of a structureless standardised matrix
    ev = r3["S"] ** 2 / M
    assert ev[-1] > (1.0 + np.sqrt(N / M)) ** 2


def test_configs1_streamed_equals_resident():
    """Out-of-core blocks (-m, FilePlink.cpp:122-218) against the resident matrix at 10k x 250k:
    same permuted input order (the structured permutation is applied at read time), same result."""
    N, M, k = 10_000, 250_000, 20
    packed = synth.torch_packed(N, M, k_pop=k + 4, seed=3, device="cuda:0", chunk=16384)
    host = packed.cpu().numpy()
    kw = dict(k=k, svd=1, maxp=6, tol=0.0, precision=3)
    a = _run(packed, N, halko.NormalRsvdOpData, **kw)
    b = _run(host, N, halko.NormalRsvdOpData, memory=0.5, **kw)
    assert np.array_equal(a["F"], b["F"])
    assert_usv_close(b["U"], b["S"], b["V"], a["U"], a["S"], a["V"], eig_rtol=1e-9)


def test_configs0_full_size_vs_compiled_reference(tmp_path):
    """configs[0]: sSVD in memory, N = 2,504 x M = 100k, k = 10 — the case the reference runs on
    the CPU today — against the unmodified reference (oracle/_ref) on the same bed and Omega."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    import os
    N, M, k = 2504, 100_000, 10
    packed = synth.torch_packed(N, M, k_pop=k + 4, seed=1, device="cuda:0", chunk=16384).cpu().numpy()
    prefix = str(tmp_path / "c1")
    synth.write_bed_from_packed(prefix, packed, N)
    th = min(16, os.cpu_count() or 1)
    r = ref.Ref(f"PCAone -b {prefix} -k {k} -d 1 -o {tmp_path}/o -n {th}", threads=th)
    r.new_op()
    r.set_flags(False, True)
    Ur, Sr, Vr = r.compute_usv(20, 1e-4)
    Fr = r.F()
    r.close()
    for prec in (0, 3):
        a = _run(packed, N, halko.NormalRsvdOpData, k=k, svd=1, maxp=20, tol=1e-4, precision=prec)
        assert np.array_equal(a["F"], Fr)
        assert_usv_close(a["U"], a["S"], a["V"], Ur, Sr, Vr)


def test_single_pass_qr_of_g_matches_cholesky_qr2(tmp_path):
    """QR(G) of the dense stage drops its second Cholesky pass when cond_F(G)^2 <= 1e5
    (orth_fused.cuh, P3); PCAONE_QR2_ALWAYS=1 keeps CholeskyQR2. Same input, both settings (the
    switch is read once per process, hence the subprocesses): identical up to rounding."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from pcaone_b200 import halko, synth\n"
        "N, M, k = 3000, 60000, 10\n"
        "packed = synth.torch_packed(N, M, k_pop=k + 4, seed=11, device='cuda:0', chunk=16384)\n"
        "p = halko.Param(k=k, svd=2, bands=16, maxp=8, tol=0.0, no_shuffle=True, precision=3)\n"
        "d = halko.FileBed(p, packed=packed, nsamples=N); d.prepare()\n"
        "op = halko.FancyRsvdOpData(d, p.k, p.oversamples); op.setFlags(False, True)\n"
        "op.computeUSV(p.maxp, p.tol)\n"
        "np.savez(sys.argv[1], U=op.U, S=op.S, V=op.V)\n" % root)
    outs = []
    for flag in ("0", "1"):
        out = str(tmp_path / f"r{flag}.npz")
        env = dict(os.environ, PCAONE_QR2_ALWAYS=flag)
        r = subprocess.run([sys.executable, "-c", code, out], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        outs.append(np.load(out))
    a, b = outs
    assert np.max(np.abs(a["S"] ** 2 - b["S"] ** 2) / b["S"] ** 2) <= 1e-11
    assert col_cos(a["U"], b["U"]).min() >= 1 - 1e-10 and col_cos(a["V"], b["V"]).min() >= 1 - 1e-10
    assert _orthonormal(a["V"], 1e-11) and _orthonormal(b["V"], 1e-11)
