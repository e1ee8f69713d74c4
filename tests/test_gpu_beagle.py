"""GPU parity tests (-m gpu) of the Beagle genotype-likelihood / PCAngsd front-end:
allele-frequency EM, expected genotypes, and the EM loop against the UNMODIFIED reference run on a
synthetic beagle.gz (tests/golden/pcangsd_small.npz) and against the numpy restatement."""
import numpy as np
import pytest

from conftest import assert_usv_close, col_cos, golden
from oracle import pcaone_oracle as orc
from pcaone_b200 import _lib, halko

pytestmark = pytest.mark.gpu


def _gl(N, M, seed, depth=2.0):
    rng = np.random.default_rng(seed)
    pop = rng.integers(0, 3, N)
    pf = np.clip(rng.uniform(0.1, 0.9, (M, 1)) + 0.15 * rng.standard_normal((M, 3)), 0.05, 0.95)[:, pop]   # (M, N)
    gt = rng.binomial(2, pf)
    d = rng.poisson(depth, gt.shape)
    alt = rng.binomial(d, np.clip(gt / 2.0, 0.01, 0.99))
    lik = [np.clip(q / 2.0, 0.01, 0.99) ** alt * (1 - np.clip(q / 2.0, 0.01, 0.99)) ** (d - alt) for q in (0, 1, 2)]
    L = np.stack(lik, -1)
    L = L / L.sum(-1, keepdims=True)
    P = np.zeros((2 * N, M))
    P[0::2] = L[:, :, 0].T
    P[1::2] = L[:, :, 1].T
    return P


def test_pcangsd_vs_reference_golden():
    g = golden("pcangsd_small")
    k = int(g["k"])
    p = halko.Param(k=k, svd=1, maxp=int(g["maxp"]), tol=0.0, maxiter=int(g["maxiter"]), precision=_lib.PREC_FP64)
    d = halko.FileBeagle(p, g["P"])
    d.prepare()
    assert d.maf_iters == int(g["maxiter"])
    assert np.abs(d.F - g["F"]).max() < 1e-13
    op = halko.NormalRsvdOpData(d, p.k, p.oversamples)
    op.setOmg(g["omega"])
    E0 = op.read_block(0, d.nsnps - 1, False)
    assert np.abs(E0 - g["E0"]).max() < 1e-13
    iters = op.runEM()
    assert iters == int(g["iters"])
    print("S rel", np.max(np.abs(op.S - g["S"]) / g["S"]), "cos U", col_cos(op.U, g["U"]).min())
    assert_usv_close(op.U, op.S, op.V, g["U"], g["S"], g["V"])
    # E of an update pass: individual allele frequencies from the final U, S, V
    op.setUSV(op.U, op.S, op.V)
    E1 = op.read_block(10, 200, False, update=True)
    Eo = orc.gl_expected(g["P"][:, 10:201], g["F"][10:201], (op.U, op.S, op.V[10:201]))
    assert np.abs(E1 - Eo).max() < 1e-12
    op.close()


@pytest.mark.parametrize("N,M,k,svd", [(300, 4000, 3, 1), (257, 3001, 4, 2)])
def test_pcangsd_vs_numpy_oracle(N, M, k, svd):
    P = _gl(N, M, N + M)
    p = halko.Param(k=k, svd=svd, bands=8, maxp=7 if svd == 2 else 4, tol=0.0, maxiter=3, no_shuffle=True,
                    precision=_lib.PREC_FP64)
    d = halko.FileBeagle(p, P)
    d.prepare()
    F, it = orc.em_maf_with_gl(P, p.maxiter, 1e-6)
    assert it == d.maf_iters and np.abs(d.F - F).max() < 1e-13
    op = halko.run_pca_with_halko(d, p)
    od = orc.OracleGLData(P, d.F)
    windows = orc.incore_windows(M, 8)[1] if svd == 2 else None
    oo = orc.OracleRsvd(od, k, winsvd=svd == 2, bands=8, omega=op.Omg, windows=windows)
    U, S, V, iters = orc.run_emu(oo, p.maxp, 0.0, maxiter=3, tolem=p.tolem, final_standardize=False)
    assert iters == op.em_iters
    print((N, M, k, svd), "eig rel", np.max(np.abs(op.S ** 2 - S ** 2) / S ** 2), "cos", col_cos(op.U, U).min())
    assert_usv_close(op.U, op.S, op.V, U, S, V)
    op.close()


def test_beagle_maf_filter_and_errors():
    P = _gl(120, 500, 1)
    p = halko.Param(k=3, svd=1, maf=0.2, maxiter=20, precision=_lib.PREC_FP64)
    d = halko.FileBeagle(p, P)
    d.prepare()
    F, _ = orc.em_maf_with_gl(P, 20, 1e-6)
    keep = np.flatnonzero(np.minimum(F, 1 - F) > 0.2)
    assert np.array_equal(d.keep, keep) and d.nsnps == len(keep) and 0 < len(keep) < 500
    with pytest.raises(RuntimeError):
        halko.FileBeagle(halko.Param(k=3, svd=1, memory=0.001, precision=_lib.PREC_FP64), P)
    with pytest.raises(RuntimeError):
        halko.FileBeagle(halko.Param(k=3, svd=1, precision=_lib.PREC_INT8X3), P)


def test_pcangsd_grm_vs_reference_golden():
    """The GRM step after the EM (Halko.cpp:320-334): pcangsd_standardize_E + N x N covariance with the Dc diagonal on
    the device, then the symmetric SVD on the device, against the unmodified reference (pcangsd_grm.npz)."""
    g, q = golden("pcangsd_small"), golden("pcangsd_grm")
    k = int(g["k"])
    p = halko.Param(k=k, svd=1, maxp=int(g["maxp"]), tol=0.0, maxiter=int(g["maxiter"]), precision=_lib.PREC_FP64)
    d = halko.FileBeagle(p, g["P"])
    d.prepare()
    op = halko.NormalRsvdOpData(d, p.k, p.oversamples)
    op.setOmg(g["omega"])
    op.runEM()
    # (1) from the reference's own U, S, V: the step in isolation
    op.setUSV(q["U"], q["S"], q["V"])
    Cm, Dc = op.grm()
    assert np.abs(Dc - q["Dc"]).max() <= 1e-12 * np.abs(q["Dc"]).max()
    assert np.abs(Cm - q["C"]).max() <= 1e-12 * np.abs(q["C"]).max()
    U2, S2 = op.symSVD(Cm)
    assert np.abs(S2 - q["S2"]).max() <= 1e-12 * q["S2"][0]
    assert col_cos(U2[:, :4], q["U2"][:, :4]).min() > 1 - 1e-10      # the separated leading vectors, up to sign
    assert np.abs(U2.T @ U2 - np.eye(U2.shape[0])).max() < 1e-12
    sg = np.sign(np.sum((q["C"] @ U2) * U2, axis=0))
    assert np.abs(q["C"] - (U2 * (S2 * sg)) @ U2.T).max() <= 1e-12 * S2[0]
    # (2) end to end from this library's own EM result
    op.runEM()
    C2, _ = op.grm()
    assert np.abs(C2 - q["C"]).max() <= 1e-9 * np.abs(q["C"]).max()
    op.close()
