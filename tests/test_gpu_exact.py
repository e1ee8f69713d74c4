"""-m gpu: the exact PCA of `--svd 3` (Main.cpp:180-217) — sample covariance GEMM + symmetric eigen-decomposition on
the device — against the unmodified reference's own Data / Eigen path (oracle/_ref) and numpy."""
import numpy as np
import pytest

from conftest import col_cos
from oracle import pcaone_oracle as orc
from pcaone_b200 import _lib, halko, synth

pytestmark = pytest.mark.gpu


def _op(packed, N, k, prec=_lib.PREC_FP64):
    p = halko.Param(k=k, svd=1, precision=prec)
    d = halko.FileBed(p, packed=packed, nsamples=N)
    d.prepare()
    return halko.NormalRsvdOpData(d, p.k, p.oversamples)


@pytest.mark.parametrize("n,kind", [(300, "psd"), (257, "psd"), (64, "indefinite"), (1, "psd"), (515, "lowrank")])
def test_sym_svd_vs_numpy(n, kind):
    rng = np.random.default_rng(n)
    B = rng.standard_normal((n, 3 if kind == "lowrank" else n + 5))
    A = B @ B.T
    if kind == "indefinite":
        A = A - 0.5 * np.trace(A) / n * np.eye(n)
    packed = np.concatenate([synth.pack_codes(c) for _, c in synth.balding_nichols_codes(40, 200, k_pop=3, seed=1)])
    op = _op(packed, 40, 2)
    U, S = op.symSVD(A)
    w = np.linalg.eigvalsh(A)
    want = np.sort(np.abs(w))[::-1]
    assert np.all(np.diff(S) <= 0)
    assert np.abs(S - want).max() <= 1e-12 * want[0]
    r = int((S > 1e-10 * S[0]).sum())
    assert np.abs(U[:, :r].T @ U[:, :r] - np.eye(r)).max() < 1e-12
    # A u = +- s u, column by column
    AU = A @ U[:, :r]
    sg = np.sign(np.sum(AU * U[:, :r], axis=0))
    assert np.abs(AU - U[:, :r] * (S[:r] * sg)).max() <= 1e-11 * S[0]
    assert op.jacobi_sweeps < 40
    op.close()


@pytest.mark.parametrize("prec", [_lib.PREC_FP64, _lib.PREC_INT8X3])
def test_sample_covariance_vs_numpy(prec):
    N, M = 333, 4100
    packed = np.concatenate([synth.pack_codes(c) for _, c in
                             synth.balding_nichols_codes(N, M, k_pop=5, seed=9, miss=0.02)])
    op = _op(packed, N, 4, prec)
    od = orc.OracleData(packed, N)
    for standardize in (False, True):
        op.setFlags(False, standardize)
        K = op.sampleCovariance()
        X = od.block(0, M - 1, standardize)
        want = X @ X.T
        assert np.abs(K - want).max() <= 1e-11 * np.abs(want).max()   # FP64 kernels whatever the context's precision
    op.close()


def test_exact_pca_vs_reference(tmp_path):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    N, M, k = 400, 3000, 10
    prefix = str(tmp_path / "s")
    packed = synth.write_bed(prefix, N, M, k_pop=6, seed=41)
    r = ref.Ref(f"PCAone -b {prefix} -k {k} -d 1 -o {tmp_path}/r -n 8", threads=8)
    Ur, Sr, Vr, Er = r.full_pca(k)
    r.close()
    op = _op(packed, N, k, _lib.PREC_INT8X3)
    E = op.exactPCA()
    assert np.abs(E - Er).max() <= 1e-11 * Er[0]
    assert np.abs(op.S - Sr).max() <= 1e-11 * Sr[0]
    # the 5 population PCs are separated; the rest of the top-10 sit in the noise bulk, where the eigenvectors
    # are still well defined to ~1e-9 at this size: compare all of them including the flip_UV sign
    assert col_cos(op.U, Ur).min() > 1 - 1e-9 and col_cos(op.V, Vr).min() > 1 - 1e-9
    assert np.abs(op.U[:, :5] - Ur[:, :5]).max() < 1e-9
    assert np.abs(op.V[:, :5] - Vr[:, :5]).max() < 1e-9
    assert np.abs(op.U.T @ op.U - np.eye(k)).max() < 1e-12
    op.close()
