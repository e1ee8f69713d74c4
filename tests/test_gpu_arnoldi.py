"""GPU parity tests (-m gpu) of the IRAM operator: pcaone_perform_op <->
ArnoldiOpData::perform_op (src/Arnoldi.cpp:18-46), y = sum_b G_b (G_b^T x)."""
import numpy as np
import pytest

from conftest import golden
from oracle import pcaone_oracle as orc
from pcaone_b200 import _lib, halko, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", [_lib.PREC_FP64, _lib.PREC_INT8X4])
def test_perform_op_vs_reference_golden(prec):
    g, a = golden("ssvd_small"), golden("arnoldi_op")
    N = int(g["N"])
    p = halko.Param(k=int(g["k"]), svd=0, memory=float(a["memory"]), precision=prec)   # --svd 0: the IRAM block plan
    d = halko.FileBed(p, packed=g["packed"], nsamples=N)
    d.prepare()
    assert np.array_equal(d.start, a["start"]) and np.array_equal(d.stop, a["stop"])
    op = halko.ArnoldiOpData(d)
    tol = 1e-12 if prec == _lib.PREC_FP64 else 2e-7      # int8x4: x is rounded once to 31 bits
    for std, key in ((True, "y_std"), (False, "y_raw")):
        op.setFlags(False, std)
        y = op.perform_op(a["x"])
        assert np.abs(y - a[key]).max() <= tol * np.abs(a[key]).max(), (prec, std)
    assert op.nops == 3
    op.close()


@pytest.mark.parametrize("N,M,miss", [(500, 3000, 0.0), (333, 5001, 0.05)])
def test_perform_op_vs_numpy_oracle_and_power_iteration(N, M, miss):
    packed = np.concatenate([synth.pack_codes(c) for _, c in synth.balding_nichols_codes(N, M, k_pop=5, seed=N, miss=miss)])
    p = halko.Param(k=1, svd=1, precision=_lib.PREC_FP64)
    d = halko.FileBed(p, packed=packed, nsamples=N)
    d.prepare()
    op = halko.ArnoldiOpData(d)
    op.setFlags(False, True)
    od = orc.OracleData(packed, N)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(N)
    y = op.perform_op(x)
    yo = orc.perform_op(od, x, None, True)
    assert np.abs(y - yo).max() <= 1e-12 * np.abs(yo).max()
    # the operator is symmetric PSD: a few power iterations converge to the top eigenvalue of X X^T
    X = od.block(0, M - 1, True)
    lam = np.linalg.eigvalsh(X @ X.T)[-1]
    for _ in range(60):
        x = op.perform_op(x)
        x /= np.linalg.norm(x)
    est = x @ op.perform_op(x)
    assert abs(est - lam) / lam < 1e-3
    op.close()


def test_perform_op_on_dosages():
    rng = np.random.default_rng(0)
    dos = np.clip(rng.normal(1.0, 0.6, (800, 150)), 0, 2).astype(np.float32)
    dos[rng.random(dos.shape) < 0.02] = np.nan
    p = halko.Param(k=1, svd=1, precision=_lib.PREC_FP64)
    d = halko.FileBgen(p, dos)
    d.prepare()
    op = halko.ArnoldiOpData(d)
    op.setFlags(False, True)
    od = orc.OracleDosageData(d.dosages)
    od.F = op.F()
    x = rng.standard_normal(150)
    y, yo = op.perform_op(x), orc.perform_op(od, x, None, True)
    assert np.abs(y - yo).max() <= 1e-12 * np.abs(yo).max()
    op.close()
