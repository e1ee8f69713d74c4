"""GPU parity tests (-m gpu) of the downstream products: projection option 1
(U = G V S^-1, src/Projection.cpp:236-241) and the per-SNP products of run_selection
(src/Selection.cpp:16-38) against the numpy restatement; the decode they sit on is pinned to the
reference by the golden tests."""
import numpy as np
import pytest

from conftest import col_cos
from oracle import pcaone_oracle as orc
from pcaone_b200 import _lib, downstream, halko, synth

pytestmark = pytest.mark.gpu


def _packed(N, M, seed, miss=0.0):
    return np.concatenate([synth.pack_codes(c) for _, c in synth.balding_nichols_codes(N, M, k_pop=6, seed=seed, miss=miss)])


@pytest.mark.parametrize("prec", [_lib.PREC_FP64, _lib.PREC_INT8X3])
@pytest.mark.parametrize("N,M,k,miss,memory", [(400, 3000, 5, 0.0, 0.0), (257, 5001, 8, 0.03, 0.0), (300, 4000, 4, 0.0, 0.003)])
def test_projection_and_selection_products(prec, N, M, k, miss, memory):
    packed = _packed(N, M, N + M, miss)
    p = halko.Param(k=k, svd=1, precision=prec, memory=memory, maxp=4, tol=0.0)
    d = halko.FileBed(p, packed=packed, nsamples=N)
    d.prepare()
    op = halko.NormalRsvdOpData(d, p.k, p.oversamples)
    op.setFlags(False, True)
    op.computeUSV(p.maxp, p.tol)
    U, S, V = op.U.copy(), op.S.copy(), op.V.copy()
    od = orc.OracleData(packed, N)
    # projecting the SAME samples with their own loadings returns U: G V S^-1 = U (exact for the
    # converged subspace; here to the accuracy of 5 epochs)
    Up = downstream.run_projection(op, V, S)
    Uo = orc.projection_scores(od, V, S)
    assert np.abs(Up - Uo).max() <= 1e-12 * np.abs(Uo).max()
    assert col_cos(Up, U).min() > 0.999
    # projection with the allele frequencies of a "reference panel" (here: perturbed F)
    rng = np.random.default_rng(0)
    Fp = np.clip(od.F + rng.normal(0, 0.01, M), 0.02, 0.98)
    Up2 = downstream.run_projection(op, V, S, ref_F=Fp)
    od2 = orc.OracleData(packed, N)
    Uo2 = orc.projection_scores(od2, V, S, ref_F=Fp)
    assert np.abs(Up2 - Uo2).max() <= 1e-12 * np.abs(Uo2).max()
    # selection products with the original allele frequencies
    op.setF(od.F)
    E = S ** 2 / M
    Vs, nrm = downstream.run_selection_products(op, U, E) if memory == 0.0 else (None, None)
    if memory == 0.0:
        Vo, no = orc.selection_products(od, U, E)
        assert np.abs(Vs - Vo).max() <= 1e-12 * np.abs(Vo).max()
        assert np.abs(nrm - no).max() <= 1e-12 * no.max()
    else:
        with pytest.raises(RuntimeError, match="squared norms"):
            downstream.run_selection_products(op, U, E)
        assert np.abs(op.xtTimes(U) - od.block(0, M - 1, True).T @ U).max() < 1e-10
    op.close()


def test_products_on_dosages_and_errors():
    rng = np.random.default_rng(2)
    dos = np.clip(rng.normal(1.0, 0.6, (900, 120)), 0, 2).astype(np.float32)
    dos[rng.random(dos.shape) < 0.02] = np.nan
    p = halko.Param(k=4, svd=1, precision=_lib.PREC_FP64)
    d = halko.FileBgen(p, dos)
    d.prepare()
    op = halko.NormalRsvdOpData(d, p.k, p.oversamples)
    od = orc.OracleDosageData(d.dosages)
    od.F = op.F()
    A = rng.standard_normal((120, 4))
    B = rng.standard_normal((900, 6))
    op.setFlags(False, True)
    X = od.block(0, 899, True)
    out, nrm = op.xtTimes(A, want_sqnorm=True)
    assert np.abs(out - X.T @ A).max() <= 1e-12 * np.abs(X.T @ A).max()
    assert np.abs(nrm - (X * X).sum(0)).max() <= 1e-12 * nrm.max()
    assert np.abs(op.xTimes(B) - X @ B).max() <= 1e-12 * np.abs(X @ B).max()
    with pytest.raises(RuntimeError, match="ncols"):
        op.xTimes(rng.standard_normal((900, 15)))      # > k + oversamples = 14
    with pytest.raises(RuntimeError, match="project"):
        downstream.run_projection(op, B[:, :4], np.ones(4), project=2)
    op.close()
