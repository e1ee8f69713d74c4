"""GPU parity tests (-m gpu) of the BGEN-dosage front-end: float dosages with NaN
for missing calls, decode + mean imputation + centring / scaling fused into the operand load of
the FP64 tensor-core products (csrc/dense_gemm.cuh k_dos_g / k_dos_h).

Pins: (1) the numpy restatement of FileBgen.cpp:15-110 (oracle.dense_from_dosage); (2) on HARD-CALL
dosages (d = 2 x half-dosage of a bed code) the dosage path must reproduce the PLINK-bed path, whose
results are pinned to the unmodified reference by tests/golden: same F bit for bit, same decoded
block, same U, S, V. The reference's own BGEN reader needs its bundled bgen library and a .bgen
container and is not compiled here, so fractional dosages are pinned by (1) only."""
import numpy as np
import pytest

from conftest import assert_usv_close, col_cos, golden
from oracle import pcaone_oracle as orc
from pcaone_b200 import _lib, halko, synth

pytestmark = pytest.mark.gpu


def _dosages(N, M, seed, miss=0.03, hard=False):
    rng = np.random.default_rng(seed)
    codes = np.concatenate([c for _, c in synth.balding_nichols_codes(N, M, k_pop=6, seed=seed)])   # (M, N)
    d = np.array([2.0, np.nan, 1.0, 0.0], dtype=np.float32)[codes]      # BED2GENO x 2 (Common.hpp:58-59)
    if not hard:
        d = np.clip(d + rng.normal(0, 0.08, d.shape).astype(np.float32), 0, 2).astype(np.float32)
    d[rng.random(d.shape) < miss] = np.nan
    return d, codes


def _op(dos, **kw):
    omega = kw.pop("omega", None)
    p = halko.Param(precision=_lib.PREC_FP64, **kw)
    d = halko.FileBgen(p, dos)
    d.prepare()
    cls = halko.FancyRsvdOpData if p.svd == 2 else halko.NormalRsvdOpData
    op = cls(d, p.k, p.oversamples)
    if omega is not None:
        op.setOmg(omega)
    return op, d, p


@pytest.mark.parametrize("N,M", [(203, 1500), (640, 3001), (1000, 777)])
def test_dosage_af_and_decode(N, M):
    dos, _ = _dosages(N, M, 5 + N)
    op, d, p = _op(dos, k=4, svd=1)
    od = orc.OracleDosageData(d.dosages)
    F = op.F()
    assert np.max(np.abs(F - od.F) / od.F) < 1e-14           # warp-tree vs sequential double sum
    assert op.missing_count() == int(np.isnan(d.dosages).sum())
    for std in (False, True):
        X = op.read_block(100, 611, std)
        od.F = F                                              # decode against the device's own F: bit-exact
        Xo = orc.dense_from_dosage(d.dosages[100:612], F[100:612], std)
        tol = 0 if not std else 4e-16                          # s_j folded into one factor: <= 1 ulp
        assert np.max(np.abs(X - Xo) / np.maximum(np.abs(Xo), 1e-300)) <= tol
    op.close()


@pytest.mark.parametrize("svd,bands", [(1, 64), (2, 8)])
def test_dosage_gandh_and_usv_vs_numpy_oracle(svd, bands):
    N, M, k = 500, 4096, 5
    dos, _ = _dosages(N, M, 77)
    maxp = 7 if svd == 2 else 4
    op, d, p = _op(dos, k=k, svd=svd, bands=bands, maxp=maxp, tol=0.0)
    op.setFlags(False, True)
    op.computeUSV(p.maxp, p.tol)
    od = orc.OracleDosageData(d.dosages)
    od.F = op.F() if d.perm is None else None
    windows = None
    if svd == 2:
        od = orc.OracleDosageData(d.dosages)
        od.permute(d.perm)
        _, windows = orc.incore_windows(M, bands)
    oo = orc.OracleRsvd(od, k, winsvd=svd == 2, bands=bands, omega=op.Omg, windows=windows)
    oo.set_flags(False, True)
    U, S, V = oo.compute_usv(maxp, 0.0)
    assert op.epochs == oo.epochs
    print("svd", svd, "eig rel", np.max(np.abs(op.S ** 2 - S ** 2) / S ** 2), "cos", col_cos(op.U, U).min())
    assert_usv_close(op.U, op.S, op.V, U, S, V)
    op.close()


def test_hard_call_dosages_equal_the_bed_path():
    """d = 2 x BED2GENO of the golden bed: the dosage path must give the reference's bed results."""
    g = golden("ssvd_small")
    N, k = int(g["N"]), int(g["k"])
    codes = orc.unpack_codes(g["packed"], N)
    dos = np.array([2.0, np.nan, 1.0, 0.0], dtype=np.float32)[codes]
    op, d, p = _op(dos, k=k, svd=1, maxp=int(g["maxp"]), tol=0.0, omega=g["omega"])
    assert d.keep is None                                   # no monomorphic variant in the golden bed
    assert np.array_equal(op.F(), g["F"])                   # half-dosage sums are exact: bit-identical F
    op.setFlags(False, True)
    op.computeUSV(p.maxp, p.tol)
    assert_usv_close(op.U, op.S, op.V, g["U"], g["S"], g["V"])
    print("vs reference bed run: S rel", np.max(np.abs(op.S - g["S"]) / g["S"]))
    op.close()


def test_dosage_maf_filter_and_errors():
    dos, _ = _dosages(120, 400, 3)
    dos[7] = 0.0                                             # af == 0 -> dropped even at --maf 0 (FileBgen.cpp:42)
    dos[9] = np.nan                                          # no call at all -> af = 0 -> dropped
    p = halko.Param(k=3, svd=1, precision=_lib.PREC_FP64)
    d = halko.FileBgen(p, dos)
    d.prepare()
    assert d.nsnps == 398 and 7 not in d.keep and 9 not in d.keep
    with pytest.raises(RuntimeError):
        halko.FileBgen(halko.Param(k=3, svd=1, precision=_lib.PREC_INT8X3), dos)
    pe = halko.Param(k=3, svd=1, precision=_lib.PREC_FP64, emu=True)
    de = halko.FileBgen(pe, dos)
    de.prepare()
    with pytest.raises(RuntimeError, match="emu"):
        halko.NormalRsvdOpData(de, 3, pe.oversamples)
