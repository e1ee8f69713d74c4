"""GPU parity tests (-m gpu) of the BGEN-dosage front-end: float dosages with NaN
for missing calls, decode + mean imputation + centring / scaling fused into the operand load of
the FP64 tensor-core products (csrc/dense_gemm.cuh k_dos_g / k_dos_h).

Pins: (1) the numpy restatement of FileBgen.cpp:15-110 (oracle.dense_from_dosage); (2) on HARD-CALL
dosages (d = 2 x half-dosage of a bed code) the dosage path must reproduce the PLINK-bed path, whose
results are pinned to the unmodified reference by tests/golden: same F bit for bit, same decoded
block, same U, S, V; (3) FRACTIONAL dosages against the unmodified reference reading a real .bgen container:
the reference's FileBgen + its vendored bgen library are compiled into oracle/_ref, the file is written with that
library's own writer, and the dosages handed to the device are the ones its reader returns
(test_fractional_dosages_vs_reference_bgen_reader)."""
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


@pytest.mark.parametrize("bit_depth", [8, 16])
def test_fractional_dosages_vs_reference_bgen_reader(tmp_path, bit_depth):
    """A layout-2 .bgen with fractional genotype probabilities and missing samples, read by the unmodified reference
    (FileBgen::read_all, FileBgen.cpp:15-72 on the vendored bgen reader): allele frequencies, the centred matrix and
    U, S, V of the same run on the device from the dosages that reader returns."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    N, M, k = 240, 1800, 4
    rng = np.random.default_rng(77)
    codes = np.concatenate([c for _, c in synth.balding_nichols_codes(N, M, k_pop=5, seed=78)])      # (M, N)
    g = np.array([2, 0, 1, 0])[codes]
    P = np.zeros((M, N, 3))
    for q in range(3):
        P[:, :, q] = np.where(g == q, 0.85, 0.075)
    P = 0.7 * P + 0.3 * rng.dirichlet([8, 8, 8], size=(M, N))
    P /= P.sum(-1, keepdims=True)
    P[rng.random((M, N)) < 0.03] = np.nan
    path = str(tmp_path / "f.bgen")
    ref.write_bgen(path, P, bit_depth=bit_depth)
    dos = ref.bgen_dosages(path, N, M)
    assert np.isnan(dos).sum() == np.isnan(P[:, :, 0]).sum()
    frac = np.abs(dos - np.round(dos))
    assert np.nanmean(frac > 0.05) > 0.5, "the case must be about fractional dosages"
    r = ref.Ref(f"PCAone --bgen {path} -k {k} -d 1 -o {tmp_path}/r -n 4 --maxp 5 --tol-rsvd 0", threads=4)
    Fr, Gr = r.F(), r.dataG()
    r.new_op()
    Ur, Sr, Vr = r.compute_usv(5, 0.0)
    r.close()
    op, d, p = _op(dos, k=k, svd=1, maxp=5, tol=0.0)
    F = op.F()
    assert np.max(np.abs(F - Fr) / Fr) < 1e-14
    X = op.read_block(0, M - 1, False)
    assert np.abs(X - Gr).max() <= 1e-14                      # dosage / 2 - F, NaN -> 0
    op.setFlags(False, True)
    op.computeUSV(5, 0.0)
    assert_usv_close(op.U, op.S, op.V, Ur, Sr, Vr, eig_rtol=1e-9, min_corr=1 - 1e-9)
    op.close()
