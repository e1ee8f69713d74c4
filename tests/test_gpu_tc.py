"""GPU parity tests (-m gpu) of the int8 tensor-core path (PCAONE_PREC_INT8X*, tc_gemm.cuh):
tcgen05 kind::i8 UMMAs fed from tensor memory, exact integer accumulation (Ozaki scheme).

What is checked, and to which tolerance:
  * the rounding contract: G~ returned by the library is within one unit of its last slice bit
    of X^T Omega~ (Omega~ = Omega rounded to 8S-1 bits against its column maxima), and
    H == X G~ to FP64 summation noise (1e-12) — i.e. the tensor-core products are exact;
  * U, S, V against the numpy oracle / golden vectors of the unmodified reference at the
    north_star tolerance (eigenvalues <= 1e-6 relative, |cos| >= 0.9999);
  * bit-reproducibility across runs (integer atomics), ragged shapes, out-of-core streaming,
    and mean imputation of missing calls by count + mask product pairs on the same kernels."""
import numpy as np
import pytest

from conftest import assert_usv_close, col_cos, golden
from oracle import pcaone_oracle as orc
from pcaone_b200 import _lib, halko, synth

pytestmark = pytest.mark.gpu


def _packed(N, M, k_pop, seed, miss=0.0):
    return np.concatenate([synth.pack_codes(c) for _, c in synth.balding_nichols_codes(N, M, k_pop=k_pop, seed=seed, miss=miss)])


def _op(packed, N, **kw):
    omega = kw.pop("omega", None)
    p = halko.Param(**kw)
    d = halko.FileBed(p, packed=packed, nsamples=N)
    d.prepare()
    cls = halko.FancyRsvdOpData if p.svd == 2 else halko.NormalRsvdOpData
    op = cls(d, p.k, p.oversamples)
    if omega is not None:
        op.setOmg(omega)
    return op, d, p


def _round_slices(X, S):
    """numpy restatement of k_tc_slice: X~ = rint(X 2^(p-e)) 2^(e-p), p = 8S-1, per-column e."""
    p = 8 * S - 1
    mx = np.abs(X).max(0)
    e = np.where(mx > 0, np.floor(np.log2(np.maximum(mx, 1e-300) * (128.0 / 126.0))) + 1, 0).astype(int)
    return np.rint(X * 2.0 ** (p - e)) * 2.0 ** (e - p), e


@pytest.mark.parametrize("S", [2, 3, 4])
@pytest.mark.parametrize("N,M,k", [(500, 3000, 5), (129, 777, 3), (1000, 2100, 20)])
def test_tc_products_are_exact(S, N, M, k):
    packed = _packed(N, M, k + 2, 100 + N)
    op, d, p = _op(packed, N, k=k, svd=1, precision=S)
    op.setFlags(False, True)
    G, H = op.computeGandH(0)
    t = op.timers()
    assert t.tc_ranges == 1 and t.fp64_ranges == 0
    od = orc.OracleData(packed, N)
    X = od.block(0, M - 1, True)                      # N x M standardized, as the reference decodes it
    Omt, _ = _round_slices(op.Omg, S)
    sd = np.sqrt(od.F * (1 - od.F))
    s = np.where(sd > 1e-9, np.sqrt(2.0) / np.maximum(sd, 1e-300), 1.0)
    W = (X.T @ Omt) * s[:, None]
    Wt, e = _round_slices(W, S)
    ulp = 2.0 ** (e - (8 * S - 1))
    # G~ = W~ / s : within one last-slice unit (ties / FP64 summation order in this numpy check)
    assert np.all(np.abs(G * s[:, None] - Wt) <= 1.01 * ulp[None, :])
    # the tensor-core product itself is exact: H == X G~ to FP64 noise
    Href = X @ G
    assert np.abs(H - Href).max() <= 1e-12 * np.abs(Href).max()
    # and G~ is close to the unrounded product
    Gref = X.T @ op.Omg
    assert np.abs(G - Gref).max() <= 2.0 ** (-(8 * S - 3)) * np.abs(Gref).max() * 8
    op.close()


def test_tc_bit_reproducible():
    N, M, k = 700, 5000, 6
    packed = _packed(N, M, 8, 5)
    outs = []
    for _ in range(2):
        op, d, p = _op(packed, N, k=k, svd=2, bands=8, precision=_lib.PREC_INT8X3, maxp=3, tol=0.0)
        op.setFlags(False, True)
        op.computeUSV(p.maxp, p.tol)
        outs.append((op.U.copy(), op.S.copy(), op.V.copy()))
        op.close()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("case", ["ssvd_small", "winsvd_small", "ssvd_ooc_small", "winsvd_ooc_small"])
@pytest.mark.parametrize("S", [3, 4])
def test_tc_usv_vs_golden(case, S):
    g = golden("ssvd_small")
    N, k = int(g["N"]), int(g["k"])
    w = golden(case)
    kw = dict(k=k, svd=2 if "win" in case else 1, maxp=int(w["maxp"]), tol=0.0, omega=g["omega"], precision=S)
    if "win" in case:
        kw["bands"] = int(w["bands"])
    if "ooc" in case:
        kw["memory"] = float(w["memory"])
    op, d, p = _op(g["packed"], N, **kw)
    op.setFlags(False, True)
    op.computeUSV(p.maxp, p.tol)
    t = op.timers()
    assert t.tc_ranges > 0 and t.fp64_ranges == 0
    assert_usv_close(op.U, op.S, op.V, w["U"], w["S"], w["V"])
    print(case, S, "S rel err", np.max(np.abs(op.S - w["S"]) / w["S"]), "min cos U", col_cos(op.U, w["U"]).min())


@pytest.mark.parametrize("N,M,k,svd,bands,S", [(500, 3000, 5, 1, 64, 3), (501, 4099, 4, 2, 8, 3), (1030, 2500, 12, 2, 4, 3),
                                               (333, 70000, 10, 1, 64, 3), (2504, 9000, 10, 2, 64, 3),
                                               (2000, 6000, 40, 2, 8, 3), (2000, 6000, 40, 1, 8, 2),
                                               (777, 5000, 30, 2, 16, 4)])
def test_tc_usv_vs_numpy_oracle(N, M, k, svd, bands, S):
    packed = _packed(N, M, k + 2, N + M)
    maxp = 7 if svd == 2 else 4
    op, d, p = _op(packed, N, k=k, svd=svd, bands=bands, maxp=maxp, tol=0.0, precision=S)
    op.setFlags(False, True)
    op.computeUSV(p.maxp, p.tol)
    t = op.timers()
    assert t.tc_ranges > 0 and t.fp64_ranges == 0
    od = orc.OracleData(packed, N)
    windows = None
    if svd == 2:
        od.permute(d.perm)
        _, windows = orc.incore_windows(M, bands)
    oo = orc.OracleRsvd(od, k, winsvd=svd == 2, bands=bands, omega=op.Omg, windows=windows)
    oo.set_flags(False, True)
    U, S_, V = oo.compute_usv(maxp, 0.0)
    assert op.epochs == oo.epochs
    print((N, M, k, svd, bands, S), "eig rel err", np.max(np.abs(op.S ** 2 - S_ ** 2) / S_ ** 2), "min cos U",
          col_cos(op.U, U).min())
    assert_usv_close(op.U, op.S, op.V, U, S_, V)


def test_tc_missing_ranges_run_count_and_mask_products():
    """Windows that contain missing calls stay on the int8 kernels: every product is run as a
    (non-missing count, missing mask) pair, which mean-imputes the missing calls (lut[01] = 0,
    FilePlink.cpp:194-197). No range falls back to the FP64 kernels."""
    N, M, k = 600, 4096, 5
    rng = np.random.default_rng(11)
    codes = np.concatenate([c for _, c in synth.balding_nichols_codes(N, M, k_pop=7, seed=9)])
    codes[rng.integers(0, 1024, 200), rng.integers(0, N, 200)] = 1      # missing only in the first quarter
    packed = synth.pack_codes(codes)
    op, d, p = _op(packed, N, k=k, svd=2, bands=8, no_shuffle=True, maxp=6, tol=0.0, precision=3)
    op.setFlags(False, True)
    op.computeUSV(p.maxp, p.tol)
    t = op.timers()
    assert t.tc_ranges > 0 and t.fp64_ranges == 0
    assert 0 < t.tc_miss_ranges < t.tc_ranges
    od = orc.OracleData(packed, N)
    _, windows = orc.incore_windows(M, 8)
    oo = orc.OracleRsvd(od, k, winsvd=True, bands=8, omega=op.Omg, windows=windows)
    oo.set_flags(False, True)
    U, S, V = oo.compute_usv(6, 0.0)
    assert_usv_close(op.U, op.S, op.V, U, S, V)


@pytest.mark.parametrize("S", [2, 3, 4])
@pytest.mark.parametrize("N,M,k,miss", [(500, 3000, 5, 0.02), (129, 777, 3, 0.3), (1000, 2100, 20, 0.001)])
def test_tc_products_with_missing_calls(S, N, M, k, miss):
    """The rounding contract with mean-imputed missing calls. The G pass stays exact (the weight
    1 - f_j of the mask term belongs to the OUTPUT row); in the H pass the mask operand
    D = (f_j - 1) W~_j is itself rounded to 8S-1 bits, so the MISSING entries of X are imputed with
    0 +- 2^-(8S) of the column scale: H == X G~ to that bound instead of FP64 noise."""
    packed = _packed(N, M, k + 2, 300 + N, miss=miss)
    op, d, p = _op(packed, N, k=k, svd=1, precision=S)
    op.setFlags(False, True)
    G, H = op.computeGandH(0)
    t = op.timers()
    assert t.tc_ranges == 1 and t.tc_miss_ranges == 1 and t.fp64_ranges == 0
    od = orc.OracleData(packed, N)
    assert np.array_equal(op.F(), od.F)
    X = od.block(0, M - 1, True)
    Omt, _ = _round_slices(op.Omg, S)
    sd = np.sqrt(od.F * (1 - od.F))
    s = np.where(sd > 1e-9, np.sqrt(2.0) / np.maximum(sd, 1e-300), 1.0)
    W = (X.T @ Omt) * s[:, None]
    Wt, e = _round_slices(W, S)
    ulp = 2.0 ** (e - (8 * S - 1))
    assert np.all(np.abs(G * s[:, None] - Wt) <= 1.01 * ulp[None, :])
    Href = X @ G
    nmiss_max = int((od.codes == 1).sum(0).max())          # missing calls of the worst sample
    bound = ulp[None, :] * nmiss_max + 1e-12 * np.abs(Href).max()
    assert np.all(np.abs(H - Href) <= bound)
    rel = np.abs(H - Href).max() / np.abs(Href).max()
    print((S, N, M, miss), "H rel err", rel)
    assert rel <= 2.0 ** (-(8 * S - 4))
    op.close()


@pytest.mark.parametrize("svd,memory", [(1, 0.0), (2, 0.0), (2, 0.004)])
def test_tc_usv_with_missing_vs_numpy_oracle(svd, memory):
    """1 % missing calls everywhere (every range takes the count + mask route), in-core and streamed."""
    N, M, k, bands = 640, 6000, 6, 8
    packed = _packed(N, M, k + 2, 4242, miss=0.01)
    kw = dict(k=k, svd=svd, bands=bands, maxp=7 if svd == 2 else 4, tol=0.0, precision=3)
    if memory:
        kw["memory"] = memory
    op, d, p = _op(packed, N, **kw)
    op.setFlags(False, True)
    op.computeUSV(p.maxp, p.tol)
    t = op.timers()
    assert t.tc_ranges > 0 and t.tc_miss_ranges == t.tc_ranges and t.fp64_ranges == 0
    # the same run on the FP64 DMMA kernels (which decode code 01 through the LUT)
    kw["precision"] = 0
    op2, d2, p2 = _op(packed, N, omega=op.Omg, **kw)
    op2.setFlags(False, True)
    op2.computeUSV(p.maxp, p.tol)
    assert op.epochs == op2.epochs
    print("svd", svd, "mem", memory, "S rel err vs fp64", np.max(np.abs(op.S - op2.S) / op2.S))
    assert_usv_close(op.U, op.S, op.V, op2.U, op2.S, op2.V)
    if not memory:
        od = orc.OracleData(packed, N)
        windows = None
        if svd == 2:
            od.permute(d.perm)
            _, windows = orc.incore_windows(M, bands)
        oo = orc.OracleRsvd(od, k, winsvd=svd == 2, bands=bands, omega=op.Omg, windows=windows)
        oo.set_flags(False, True)
        U, S_, V = oo.compute_usv(p.maxp, 0.0)
        assert_usv_close(op.U, op.S, op.V, U, S_, V)


def test_tc_ooc_equals_incore():
    """Streamed blocks and the resident shard: G is bit-identical (integer sums do not depend on how
    the work is cut), H agrees to FP64 noise (the rank-1 centring term sum_j f_j W~_j is an FP64 sum
    whose partial-block boundaries follow the k-block alignment of the launch)."""
    N, M, k = 640, 4000, 6
    packed = _packed(N, M, 8, 77)
    op, d, p = _op(packed, N, k=k, svd=1, memory=0.004, precision=3)
    assert d.nblocks > 1
    op.setFlags(False, True)
    G1, H1 = op.computeGandH(0)
    # the same block plan walked over a resident shard
    p2 = halko.Param(k=k, svd=1, precision=3)
    d2 = halko.FileBed(p2, packed=packed, nsamples=N)
    d2.prepare()
    d2.start, d2.stop, d2.bandFactor = d.start, d.stop, 1
    op2 = halko.NormalRsvdOpData(d2, k, p2.oversamples)
    op2.setFlags(False, True)
    G2, H2 = op2.computeGandH(0)
    assert np.array_equal(G1, G2)
    assert np.abs(H1 - H2).max() <= 1e-13 * np.abs(H1).max()
