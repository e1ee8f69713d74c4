"""GPU parity tests proper (-m gpu): the CUDA path, called through the C-ABI via the host
mirror in pcaone_b200/halko.py, against (a) golden vectors from the unmodified reference,
(b) the numpy oracle on seeded inputs, (c) the compiled reference (oracle/_ref) when it
travelled to the box, (d) size-independent properties.

Tolerances (BASELINE.json north_star): decode / allele frequency / masks bit-exact;
top-k eigenvalues <= 1e-6 relative; PCs |corr| >= 0.9999 up to sign. G/H per epoch are
compared at 1e-11 relative to the matrix scale (FP64 summation-order noise only)."""
import numpy as np
import pytest

from conftest import assert_usv_close, col_cos, golden
from oracle import pcaone_oracle as orc
from pcaone_b200 import halko, synth

pytestmark = pytest.mark.gpu


def _g():
    g = golden("ssvd_small")
    return g, int(g["N"]), int(g["M"]), int(g["k"])


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


def _close(A, B, tol=1e-11):
    scale = np.abs(B).max()
    assert np.abs(A - B).max() <= tol * scale, np.abs(A - B).max() / scale


def test_decode_af_bit_exact_vs_golden():
    g, N, M, k = _g()
    op, d, p = _op(g["packed"], N, k=k, svd=1)
    assert np.array_equal(op.F(), g["F"])
    lut = orc.centered_lookup(g["F"])
    assert np.array_equal(op.lookup(), lut)
    assert np.array_equal(op.read_block(0, M - 1, False), g["X_centered"])
    assert np.array_equal(op.read_block(0, M - 1, True), g["X_standardized"])
    assert np.array_equal(op.read_block(17, 93, True), g["X_standardized"][:, 17:94])
    assert op.missing_count() == 0


def test_omega_matches_reference_stream():
    g, N, M, k = _g()
    op, d, p = _op(g["packed"], N, k=k, svd=1)
    assert np.array_equal(op.Omg, g["omega"])
    assert np.array_equal(op.omega(), g["omega"])  # host -> device -> host round trip


def test_ssvd_gandh_epochs_vs_golden():
    g, N, M, k = _g()
    op, d, p = _op(g["packed"], N, k=k, svd=1, omega=g["omega"])
    op.setFlags(False, True)
    G, H = op.computeGandH(0)
    _close(G, g["G0"])
    _close(H, g["H0"])
    G, H = op.computeGandH(1)
    sg = np.sign((G * g["G1"]).sum(0))
    _close(G * sg, g["G1"], 1e-9)
    _close(H * sg, g["H1"], 1e-9)


@pytest.mark.parametrize("case", ["ssvd_small", "winsvd_small", "ssvd_ooc_small", "winsvd_ooc_small"])
def test_usv_vs_golden(case):
    g, N, M, k = _g()
    w = golden(case)
    kw = dict(k=k, svd=2 if "win" in case else 1, maxp=int(w["maxp"]), tol=0.0, omega=g["omega"])
    if "win" in case:
        kw["bands"] = int(w["bands"])
    if "ooc" in case:
        kw["memory"] = float(w["memory"])
    op, d, p = _op(g["packed"], N, **kw)
    op.setFlags(False, True)
    op.computeUSV(p.maxp, p.tol)
    if "win" in case:
        assert np.array_equal(d.perm, w["perm"])
    assert_usv_close(op.U, op.S, op.V, w["U"], w["S"], w["V"])
    print(case, "S rel err", np.max(np.abs(op.S - w["S"]) / w["S"]), "min cos U", col_cos(op.U, w["U"]).min())


def test_emu_vs_golden():
    g, N, M, k = _g()
    e = golden("emu_small")
    op, d, p = _op(e["packed"], N, k=k, svd=1, emu=True, maxp=int(e["maxp"]), tol=0.0, maxiter=int(e["maxiter"]),
                   omega=g["omega"])
    assert np.array_equal(op.F(), e["F"])
    assert op.missing_count() == int(e["mask"].sum())
    iters = op.runEM()
    assert iters == int(e["iters"])
    assert_usv_close(op.U, op.S, op.V, e["U"], e["S"], e["V"])
    b0, b1 = [int(x) for x in e["b0"]]
    op.setUSV(e["U"], e["S"], e["V"])
    blk = op.read_block(b0, b1, True, update=True)
    np.testing.assert_allclose(blk, e["block0_update"], rtol=1e-12, atol=1e-14)
    obs = (orc.unpack_codes(e["packed"], N)[b0:b1 + 1] != 1).T
    assert np.array_equal(blk[obs], e["block0_update"][obs])


@pytest.mark.parametrize("N,M,k,svd,bands", [(500, 3000, 5, 1, 64), (501, 4099, 4, 2, 8), (1030, 2500, 12, 2, 4),
                                             (333, 70000, 10, 1, 64), (2504, 9000, 10, 2, 64)])
def test_usv_vs_numpy_oracle(N, M, k, svd, bands):
    packed = np.concatenate([pk for _, pk in ((s, synth.pack_codes(c)) for s, c in
                                              synth.balding_nichols_codes(N, M, k_pop=k + 2, seed=N + M))])
    maxp = 7 if svd == 2 else 4
    op, d, p = _op(packed, N, k=k, svd=svd, bands=bands, maxp=maxp, tol=0.0)
    op.setFlags(False, True)
    op.computeUSV(p.maxp, p.tol)
    od = orc.OracleData(packed, N)
    assert np.array_equal(op.F(), od.F if d.perm is None else od.F[d.perm])
    windows = None
    if svd == 2:
        od.permute(d.perm)
        _, windows = orc.incore_windows(M, bands)
    oo = orc.OracleRsvd(od, k, winsvd=svd == 2, bands=bands, omega=op.Omg, windows=windows)
    oo.set_flags(False, True)
    U, S, V = oo.compute_usv(maxp, 0.0)
    assert op.epochs == oo.epochs
    print((N, M, k, svd, bands), "S rel err", np.max(np.abs(op.S - S) / S), "min cos U", col_cos(op.U, U).min())
    assert_usv_close(op.U, op.S, op.V, U, S, V)


def test_ragged_and_missing_edge_cases():
    rng = np.random.default_rng(3)
    for N in (5, 64, 65, 127, 129):
        M = 257
        codes = rng.integers(0, 4, size=(M, N)).astype(np.uint8)
        codes[0] = 1          # an all-missing SNP: F = 0
        codes[1] = 0          # monomorphic: F = 1, sd = 0 -> scale stays 1
        codes[2] = 3          # monomorphic: F = 0
        packed = synth.pack_codes(codes, pad_code=int(rng.integers(0, 4)))
        op, d, p = _op(packed, N, k=2, oversamples=2, svd=1)
        od = orc.OracleData(packed, N)
        assert np.array_equal(op.F(), od.F)
        assert np.array_equal(op.read_block(0, M - 1, True), od.block(0, M - 1, True))
        assert op.missing_count() == int((codes == 1).sum())
        op.setFlags(False, True)
        G, H = op.computeGandH(0)
        X = od.block(0, M - 1, True)
        _close(G, X.T @ op.Omg)
        _close(H, X @ (X.T @ op.Omg))
        op.close()


def test_linearity_and_ooc_equals_incore():
    """Size-independent properties: H is linear in Omega; streamed blocks == resident."""
    N, M, k = 700, 5000, 6
    packed = np.concatenate([synth.pack_codes(c) for _, c in synth.balding_nichols_codes(N, M, k_pop=8, seed=5)])
    op, d, p = _op(packed, N, k=k, svd=1)
    op.setFlags(False, True)
    O1 = op.Omg.copy()
    G1, H1 = op.computeGandH(0)
    rng = np.random.default_rng(0)
    O2 = np.asfortranarray(rng.standard_normal(O1.shape))
    op.setOmg(O2)
    G2, H2 = op.computeGandH(0)
    op.setOmg(2.0 * O1 - 3.0 * O2)
    G3, H3 = op.computeGandH(0)
    _close(G3, 2.0 * G1 - 3.0 * G2, 1e-12)
    _close(H3, 2.0 * H1 - 3.0 * H2, 1e-12)
    op2, d2, p2 = _op(packed, N, k=k, svd=1, memory=0.004)
    assert d2.nblocks > 1
    op2.setFlags(False, True)
    op2.setOmg(O1)
    Gs, Hs = op2.computeGandH(0)
    assert np.array_equal(op2.F(), op.F())
    _close(Gs, G1, 1e-13)
    _close(Hs, H1, 1e-12)


def test_vs_compiled_reference_if_present(tmp_path):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    N, M, k = 1200, 20000, 8
    prefix = str(tmp_path / "r")
    packed = synth.write_bed(prefix, N, M, k_pop=10, seed=21)
    for svd, maxp in ((1, 5), (2, 6)):
        r = ref.Ref(f"PCAone -b {prefix} -k {k} -d {svd} -o {tmp_path}/o{svd} --maxp {maxp} --tol-rsvd 0 -n 8",
                    threads=8)
        r.new_op()
        Ur, Sr, Vr = r.compute_usv(maxp, 0.0)
        Fr = r.F()
        r.close()
        op, d, p = _op(packed, N, k=k, svd=svd, maxp=maxp, tol=0.0)
        assert np.array_equal(op.F() if d.perm is None else op.F()[np.argsort(d.perm)], Fr)
        op.setFlags(False, True)
        op.computeUSV(maxp, 0.0)
        assert_usv_close(op.U, op.S, op.V, Ur, Sr, Vr)
        op.close()
