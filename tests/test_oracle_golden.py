"""CPU: the numpy oracle (oracle/pcaone_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py). This is what pins the oracle."""
import numpy as np
import pytest

from conftest import assert_usv_close, col_cos, golden
from oracle import pcaone_oracle as orc


def _ssvd():
    g = golden("ssvd_small")
    return g, int(g["N"]), int(g["M"]), int(g["k"])


def test_decode_af_lookup_bit_exact():
    g, N, M, k = _ssvd()
    codes = orc.unpack_codes(g["packed"], N)
    F = orc.allele_freq(codes)
    assert np.array_equal(F, g["F"])  # bit-exact
    # the reference only fills centered_geno_lookup in out-of-core mode; in-core keeps zeros
    Xc = orc.dense_from_codes(codes, F, standardize=False)
    assert np.array_equal(Xc, g["X_centered"])
    Xs = orc.dense_from_codes(codes, F, standardize=True)
    assert np.array_equal(Xs, g["X_standardized"])


def test_padding_bits_ignored():
    g, N, M, k = _ssvd()
    packed = g["packed"].copy()
    assert N % 4 != 0
    last = packed[:, -1]
    keep = (1 << (2 * (N % 4))) - 1
    packed[:, -1] = (last & keep) | (~np.uint8(keep) & 0xFF & 0b10101010)
    F = orc.allele_freq(orc.unpack_codes(packed, N))
    assert np.array_equal(F, g["F"])


def test_ssvd_epochs_G_H():
    g, N, M, k = _ssvd()
    d = orc.OracleData(g["packed"], N)
    op = orc.OracleRsvd(d, k, omega=g["omega"])
    op.set_flags(False, True)
    G = np.zeros((M, op.l))
    H = np.zeros((N, op.l))
    op.gandh(G, H, 0)
    np.testing.assert_allclose(G, g["G0"], rtol=1e-12, atol=1e-12 * np.abs(g["G0"]).max())
    np.testing.assert_allclose(H, g["H0"], rtol=1e-12, atol=1e-12 * np.abs(g["H0"]).max())
    op.gandh(G, H, 1)
    # Omega = thinQ(H) is sign-ambiguous per column before flipOmg; after flipOmg both sides
    # are aligned to the previous Omega, so G/H agree column-wise up to sign.
    c = col_cos(G, g["G1"])
    assert c.min() > 1 - 1e-10
    sg = np.sign((G * g["G1"]).sum(0))
    np.testing.assert_allclose(G * sg, g["G1"], rtol=0, atol=1e-9 * np.abs(g["G1"]).max())
    np.testing.assert_allclose(H * sg, g["H1"], rtol=0, atol=1e-9 * np.abs(g["H1"]).max())


def test_ssvd_usv():
    g, N, M, k = _ssvd()
    d = orc.OracleData(g["packed"], N)
    op = orc.OracleRsvd(d, k, omega=g["omega"])
    op.set_flags(False, True)
    U, S, V = op.compute_usv(int(g["maxp"]), 0.0)
    assert op.epochs == int(g["maxp"]) + 1
    assert_usv_close(U, S, V, g["U"], g["S"], g["V"], eig_rtol=1e-9, min_corr=1 - 1e-9)


def test_winsvd_incore_usv_and_schedule():
    g, N, M, k = _ssvd()
    w = golden("winsvd_small")
    bands = int(w["bands"])
    d = orc.OracleData(g["packed"], N)
    _, windows = orc.incore_windows(M, bands)
    op = orc.OracleRsvd(d, k, winsvd=True, bands=bands, omega=g["omega"], windows=windows)
    op.set_flags(False, True)
    d.permute(w["perm"])
    U, S, V = op.compute_usv(int(w["maxp"]), 0.0)
    # Omega updates per epoch for 8 windows: 7, 4, 2, then 1 per epoch
    assert op.n_omega_updates[:4] == [7, 4, 2, 1]
    assert_usv_close(U, S, V, w["U"], w["S"], w["V"], eig_rtol=1e-9, min_corr=1 - 1e-9)


def test_winsvd_schedule_64_bands():
    """63, 32, 16, 8, 4, 2, then 1 Omega updates per epoch at 64 windows."""
    rng = np.random.default_rng(0)
    N, M = 16, 256
    packed = rng.integers(0, 256, size=(M, 4), dtype=np.uint8)
    packed[packed == 0b01010101] = 0
    d = orc.OracleData(packed, N)
    _, windows = orc.incore_windows(M, 64)
    op = orc.OracleRsvd(d, 2, oversamples=2, winsvd=True, bands=64, omega=rng.standard_normal((N, 4)),
                        windows=windows)
    op.set_flags(False, False)
    G = np.zeros((M, 4))
    H = np.zeros((N, 4))
    for pi in range(8):
        op.gandh(G, H, pi)
    assert op.n_omega_updates == [63, 32, 16, 8, 4, 2, 1, 1]


def test_ooc_block_plan_and_permute_plink():
    g, N, M, k = _ssvd()
    w = golden("winsvd_ooc_small")
    bs, nb, bf, start, stop = orc.ooc_block_plan(N, M, k, 10, float(w["memory"]), True, int(w["bands"]))
    assert [bs, nb, bf] == list(w["plan"])
    assert np.array_equal(start, w["start"]) and np.array_equal(stop, w["stop"])
    assert np.array_equal(orc.permute_plink_indices(M, N, int(w["bands"])), w["perm"])
    s = golden("ssvd_ooc_small")
    bs, nb, bf, start, stop = orc.ooc_block_plan(N, M, k, 10, float(s["memory"]), False)
    assert [bs, nb, bf] == list(s["plan"])
    assert np.array_equal(start, s["start"]) and np.array_equal(stop, s["stop"])
    h = golden("helpers")
    assert np.array_equal(orc.permute_plink_indices(203, 13, 8), h["plink_perm"])


def test_winsvd_ooc_usv():
    g, N, M, k = _ssvd()
    w = golden("winsvd_ooc_small")
    d = orc.OracleData(g["packed"], N)
    d.permute(w["perm"])
    assert np.array_equal(d.F, w["F"])
    blk = d.block(int(w["start"][0]), int(w["stop"][0]), True)
    assert np.array_equal(blk, w["block0"])
    windows = list(zip(w["start"].tolist(), w["stop"].tolist()))
    op = orc.OracleRsvd(d, k, winsvd=True, bands=int(w["bands"]), omega=g["omega"], windows=windows,
                        band_factor=int(w["plan"][2]), out_of_core=True)
    op.set_flags(False, True)
    U, S, V = op.compute_usv(int(w["maxp"]), 0.0)
    assert_usv_close(U, S, V, w["U"], w["S"], w["V"], eig_rtol=1e-9, min_corr=1 - 1e-9)


def test_ssvd_ooc_usv():
    g, N, M, k = _ssvd()
    s = golden("ssvd_ooc_small")
    d = orc.OracleData(g["packed"], N)
    windows = list(zip(s["start"].tolist(), s["stop"].tolist()))
    op = orc.OracleRsvd(d, k, omega=g["omega"], windows=windows, out_of_core=True)
    op.set_flags(False, True)
    U, S, V = op.compute_usv(int(s["maxp"]), 0.0)
    assert_usv_close(U, S, V, s["U"], s["S"], s["V"], eig_rtol=1e-9, min_corr=1 - 1e-9)


def test_emu():
    g, N, M, k = _ssvd()
    e = golden("emu_small")
    d = orc.OracleData(e["packed"], N)
    assert np.array_equal(d.F, e["F"])
    assert np.array_equal((d.codes == 1).astype(np.uint8), e["mask"])
    op = orc.OracleRsvd(d, k, omega=g["omega"])
    U, S, V, iters = orc.run_emu(op, int(e["maxp"]), 0.0, maxiter=int(e["maxiter"]))
    assert iters == int(e["iters"])
    assert_usv_close(U, S, V, e["U"], e["S"], e["V"], eig_rtol=1e-8, min_corr=1 - 1e-8)
    # read_block_update: fused fill + clamp + scale for one block
    b0, b1 = [int(x) for x in e["b0"]]
    blk = d.block(b0, b1, True, usv=(e["U"], e["S"], e["V"]), emu=True)
    np.testing.assert_allclose(blk, e["block0_update"], rtol=1e-13, atol=1e-15)
    obs = (d.codes[b0:b1 + 1] != 1).T
    assert np.array_equal(blk[obs], e["block0_update"][obs])  # observed entries bit-exact


def test_ld_r2():
    ld = golden("ld_small")
    g, N, M, k = _ssvd()
    import os, tempfile
    p = os.path.join(tempfile.mkdtemp(), "r.residuals")
    ld["residuals_file"].tofile(p)
    G = orc.read_residuals(p)
    np.testing.assert_allclose(G, ld["G"].astype(np.float64), rtol=0, atol=1e-6)
    per_chr = (M + 21) // 22
    chrom = [j // per_chr + 1 for j in range(M)]
    pos = [(j % per_chr + 1) * 100 for j in range(M)]
    ws, we = orc.ld_windows(chrom, pos, int(ld["ld_bp"]))
    assert np.array_equal(ws, ld["ws"]) and np.array_equal(we, ld["we"])
    r2 = orc.ld_r2(G, ws, we)
    np.testing.assert_allclose(r2, ld["r2"], rtol=1e-10, atol=1e-14)


def test_helpers():
    h = golden("helpers")
    A, B = h["A"], h["B"]
    assert abs(orc.mev(A, B) - float(h["mev"])) < 1e-14
    O2, O1 = orc.flip_omg(A, -A + 0.01 * B)
    assert np.array_equal(O1, h["flip_omg"]) and np.array_equal(O2, h["flip_omg2"])
    U, V = orc.flip_uv(A, B)
    assert np.array_equal(U, h["flipU"]) and np.array_equal(V, h["flipV"])


@pytest.mark.parametrize("name", ["tall", "wide"])
@pytest.mark.parametrize("p,w", [(3, 0), (5, 4), (3, 8)])
def test_rsvd_one_restatement_vs_reference(name, p, w):
    """oracle.rsvd_one (numpy restatement of RSVD.hpp:92-362) against PCAone::RsvdOne<MatrixXd>
    outputs generated from the unmodified reference (tests/golden/rsvd_one.npz)."""
    g = golden("rsvd_one")
    A, k, os_ = g[f"A_{name}"], int(g["k"]), int(g["os"])
    U, S, V = orc.rsvd_one(A, k, os_, g[f"omega_{name}"], p, w)
    Ur, Sr, Vr = g[f"{name}_p{p}_w{w}_U"], g[f"{name}_p{p}_w{w}_S"], g[f"{name}_p{p}_w{w}_V"]
    assert U.shape == Ur.shape and V.shape == Vr.shape
    assert np.max(np.abs(S - Sr) / Sr) < 1e-12
    assert col_cos(U, Ur).min() > 1 - 1e-12 and col_cos(V, Vr).min() > 1 - 1e-12


def test_rsvd_one_restatement_argument_checks():
    A = np.random.default_rng(0).standard_normal((40, 12))
    om = np.zeros((12, 6))
    with pytest.raises(RuntimeError):
        orc.rsvd_one(A, 3, 3, om, 3, 3)      # windows must be even
    with pytest.raises(RuntimeError):
        orc.rsvd_one(A, 3, 3, om, 1, 4)      # 2^p >= windows
    with pytest.raises(RuntimeError):
        orc.rsvd_one(A, 3, 3, om, 4, 16)     # block smaller than the number of windows


def test_perform_op_restatement_vs_reference():
    """oracle.perform_op against ArnoldiOpData::perform_op of the unmodified reference."""
    g, a = golden("ssvd_small"), golden("arnoldi_op")
    od = orc.OracleData(g["packed"], int(g["N"]))
    blocks = list(zip(a["start"], a["stop"]))
    for std, key in ((True, "y_std"), (False, "y_raw")):
        y = orc.perform_op(od, a["x"], blocks, std)
        assert np.abs(y - a[key]).max() <= 1e-13 * np.abs(a[key]).max()


def test_ld_prune_restatement_vs_reference(tmp_path):
    """oracle.ld_prune against the keep masks recovered from ld_prune_big's own .ld.prune.in."""
    ld, pr = golden("ld_small"), golden("ld_prune_small")
    p = str(tmp_path / "r.residuals")
    ld["residuals_file"].tofile(p)
    G = orc.read_residuals(p)
    for tol in pr["tols"]:
        assert np.array_equal(orc.ld_prune(G, ld["ws"], ld["we"], float(tol), pr["af"]), pr[f"keep_af_{tol}"])
        assert np.array_equal(orc.ld_prune(G, ld["ws"], ld["we"], float(tol), None), pr[f"keep_noaf_{tol}"])


def test_pcangsd_restatement_vs_reference():
    """emMAF_with_GL, the initial E and the PCAngsd EM loop restated in numpy against the unmodified
    reference driven on a synthetic beagle.gz (tests/golden/pcangsd_small.npz)."""
    g = golden("pcangsd_small")
    P, k = g["P"], int(g["k"])
    F, it = orc.em_maf_with_gl(P, int(g["maxiter"]), 1e-6)       # --maxiter also bounds the MAF EM (FileBeagle.cpp:52)
    assert np.abs(F - g["F"]).max() < 1e-13
    assert np.abs(orc.gl_expected(P, g["F"]) - g["E0"]).max() < 1e-13
    od = orc.OracleGLData(P, g["F"])
    oo = orc.OracleRsvd(od, k, omega=g["omega"])
    U, S, V, iters = orc.run_emu(oo, int(g["maxp"]), 0.0, maxiter=int(g["maxiter"]), tolem=1e-5, final_standardize=False)
    assert iters == int(g["iters"])
    assert_usv_close(U, S, V, g["U"], g["S"], g["V"], eig_rtol=1e-10, min_corr=1 - 1e-10)


def test_pcangsd_grm_restatement_vs_reference():
    """numpy restatement of the GRM step against the unmodified reference (tests/golden/pcangsd_grm.npz, written by
    tests/golden/make_golden_grm.py from oracle/_ref on the beagle.gz of pcangsd_small.npz)."""
    g, q = golden("pcangsd_small"), golden("pcangsd_grm")
    Cm, Dc = orc.pcangsd_grm(g["P"], g["F"], q["U"], q["S"], q["V"])
    assert np.abs(Dc - q["Dc"]).max() <= 1e-12 * np.abs(q["Dc"]).max()
    assert np.abs(Cm - q["C"]).max() <= 1e-12 * np.abs(q["C"]).max()
    w = np.sort(np.abs(np.linalg.eigvalsh(q["C"])))[::-1]
    assert np.abs(w - q["S2"]).max() <= 1e-12 * w[0]
