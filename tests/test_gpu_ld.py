"""GPU parity tests (-m gpu) of LD r2 (pcaone_ld_r2, csrc/ld.cuh) against the golden vectors of
the unmodified reference (ld_r2_big, src/LD.cpp:450-473) and the numpy oracle.

Tolerance: r2 within 1e-12 absolute of the FP64 reference (only the summation order of the dot
products differs); window tables (integer logic) must be identical."""
import os

import numpy as np
import pytest

from conftest import golden
from oracle import pcaone_oracle as orc
from pcaone_b200 import halko, ld, synth

pytestmark = pytest.mark.gpu


def _ctx(packed, N, k=2):
    p = halko.Param(k=k, svd=1)
    d = halko.FileBed(p, packed=packed, nsamples=N)
    d.prepare()
    return halko.NormalRsvdOpData(d, p.k, p.oversamples)


def _bim(M, nchr=22, step=100):
    per_chr = (M + nchr - 1) // nchr
    chrom = np.array([j // per_chr + 1 for j in range(M)])
    pos = np.array([(j % per_chr + 1) * step for j in range(M)])
    return chrom, pos


def test_ld_windows_match_reference():
    g = golden("ld_small")
    M = g["G"].shape[1]
    chrom, pos = _bim(M)
    ws, we = ld.divide_pos_by_window(chrom, pos, int(g["ld_bp"]))
    assert np.array_equal(ws, g["ws"]) and np.array_equal(we, g["we"])


def test_ld_r2_vs_golden():
    g = golden("ld_small")
    s = golden("ssvd_small")
    N = int(s["N"])
    tmp = os.path.join(os.environ.get("TMPDIR", "/tmp"), "ld_golden.residuals")
    g["residuals_file"].tofile(tmp)
    G = orc.read_residuals(tmp)  # what FileBin::read_all hands to ld_r2_big
    op = _ctx(s["packed"], N)
    r2 = ld.ld_r2_big(op, G, g["ws"], g["we"])
    np.testing.assert_allclose(r2, g["r2"], rtol=0, atol=1e-12)
    t = op.timers()
    assert t.ld_pairs == len(g["r2"]) and t.ld_tiles > 0
    op.close()


@pytest.mark.parametrize("N,M,bp,chunk", [(301, 900, 2500, 0), (1000, 3000, 12000, 0), (257, 1500, 7000, 256)])
def test_ld_r2_dense_vs_oracle(N, M, bp, chunk, monkeypatch):
    rng = np.random.default_rng(N + M)
    packed = np.concatenate([synth.pack_codes(c) for _, c in synth.balding_nichols_codes(N, M, k_pop=4, seed=M)])
    od = orc.OracleData(packed, N)
    G = od.block(0, M - 1, False) + 0.01 * rng.standard_normal((N, M))   # residual-like dense input
    G -= G.mean(0, keepdims=True)
    chrom, pos = _bim(M, nchr=3)
    ws, we = ld.divide_pos_by_window(chrom, pos, bp)
    if chunk:
        monkeypatch.setenv("PCAONE_LD_CHUNK", str(chunk))  # several lead chunks + halos
    op = _ctx(packed, N)
    r2 = ld.ld_r2_big(op, G, ws, we)
    ref = orc.ld_r2(G, ws, we)
    assert len(r2) == len(ref) and len(ref) > 0
    np.testing.assert_allclose(r2, ref, rtol=0, atol=1e-12)
    op.close()


def test_ld_r2_from_resident_bed():
    """G == NULL: centred, unscaled genotypes decoded from the resident packed shard
    (read_block_initial with standardize = false, src/LD.cpp:403-424)."""
    N, M = 500, 2000
    packed = np.concatenate([synth.pack_codes(c) for _, c in synth.balding_nichols_codes(N, M, k_pop=5, seed=3, miss=0.02)])
    od = orc.OracleData(packed, N)
    G = od.block(0, M - 1, False)
    chrom, pos = _bim(M, nchr=2)
    ws, we = ld.divide_pos_by_window(chrom, pos, 5000)
    op = _ctx(packed, N)
    r2 = ld.ld_r2_big(op, None, ws, we)
    ref = orc.ld_r2(G, ws, we)
    ok = np.isfinite(ref)
    np.testing.assert_allclose(r2[ok], ref[ok], rtol=0, atol=1e-12)
    op.close()


def test_ld_prune_vs_reference_golden():
    """pcaone_ld_prune against the keep masks of the unmodified ld_prune_big (LD.cpp:240-268)."""
    g, pr, s = golden("ld_small"), golden("ld_prune_small"), golden("ssvd_small")
    N = int(s["N"])
    tmp = os.path.join(os.environ.get("TMPDIR", "/tmp"), "ld_golden_prune.residuals")
    g["residuals_file"].tofile(tmp)
    G = orc.read_residuals(tmp)
    op = _ctx(s["packed"], N)
    for tol in pr["tols"]:
        assert np.array_equal(ld.ld_prune_big(op, G, g["ws"], g["we"], float(tol), pr["af"]), pr[f"keep_af_{tol}"])
        assert np.array_equal(ld.ld_prune_big(op, G, g["ws"], g["we"], float(tol), None), pr[f"keep_noaf_{tol}"])
    op.close()


@pytest.mark.parametrize("N,M,bp,chunk,tol", [(301, 900, 2500, 0, 0.02), (600, 4000, 30000, 0, 0.05),
                                              (257, 1500, 7000, 256, 0.03)])
def test_ld_prune_vs_oracle(N, M, bp, chunk, tol, monkeypatch):
    """Larger windows, several chromosomes, chunked leads with halos; with and without af.
    Pairs whose r2 sits within 1e-12 of the threshold are excluded from the comparison."""
    rng = np.random.default_rng(N + M)
    packed = np.concatenate([synth.pack_codes(c) for _, c in synth.balding_nichols_codes(N, M, k_pop=4, seed=M)])
    od = orc.OracleData(packed, N)
    G = od.block(0, M - 1, False) + 0.01 * rng.standard_normal((N, M))
    G -= G.mean(0, keepdims=True)
    chrom, pos = _bim(M, nchr=3)
    ws, we = ld.divide_pos_by_window(chrom, pos, bp)
    assert np.abs(orc.ld_r2(G, ws, we) - tol).min() > 1e-10     # no knife-edge pair in this seed
    if chunk:
        monkeypatch.setenv("PCAONE_LD_CHUNK", str(chunk))
    op = _ctx(packed, N)
    for af in (od.F, None):
        keep = ld.ld_prune_big(op, G, ws, we, tol, af)
        ref = orc.ld_prune(G, ws, we, tol, af)
        assert 0 < ref.sum() < M
        assert np.array_equal(keep, ref)
    # from the resident bed (G == NULL)
    keep = ld.ld_prune_big(op, None, ws, we, tol, od.F)
    assert np.array_equal(keep, orc.ld_prune(od.block(0, M - 1, False), ws, we, tol, od.F))
    op.close()


@pytest.mark.ref
@pytest.mark.parametrize("ld_stats", [0, 1])
def test_residuals_and_adjusted_ld_vs_live_reference(tmp_path, ld_stats):
    """The ancestry-adjusted LD path end to end against the UNMODIFIED reference run here (oracle/_ref):
    `PCAone -b X -k K --ld --ld-stats s` -> Data::write_residuals (Data.cpp:242-291) -> `PCAone -B X.residuals
    --print-r2` (FileBin::read_all + ld_r2_big). On the device: pcaone_residuals_block (the float32 rows of the
    file), pcaone_ld_r2_ex from those floats (-B) and straight from the bed without the file (PACKED_RESID)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref was not built")
    N, M, k = 211, 1500, 3
    prefix = str(tmp_path / "g")
    packed = synth.write_bed(prefix, N, M, k_pop=4, seed=17, miss=0.01)
    out = str(tmp_path / "a")
    r = ref.Ref(f"PCAone -b {prefix} -k {k} -d 1 --ld --ld-stats {ld_stats} -o {out} --maxp 3 --tol-rsvd 0 -n 2", threads=2)
    r.new_op()
    U, S, V = r.compute_usv(3, 0.0)
    r.write_residuals()
    r.close()
    raw = np.fromfile(out + ".residuals", dtype=np.uint8)
    hdr = raw[:8].view(np.uint32)
    assert tuple(hdr) == (M, N)
    resid_ref = raw[8:].view(np.float32).reshape(M, N)
    r2_ = ref.Ref(f"PCAone -B {out}.residuals -F {out}.mbim --print-r2 --ld-bp 1500 -o {out}2 -n 2", threads=2)
    r2_ref, ws, we = r2_.ld_r2(out + ".mbim", 1500)
    r2_.close()
    assert len(r2_ref) > 1000

    p = halko.Param(k=k, svd=1, ld=True)
    d = halko.FileBed(p, packed=packed, nsamples=N)
    d.prepare()
    op = halko.NormalRsvdOpData(d, p.k, p.oversamples)
    op.setFlags(False, False)
    op.setUSV(U, S, V)          # the same U, S, V the reference subtracted
    # (1) the float32 rows of the .residuals file
    path = str(tmp_path / "ours.residuals")
    ld.write_residuals(op, path, ld_stats=ld_stats, chunk=400)
    mine = np.fromfile(path, dtype=np.uint8)
    assert np.array_equal(mine[:8], raw[:8])
    resid = mine[8:].view(np.float32).reshape(M, N)
    # identical up to float32 rounding of doubles that differ in the last bits (summation order of U S V^T / the mean)
    scale = np.abs(resid_ref).max()
    assert np.abs(resid - resid_ref).max() <= 2.5e-7 * scale
    assert (resid == resid_ref).mean() > 0.98
    # (2) -B: r2 from the reference's own file content
    r2_b = ld.ld_from_residuals_file(op, resid_ref, ws, we)
    np.testing.assert_allclose(r2_b, r2_ref, rtol=0, atol=1e-12)
    # (3) straight from the bed, no file
    r2_d = ld.ld_adjusted_from_bed(op, ws, we, ld_stats=ld_stats)
    np.testing.assert_allclose(r2_d, r2_ref, rtol=0, atol=5e-7)
    # pruning on the same operand: equal keep masks away from knife-edge pairs
    tol = 0.1
    if np.abs(r2_ref - tol).min() > 1e-5:
        keep_b = ld.ld_from_residuals_file(op, resid_ref, ws, we, r2_tol=tol)
        keep_d = ld.ld_adjusted_from_bed(op, ws, we, ld_stats=ld_stats, r2_tol=tol)
        assert np.array_equal(keep_b, keep_d)
    op.close()


def test_ld_projected_from_bed_vs_oracle():
    """bed + --USV: data->G = (I - U U^T) G (LD.cpp:491-496) on the device against numpy."""
    N, M, k = 300, 1800, 4
    packed = np.concatenate([synth.pack_codes(c) for _, c in synth.balding_nichols_codes(N, M, k_pop=5, seed=8)])
    od = orc.OracleData(packed, N)
    G = od.block(0, M - 1, False)
    U, _ = np.linalg.qr(np.random.default_rng(1).standard_normal((N, k)))
    Gp = G - U @ (U.T @ G)
    chrom, pos = _bim(M, nchr=3)
    ws, we = ld.divide_pos_by_window(chrom, pos, 4000)
    op = _ctx(packed, N)
    r2 = ld.ld_projected_from_bed(op, U, ws, we)
    np.testing.assert_allclose(r2, orc.ld_r2(Gp, ws, we), rtol=0, atol=1e-11)
    op.close()
