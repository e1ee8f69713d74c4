"""GPU parity tests (-m gpu) of the downstream products: projection option 1
(U = G V S^-1, src/Projection.cpp:236-241) and the per-SNP products of run_selection
(src/Selection.cpp:16-38) against the numpy restatement; the decode they sit on is pinned to the
reference by the golden tests."""
import os

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
    with pytest.raises(RuntimeError, match="packed"):
        downstream.run_projection(op, A[:, :4], np.ones(4), project=2)   # the missing-call indicator needs 2-bit codes
    with pytest.raises(RuntimeError, match="project 3"):
        downstream.run_projection(op, A[:, :4], np.ones(4), project=3)
    op.close()


def _write_panel(prefix, bim, U, S, V, F):
    N, M = U.shape[0], V.shape[0]
    open(prefix + ".sigvals", "w").write(f"#{N},{M}\n" + "".join(f"{x:.10g}\n" for x in S))
    np.savetxt(prefix + ".loadings", V, fmt="%.10g", delimiter="\t")
    np.savetxt(prefix + ".eigvecs", U, fmt="%.10g", delimiter="\t")
    np.savetxt(prefix + ".eigvals", S ** 2 / M, fmt="%.10g")
    with open(prefix + ".mbim", "w") as f:
        for ln, af in zip(open(bim), F):
            f.write(ln.rstrip("\n") + f"\t{af:.10g}\n")


@pytest.mark.parametrize("prec", [_lib.PREC_FP64, _lib.PREC_INT8X3])
def test_projection_options_1_and_2_vs_reference(tmp_path, prec):
    """--project 1 and 2 (Projection.cpp:188-246) of NEW samples with missing calls onto a panel's PCs, with the
    panel's allele frequencies: against the unmodified reference's run_projection on the same files (six-digit text)
    and against a numpy least-squares restatement (1e-9). Option 2 = one product + the missing-call indicator product
    on the device, N small solves on the host."""
    import shutil
    import subprocess
    from conftest import ROOT
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    Np, Nt, M, k = 260, 170, 2100, 4
    panel = str(tmp_path / "panel")
    pk = synth.write_bed(panel, Np, M, k_pop=5, seed=61)
    p = halko.Param(k=k, svd=1, precision=prec, maxp=8, tol=0.0)
    d = halko.FileBed(p, packed=pk, nsamples=Np)
    d.prepare()
    op = halko.NormalRsvdOpData(d, p.k, p.oversamples)
    op.setFlags(False, True)
    op.computeUSV(p.maxp, p.tol)
    F = op.F()
    pc = str(tmp_path / "pc")
    _write_panel(pc, panel + ".bim", op.U, op.S, op.V, F)
    op.close()
    V = np.loadtxt(pc + ".loadings", ndmin=2)
    S = np.loadtxt(pc + ".sigvals", ndmin=1)
    Fp = np.array([float(l.split()[6]) for l in open(pc + ".mbim")])
    # target: other samples of the same populations, 6 % of the calls missing
    codes = np.concatenate([c for _, c in synth.balding_nichols_codes(Nt, M, k_pop=5, seed=61)]).copy()
    codes[np.random.default_rng(3).random(codes.shape) < 0.06] = 1
    pt = synth.pack_codes(codes)
    tgt = str(tmp_path / "tgt")
    synth.write_bed_from_packed(tgt, pt, Nt)
    shutil.copy(panel + ".bim", tgt + ".bim")
    pq = halko.Param(k=k, svd=1, precision=prec)
    dt = halko.FileBed(pq, packed=pt, nsamples=Nt)
    dt.prepare()
    opt = halko.NormalRsvdOpData(dt, pq.k, pq.oversamples)
    od = orc.OracleData(pt, Nt)
    od.F = Fp
    X = od.block(0, M - 1, True)
    miss = (od.codes == 1).T
    W = V * S[None, :]
    for mode in (1, 2):
        ref.run_projection(f"PCAone -b {tgt} --USV {pc} --project {mode} -k {k} -o {tmp_path}/rp{mode} -n 4")
        Ur = np.loadtxt(str(tmp_path / f"rp{mode}.eigvecs"), ndmin=2)
        Ud = downstream.run_projection(opt, V, S, ref_F=Fp, project=mode)
        if mode == 1:
            Un = X @ (V / S[None, :])
        else:
            Un = np.stack([np.linalg.lstsq(W[~miss[i]], X[i, ~miss[i]], rcond=None)[0] for i in range(Nt)])
        assert np.abs(Ud - Un).max() <= 1e-9 * np.abs(Un).max()
        assert np.abs(Ud - Ur).max() <= 2e-5 * np.abs(Ur).max()
        # the front-end
        out = str(tmp_path / f"o{mode}")
        r = subprocess.run([os.path.join(ROOT, "pcaone_b200", "bin", "PCAone-b200"), "-b", tgt, "--USV", pc, "--project", str(mode),
                            "-k", str(k), "-o", out], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        Uc = np.loadtxt(out + ".eigvecs", ndmin=2)
        assert np.abs(Uc - Ur).max() <= 2e-5 * np.abs(Ur).max()
    assert np.abs(downstream.run_projection(opt, V, S, ref_F=Fp, project=2)
                  - downstream.run_projection(opt, V, S, ref_F=Fp, project=1)).max() > 1e-3 * np.abs(Un).max()
    opt.close()
