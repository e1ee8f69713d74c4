"""CPU tests (no GPU): numpy restatements used as checkers by the -m gpu tests, pinned here against the UNMODIFIED
reference run live through oracle/_ref (skipped where it was not built): the exact PCA of `--svd 3`, the projection
options 1 and 2, the BGEN dosage semantics. The -m gpu tests compare the device with the same reference calls; these
keep the checkers honest on a box without a GPU."""
import shutil

import numpy as np
import pytest

from oracle import pcaone_oracle as orc
from pcaone_b200 import synth


def _ref():
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    return ref


def test_exact_pca_restatement_vs_reference(tmp_path):
    """Main.cpp:180-217: K = G G^T / M on the standardised genotypes, eigh, V = G^T U / sqrt(eval M), flip_UV."""
    ref = _ref()
    N, M, k = 150, 1200, 5
    prefix = str(tmp_path / "s")
    pk = synth.write_bed(prefix, N, M, k_pop=4, seed=71)
    r = ref.Ref(f"PCAone -b {prefix} -k {k} -d 1 -o {tmp_path}/r -n 2", threads=2)
    Ur, Sr, Vr, Er = r.full_pca(k)
    r.close()
    od = orc.OracleData(pk, N)
    X = od.block(0, M - 1, True)
    w, Q = np.linalg.eigh(X @ X.T / M)
    E = np.maximum(0.0, w[::-1][:k])
    U = Q[:, ::-1][:, :k]
    S = np.sqrt(E * M)
    V = X.T @ U / S[None, :]
    sg = np.where(U[np.abs(U).argmax(axis=0), np.arange(k)] < 0, -1.0, 1.0)   # flip_UV(U, V), U-based (Utils.cpp:118-133)
    U, V = U * sg, V * sg
    assert np.abs(E - Er).max() <= 1e-12 * Er[0] and np.abs(S - Sr).max() <= 1e-12 * Sr[0]
    assert np.abs(U[:, :3] - Ur[:, :3]).max() < 1e-9 and np.abs(V[:, :3] - Vr[:, :3]).max() < 1e-9


@pytest.mark.parametrize("mode", [1, 2])
def test_projection_restatement_vs_reference(tmp_path, mode):
    """Projection.cpp:188-246: option 1 = G V S^-1, option 2 = per-sample least squares over the called SNPs."""
    ref = _ref()
    Np, Nt, M, k = 120, 90, 900, 3
    panel = str(tmp_path / "panel")
    pk = synth.write_bed(panel, Np, M, k_pop=4, seed=72)
    r = ref.Ref(f"PCAone -b {panel} -k {k} -d 1 -o {tmp_path}/pcr -n 2", threads=2)
    F = r.F()
    r.new_op()
    U, S, V = r.compute_usv(20, 1e-4)
    r.close()
    pc = str(tmp_path / "pc")
    open(pc + ".sigvals", "w").write(f"#{Np},{M}\n" + "".join(f"{x:.10g}\n" for x in S))
    np.savetxt(pc + ".loadings", V, fmt="%.10g", delimiter="\t")
    np.savetxt(pc + ".eigvecs", U, fmt="%.10g", delimiter="\t")
    np.savetxt(pc + ".eigvals", S ** 2 / M, fmt="%.10g")
    with open(pc + ".mbim", "w") as f:
        for ln, af in zip(open(panel + ".bim"), F):
            f.write(ln.rstrip("\n") + f"\t{af:.10g}\n")
    codes = np.concatenate([c for _, c in synth.balding_nichols_codes(Nt, M, k_pop=4, seed=72)]).copy()
    codes[np.random.default_rng(4).random(codes.shape) < 0.05] = 1
    pt = synth.pack_codes(codes)
    tgt = str(tmp_path / "tgt")
    synth.write_bed_from_packed(tgt, pt, Nt)
    shutil.copy(panel + ".bim", tgt + ".bim")
    ref.run_projection(f"PCAone -b {tgt} --USV {pc} --project {mode} -k {k} -o {tmp_path}/rp -n 2", threads=2)
    Ur = np.loadtxt(str(tmp_path / "rp.eigvecs"), ndmin=2)
    Vt, St = np.loadtxt(pc + ".loadings", ndmin=2), np.loadtxt(pc + ".sigvals", ndmin=1)
    od = orc.OracleData(pt, Nt)
    od.F = np.array([float(l.split()[6]) for l in open(pc + ".mbim")])
    X = od.block(0, M - 1, True)
    if mode == 1:
        Un = X @ (Vt / St[None, :])
    else:
        miss = (od.codes == 1).T
        W = Vt * St[None, :]
        Un = np.stack([np.linalg.lstsq(W[~miss[i]], X[i, ~miss[i]], rcond=None)[0] for i in range(Nt)])
    assert np.abs(Un - Ur).max() <= 2e-5 * np.abs(Ur).max()      # the reference writes six significant digits


def test_bgen_dosage_semantics_vs_reference(tmp_path):
    """FileBgen::read_all (FileBgen.cpp:15-72) on a real container: F and the centred matrix from the dosages the
    vendored reader returns, restated in numpy (oracle.dense_from_dosage)."""
    ref = _ref()
    N, M = 90, 300
    rng = np.random.default_rng(9)
    P = rng.dirichlet([2, 2, 2], size=(M, N))
    P[rng.random((M, N)) < 0.04] = np.nan
    path = str(tmp_path / "t.bgen")
    ref.write_bgen(path, P, bit_depth=16)
    dos = ref.bgen_dosages(path, N, M)
    r = ref.Ref(f"PCAone --bgen {path} -k 2 -d 1 -o {tmp_path}/r -n 1", threads=1)
    Fr, Gr = r.F(), r.dataG()
    r.close()
    od = orc.OracleDosageData(dos)
    assert np.abs(od.F - Fr).max() <= 1e-15
    X = od.block(0, M - 1, False)
    assert np.abs(X - Gr).max() <= 1e-15


def test_clump_restatement_vs_reference(tmp_path):
    """oracle.ld_clump against the .clump file the unmodified reference writes (LD.cpp:323-401, :505-540)."""
    ref = _ref()
    N, M = 200, 2200
    prefix = str(tmp_path / "s")
    pk = synth.write_bed(prefix, N, M, k_pop=4, seed=36)
    rng = np.random.default_rng(3)
    keep = set(np.sort(rng.choice(M, size=1800, replace=False)).tolist())
    pv = rng.uniform(0, 0.03, size=M) ** 2
    assoc = str(tmp_path / "gwas.tsv")
    lines, a_chr, a_bp, a_p = [], [], [], []
    bim = [l.split() for l in open(prefix + ".bim")]
    with open(assoc, "w") as f:
        f.write("SNP\tCHR\tBP\tBETA\tP\n")
        for j, t in enumerate(bim):
            if j in keep:
                ln = f"{t[1]}\t{t[0]}\t{t[3]}\t{rng.normal():.4f}\t{pv[j]:.10g}"
                f.write(ln + "\n")
                lines.append(ln)
                a_chr.append(t[0])
                a_bp.append(int(t[3]))
                a_p.append(float(f"{pv[j]:.10g}"))
    args = dict(clump_bp=4000, clump_r2=0.02, p1=1e-4, p2=5e-4)
    r = ref.Ref(f"PCAone -b {prefix} -k 2 -d 1 -o {tmp_path}/r -n 2", threads=2)
    r.ld_clump(prefix + ".bim", assoc, str(tmp_path / "ref.clump"), **args)
    r.close()
    want = open(tmp_path / "ref.clump").read().splitlines()
    od = orc.OracleData(pk, N)
    G = od.block(0, M - 1, False)
    got = orc.ld_clump(G, [t[0] for t in bim], [int(t[3]) for t in bim], lines, a_chr, a_bp, a_p, args["clump_bp"],
                       args["clump_r2"], args["p1"], args["p2"])
    assert len(want) > 20 and sum(1 for l in want[1:] if not l.endswith("NONE")) > 5
    assert got == want[1:]
