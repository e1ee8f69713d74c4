"""The C++ front-end `PCAone-b200` (pcaone_b200/host): same flags, same output files as the
reference CLI (src/Main.cpp, src/Cmd.cpp, Data::write_eigs_files). CPU: option handling and the
loud failure without a GPU. GPU (-m gpu): the files it writes against the compiled reference
(oracle/_ref) on the same bed — eigenvalues <= 1e-5 relative (the files carry 6 significant
digits), PCs / loadings |cos| >= 0.9999 up to sign."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, col_cos
from pcaone_b200 import synth

BIN = os.path.join(ROOT, "pcaone_b200", "bin", "PCAone-b200")


def _run(args, cwd=None, ok=True):
    r = subprocess.run([BIN] + [str(a) for a in args], capture_output=True, text=True, cwd=cwd, timeout=600)
    if ok:
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r


def test_cli_binary_is_built_and_prints_reference_flags():
    assert os.path.exists(BIN), "run __graft_entry__.build()"
    out = _run(["--help"]).stdout
    for flag in ("--bfile", "--pc", "--svd", "--memory", "--batches", "--no-shuffle", "--emu", "--maxp", "--tol-rsvd",
                 "--print-r2", "--ld-bp", "--printv", "--out", "--seed", "--oversamples", "--gpus", "--precision", "--beagle",
                 "--pcangsd", "--tol-maf"):
        assert flag in out, flag


def test_cli_rejects_bad_options(tmp_path):
    assert _run(["-b", "x", "-w", "3"], ok=False).returncode != 0            # Cmd.cpp:228 bands rule
    assert _run(["-b", "x", "--svd", "0"], ok=False).returncode != 0          # IRAM is off the GPU path
    r = _run(["-b", "x", "--pgen", "y"], ok=False)
    assert r.returncode != 0 and "outside the B200" in r.stderr
    assert _run(["--beagle", "x.gz", "--emu"], ok=False).returncode != 0
    assert _run(["-b", "x", "-k", "abc"], ok=False).returncode != 0
    assert _run(["--nope"], ok=False).returncode != 0
    # a missing / corrupt bed is an error, not a silent fallback
    r = _run(["-b", str(tmp_path / "missing"), "-o", str(tmp_path / "o")], ok=False)
    assert r.returncode != 0


def test_cli_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    prefix = str(tmp_path / "s")
    synth.write_bed(prefix, 50, 300, k_pop=3, seed=1)
    r = _run(["-b", prefix, "-k", "2", "-o", str(tmp_path / "o")], ok=False)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr


# ------------------------------------------------------------------------------- GPU
def _ref(cmd, maxp, em=False):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    r = ref.Ref(cmd, threads=8)
    r.new_op()
    if em:
        U, S, V, _ = r.run_em()
    else:
        U, S, V = r.compute_usv(maxp, 0.0)
    perm = r.perm_indices() if r.perm else None
    r.close()
    return U, S, V, perm


def _load(prefix, k, M):
    U = np.loadtxt(prefix + ".eigvecs", ndmin=2)
    S = np.loadtxt(prefix + ".sigvals", ndmin=1)
    E = np.loadtxt(prefix + ".eigvals", ndmin=1)
    V = np.loadtxt(prefix + ".loadings", ndmin=2) if os.path.exists(prefix + ".loadings") else None
    assert U.shape[1] == k and S.shape == (k,) and E.shape == (k,)
    assert open(prefix + ".sigvals").readline().startswith("#")
    np.testing.assert_allclose(E, S ** 2 / M, rtol=2e-5)
    return U, S, V


def _check(prefix, k, N, M, Ur, Sr, Vr, perm):
    U, S, V = _load(prefix, k, M)
    assert U.shape[0] == N
    assert np.max(np.abs(S - Sr) / Sr) < 1e-5, (S, Sr)
    assert col_cos(U, Ur).min() > 0.9999, col_cos(U, Ur)
    if V is not None:
        Vo = Vr
        if perm is not None:  # the reference's V is in permuted order; .loadings are un-permuted
            Vo = np.zeros_like(Vr)
            Vo[perm] = Vr
        assert V.shape == Vo.shape
        assert col_cos(V, Vo).min() > 0.9999, col_cos(V, Vo)
    first = open(prefix + ".eigvecs2").readline().split()
    assert first[:3] == ["#FID", "IID", "PC1"] and len(first) == 2 + k
    assert sum(1 for _ in open(prefix + ".eigvecs2")) == N + 1


@pytest.mark.gpu
@pytest.mark.parametrize("svd,extra", [(1, []), (2, ["-S"]), (2, []), (1, ["--precision", "fp64"])])
def test_cli_incore_vs_reference(tmp_path, svd, extra):
    N, M, k, maxp = 600, 9000, 5, 6
    prefix = str(tmp_path / "s")
    synth.write_bed(prefix, N, M, k_pop=7, seed=31)
    ref_extra = " ".join(e for e in extra if e == "-S")
    Ur, Sr, Vr, perm = _ref(f"PCAone -b {prefix} -k {k} -d {svd} -o {tmp_path}/r --maxp {maxp} --tol-rsvd 0 -n 8 {ref_extra}", maxp)
    out = str(tmp_path / "o")
    _run(["-b", prefix, "-k", k, "-d", svd, "-o", out, "--maxp", maxp, "--tol-rsvd", 0, "-V"] + extra)
    _check(out, k, N, M, Ur, Sr, Vr, perm)
    mb = [l.split() for l in open(out + ".mbim")]
    assert len(mb) == M and len(mb[0]) == 7
    # the 7th column is the allele frequency of the ORIGINAL SNP order
    codes = synth.unpack_codes(synth.read_bed(prefix)[0], N)
    g = np.array([1.0, np.nan, 0.5, 0.0])[codes[:50]]
    np.testing.assert_allclose([float(x[6]) for x in mb[:50]], np.nanmean(g, axis=1), rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("svd", [1, 2])
def test_cli_out_of_core_vs_reference(tmp_path, svd):
    N, M, k, maxp, mem = 400, 30000, 4, 7, 0.02
    prefix = str(tmp_path / "s")
    synth.write_bed(prefix, N, M, k_pop=6, seed=32)
    # the reference permutes on disk (<out>.perm.bed); the GPU host applies the same map at read time
    Ur, Sr, Vr, perm = _ref(f"PCAone -b {prefix} -k {k} -d {svd} -m {mem} -o {tmp_path}/r --maxp {maxp} --tol-rsvd 0 -n 8 -w 16", maxp)
    out = str(tmp_path / "o")
    r = _run(["-b", prefix, "-k", k, "-d", svd, "-m", mem, "-o", out, "--maxp", maxp, "--tol-rsvd", 0, "-V", "-w", 16])
    assert "blocksize" in r.stdout
    _check(out, k, N, M, Ur, Sr, Vr, perm)
    assert not os.path.exists(out + ".perm.bed")


@pytest.mark.gpu
def test_cli_emu_vs_reference(tmp_path):
    N, M, k = 500, 6000, 3
    prefix = str(tmp_path / "s")
    synth.write_bed(prefix, N, M, k_pop=5, seed=33, miss=0.08)
    Ur, Sr, Vr, perm = _ref(f"PCAone -b {prefix} -k {k} -d 1 --emu -o {tmp_path}/r -n 8", 20, em=True)
    out = str(tmp_path / "o")
    r = _run(["-b", prefix, "-k", k, "-d", 1, "--emu", "-o", out, "-V"])
    assert "missingness" in r.stdout
    U, S, V = _load(out, k, M)
    assert np.max(np.abs(S - Sr) / Sr) < 1e-4
    assert col_cos(U, Ur).min() > 0.9999


@pytest.mark.gpu
def test_cli_print_r2_vs_reference(tmp_path):
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    N, M = 300, 1500
    prefix = str(tmp_path / "s")
    synth.write_bed(prefix, N, M, k_pop=4, seed=34)
    out = str(tmp_path / "o")
    _run(["-b", prefix, "--print-r2", "--ld-bp", 3000, "-o", out])
    lines = gzip.open(out + ".ld.gz", "rt").read().splitlines()
    assert lines[0].split("\t") == ["CHR_A", "BP_A", "SNP_A", "CHR_B", "BP_B", "SNP_B", "R2"]
    r2 = np.array([float(l.split("\t")[6]) for l in lines[1:]])
    # reference: in-core bed, centred genotypes, ld_r2_big windows from the same .bim
    r = ref.Ref(f"PCAone -b {prefix} -k 2 -d 1 -o {tmp_path}/r -n 4", threads=4)
    want, ws, we = r.ld_r2(prefix + ".bim", 3000)
    r.close()
    assert r2.shape == want.shape
    assert np.abs(r2 - want).max() < 2e-6  # std::to_string keeps 6 decimals


@pytest.mark.gpu
def test_cli_ld_prune_vs_reference(tmp_path):
    """--ld-r2: .ld.prune.in / .ld.prune.out against the reference's ld_prune_big on the same bed
    (centred genotypes read in core), without and with the allele-frequency column."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    N, M, tol = 300, 1500, 0.03
    prefix = str(tmp_path / "s")
    synth.write_bed(prefix, N, M, k_pop=4, seed=35)
    r = ref.Ref(f"PCAone -b {prefix} -k 2 -d 1 -o {tmp_path}/r -n 4", threads=4)
    F = r.F()
    mbim = str(tmp_path / "s.mbim")
    with open(mbim, "w") as f:
        for ln, af in zip(open(prefix + ".bim"), F):
            f.write(ln.rstrip("\n") + "\t%.6g\n" % af)
    for bim, tag in ((prefix + ".bim", "noaf"), (mbim, "af")):
        want = r.ld_prune(bim, 3000, tol, str(tmp_path / ("ref_" + tag)))
        out = str(tmp_path / ("o_" + tag))
        _run(["-b", prefix, "--ld-r2", tol, "--ld-bp", 3000, "-F", bim, "-o", out])
        kept = [l.split("\t")[1] for l in open(out + ".ld.prune.in")]
        dropped = [l.split("\t")[1] for l in open(out + ".ld.prune.out")]
        ids = [l.split()[1] for l in open(prefix + ".bim")]
        assert len(kept) + len(dropped) == M and 0 < len(dropped) < M
        got = np.isin(ids, kept)
        assert np.array_equal(got, want), (tag, int((got != want).sum()))
    r.close()


def _write_beagle(path, P):
    """BEAGLE text (gz) from the 2N x M likelihood matrix: the columns parse_beagle_file reads."""
    N, M = P.shape[0] // 2, P.shape[1]
    with gzip.open(path, "wt") as f:
        f.write("marker\tallele1\tallele2" + "".join(f"\tInd{i}\tInd{i}\tInd{i}" for i in range(N)) + "\n")
        for j in range(M):
            p0, p1 = P[0::2, j], P[1::2, j]
            f.write(f"chr1_{j + 1}\t0\t1" + "".join("\t%.6f\t%.6f\t%.6f" % (a, b, max(1 - a - b, 0.0)) for a, b in zip(p0, p1)) + "\n")


@pytest.mark.gpu
def test_cli_beagle_pcangsd_vs_reference_golden(tmp_path):
    """--beagle: gz parsing on the host, PCAngsd EM on the device; against U, S of the unmodified
    reference on the same likelihoods (tests/golden/pcangsd_small.npz, six-decimal text round trip)."""
    from conftest import golden
    g = golden("pcangsd_small")
    P, k = g["P"], int(g["k"])
    N, M = P.shape[0] // 2, P.shape[1]
    bgl = str(tmp_path / "g.beagle.gz")
    _write_beagle(bgl, P)
    out = str(tmp_path / "o")
    r = _run(["--beagle", bgl, "-k", k, "-d", 1, "-o", out, "--maxp", int(g["maxp"]), "--tol-rsvd", 0, "--maxiter",
              int(g["maxiter"]), "-V"])
    U, S, V = _load(out, k, M)
    assert U.shape[0] == N and V.shape == (M, k)
    assert np.max(np.abs(S - g["S"]) / g["S"]) < 1e-5
    assert col_cos(U, g["U"]).min() > 0.9999 and col_cos(V, g["V"]).min() > 0.9999
    # the GRM step (Halko.cpp:320-334): <out>.cov and the N eigenvectors of it in <out>.eigvecs2
    q = golden("pcangsd_grm")
    Cm = np.loadtxt(out + ".cov")
    assert Cm.shape == (N, N) and np.abs(Cm - q["C"]).max() < 2e-5 * np.abs(q["C"]).max()
    rows = open(out + ".eigvecs2").read().splitlines()
    assert rows[0].split("\t")[:3] == ["#FID", "IID", "PC1"] and len(rows) == N + 1 and rows[1].split("\t")[0] == "Ind0"
    E2 = np.array([[float(x) for x in r.split("\t")[2].split()] for r in rows[1:]])
    assert E2.shape == (N, N) and col_cos(E2[:, :3], q["U2"][:, :3]).min() > 0.9999
    # winSVD with the in-core shuffle runs too and finds the same top PCs
    out2 = str(tmp_path / "o2")
    _run(["-G", bgl, "-k", k, "-d", 2, "-w", 8, "-o", out2, "--maxiter", 4, "-V"])
    U2, S2, V2 = _load(out2, k, M)
    assert col_cos(U2[:, :1], U[:, :1]).min() > 0.99 and V2.shape == (M, k)
    assert _run(["--beagle", bgl, "-m", "0.001", "-o", out], ok=False).returncode != 0   # Cmd.cpp:233


@pytest.mark.gpu
@pytest.mark.parametrize("mem,svd", [(0, 1), (0.02, 1), (0.02, 2)])
def test_cli_ld_residuals_vs_reference(tmp_path, mem, svd):
    """--ld [-m]: <out>.residuals (Data::write_residuals, Data.cpp:242-291 — float32 rows at their ORIGINAL SNP
    positions, un-permuted by seek) against the file the unmodified reference writes for the same command.
    The float rows come from pcaone_residuals_block (resident, or streamed through the block plan's buffers)."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    N, M, k, maxp = 300, 12000, 3, 7
    prefix = str(tmp_path / "s")
    synth.write_bed(prefix, N, M, k_pop=5, seed=41, miss=0.01)
    memf = f"-m {mem}" if mem else ""
    r = ref.Ref(f"PCAone -b {prefix} -k {k} -d {svd} {memf} --ld -o {tmp_path}/r --maxp {maxp} --tol-rsvd 0 -n 8 -w 16", threads=8)
    r.new_op()
    r.compute_usv(maxp, 0.0)
    r.write_residuals()
    r.close()
    out = str(tmp_path / "o")
    _run(["-b", prefix, "-k", k, "-d", svd, "--ld", "-o", out, "--maxp", maxp, "--tol-rsvd", 0, "-w", 16,
          "--precision", "fp64"] + (["-m", mem] if mem else []))
    a = np.fromfile(f"{tmp_path}/r.residuals", dtype=np.uint8)
    b = np.fromfile(out + ".residuals", dtype=np.uint8)
    assert a.size == b.size == 8 + 4 * N * M and np.array_equal(a[:8], b[:8])
    ra, rb = a[8:].view(np.float32).reshape(M, N), b[8:].view(np.float32).reshape(M, N)
    # U S V^T of two runs of a randomized SVD agree to ~1e-9 relative (7 epochs, tol 0): compare the residual rows
    assert np.abs(ra - rb).max() <= 1e-4 * np.abs(ra).max()
    assert np.array_equal(open(f"{tmp_path}/r.mbim").read().split()[:14], open(out + ".mbim").read().split()[:14])


@pytest.mark.gpu
def test_cli_large_k_falls_back_to_fp64(tmp_path):
    """-k 45 with the default --precision int8x3 needs 3 x 90 > 256 UMMA columns: the library runs the FP64
    kernels instead of refusing (the reference has no bound on k) and the front-end says so."""
    N, M, k = 300, 4000, 45
    prefix = str(tmp_path / "s")
    synth.write_bed(prefix, N, M, k_pop=50, seed=43)
    out = str(tmp_path / "o")
    r = _run(["-b", prefix, "-k", k, "-d", 1, "-o", out, "--maxp", 3, "--tol-rsvd", 0])
    assert "FP64" in r.stdout + r.stderr
    S = np.loadtxt(out + ".sigvals", ndmin=1)
    assert S.shape == (k,) and np.all(np.diff(S) <= 0)


@pytest.mark.gpu
def test_cli_two_gpus_in_one_process(tmp_path):
    """--gpus 2: one host thread per GPU in ONE process (per-context kernel attributes, library-side NCCL
    through pcaone_comm_attach) equals the one-GPU run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    N, M, k, maxp = 500, 20000, 5, 6
    prefix = str(tmp_path / "s")
    synth.write_bed(prefix, N, M, k_pop=7, seed=44)
    one, two = str(tmp_path / "o1"), str(tmp_path / "o2")
    _run(["-b", prefix, "-k", k, "-d", 2, "-S", "-o", one, "--maxp", maxp, "--tol-rsvd", 0, "-V"])
    _run(["-b", prefix, "-k", k, "-d", 2, "-S", "-o", two, "--maxp", maxp, "--tol-rsvd", 0, "-V", "--gpus", 2])
    U1, S1, V1 = _load(one, k, M)
    U2, S2, V2 = _load(two, k, M)
    assert np.max(np.abs(S1 - S2) / S1) < 1e-5
    assert col_cos(U1, U2).min() > 0.9999 and col_cos(V1, V2).min() > 0.9999


@pytest.mark.gpu
def test_cli_ld_two_step_through_residual_file(tmp_path):
    """config-5 flow of the reference, both commands on the GPU front-end: `--ld` writes <out>.residuals + .mbim,
    `-B <out>.residuals -F <out>.mbim --print-r2` reads them back (FileBin) — against the reference's r2 for the same
    two commands (the residual rows come from two independent randomized SVDs: loose tolerance on r2)."""
    import gzip
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    N, M, k = 250, 3000, 3
    prefix = str(tmp_path / "s")
    synth.write_bed(prefix, N, M, k_pop=5, seed=45)
    r = ref.Ref(f"PCAone -b {prefix} -k {k} -d 1 --ld -o {tmp_path}/r --maxp 8 --tol-rsvd 0 -n 4", threads=4)
    r.new_op()
    r.compute_usv(8, 0.0)
    r.write_residuals()
    r.close()
    r2 = ref.Ref(f"PCAone -B {tmp_path}/r.residuals -F {tmp_path}/r.mbim --print-r2 --ld-bp 2000 -o {tmp_path}/r2 -n 4", threads=4)
    want, ws, we = r2.ld_r2(f"{tmp_path}/r.mbim", 2000)
    r2.close()
    out = str(tmp_path / "o")
    _run(["-b", prefix, "-k", k, "-d", 1, "--ld", "-o", out, "--maxp", 8, "--tol-rsvd", 0, "--precision", "fp64"])
    _run(["-B", out + ".residuals", "-F", out + ".mbim", "--print-r2", "--ld-bp", 2000, "-o", out + "2"])
    lines = gzip.open(out + "2.ld.gz", "rt").read().splitlines()
    assert lines[0].split("\t")[-1] == "R2" and len(lines) == len(want) + 1
    got = np.array([float(x.split("\t")[-1]) for x in lines[1:]])
    assert np.abs(got - want).max() < 2e-5   # 6 decimals in the text + two independent RSVD runs behind the residuals


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["bed", "residuals"])
def test_cli_ld_clump_vs_reference(tmp_path, source):
    """--clump: <out>.p0.clump against the file the unmodified reference writes (ld_clump_single_pheno, LD.cpp:323-401,
    with its bookkeeping of chromosome runs) for the same association table — from the centred genotypes of the bed and
    from a `-B` residual file. The r2 values come from one banded tile Gram on the device (one forward window per
    candidate), the greedy pass runs on the host."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    N, M = 300, 3000
    prefix = str(tmp_path / "s")
    synth.write_bed(prefix, N, M, k_pop=4, seed=36)
    rng = np.random.default_rng(3)
    assoc = str(tmp_path / "gwas.tsv")
    keep = np.sort(rng.choice(M, size=2400, replace=False))        # the table lists a subset of the variants ...
    pv = rng.uniform(0, 0.03, size=M) ** 2                          # ... about half of them under p2, a tenth under p1
    with open(assoc, "w") as f:
        f.write("SNP\tCHR\tBP\tBETA\tP\n")
        for j, ln in enumerate(open(prefix + ".bim")):
            if j in set(keep.tolist()):
                c, rs, _, bp, _, _ = ln.split()
                f.write(f"{rs}\t{c}\t{bp}\t{rng.normal():.4f}\t{pv[j]:.10g}\n")
        f.write(f"rsX\t22\t999999999\t0.1\t1e-9\n")           # a position the bim does not have
    args = dict(clump_bp=4000, clump_r2=0.02, p1=1e-4, p2=5e-4)
    out = str(tmp_path / "o")
    if source == "bed":
        r = ref.Ref(f"PCAone -b {prefix} -k 2 -d 1 -o {tmp_path}/r -n 4", threads=4)
        r.ld_clump(prefix + ".bim", assoc, str(tmp_path / "ref.clump"), **args)
        r.close()
        _run(["-b", prefix, "--clump", assoc, "--clump-bp", 4000, "--clump-r2", 0.02, "--clump-p1", 1e-4, "--clump-p2", 5e-4,
              "-o", out])
    else:
        _run(["-b", prefix, "-k", 3, "-d", 1, "--ld", "-o", str(tmp_path / "pc")])
        r = ref.Ref(f"PCAone -B {tmp_path}/pc.residuals -F {tmp_path}/pc.mbim --print-r2 -o {tmp_path}/r -n 4", threads=4)
        r.ld_clump(str(tmp_path / "pc.mbim"), assoc, str(tmp_path / "ref.clump"), **args)
        r.close()
        _run(["-B", str(tmp_path / "pc.residuals"), "-F", str(tmp_path / "pc.mbim"), "--clump", assoc, "--clump-bp", 4000,
              "--clump-r2", 0.02, "--clump-p1", 1e-4, "--clump-p2", 5e-4, "-o", out])
    want = open(tmp_path / "ref.clump").read().splitlines()
    got = open(out + ".p0.clump").read().splitlines()
    assert want[0] == got[0] and want[0].endswith("\tSP2")
    n_clumped = sum(1 for l in want[1:] if not l.endswith("NONE"))
    assert len(want) > 30 and n_clumped > 5, (len(want), n_clumped)
    assert got == want


@pytest.mark.gpu
def test_cli_print_r2_with_usv_projection(tmp_path):
    """-b + --USV + --print-r2: LD of (I - U U^T) G (LD.cpp:491-496) with U read from <prefix>.eigvecs, the projection
    applied on the device; against the numpy restatement on the same U."""
    from oracle import pcaone_oracle as orc
    N, M = 300, 1500
    prefix = str(tmp_path / "s")
    packed = synth.write_bed(prefix, N, M, k_pop=4, seed=37)
    pc = str(tmp_path / "pc")
    _run(["-b", prefix, "-k", 3, "-d", 1, "-o", pc])
    out = str(tmp_path / "o")
    _run(["-b", prefix, "--USV", pc, "-F", prefix + ".bim", "--print-r2", "--ld-bp", 3000, "-o", out])
    lines = gzip.open(out + ".ld.gz", "rt").read().splitlines()
    r2 = np.array([float(l.split("\t")[6]) for l in lines[1:]])
    U = np.loadtxt(pc + ".eigvecs", ndmin=2)
    od = orc.OracleData(packed, N)
    G = od.block(0, M - 1, False)
    Gp = G - U @ (U.T @ G)
    chrom = np.array([l.split()[0] for l in open(prefix + ".bim")])
    pos = np.array([int(l.split()[3]) for l in open(prefix + ".bim")])
    ws, we = orc.ld_windows(chrom, pos, 3000)
    want = orc.ld_r2(Gp, ws, we)
    assert r2.shape == want.shape
    assert np.abs(r2 - want).max() < 2e-6  # std::to_string keeps 6 decimals
    plain = orc.ld_r2(G, ws, we)
    assert np.abs(plain - want).max() > 1e-3, "the projection must matter in this case"


@pytest.mark.gpu
def test_cli_svd3_exact_pca_vs_reference(tmp_path):
    """--svd 3: .eigvals / .sigvals / .eigvecs / .loadings of the exact PCA against the reference's FULL branch."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    N, M, k = 320, 2500, 6
    prefix = str(tmp_path / "s")
    synth.write_bed(prefix, N, M, k_pop=5, seed=42)
    r = ref.Ref(f"PCAone -b {prefix} -k {k} -d 1 -o {tmp_path}/r -n 8", threads=8)
    Ur, Sr, Vr, Er = r.full_pca(k)
    r.close()
    out = str(tmp_path / "o")
    _run(["-b", prefix, "-k", k, "--svd", 3, "-V", "-o", out])
    U, S, V = _load(out, k, M)
    np.testing.assert_allclose(S, Sr, rtol=2e-5)
    np.testing.assert_allclose(np.loadtxt(out + ".eigvals"), Er, rtol=2e-5)
    assert np.abs(U[:, :4] - Ur[:, :4]).max() < 2e-5 and np.abs(V[:, :4] - Vr[:, :4]).max() < 2e-5   # 6 significant digits
    assert col_cos(U, Ur).min() > 0.99999 and col_cos(V, Vr).min() > 0.99999


@pytest.mark.gpu
def test_cli_bgen_vs_reference(tmp_path):
    """--bgen: the front-end parses the container itself (host/bgen.cpp) and runs the dosage path on the device;
    U, S, V against the unmodified reference reading the same file with its vendored bgen library."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    N, M, k = 220, 1600, 4
    rng = np.random.default_rng(91)
    codes = np.concatenate([c for _, c in synth.balding_nichols_codes(N, M, k_pop=5, seed=92)])
    g = np.array([2, 0, 1, 0])[codes]
    P = np.zeros((M, N, 3))
    for q in range(3):
        P[:, :, q] = np.where(g == q, 0.85, 0.075)
    P = 0.7 * P + 0.3 * rng.dirichlet([8, 8, 8], size=(M, N))
    P /= P.sum(-1, keepdims=True)
    P[rng.random((M, N)) < 0.02] = np.nan
    path = str(tmp_path / "f.bgen")
    ref.write_bgen(path, P, bit_depth=16, compression=2)
    r = ref.Ref(f"PCAone --bgen {path} -k {k} -d 1 -o {tmp_path}/r -n 4 --maxp 6 --tol-rsvd 0", threads=4)
    r.new_op()
    Ur, Sr, Vr = r.compute_usv(6, 0.0)
    r.close()
    out = str(tmp_path / "o")
    _run(["--bgen", path, "-k", k, "-d", 1, "--maxp", 6, "--tol-rsvd", 0, "-V", "-o", out])
    U, S, V = _load(out, k, M)
    np.testing.assert_allclose(S, Sr, rtol=2e-5)
    assert col_cos(U, Ur).min() > 0.99999 and col_cos(V, Vr).min() > 0.99999
    assert _run(["--bgen", path, "--emu", "-o", out], ok=False).returncode != 0


@pytest.mark.gpu
@pytest.mark.parametrize("scale,svd", [(0, 1), (1, 2), (2, 1), (2, 2)])
def test_cli_csv_vs_reference(tmp_path, scale, svd):
    """--csv: a zstd-compressed count matrix (features x samples), the -C normalisations and the per-feature
    standardisation on the host, the dense FP64 products on the device; U, S, V against the unmodified reference's
    FileCsv on the same file. More samples than features too (the operand is never transposed)."""
    import pyarrow as pa
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    for N, M in ((120, 900), (700, 300)):
        rng = np.random.default_rng(7 + scale + N)
        grp = rng.integers(0, 3, size=N)
        prog = rng.gamma(2.0, 2.0, size=(M, 3))
        cnt = rng.poisson(prog[:, grp] * np.exp(rng.normal(0, 0.4, size=(1, N))))
        path = str(tmp_path / f"c{N}.csv.zst")
        text = "".join(",".join(str(int(x)) for x in row) + "\n" for row in cnt)
        open(path, "wb").write(pa.Codec("zstd").compress(text.encode(), asbytes=True))
        k = 3
        r = ref.Ref(f"PCAone --csv {path} -k {k} -d {svd} -S -w 8 -C {scale} -o {tmp_path}/r -n 4 --maxp 7 --tol-rsvd 0", threads=4)
        r.new_op()
        Ur, Sr, Vr = r.compute_usv(7, 0.0)
        r.close()
        out = str(tmp_path / f"o{N}")
        _run(["--csv", path, "-k", k, "-d", svd, "-S", "-w", 8, "-C", scale, "--maxp", 7, "--tol-rsvd", 0, "-V", "-o", out])
        U, S, V = np.loadtxt(out + ".eigvecs", ndmin=2), np.loadtxt(out + ".sigvals", ndmin=1), np.loadtxt(out + ".loadings", ndmin=2)
        assert U.shape == (N, k) and V.shape == (M, k)
        np.testing.assert_allclose(S, Sr, rtol=2e-5)
        assert col_cos(U, Ur).min() > 0.99999 and col_cos(V, Vr).min() > 0.99999
    assert _run(["--csv", path, "-C", 3, "-o", out], ok=False).returncode != 0
    assert _run(["--csv", path, "--emu", "-o", out], ok=False).returncode != 0
