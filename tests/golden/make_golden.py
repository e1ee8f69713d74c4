"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
(oracle/_ref/libpcaone_ref.so, built from /root/reference by oracle/Makefile).

Run in the build container (where /root/reference exists):
    make -C oracle ref && python tests/golden/make_golden.py
The .npz files are committed; tests never need /root/reference at run time.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from pcaone_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
N, M, K = 67, 320, 3  # N % 4 != 0 on purpose (padding bits), l = 13


def main():
    tmp = tempfile.mkdtemp(prefix="golden_")
    bed = os.path.join(tmp, "g")
    # padding bits set to garbage (code 1 = "missing") to pin that they are ignored
    packed = synth.write_bed(bed, N, M, k_pop=5, seed=11, pad_code=1)
    bedm = os.path.join(tmp, "gm")
    packed_m = synth.write_bed(bedm, N, M, k_pop=5, seed=12, miss=0.1, pad_code=2)
    thr = 1  # single thread: Eigen's summation order is then fixed

    # ---- A: sSVD in-core, per-epoch G/H and final U,S,V at a fixed epoch count
    r = ref.Ref(f"PCAone -b {bed} -k {K} -d 1 -o {tmp}/a --maxp 4 --tol-rsvd 0 -n 1", threads=thr)
    r.new_op()
    omega = ref.init_omega(r.N, r.l, 112, True)
    F, lut = r.F(), r.lookup()
    Xc = r.dataG()  # centred, not yet standardised (read_all)
    G0, H0 = r.gandh(0)
    assert np.array_equal(omega, r.omega())
    Xs = r.dataG()  # standardised in place at pi == 0
    G1, H1 = r.gandh(1)
    omega1 = r.omega()
    r.close()
    r = ref.Ref(f"PCAone -b {bed} -k {K} -d 1 -o {tmp}/a --maxp 4 --tol-rsvd 0 -n 1", threads=thr)
    r.new_op()
    U, S, V = r.compute_usv(4, 0.0)
    r.close()
    np.savez_compressed(os.path.join(OUT, "ssvd_small.npz"), packed=packed, N=N, M=M, k=K, omega=omega, F=F,
                        lookup=lut, X_centered=Xc, X_standardized=Xs, G0=G0, H0=H0, G1=G1, H1=H1,
                        omega1=omega1, U=U, S=S, V=V, maxp=4)

    # ---- B: winSVD in-core, 8 windows, default stopping (runs to 2^pi >= bands)
    r = ref.Ref(f"PCAone -b {bed} -k {K} -d 2 -w 8 -o {tmp}/b --maxp 6 --tol-rsvd 0 -n 1", threads=thr)
    r.new_op()
    U, S, V = r.compute_usv(6, 0.0)
    perm = r.perm_indices()
    r.close()
    assert np.array_equal(perm, ref.permute_indices(M))
    np.savez_compressed(os.path.join(OUT, "winsvd_small.npz"), U=U, S=S, V=V, perm=perm, bands=8, maxp=6)

    # ---- C: winSVD out-of-core (permute_plink + block plan + read_block_initial)
    r = ref.Ref(f"PCAone -b {bed} -k {K} -d 2 -w 8 -m 0.00012 -o {tmp}/c --maxp 6 --tol-rsvd 0 -n 1",
                threads=thr)
    r.new_op()
    start, stop = r.block_plan()
    plan = np.array([r.blocksize, r.nblocks, r.bandFactor])
    blk = r.read_block_initial(int(start[0]), int(stop[0]), True)
    r.close()
    r = ref.Ref(f"PCAone -b {bed} -k {K} -d 2 -w 8 -m 0.00012 -o {tmp}/c --maxp 6 --tol-rsvd 0 -n 1",
                threads=thr)
    r.new_op()
    U, S, V = r.compute_usv(6, 0.0)
    perm_ooc = r.perm_indices()
    F_ooc = r.F()
    r.close()
    np.savez_compressed(os.path.join(OUT, "winsvd_ooc_small.npz"), U=U, S=S, V=V, perm=perm_ooc, F=F_ooc,
                        plan=plan, start=start, stop=stop, block0=blk, memory=0.00012, bands=8, maxp=6)

    # ---- C2: sSVD out-of-core (no permutation), block plan from -m
    r = ref.Ref(f"PCAone -b {bed} -k {K} -d 1 -m 0.00012 -o {tmp}/c2 --maxp 3 --tol-rsvd 0 -n 1", threads=thr)
    r.new_op()
    start, stop = r.block_plan()
    plan = np.array([r.blocksize, r.nblocks, r.bandFactor])
    U, S, V = r.compute_usv(3, 0.0)
    r.close()
    np.savez_compressed(os.path.join(OUT, "ssvd_ooc_small.npz"), U=U, S=S, V=V, plan=plan, start=start,
                        stop=stop, memory=0.00012, maxp=3)

    # ---- D: EMU, sSVD in-core, 10 % missing
    cmd = f"PCAone -b {bedm} -k {K} -d 1 --emu -o {tmp}/d --maxp 3 --tol-rsvd 0 --maxiter 3 -n 1"
    r = ref.Ref(cmd, threads=thr)
    r.new_op()
    mask = r.mask()
    Fm = r.F()
    U, S, V, iters = r.run_em()
    r.close()
    # one read_block_update through the out-of-core reader with a known U,S,V
    r = ref.Ref(f"PCAone -b {bedm} -k {K} -d 1 --emu -m 0.00012 -o {tmp}/d2 --maxp 3 -n 1", threads=thr)
    r.new_op()
    s2, e2 = r.block_plan()
    r.read_block_initial(int(s2[0]), int(e2[0]), False)  # estimates F for block 0
    for b in range(1, r.nblocks):
        r.read_block_initial(int(s2[b]), int(e2[b]), False)
    blk_upd = r.read_block_update(int(s2[0]), int(e2[0]), U, S, V, True)
    r.close()
    np.savez_compressed(os.path.join(OUT, "emu_small.npz"), packed=packed_m, mask=mask, F=Fm, U=U, S=S, V=V,
                        iters=iters, maxp=3, maxiter=3, block0_update=blk_upd, b0=np.array([s2[0], e2[0]]))

    # ---- E: LD r2 on standardized genotypes (--ld --ld-stats 1 -> residuals -> r2)
    r = ref.Ref(f"PCAone -b {bed} -k {K} -d 1 --ld --ld-stats 1 -o {tmp}/e --maxp 2 --tol-rsvd 0 -n 1",
                threads=thr)
    r.new_op()
    r.compute_usv(2, 0.0)
    from oracle.ref import lib
    assert lib().ref_write_residuals(r.h) == 0
    r.close()
    resid = np.fromfile(f"{tmp}/e.residuals", dtype=np.uint8)
    r = ref.Ref(f"PCAone -B {tmp}/e.residuals -F {tmp}/e.mbim --print-r2 --ld-bp 1000 -o {tmp}/e2 -n 1", threads=thr)
    r2, ws, we = r.ld_r2(f"{tmp}/e.mbim", 1000)
    Gres = r.dataG()
    r.close()
    np.savez_compressed(os.path.join(OUT, "ld_small.npz"), residuals_file=resid, r2=r2, ws=ws, we=we,
                        G=Gres.astype(np.float32), ld_bp=1000)

    # ---- F: helper known-answer vectors
    rng = np.random.default_rng(5)
    A = rng.standard_normal((40, 5))
    B = rng.standard_normal((40, 5))
    O2, O1 = ref.flip_omg(A, -A + 0.01 * B)
    Uf, Vf = ref.flip_uv(A, B)
    np.savez_compressed(os.path.join(OUT, "helpers.npz"), A=A, B=B, mev=ref.mev(A, B), flip_omg2=O2,
                        flip_omg=O1, flipU=Uf, flipV=Vf, omega_7x3_seed9=ref.init_omega(7, 3, 9, True),
                        omega_uniform=ref.init_omega(7, 3, 9, False), shuffle10=ref.permute_indices(10),
                        plink_perm=r_perm())
    # ---- G: dense front-end RsvdOne (RSVD.hpp:327-362): tall, wide, plain and windowed
    rng = np.random.default_rng(21)
    def lowrank(r, c, kk):
        return (rng.standard_normal((r, kk)) * np.linspace(20, 5, kk)) @ rng.standard_normal((kk, c)) \
            + 0.05 * rng.standard_normal((r, c))
    At, Aw = lowrank(301, 57, 6), lowrank(45, 260, 5)
    out = {"A_tall": At, "A_wide": Aw, "k": 4, "os": 6,
           "omega_tall": ref.init_omega(57, 10, 1, True), "omega_wide": ref.init_omega(45, 10, 1, True)}
    for name, A in (("tall", At), ("wide", Aw)):
        for p_, w_ in ((3, 0), (5, 4), (3, 8)):
            U, S, V = ref.rsvd_one(A, 4, 6, 1, p=p_, windows=w_)
            out[f"{name}_p{p_}_w{w_}_U"], out[f"{name}_p{p_}_w{w_}_S"], out[f"{name}_p{p_}_w{w_}_V"] = U, S, V
    np.savez_compressed(os.path.join(OUT, "rsvd_one.npz"), **out)
    # ---- E2: LD pruning (ld_prune_big, LD.cpp:240-268) on the same residuals, with the .mbim allele
    # frequencies (lower-MAF SNP of a pair goes) and with a 6-column bim (the partner goes)
    r = ref.Ref(f"PCAone -B {tmp}/e.residuals -F {tmp}/e.mbim --print-r2 --ld-bp 1000 -o {tmp}/e3 -n 1", threads=thr)
    af = np.array([float(ln.split()[6]) for ln in open(f"{tmp}/e.mbim")])
    with open(f"{tmp}/e6.bim", "w") as f6:
        for ln in open(f"{tmp}/e.mbim"):
            f6.write("\t".join(ln.split()[:6]) + "\n")
    pr = {}
    for tol_ in (0.02, 0.1):
        pr[f"keep_af_{tol_}"] = r.ld_prune(f"{tmp}/e.mbim", 1000, tol_, f"{tmp}/p_af_{tol_}")
        pr[f"keep_noaf_{tol_}"] = r.ld_prune(f"{tmp}/e6.bim", 1000, tol_, f"{tmp}/p_no_{tol_}")
    r.close()
    np.savez_compressed(os.path.join(OUT, "ld_prune_small.npz"), af=af, tols=np.array([0.02, 0.1]), **pr)

    # ---- I: Beagle genotype likelihoods, PCAngsd EM (FileBeagle.cpp, Utils.cpp:745-775, Data.cpp:296-316,
    # Halko.cpp:290-311) on a synthetic low-depth beagle.gz written here
    import gzip
    Nb, Mb = 61, 300
    rngb = np.random.default_rng(8)
    codes_b = np.concatenate([c for _, c in synth.balding_nichols_codes(Nb, Mb, k_pop=4, seed=21)])
    gt = np.array([2, 0, 1, 0])[codes_b]
    depth = rngb.poisson(2.0, size=gt.shape)
    alt = rngb.binomial(depth, np.clip(gt / 2.0, 0.01, 0.99))
    lik = [np.clip(q / 2.0, 0.01, 0.99) ** alt * (1 - np.clip(q / 2.0, 0.01, 0.99)) ** (depth - alt) for q in (0, 1, 2)]
    Lb = np.stack(lik, axis=-1)
    Lb = Lb / Lb.sum(-1, keepdims=True)
    bgl = os.path.join(tmp, "g.beagle.gz")
    with gzip.open(bgl, "wt") as f:
        f.write("marker\tallele1\tallele2" + "".join(f"\tInd{i}\tInd{i}\tInd{i}" for i in range(Nb)) + "\n")
        for j in range(Mb):
            f.write(f"chr1_{j + 1}\t0\t1" + "".join("\t%.6f\t%.6f\t%.6f" % tuple(Lb[j, i]) for i in range(Nb)) + "\n")
    r = ref.Ref(f"PCAone --beagle {bgl} -k {K} -d 1 -o {tmp}/b --maxp 3 --tol-rsvd 0 --maxiter 4 -n 1", threads=thr)
    Pb, Fb, E0 = r.P(), r.F(), r.dataG()
    r.new_op()
    omega_b = ref.init_omega(r.N, r.l, 112, True)
    Ub, Sb, Vb, itb = r.run_em()
    r.close()
    np.savez_compressed(os.path.join(OUT, "pcangsd_small.npz"), P=Pb, F=Fb, E0=E0, omega=omega_b, U=Ub, S=Sb, V=Vb,
                        iters=itb, k=K, maxp=3, maxiter=4)

    # ---- H: IRAM operator ArnoldiOpData::perform_op (Arnoldi.cpp:18-46) on the out-of-core plan
    r = ref.Ref(f"PCAone -b {bed} -k {K} -d 0 -m 0.00012 -o {tmp}/h -n 1", threads=thr)
    s3, e3 = r.block_plan()
    xa = np.random.default_rng(4).standard_normal(N)
    np.savez_compressed(os.path.join(OUT, "arnoldi_op.npz"), x=xa, y_std=r.perform_op(xa, False, True),
                        y_raw=r.perform_op(xa, False, False), start=s3, stop=e3, memory=0.00012)
    r.close()
    print("golden written to", OUT)


def r_perm():
    """permute_plink's PermMat for a ragged case (M % bands != 0)."""
    tmp = tempfile.mkdtemp(prefix="golden_p_")
    bed = os.path.join(tmp, "p")
    synth.write_bed(bed, 13, 203, k_pop=2, seed=2)
    r = ref.Ref(f"PCAone -b {bed} -k 2 -d 2 -w 8 -m 0.00001 -o {tmp}/o -n 1", threads=1)
    p = r.perm_indices()
    r.close()
    return p


if __name__ == "__main__":
    main()
