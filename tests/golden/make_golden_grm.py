"""Golden vectors of the PCAngsd GRM step (Halko.cpp:320-334) from the UNMODIFIED reference (oracle/_ref), on the
same synthetic beagle.gz as tests/golden/pcangsd_small.npz (section I of make_golden.py: same seeds, same command).
Run here (needs /root/reference at oracle build time):  python tests/golden/make_golden_grm.py
Writes tests/golden/pcangsd_grm.npz: C (N x N covariance with the Dc diagonal), U2 / S2 (JacobiSVD of C), Dc."""
import gzip
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from pcaone_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
K = 3


def main():
    base = np.load(os.path.join(OUT, "pcangsd_small.npz"))
    tmp = tempfile.mkdtemp(prefix="golden_grm_")
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
    k = int(base["k"])
    r = ref.Ref(f"PCAone --beagle {bgl} -k {k} -d 1 -o {tmp}/b --maxp 3 --tol-rsvd 0 --maxiter 4 -n 1", threads=1)
    assert np.array_equal(r.P(), base["P"]), "the beagle file of this script must be the one of pcangsd_small.npz"
    r.new_op()
    U, S, V, it = r.run_em()
    assert np.allclose(S, base["S"], rtol=1e-12) and it == int(base["iters"])
    Cm, U2, S2, Dc = r.pcangsd_grm()
    r.close()
    np.savez_compressed(os.path.join(OUT, "pcangsd_grm.npz"), C=Cm, U2=U2, S2=S2, Dc=Dc, U=U, S=S, V=V)
    print("wrote pcangsd_grm.npz: N =", Nb, "top singular values", S2[:4])


if __name__ == "__main__":
    main()
