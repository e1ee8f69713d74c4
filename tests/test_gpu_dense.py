"""GPU parity tests (-m gpu) of the generic dense-matrix front-end:
pcaone_b200.rsvd.RsvdOne  <->  PCAone::RsvdOne<MatrixXd> (src/RSVD.hpp:327-362).

Same Omega as the reference (the default-seeded libstdc++ engine reproduced by
pcaone_init_omega), tolerance of north_star: singular values^2 <= 1e-6 relative,
|cos| >= 0.9999 per singular vector — the observed agreement is ~1e-12."""
import numpy as np
import pytest

from conftest import col_cos, golden
from oracle import pcaone_oracle as orc
from pcaone_b200.rsvd import RsvdOne

pytestmark = pytest.mark.gpu


def _close(U, S, V, Ur, Sr, Vr):
    assert U.shape == Ur.shape and V.shape == Vr.shape and S.shape == Sr.shape
    rel = np.max(np.abs(S ** 2 - Sr ** 2) / Sr ** 2)
    cu, cv = col_cos(U, Ur).min(), col_cos(V, Vr).min()
    assert rel <= 1e-6 and cu >= 0.9999 and cv >= 0.9999, (rel, cu, cv)
    return rel, cu, cv


@pytest.mark.parametrize("name", ["tall", "wide"])
@pytest.mark.parametrize("p,w", [(3, 0), (5, 4), (3, 8)])
def test_rsvd_one_vs_reference_golden(name, p, w):
    g = golden("rsvd_one")
    A, k, os_ = g[f"A_{name}"], int(g["k"]), int(g["os"])
    r = RsvdOne(A, k, os_, 1)
    assert np.array_equal(r.Omg, g[f"omega_{name}"])     # bit-identical default-engine stream
    r.compute(p, w)
    res = _close(r.matrixU(), r.singularValues(), r.matrixV(),
                 g[f"{name}_p{p}_w{w}_U"], g[f"{name}_p{p}_w{w}_S"], g[f"{name}_p{p}_w{w}_V"])
    print(name, p, w, res)


@pytest.mark.parametrize("rows,cols,k,os_,p,w", [(3000, 517, 10, 10, 4, 0), (2500, 800, 8, 12, 6, 16),
                                                  (640, 4000, 12, 8, 5, 8), (1031, 1031, 5, 5, 3, 0),
                                                  (5000, 300, 20, 20, 7, 64)])
def test_rsvd_one_vs_numpy_oracle(rows, cols, k, os_, p, w):
    rng = np.random.default_rng(rows + cols)
    kk = k + 3
    A = (rng.standard_normal((rows, kk)) * np.linspace(30, 4, kk)) @ rng.standard_normal((kk, cols)) \
        + 0.1 * rng.standard_normal((rows, cols))
    r = RsvdOne(A, k, os_, 1)
    r.compute(p, w)
    U, S, V = orc.rsvd_one(A, k, os_, r.Omg, p, w)
    res = _close(r.matrixU(), r.singularValues(), r.matrixV(), U, S, V)
    # and against the exact SVD: the matrix has k+3 strong directions, so k of them are recovered
    s = np.linalg.svd(A, compute_uv=False)[:k]
    assert np.max(np.abs(r.singularValues() - s) / s) < 1e-6
    print((rows, cols, k, os_, p, w), res)


def test_rsvd_one_uniform_omega_and_injected_omega():
    rng = np.random.default_rng(3)
    A = rng.standard_normal((700, 90))
    r = RsvdOne(A, 6, 6, 0)                   # rand != 1 -> UniformRandom (RSVD.hpp:126-128)
    assert np.abs(r.Omg).max() <= 1.0
    r.compute(3)
    U, S, V = orc.rsvd_one(A, 6, 6, r.Omg, 3, 0)
    _close(r.matrixU(), r.singularValues(), r.matrixV(), U, S, V)
    om = rng.standard_normal((90, 12))
    r2 = RsvdOne(A, 6, 6, 1, omega=om)
    r2.compute(2)
    U, S, V = orc.rsvd_one(A, 6, 6, om, 2, 0)
    _close(r2.matrixU(), r2.singularValues(), r2.matrixV(), U, S, V)


def test_rsvd_one_argument_errors():
    A = np.random.default_rng(0).standard_normal((400, 60))
    r = RsvdOne(A, 3, 3, 1)
    with pytest.raises(RuntimeError, match="power of 2"):
        r.compute(3, 3)
    with pytest.raises(RuntimeError, match="pow"):
        r.compute(1, 4)
    with pytest.raises(RuntimeError, match="window size"):
        r.compute(6, 64)
    r.setRangeFinder(2)
    with pytest.raises(RuntimeError, match="finder"):
        r.compute(3)
