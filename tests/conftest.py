import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libpcaone_ref.so (the compiled reference)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def col_corr(A, B):
    """|corr| per column (PCs compared up to sign, BASELINE.json north_star)."""
    A = A - A.mean(0)
    B = B - B.mean(0)
    num = np.abs((A * B).sum(0))
    den = np.linalg.norm(A, axis=0) * np.linalg.norm(B, axis=0)
    return num / den


def col_cos(A, B):
    """|cosine| per column without centring (unit singular vectors)."""
    num = np.abs((A * B).sum(0))
    return num / (np.linalg.norm(A, axis=0) * np.linalg.norm(B, axis=0))


def assert_usv_close(U, S, V, Ur, Sr, Vr, eig_rtol=1e-6, min_corr=0.9999):
    """The tolerance BASELINE.json's north_star states: top-k eigenvalues (S^2/M) <= 1e-6
    relative, PCs |corr| >= 0.9999 up to sign."""
    ev, evr = S ** 2, Sr ** 2
    assert np.max(np.abs(ev - evr) / evr) <= eig_rtol, (ev, evr)
    assert col_cos(U, Ur).min() >= min_corr, col_cos(U, Ur)
    assert col_cos(V, Vr).min() >= min_corr, col_cos(V, Vr)


@pytest.fixture(scope="session")
def tmpdir_session(tmp_path_factory):
    return str(tmp_path_factory.mktemp("pcaone"))
