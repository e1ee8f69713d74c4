"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/_ref/libpcaone_ref.so (the
unmodified reference compiled by oracle/Makefile + our driver oracle/ref_shim.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libpcaone_ref.so")

_dp = C.POINTER(C.c_double)
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_open.restype = C.c_void_p
        L.ref_open.argtypes = [C.c_char_p]
        L.ref_last_error.restype = C.c_char_p
        L.ref_time_gandh.restype = C.c_double
        L.ref_mev.restype = C.c_double
        L.ref_dataG_cols.restype = C.c_longlong
        L.ref_perm_size.restype = C.c_longlong
        L.ref_missing_count.restype = C.c_longlong
        L.ref_ld_r2.restype = C.c_longlong
        L.ref_get_P.restype = C.c_longlong
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f(shape):
    return np.zeros(shape, dtype=np.float64, order="F")


class Ref:
    """One reference run: Param + FileBed (+permutation) + prepare (+ op)."""

    def __init__(self, cmdline: str, threads: int | None = None):
        L = lib()
        if threads:
            L.ref_set_threads(int(threads))
        self.h = L.ref_open(cmdline.encode())
        if not self.h:
            raise RuntimeError("reference failed: " + L.ref_last_error().decode())
        self.h = C.c_void_p(self.h)
        d = (C.c_longlong * 10)()
        L.ref_dims(self.h, d)
        (self.N, self.M, self.k, self.l, self.nblocks, self.blocksize, self.bandFactor,
         self.bands, self.out_of_core, self.perm) = [int(x) for x in d]

    def _chk(self, rc):
        if rc:
            raise RuntimeError("reference failed: " + lib().ref_last_error().decode())

    def close(self):
        if self.h:
            lib().ref_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def new_op(self):
        self._chk(lib().ref_new_op(self.h))
        d = (C.c_longlong * 10)()
        lib().ref_dims(self.h, d)
        self.N, self.M = int(d[0]), int(d[1])

    def block_plan(self):
        s = np.zeros(self.nblocks, dtype=np.uint32)
        e = np.zeros(self.nblocks, dtype=np.uint32)
        lib().ref_block_plan(self.h, _p(s), _p(e))
        return s, e

    def F(self):
        out = np.zeros(self.M)
        lib().ref_get_F(self.h, _p(out))
        return out

    def lookup(self):
        out = _f((4, self.M))
        lib().ref_get_lookup(self.h, _p(out))
        return out

    def dataG(self):
        cols = int(lib().ref_dataG_cols(self.h))
        out = _f((self.N, cols))
        lib().ref_get_dataG(self.h, _p(out))
        return out

    def perm_indices(self):
        n = int(lib().ref_perm_size(self.h))
        out = np.zeros(n, dtype=np.int32)
        if n:
            lib().ref_get_perm(self.h, _p(out))
        return out

    def mask(self):
        out = np.zeros(self.M * self.N, dtype=np.uint8)
        lib().ref_get_mask(self.h, _p(out))
        return out.reshape(self.M, self.N)

    def set_flags(self, update, standardize):
        lib().ref_set_flags(self.h, int(update), int(standardize))

    def omega(self):
        out = _f((self.N, self.l))
        lib().ref_get_omg(self.h, _p(out))
        return out

    def gandh(self, pi):
        G = _f((self.M, self.l))
        H = _f((self.N, self.l))
        self._chk(lib().ref_gandh(self.h, int(pi), _p(G), _p(H)))
        return G, H

    def time_gandh(self, pi):
        t = lib().ref_time_gandh(self.h, int(pi))
        if t < 0:
            self._chk(1)
        return t

    def compute_usv(self, maxp, tol, want=True):
        if not want:  # timing runs: leave U, S, V inside the reference object
            self._chk(lib().ref_compute_usv(self.h, int(maxp), C.c_double(tol), None, None, None))
            return None
        U, S, V = _f((self.N, self.k)), np.zeros(self.k), _f((self.M, self.k))
        self._chk(lib().ref_compute_usv(self.h, int(maxp), C.c_double(tol), _p(U), _p(S), _p(V)))
        return U, S, V

    def write_residuals(self):
        """Data::write_residuals (Data.cpp:242-291) with the op's U, S, V -> <fileout>.residuals + .mbim."""
        self._chk(lib().ref_write_residuals(self.h))

    def last_epochs(self):
        return int(lib().ref_last_epochs(self.h))

    def compute_u(self, G, H):
        G = np.asfortranarray(G, dtype=np.float64)
        H = np.asfortranarray(H, dtype=np.float64)
        U = _f((self.N, self.k))
        self._chk(lib().ref_compute_u(self.h, _p(G), _p(H), _p(U)))
        return U

    def run_em(self):
        U, S, V = _f((self.N, self.k)), np.zeros(self.k), _f((self.M, self.k))
        it = C.c_int(0)
        self._chk(lib().ref_run_em(self.h, _p(U), _p(S), _p(V), C.byref(it)))
        return U, S, V, it.value

    def set_usv(self, U, S, V):
        U = np.asfortranarray(U, dtype=np.float64)
        V = np.asfortranarray(V, dtype=np.float64)
        S = np.ascontiguousarray(S, dtype=np.float64)
        lib().ref_set_usv(self.h, _p(U), _p(S), _p(V))

    def read_block_initial(self, start, stop, standardize):
        out = _f((self.N, stop - start + 1))
        self._chk(lib().ref_read_block_initial(self.h, C.c_ulonglong(start), C.c_ulonglong(stop),
                                               int(standardize), _p(out)))
        return out

    def read_block_update(self, start, stop, U, S, V, standardize):
        U = np.asfortranarray(U, dtype=np.float64)
        V = np.asfortranarray(V, dtype=np.float64)
        S = np.ascontiguousarray(S, dtype=np.float64)
        out = _f((self.N, stop - start + 1))
        self._chk(lib().ref_read_block_update(self.h, C.c_ulonglong(start), C.c_ulonglong(stop), _p(U), _p(S),
                                              _p(V), int(S.shape[0]), int(standardize), _p(out)))
        return out

    def fit_with_pi(self, U, S, V):
        U = np.asfortranarray(U, dtype=np.float64)
        V = np.asfortranarray(V, dtype=np.float64)
        S = np.ascontiguousarray(S, dtype=np.float64)
        self._chk(lib().ref_fit_with_pi(self.h, _p(U), _p(S), _p(V), int(S.shape[0])))

    def standardize_E(self):
        self._chk(lib().ref_standardize_E(self.h))

    def ld_prune(self, filebim, ld_bp, r2_tol, fileout):
        """ld_prune_big on data->G (LD.cpp:240-268) -> boolean keep mask."""
        keep = np.zeros(int(self.M), dtype=np.uint8)
        self._chk(lib().ref_ld_prune(self.h, filebim.encode(), int(ld_bp), C.c_double(r2_tol), fileout.encode(),
                                     _p(keep), C.c_longlong(int(self.M))))
        return keep.astype(bool)

    def ld_clump(self, filebim, assoc, fileout, colnames="CHR,BP,P", clump_bp=250000, clump_r2=0.5, p1=1e-4, p2=1e-2):
        """the clump branch of run_ld_stuff (LD.cpp:505-540) on data->G -> writes `fileout` like the reference."""
        self._chk(lib().ref_ld_clump(self.h, filebim.encode(), assoc.encode(), colnames.encode(), int(clump_bp),
                                     C.c_double(clump_r2), C.c_double(p1), C.c_double(p2), fileout.encode()))

    def full_pca(self, k):
        """the FULL branch of main() (Main.cpp:180-217, --svd 3) -> U (N x k), svals, V (M x k), evals."""
        k = min(int(k), int(self.N), int(self.M))
        U, V = _f((int(self.N), k)), _f((int(self.M), k))
        S, E = np.zeros(k), np.zeros(k)
        self._chk(lib().ref_full_pca(self.h, k, _p(U), _p(S), _p(V), _p(E)))
        return U, S, V, E

    def pcangsd_grm(self):
        """the GRM step after run_em on a Beagle run (Halko.cpp:320-334) -> C (N x N), U2 (JacobiSVD matrixU), S2, Dc."""
        n = int(self.N)
        Cm, U2, S2, Dc = _f((n, n)), _f((n, n)), np.zeros(n), np.zeros(n)
        self._chk(lib().ref_pcangsd_grm(self.h, _p(Cm), _p(U2), _p(S2), _p(Dc)))
        return Cm, U2, S2, Dc

    def P(self):
        """Beagle input: the 2N x M likelihood matrix parse_beagle_file filled."""
        n = lib().ref_get_P(self.h, None)
        out = _f((2 * self.N, n // (2 * self.N)))
        lib().ref_get_P(self.h, _p(out))
        return out

    def perform_op(self, x, update=False, standardize=True):
        """ArnoldiOpData(data).perform_op(x) on this (out-of-core) run -> y (N)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        self._chk(lib().ref_perform_op(self.h, int(update), int(standardize), _p(x), _p(y)))
        return y

    def ld_r2(self, filebim, ld_bp):
        nwin = C.c_longlong(0)
        n = lib().ref_ld_r2(self.h, filebim.encode(), int(ld_bp), None, C.c_longlong(0), None, None,
                            C.byref(nwin))
        if n < 0:
            self._chk(1)
        out = np.zeros(n)
        ws = np.zeros(nwin.value, dtype=np.int32)
        we = np.zeros(nwin.value, dtype=np.int32)
        lib().ref_ld_r2(self.h, filebim.encode(), int(ld_bp), _p(out), C.c_longlong(n), _p(ws), _p(we),
                        C.byref(nwin))
        return out, ws, we


def run_projection(cmdline, threads=4):
    """the projection branch of main() (Main.cpp:95-107) on a PLINK input: writes <out>.eigvecs like the reference."""
    lib().ref_set_threads(int(threads))
    if lib().ref_run_projection(cmdline.encode()):
        raise RuntimeError("reference failed: " + lib().ref_last_error().decode())


def write_bgen(path, probs, bit_depth=8, compression=1, layout=2):
    """probs: (M, N, 3) genotype probabilities (NaN triple = missing) -> a BGEN file (layout 1 / 2, compression
    0 none / 1 zlib / 2 zstd) written by the writer of the reference's vendored bgen library."""
    probs = np.ascontiguousarray(probs, dtype=np.float64)
    M, N, _ = probs.shape
    if lib().ref_write_bgen(path.encode(), C.c_longlong(N), C.c_longlong(M), _p(probs), int(bit_depth), int(compression),
                            int(layout)):
        raise RuntimeError(lib().ref_last_error().decode())


def bgen_dosages(path, N, M):
    """the minor-allele dosages FileBgen::read_all sees, (M, N) float32, NaN = missing."""
    out = np.zeros((M, N), dtype=np.float32)
    if lib().ref_bgen_dosages(path.encode(), _p(out), C.c_longlong(N), C.c_longlong(M)):
        raise RuntimeError(lib().ref_last_error().decode())
    return out


def init_omega(rows, cols, seed=112, gaussian=True):
    out = _f((rows, cols))
    lib().ref_init_omega(C.c_longlong(rows), C.c_longlong(cols), int(seed), int(gaussian), _p(out))
    return out


def permute_indices(n):
    out = np.zeros(n, dtype=np.int32)
    lib().ref_permute_indices(C.c_longlong(n), _p(out))
    return out


def mev(X, Y):
    X = np.asfortranarray(X, dtype=np.float64)
    Y = np.asfortranarray(Y, dtype=np.float64)
    return float(lib().ref_mev(_p(X), _p(Y), C.c_longlong(X.shape[0]), C.c_longlong(X.shape[1])))


def flip_uv(U, V):
    U = np.asfortranarray(U, dtype=np.float64).copy(order="F")
    V = np.asfortranarray(V, dtype=np.float64).copy(order="F")
    lib().ref_flip_uv(_p(U), C.c_longlong(U.shape[0]), _p(V), C.c_longlong(V.shape[0]), C.c_longlong(U.shape[1]))
    return U, V


def flip_omg(Omg2, Omg):
    A = np.asfortranarray(Omg2, dtype=np.float64).copy(order="F")
    B = np.asfortranarray(Omg, dtype=np.float64).copy(order="F")
    lib().ref_flip_omg(_p(A), _p(B), C.c_longlong(A.shape[0]), C.c_longlong(A.shape[1]))
    return A, B


def householder_q(A):
    A = np.asfortranarray(A, dtype=np.float64)
    Q = _f(A.shape)
    lib().ref_householder_q(_p(A), C.c_longlong(A.shape[0]), C.c_longlong(A.shape[1]), _p(Q))
    return Q


def rsvd_one(A, k, os_=10, rand=1, p=3, windows=0, finder=1):
    """PCAone::RsvdOne<MatrixXd>(A, k, os, rand); setRangeFinder(finder); compute(p, windows)
    (RSVD.hpp:327-362) -> U (rows x k), S (k), V (cols x k)."""
    A = np.asfortranarray(A, dtype=np.float64)
    r, c = A.shape
    U, S, V = _f((r, k)), np.zeros(k), _f((c, k))
    rc = lib().ref_rsvd_one(_p(A), C.c_longlong(r), C.c_longlong(c), int(k), int(os_), int(rand), int(p), int(windows),
                            int(finder), _p(U), _p(S), _p(V))
    if rc:
        raise RuntimeError("reference failed: " + lib().ref_last_error().decode())
    return U, S, V


# ---- the compiled drop-in: reference host code + integration/HalkoGpu.hpp + libpcaone_b200.so ----------
GPU_LIB_PATH = os.path.join(_HERE, "_ref", "libpcaone_ref_gpu.so")
_gpu_lib = None


def gpu_available() -> bool:
    return os.path.exists(GPU_LIB_PATH)


def gpu_run(cmdline: str, precision: int, on_device: bool = False):
    """Run a PCAone command line through the UNMODIFIED reference host code with the GPU ops of
    integration/HalkoGpu.hpp (oracle/ref_gpu_shim.cpp). Returns U, S, V (V in run order) and perm."""
    global _gpu_lib
    if _gpu_lib is None:
        _gpu_lib = C.CDLL(GPU_LIB_PATH)
        _gpu_lib.refgpu_last_error.restype = C.c_char_p
    dims = (C.c_longlong * 3)()
    # sizes first (cheap: the shim only needs the .fam / .bim line counts, so ask the host side)
    toks = cmdline.split()
    prefix = toks[toks.index("-b") + 1]
    n = sum(1 for _ in open(prefix + ".fam"))
    m = sum(1 for _ in open(prefix + ".bim"))
    k = int(toks[toks.index("-k") + 1])
    U, S, V = _f((n, k)), np.zeros(k), _f((m, k))
    perm = np.full(m, -1, dtype=np.int32)
    rc = _gpu_lib.refgpu_run(cmdline.encode(), int(precision), int(on_device), _p(U), _p(S), _p(V), dims, _p(perm))
    if rc:
        raise RuntimeError("reference-on-GPU run failed: " + _gpu_lib.refgpu_last_error().decode())
    assert (dims[0], dims[1], dims[2]) == (n, m, k)
    return U, S, V, (perm if perm[0] >= 0 or (perm >= 0).all() else None)
