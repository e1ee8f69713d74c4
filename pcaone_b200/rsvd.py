"""Host-side mirror of the generic dense-matrix front-end of the reference (src/RSVD.hpp:92-362,
what PCAoneR binds): `RsvdOne(mat, k, os, rand)` with `setRangeFinder`, `compute(p, windows)`,
`matrixU()`, `matrixV()`, `singularValues()`. The matrix is uploaded once as doubles and the two
power-iteration products run on the FP64 tensor cores (csrc/dense_gemm.cuh); Omega updates, QR(G) x2
and the SVD of the l x l core are the same device code as for genotypes. No CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .halko import _f, _vp


class RsvdOne:
    """RSVD.hpp:327-362. `rand` 1 = standard normal Omega, otherwise uniform(-1, 1); the engine is the
    default-seeded `std::default_random_engine{}` of RSVD.hpp:123 (pcaone_init_omega reproduces
    the libstdc++ stream bit for bit), or pass `omega` (ncol x (k + os)) to inject one."""

    def __init__(self, mat, k, os_=10, rand=1, *, omega=None, device=0):
        L = _lib.load()
        self.L = L
        mat = np.asarray(mat, dtype=np.float64)
        if mat.ndim != 2:
            raise ValueError("mat must be a matrix")
        self.mat = np.asfortranarray(mat)
        self.k, self.os, self.rand = int(k), int(os_), int(rand)
        self.size = self.k + self.os
        r, c = self.mat.shape
        self.trans = r < c                      # RSVD.hpp:337-340
        self.nrow, self.ncol = (c, r) if self.trans else (r, c)
        self.finder = 1
        self.device = device
        self.h = None
        if omega is None:
            omega = _f((self.ncol, self.size))
            # default-constructed std::default_random_engine == minstd_rand0 seeded with 1
            if L.pcaone_init_omega(self.ncol, self.size, 1, int(self.rand == 1), _vp(omega)):
                raise RuntimeError("pcaone_init_omega failed")
        self.Omg = np.asfortranarray(omega, dtype=np.float64)
        if self.Omg.shape != (self.ncol, self.size):
            raise ValueError("omega must be ncol x (k + os)")
        self._U = self._S = self._V = None

    def setRangeFinder(self, flag):
        self.finder = int(flag)

    def _chk(self, rc):
        if rc:
            raise RuntimeError(self.L.pcaone_last_error(self.h).decode())

    def compute(self, p, windows=0):
        """RsvdOnePass::computeUSV(p, windows), RSVD.hpp:281-313."""
        L = self.L
        cfg = _lib.Config(nsamples=self.ncol, nsnps=self.nrow, nsnps_total=self.nrow, k=self.k, oversamples=self.os,
                          svd=2 if windows > 0 else 1, bands=max(int(windows), 2), maxp=int(p), tol=0.0, ploidy=2,
                          scale=0, emu=0, out_of_core=1, precision=_lib.PREC_FP64, device=self.device, rank=0, world=1,
                          maxiter=0, tolem=0.0)
        h = C.c_void_p()
        if L.pcaone_create(C.byref(cfg), C.byref(h)):
            raise RuntimeError(L.pcaone_last_error(None).decode())
        self.h = h
        try:
            self._chk(L.pcaone_upload_dense(h, _vp(self.mat), self.mat.shape[0], self.mat.shape[1]))
            self._chk(L.pcaone_set_omega(h, _vp(self.Omg)))
            self._chk(L.pcaone_dense_rsvd(h, int(p), int(windows), self.finder))
            U, S, V = _f((self.ncol, self.k)), np.zeros(self.k), _f((self.nrow, self.k))
            self._chk(L.pcaone_get_usv(h, _vp(U), _vp(S), _vp(V)))
        finally:
            L.pcaone_destroy(h)
            self.h = None
        # b_leftSingularVectors = G * svd.matrixU() (nrow x k), b_rightSingularVectors = svd.matrixV()
        self._left, self._right, self._S = V, U, S

    def matrixU(self):
        return self._right if self.trans else self._left

    def matrixV(self):
        return self._left if self.trans else self._right

    def singularValues(self):
        return self._S
