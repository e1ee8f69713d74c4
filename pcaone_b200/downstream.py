"""Host-side mirror of the two downstream consumers of U, S, V that reuse the decode + GEMM
products. Only the products run on the device; file matching, statistics and
writers stay host code in PCAone.

    run_projection (option 1)     src/Projection.cpp:188-241   U = G (V S^-1)
    run_selection  (products)     src/Selection.cpp:5-38       V = U^T G per SNP, ||g_j||^2
"""
from __future__ import annotations

import numpy as np


def run_projection(op, V, S, ref_F=None, project=1):
    """`op`: an RsvdOpData context on the target genotypes; V (M x K), S (K): the reference panel's
    loadings and singular values; ref_F: the panel's allele frequencies (used for centring and
    scaling, as Data::prepare does for --project). Returns U (N x K)."""
    if project != 1:
        raise RuntimeError("--project 2/3 (per-sample least squares with missing calls, GL-aware EM) are not on the GPU path")
    if ref_F is not None:
        op.setF(ref_F)
    op.setFlags(False, True)                       # data->standardize_E(), Projection.cpp:216
    V = np.asarray(V, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    return op.xTimes(V / S[None, :])               # V * diag(1 / S), then U = G V (:239-241)


def run_selection_products(op, U, E):
    """V.row(j) = U^T G.col(j), y_norm2(j) = ||G.col(j)||^2 on standardized genotypes, then the
    downscaling of Selection.cpp:37-38 (`E = E.head(K) * M; V /= sqrt(E)`). Returns (V, y_norm2)."""
    U = np.asarray(U, dtype=np.float64)
    E = np.asarray(E, dtype=np.float64)[: U.shape[1]]
    op.setFlags(False, True)                       # data->standardize_E(), Selection.cpp:17
    V, nrm = op.xtTimes(U, want_sqnorm=True)
    return V / np.sqrt(E * V.shape[0])[None, :], nrm
