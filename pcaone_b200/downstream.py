"""Host-side mirror of the two downstream consumers of U, S, V that reuse the decode + GEMM
products. Only the products run on the device; file matching, statistics and
writers stay host code in PCAone.

    run_projection (options 1, 2) src/Projection.cpp:159-246   U = G (V S^-1); per-sample least squares over the called SNPs
    run_selection  (products)     src/Selection.cpp:5-38       V = U^T G per SNP, ||g_j||^2
"""
from __future__ import annotations

import numpy as np


def run_projection(op, V, S, ref_F=None, project=1):
    """`op`: an RsvdOpData context on the target genotypes; V (M x K), S (K): the reference panel's
    loadings and singular values; ref_F: the panel's allele frequencies (used for centring and
    scaling, as Data::prepare does for --project). Returns U (N x K)."""
    if project not in (1, 2):
        raise RuntimeError("--project 3 (the GL-aware EM projection of BEAGLE input) is not on the GPU path")
    if ref_F is not None:
        op.setF(ref_F)
    op.setFlags(False, True)                       # data->standardize_E(), Projection.cpp:216
    V = np.asarray(V, dtype=np.float64)
    S = np.asarray(S, dtype=np.float64)
    if project == 1:
        return op.xTimes(V / S[None, :])           # V * diag(1 / S), then U = G V (:239-241)
    # --project 2 (:242-246, solve_projection_scores :159-179): per sample the least-squares solution of
    # g_i = (V S) x over the SNPs sample i was called at. Normal equations A_i x = b_i with
    #   b_i = (G W)_i        (missing calls are 0 in the standardised G: one product on the device)
    #   A_i = W^T W - sum_{j missing at i} w_j w_j^T   (the missing-call indicator times the K (K + 1) / 2 products
    #                                                   of the columns of W: pcaone_mask_times on the device)
    W = V * S[None, :]
    K = W.shape[1]
    b = op.xTimes(W)
    iu = np.triu_indices(K)
    Z = W[:, iu[0]] * W[:, iu[1]]                  # M x K (K + 1) / 2
    step = op.size()
    Mz = np.concatenate([op.maskTimes(Z[:, c:c + step]) for c in range(0, Z.shape[1], step)], axis=1)
    A = np.empty((b.shape[0], K, K))
    A0 = W.T @ W
    A[:, iu[0], iu[1]] = A0[iu][None, :] - Mz
    A[:, iu[1], iu[0]] = A[:, iu[0], iu[1]]
    return np.linalg.solve(A, b[:, :, None])[:, :, 0]


def run_selection_products(op, U, E):
    """V.row(j) = U^T G.col(j), y_norm2(j) = ||G.col(j)||^2 on standardized genotypes, then the
    downscaling of Selection.cpp:37-38 (`E = E.head(K) * M; V /= sqrt(E)`). Returns (V, y_norm2)."""
    U = np.asarray(U, dtype=np.float64)
    E = np.asarray(E, dtype=np.float64)[: U.shape[1]]
    op.setFlags(False, True)                       # data->standardize_E(), Selection.cpp:17
    V, nrm = op.xtTimes(U, want_sqnorm=True)
    return V / np.sqrt(E * V.shape[0])[None, :], nrm
