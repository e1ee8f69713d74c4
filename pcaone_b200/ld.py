"""Host-side mirror of the reference's LD r2 path on top of the C-ABI (pcaone_ld_r2):

    get_snp_pos_bim / divide_pos_by_window    src/LD.cpp:79-103, 154-168  (host integer logic)
    ld_r2_big                                 src/LD.cpp:450-473          (device: ld.cuh)

The window table is planned on the host exactly as the reference does; the pairwise
correlations are one banded tile Gram on the FP64 tensor cores. No CPU arithmetic fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .halko import _f, _vp


def divide_pos_by_window(chrom, pos, ld_window_bp):
    """ws (lead SNP index) and we (#SNPs in the window incl. the lead) of src/LD.cpp:154-168;
    `end_pos` = last index of every chromosome as get_snp_pos_bim builds it (:79-103).
    The last SNP of each chromosome leads no window."""
    chrom = np.asarray(chrom)
    pos = np.asarray(pos, dtype=np.int64)
    n = len(pos)
    end_pos = np.flatnonzero(chrom[1:] != chrom[:-1]).tolist() + [n - 1]
    ws, we = [], []
    c = 0
    for i in range(n):
        if pos[i] == pos[end_pos[c]]:
            c += 1
            continue
        e = end_pos[c]
        # first j in (i, e] with pos[j] - pos[i] > bp (positions ascend inside a chromosome)
        j = i + int(np.searchsorted(pos[i:e + 1] - pos[i], ld_window_bp, side="right"))
        ws.append(i)
        we.append(j - i)
    return np.asarray(ws, dtype=np.int32), np.asarray(we, dtype=np.int32)


def ld_r2_big(op, G, ws, we):
    """r2 of every (lead, partner) pair in the reference's output order (src/LD.cpp:450-473).
    `op` is any RsvdOpData context (it owns the GPU and, for G is None, the resident packed
    shard whose centred genotypes are used); G is N x M column-centred doubles or None."""
    ws = np.ascontiguousarray(ws, dtype=np.int32)
    we = np.ascontiguousarray(we, dtype=np.int32)
    n = int((we.astype(np.int64) - 1).sum())
    out = np.zeros(n, dtype=np.float64)
    if G is not None:
        G = np.asfortranarray(G, dtype=np.float64)
        if G.shape[0] != op.cols():
            raise RuntimeError("ld_r2_big: G must have one row per sample")
        nsnps = G.shape[1]
    else:
        nsnps = op.rows()
    op._chk(op.L.pcaone_ld_r2(op.h, _vp(G), C.c_uint64(nsnps), _vp(ws), _vp(we), C.c_uint64(len(ws)), _vp(out)))
    return out


def ld_prune_big(op, G, ws, we, r2_tol, af=None):
    """Greedy LD pruning (src/LD.cpp:240-268): returns the boolean keep mask that
    write_pruned_snp_ids (:170-190) splits into .ld.prune.in / .ld.prune.out. `af` is the 7th
    column of the .mbim (then the lower-MAF SNP of a pair goes) or None (the partner goes)."""
    ws = np.ascontiguousarray(ws, dtype=np.int32)
    we = np.ascontiguousarray(we, dtype=np.int32)
    if G is not None:
        G = np.asfortranarray(G, dtype=np.float64)
        if G.shape[0] != op.cols():
            raise RuntimeError("ld_prune_big: G must have one row per sample")
        nsnps = G.shape[1]
    else:
        nsnps = op.rows()
    if af is not None:
        af = np.ascontiguousarray(af, dtype=np.float64)
        if af.shape != (nsnps,):
            raise RuntimeError("ld_prune_big: af must have one entry per SNP")
    keep = np.zeros(nsnps, dtype=np.uint8)
    op._chk(op.L.pcaone_ld_prune(op.h, _vp(G), C.c_uint64(nsnps), _vp(ws), _vp(we), C.c_uint64(len(ws)), _vp(af),
                                 C.c_double(r2_tol), _vp(keep)))
    return keep.astype(bool)
