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
    The reference walks i upward with a chromosome counter c that advances whenever
    pos[i] == pos[end_pos[c]] (such a SNP leads no window). Vectorised: one searchsorted per
    chromosome instead of a Python loop over the SNPs (0.57 s of the 0.71 s config-5 run)."""
    chrom = np.asarray(chrom)
    pos = np.asarray(pos, dtype=np.int64)
    n = len(pos)
    end_pos = np.flatnonzero(chrom[1:] != chrom[:-1]).tolist() + [n - 1]
    ws_all, we_all = [], []
    i, c = 0, 0
    while i < n and c < len(end_pos):
        e = end_pos[c]
        # SNPs i..e are candidates under counter c; the first one whose position equals pos[e] ends the run
        seg = pos[i:e + 1]
        hit = np.flatnonzero(seg == pos[e])
        stop = i + int(hit[0])                     # this SNP bumps c and leads no window
        lead = np.arange(i, stop)
        if len(lead):
            # first j in (lead, e] with pos[j] - pos[lead] > bp (positions ascend inside a chromosome)
            j = i + np.searchsorted(seg, pos[lead] + ld_window_bp, side="right")
            # searchsorted on the whole segment counts SNPs before `lead` too: they all satisfy pos <= pos[lead] + bp
            ws_all.append(lead)
            we_all.append(j - lead)
        i = stop + 1
        c += 1
    if not ws_all:
        return np.zeros(0, np.int32), np.zeros(0, np.int32)
    return np.concatenate(ws_all).astype(np.int32), np.concatenate(we_all).astype(np.int32)


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


def _ld_ex(op, src, nsnps, ws, we, r2_tol=None, af=None):
    ws = np.ascontiguousarray(ws, dtype=np.int32)
    we = np.ascontiguousarray(we, dtype=np.int32)
    if r2_tol is None:
        out = np.zeros(int((we.astype(np.int64) - 1).sum()), dtype=np.float64)
        op._chk(op.L.pcaone_ld_r2_ex(op.h, C.byref(src), C.c_uint64(nsnps), _vp(ws), _vp(we), C.c_uint64(len(ws)),
                                     _vp(out), None, C.c_double(0.0), None))
        return out
    keep = np.zeros(nsnps, dtype=np.uint8)
    if af is not None:
        af = np.ascontiguousarray(af, dtype=np.float64)
    op._chk(op.L.pcaone_ld_r2_ex(op.h, C.byref(src), C.c_uint64(nsnps), _vp(ws), _vp(we), C.c_uint64(len(ws)), None,
                                 _vp(af), C.c_double(r2_tol), _vp(keep)))
    return keep.astype(bool)


def ld_from_residuals_file(op, resid_f32, ws, we, r2_tol=None, af=None):
    """`-B file.residuals`: resid_f32 = the float32 rows of the file ([M][N], after its 8-byte header).
    FileBin::read_all (src/FileBinary.cpp:21-30) + ld_r2_big / ld_prune_big, streamed through the device."""
    R = np.ascontiguousarray(resid_f32, dtype=np.float32)
    if R.ndim != 2 or R.shape[1] != op.cols():
        raise RuntimeError("residuals must be [nsnps][nsamples] float32")
    src = _lib.LdSource(kind=_lib.LD_RESID_F32, ld_stats=0, data=R.ctypes.data, ncols=0)
    return _ld_ex(op, src, R.shape[0], ws, we, r2_tol, af)


def ld_adjusted_from_bed(op, ws, we, ld_stats=0, r2_tol=None, af=None):
    """`--ld` followed by `-B`: in ONE step on the device: the residuals Data::write_residuals would write
    (src/Data.cpp:242-291; ld_stats 0 subtracts U S V^T of the PCA the context just ran) feed the r2 tiles
    directly, rounded through float32 exactly like the file round trip."""
    src = _lib.LdSource(kind=_lib.LD_PACKED_RESID, ld_stats=int(ld_stats), data=None, ncols=0)
    return _ld_ex(op, src, op.rows(), ws, we, r2_tol, af)


def ld_projected_from_bed(op, U, ws, we, r2_tol=None, af=None):
    """bed + `--USV`: data->G = (I - U U^T) G (src/LD.cpp:491-496), then ld_r2_big / ld_prune_big."""
    U = np.asfortranarray(U, dtype=np.float64)
    if U.shape[0] != op.cols():
        raise RuntimeError("U must have one row per sample")
    src = _lib.LdSource(kind=_lib.LD_PACKED_PROJECT, ld_stats=0, data=U.ctypes.data, ncols=U.shape[1])
    return _ld_ex(op, src, op.rows(), ws, we, r2_tol, af)


def write_residuals(op, path, ld_stats=0, perm=None, chunk=8192):
    """Data::write_residuals (src/Data.cpp:242-291): `<out>.residuals` = [uint32 M][uint32 N][M x N float32],
    rows un-permuted by seeking (Data.cpp:264-267). The float rows come from the device."""
    M, N = op.rows(), op.cols()
    with open(path, "wb") as f:
        f.write(np.array([M, N], dtype=np.uint32).tobytes())
        if perm is not None:
            f.truncate(8 + 4 * M * N)
        for s in range(0, M, chunk):
            e = min(M, s + chunk) - 1
            blk = np.zeros((e - s + 1, N), dtype=np.float32)
            op._chk(op.L.pcaone_residuals_block(op.h, C.c_uint64(s), C.c_uint64(e), int(ld_stats), _vp(blk)))
            if perm is None:
                f.write(blk.tobytes())
            else:
                for i in range(e - s + 1):
                    f.seek(8 + int(perm[s + i]) * 4 * N)
                    f.write(blk[i].tobytes())
