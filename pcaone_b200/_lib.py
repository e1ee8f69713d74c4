"""ctypes binding of the C-ABI in include/pcaone_b200.h (libpcaone_b200.so, built in-tree by
__graft_entry__.build()). There is no CPU fallback: if the library is missing or no CUDA device
is usable the product path raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpcaone_b200.so")

SVD_SSVD, SVD_WINSVD = 1, 2
PREC_FP64, PREC_INT8X2, PREC_INT8X3, PREC_INT8X4 = 0, 2, 3, 4


class Config(C.Structure):
    _fields_ = [
        ("nsamples", C.c_uint64), ("nsnps", C.c_uint64), ("nsnps_total", C.c_uint64),
        ("k", C.c_uint32), ("oversamples", C.c_uint32), ("svd", C.c_uint32), ("bands", C.c_uint32),
        ("maxp", C.c_uint32), ("tol", C.c_double), ("ploidy", C.c_int32), ("scale", C.c_int32),
        ("emu", C.c_int32), ("out_of_core", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32),
        ("rank", C.c_int32), ("world", C.c_int32), ("maxiter", C.c_uint32), ("tolem", C.c_double),
        ("shard_samples", C.c_int32), ("nsamples_total", C.c_uint64), ("sample_offset", C.c_uint64),
    ]


class LdSource(C.Structure):
    """pcaone_ld_source: which operand pcaone_ld_r2_ex builds its tiles from."""
    _fields_ = [("kind", C.c_int32), ("ld_stats", C.c_int32), ("data", C.c_void_p), ("ncols", C.c_uint32)]


LD_DENSE_F64, LD_PACKED, LD_RESID_F32, LD_PACKED_RESID, LD_PACKED_PROJECT = 0, 1, 2, 3, 4


class Timers(C.Structure):
    _fields_ = [
        ("gemm_g_ms", C.c_double), ("gemm_h_ms", C.c_double), ("orth_ms", C.c_double), ("small_ms", C.c_double),
        ("h2d_ms", C.c_double), ("allreduce_ms", C.c_double), ("decode_ms", C.c_double),
        ("gemm_g_launches", C.c_uint64), ("gemm_h_launches", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("omega_updates", C.c_uint64),
        ("tc_ranges", C.c_uint64), ("fp64_ranges", C.c_uint64), ("tc_g_ms", C.c_double), ("tc_h_ms", C.c_double),
        ("ld_ms", C.c_double), ("ld_tiles", C.c_uint64), ("ld_pairs", C.c_uint64),
        ("tc_miss_ranges", C.c_uint64), ("cache_hits", C.c_uint64), ("tc_emu_ranges", C.c_uint64), ("emu_fix_ms", C.c_double),
    ]


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p)
ALLREDUCE2_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p)
READ_BLOCK_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p)

# every symbol include/pcaone_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "pcaone_create", "pcaone_destroy", "pcaone_last_error", "pcaone_abi_version", "pcaone_stream", "pcaone_sync",
    "pcaone_set_allreduce", "pcaone_upload_bed", "pcaone_set_host_source", "pcaone_set_reader_source",
    "pcaone_open_bed", "pcaone_set_blocks", "pcaone_permute_resident", "pcaone_allele_freq", "pcaone_get_F",
    "pcaone_set_F", "pcaone_get_lookup", "pcaone_get_scale", "pcaone_missing_count", "pcaone_decode_block",
    "pcaone_set_flags", "pcaone_set_omega", "pcaone_get_omega", "pcaone_set_usv", "pcaone_get_usv", "pcaone_get_GH",
    "pcaone_set_H", "pcaone_compute_gandh", "pcaone_small_stage", "pcaone_compute_usv", "pcaone_run_em",
    "pcaone_orth_omega", "pcaone_mev", "pcaone_init_omega", "pcaone_shuffle_indices", "pcaone_ld_r2",
    "pcaone_get_timers", "pcaone_enable_timing", "pcaone_alloc_pinned", "pcaone_free_pinned", "pcaone_device_count",
    "pcaone_upload_dense", "pcaone_dense_rsvd", "pcaone_upload_dosage", "pcaone_perform_op", "pcaone_ld_prune", "pcaone_xt_times", "pcaone_x_times",
    "pcaone_upload_gl", "pcaone_gl_em_maf",
    "pcaone_comm_unique_id", "pcaone_comm_init", "pcaone_comm_attach", "pcaone_set_host_source2", "pcaone_set_allreduce2",
    "pcaone_ld_r2_ex", "pcaone_residuals_block", "pcaone_precision", "pcaone_comm_peer_export", "pcaone_comm_peer_import",
    "pcaone_sample_covariance", "pcaone_sym_svd", "pcaone_gl_grm", "pcaone_upload_dense_data", "pcaone_mask_times",
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "pcaone_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.pcaone_last_error.restype = C.c_char_p
    L.pcaone_last_error.argtypes = [C.c_void_p]
    L.pcaone_stream.restype = C.c_void_p
    L.pcaone_stream.argtypes = [C.c_void_p]
    L.pcaone_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.pcaone_destroy.argtypes = [C.c_void_p]
    L.pcaone_destroy.restype = None
    vp, u64, i32, u32, dbl = C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, C.c_double
    sig = {
        "pcaone_sync": [vp], "pcaone_set_allreduce": [vp, ALLREDUCE_FN, vp],
        "pcaone_upload_bed": [vp, vp, u64, i32], "pcaone_set_host_source": [vp, vp, u64],
        "pcaone_set_reader_source": [vp, READ_BLOCK_FN, vp], "pcaone_open_bed": [vp, C.c_char_p, u64],
        "pcaone_set_blocks": [vp, vp, vp, u32, u32], "pcaone_permute_resident": [vp, vp],
        "pcaone_allele_freq": [vp], "pcaone_get_F": [vp, vp], "pcaone_set_F": [vp, vp],
        "pcaone_get_lookup": [vp, vp], "pcaone_get_scale": [vp, vp], "pcaone_missing_count": [vp, vp],
        "pcaone_decode_block": [vp, u64, u64, i32, i32, vp], "pcaone_set_flags": [vp, i32, i32],
        "pcaone_set_omega": [vp, vp], "pcaone_get_omega": [vp, vp], "pcaone_set_usv": [vp, vp, vp, vp],
        "pcaone_get_usv": [vp, vp, vp, vp], "pcaone_get_GH": [vp, vp, vp], "pcaone_set_H": [vp, vp],
        "pcaone_compute_gandh": [vp, i32], "pcaone_small_stage": [vp],
        "pcaone_compute_usv": [vp, i32, dbl, vp, vp], "pcaone_run_em": [vp, vp], "pcaone_orth_omega": [vp, i32],
        "pcaone_mev": [vp, vp, vp, u64, u32, vp], "pcaone_init_omega": [u64, u32, i32, i32, vp],
        "pcaone_shuffle_indices": [u64, vp], "pcaone_ld_r2": [vp, vp, u64, vp, vp, u64, vp],
        "pcaone_get_timers": [vp, C.POINTER(Timers), i32], "pcaone_enable_timing": [vp, i32],
        "pcaone_upload_dense": [vp, vp, u64, u64], "pcaone_dense_rsvd": [vp, u32, u32, i32],
        "pcaone_upload_dosage": [vp, vp, u64, i32], "pcaone_perform_op": [vp, vp, vp], "pcaone_xt_times": [vp, vp, u32, vp, vp], "pcaone_x_times": [vp, vp, u32, vp], "pcaone_mask_times": [vp, vp, u32, vp],
        "pcaone_upload_gl": [vp, vp, u64, i32], "pcaone_gl_em_maf": [vp, u32, dbl, vp],
        "pcaone_ld_prune": [vp, vp, u64, vp, vp, u64, vp, dbl, vp],
        "pcaone_comm_unique_id": [vp], "pcaone_comm_init": [vp, vp, i32, i32], "pcaone_comm_attach": [vp, vp],
        "pcaone_ld_r2_ex": [vp, C.POINTER(LdSource), u64, vp, vp, u64, vp, vp, dbl, vp],
        "pcaone_residuals_block": [vp, u64, u64, i32, vp], "pcaone_precision": [vp], "pcaone_comm_peer_export": [vp, vp], "pcaone_comm_peer_import": [vp, vp, i32],
        "pcaone_set_host_source2": [vp, vp, u64, u64], "pcaone_set_allreduce2": [vp, ALLREDUCE2_FN, vp],
        "pcaone_sample_covariance": [vp, vp], "pcaone_upload_dense_data": [vp, vp], "pcaone_gl_grm": [vp, vp, vp], "pcaone_sym_svd": [vp, vp, u64, vp, vp, vp],
    }
    L.pcaone_alloc_pinned.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    L.pcaone_alloc_pinned.restype = C.c_int
    L.pcaone_free_pinned.argtypes = [C.c_void_p]
    L.pcaone_free_pinned.restype = None
    L.pcaone_device_count.argtypes = []
    L.pcaone_device_count.restype = C.c_int
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = L
    return L
