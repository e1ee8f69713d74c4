"""SNP-sharded multi-GPU plumbing: one process per GPU, torch.distributed for the
rendezvous and the collectives. The only exchange step on the path is the sum of the N x l
partial products H (and the l x l Gram of the SNP-sharded G); both go through the C-ABI's
allreduce hook, which this module implements with torch.distributed.all_reduce (NCCL on GPUs,
gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous split of 0..n_items-1: the first (n_items % world) ranks get one extra."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_windows(n_snps_total: int, bands: int, rank: int, world: int):
    """winSVD: every window is split across the ranks (each GPU walks the same schedule on 1/world
    of each window). Returns this rank's global SNP index array, window-major, and
    the per-window local (start, stop) inclusive ranges."""
    bs = -(-n_snps_total // bands)
    idx, start, stop = [], [], []
    pos = 0
    for b in range(bands):
        s = b * bs
        e = min((b + 1) * bs, n_snps_total)
        if s >= e:
            continue
        ls, le = shard_range(e - s, rank, world)
        if le > ls:
            idx.append(np.arange(s + ls, s + le, dtype=np.int64))
            start.append(pos)
            pos += le - ls
            stop.append(pos - 1)
    return (np.concatenate(idx) if idx else np.zeros(0, np.int64),
            np.array(start, dtype=np.uint64), np.array(stop, dtype=np.uint64))


class _DevBuf:
    """Minimal __cuda_array_interface__ carrier for a raw device pointer."""

    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def make_allreduce_hook(group=None, device_buffers=True):
    """Return a Python callable with the pcaone_allreduce_fn signature. With device_buffers the
    pointer is device memory and the all_reduce is enqueued on the library's stream; otherwise
    (CPU tests, gloo) the pointer is host memory."""
    import torch
    import torch.distributed as dist

    cache = {}

    def hook(user, buf, count, stream):
        try:
            key = (int(buf), int(count))
            t = cache.get(key)
            if t is None:
                if device_buffers:
                    t = torch.as_tensor(_DevBuf(buf, count), device="cuda")
                else:
                    arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_double)), shape=(int(count),))
                    t = torch.from_numpy(arr)
                cache[key] = t
            if device_buffers:
                ext = torch.cuda.ExternalStream(int(stream))
                with torch.cuda.stream(ext):
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception as e:  # never let an exception cross the C boundary
            print("allreduce hook failed:", repr(e), flush=True)
            return 1

    return hook


def init_process_group_from_env(backend=None):
    """torchrun-style rendezvous (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local
