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
            # empty trailing window (M < bands * blocksize): kept as a placeholder so that every rank
            # walks the same `bands`-step schedule as one GPU does (start > stop = empty)
            start.append(pos + 1 if pos else 1)
            stop.append(pos if pos else 0)
            continue
        if e - s < world:
            raise RuntimeError(f"window {b} has {e - s} SNPs, fewer than the {world} ranks")
        ls, le = shard_range(e - s, rank, world)
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


def make_allreduce2_hook(group=None):
    """Typed collective hook (pcaone_allreduce2_fn) over ANY torch.distributed backend: the buffer is
    device memory; with a CPU backend (gloo) it is staged through the host. Slow, but it lets two
    ranks that time-share ONE GPU run the sharded schedules (NCCL refuses two ranks per device), which
    is how the sharded paths are tested on a one-GPU box."""
    import torch
    import torch.distributed as dist

    kinds = {0: ("<f8", torch.float64, dist.ReduceOp.SUM), 1: ("<i8", torch.int64, dist.ReduceOp.SUM),
             2: ("<i8", torch.int64, dist.ReduceOp.MAX),   # bit patterns of non-negative doubles: order-preserving as int64
             3: ("<i4", torch.int32, dist.ReduceOp.SUM)}   # genotype counts < 2^31

    class Buf:
        def __init__(self, ptr, count, typestr):
            self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                             "version": 3, "strides": None}

    def hook(user, buf, count, kind, stream):
        try:
            typestr, _, op = kinds[int(kind)]
            t = torch.as_tensor(Buf(buf, count, typestr), device="cuda")
            ext = torch.cuda.ExternalStream(int(stream))
            if dist.get_backend(group) == "nccl":
                with torch.cuda.stream(ext):
                    dist.all_reduce(t, op=op, group=group)
            else:
                ext.synchronize()
                h = t.cpu()
                dist.all_reduce(h, op=op, group=group)
                with torch.cuda.stream(ext):
                    t.copy_(h)
                ext.synchronize()
            return 0
        except Exception as e:  # never let an exception cross the C boundary
            print("allreduce2 hook failed:", repr(e), flush=True)
            return 1

    return hook


def init_library_comm(L, handle, rank, world, peer_mailboxes=False):
    """Give the context its own NCCL communicator (include/pcaone_b200.h: pcaone_comm_unique_id /
    pcaone_comm_init): rank 0 makes the id, torch.distributed carries the 128 bytes to the others.
    From then on every exchange step runs inside the library on its stream."""
    import torch
    import torch.distributed as dist

    ident = (C.c_uint8 * 128)()
    if rank == 0 and L.pcaone_comm_unique_id(ident):
        raise RuntimeError("pcaone_comm_unique_id failed (libnccl.so.2 not loadable?)")
    t = torch.tensor(list(ident), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
        dist.broadcast(t, 0)
        t = t.cpu()
    else:
        dist.broadcast(t, 0)
    ident = (C.c_uint8 * 128)(*t.tolist())
    if L.pcaone_comm_init(handle, ident, rank, world):
        raise RuntimeError(L.pcaone_last_error(handle).decode())
    if peer_mailboxes and dist.get_backend() == "nccl":
        # one GPU per rank: map every rank's mailbox into every other rank (CUDA IPC) so that the small
        # exchanges of the row-sharded Omega update run inside its kernel over NVLink
        mine = (C.c_uint8 * 64)()
        if L.pcaone_comm_peer_export(handle, mine):
            raise RuntimeError(L.pcaone_last_error(handle).decode())
        allh = torch.zeros(world * 64, dtype=torch.uint8, device="cuda")
        allh[rank * 64:(rank + 1) * 64] = torch.tensor(list(mine), dtype=torch.uint8, device="cuda")
        dist.all_reduce(allh)   # disjoint slots: the sum is an all-gather
        blob = (C.c_uint8 * (world * 64))(*allh.cpu().tolist())
        if L.pcaone_comm_peer_import(handle, blob, world):
            raise RuntimeError(L.pcaone_last_error(handle).decode())


def shard_samples_range(n_samples: int, rank: int, world: int):
    """Sample shard of a sample-sharded job: contiguous, starting at a multiple of 4 samples (whole bed bytes)."""
    per = -(-n_samples // world)
    per = (per + 3) // 4 * 4
    s = min(rank * per, n_samples)
    return s, min(s + per, n_samples)


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
