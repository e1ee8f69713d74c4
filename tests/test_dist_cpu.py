"""CPU, world_size 2, gloo: the host-side multi-GPU logic — shard plans and the allreduce hook
that the C-ABI calls — checked against the numpy oracle (sum of per-shard partial H == H)."""
import ctypes as C
import os

import numpy as np
import torch.multiprocessing as mp

from pcaone_b200 import dist as pdist


def test_shard_plans():
    for n, w in ((10, 3), (7, 8), (1000, 4)):
        got = [pdist.shard_range(n, r, w) for r in range(w)]
        assert got[0][0] == 0 and got[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
    M, bands, world = 1003, 8, 3
    seen = []
    for r in range(world):
        idx, start, stop = pdist.shard_windows(M, bands, r, world)
        assert len(start) == bands and np.all(stop >= start)
        assert int(stop[-1]) + 1 == len(idx)
        seen.append(idx)
    allidx = np.sort(np.concatenate(seen))
    assert np.array_equal(allidx, np.arange(M))
    # every rank holds a slice of every window
    bs = -(-M // bands)
    for r in range(world):
        idx, start, stop = pdist.shard_windows(M, bands, r, world)
        for b in range(bands):
            w = idx[int(start[b]):int(stop[b]) + 1]
            assert w.min() >= b * bs and w.max() < min((b + 1) * bs, M)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import pcaone_oracle as orc
    from pcaone_b200 import synth

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = pdist.init_process_group_from_env("gloo")
    N, M, l = 40, 300, 6
    packed = np.concatenate([synth.pack_codes(c) for _, c in synth.balding_nichols_codes(N, M, k_pop=3, seed=2)])
    d = orc.OracleData(packed, N)
    X = d.block(0, M - 1, True)
    rng = np.random.default_rng(0)
    Om = rng.standard_normal((N, l))
    idx, start, stop = pdist.shard_windows(M, 4, r, w)
    Xs = X[:, idx]
    H = np.ascontiguousarray(Xs @ (Xs.T @ Om))  # this rank's partial H
    hook = pdist.make_allreduce_hook(device_buffers=False)
    rc = hook(None, H.ctypes.data_as(C.c_void_p).value, H.size, None)
    ok = rc == 0 and np.allclose(H, X @ (X.T @ Om), rtol=1e-12, atol=1e-9)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, bool(ok)))


def test_allreduce_hook_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res


def test_sample_shard_ranges_and_placeholder_windows():
    """sample-sharded jobs cut the sample axis on whole bed bytes; SNP-sharded winSVD keeps every rank on the same
    `bands`-step schedule (empty trailing windows stay as start = stop + 1 placeholders) and refuses a
    window with fewer SNPs than ranks instead of letting the ranks issue different numbers of collectives."""
    import pytest
    for n, w in ((900, 2), (900, 8), (500_000, 8), (1003, 4)):
        got = [pdist.shard_samples_range(n, r, w) for r in range(w)]
        assert got[0][0] == 0 and got[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
        assert all(s % 4 == 0 for s, _ in got)
    # M < bands * blocksize: the trailing window is empty on every rank, the schedule keeps `bands` entries
    M, bands, world = 130, 16, 2          # blocksize 9 -> 15 windows of SNPs, window 15 is empty
    lens = []
    for r in range(world):
        idx, start, stop = pdist.shard_windows(M, bands, r, world)
        assert len(start) == len(stop) == bands
        assert int(start[-1]) == int(stop[-1]) + 1      # placeholder
        lens.append(len(idx))
    assert sum(lens) == M
    with pytest.raises(RuntimeError):
        pdist.shard_windows(40, 16, 0, 8)               # 3-SNP windows cannot be cut 8 ways
