"""Multi-GPU parity script, launched by torchrun (see tests/test_gpu_multi.py):
SNP-sharded sSVD / winSVD over WORLD_SIZE GPUs must reproduce the single-GPU result on the
same bed and the same Omega (H all-reduced at every Omega update, Gram of G per epoch)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from conftest import assert_usv_close  # noqa: E402
from pcaone_b200 import dist as pdist  # noqa: E402
from pcaone_b200 import halko, synth  # noqa: E402


def run(svd, packed, N, k, bands, maxp, rank, world, local, hook, emu=False):
    M = packed.shape[0]
    p = halko.Param(k=k, svd=svd, bands=bands, maxp=maxp, tol=0.0, no_shuffle=True, device=local, emu=emu, maxiter=3)
    if svd == 2:
        idx, start, stop = pdist.shard_windows(M, bands, rank, world)
    else:
        s, e = pdist.shard_range(M, rank, world)
        idx, start, stop = np.arange(s, e), None, None
    d = halko.FileBed(p, packed=np.ascontiguousarray(packed[idx]), nsamples=N)
    d.start, d.stop = start, stop
    cls = halko.FancyRsvdOpData if svd == 2 else halko.NormalRsvdOpData
    op = cls(d, p.k, p.oversamples, rank=rank, world=world, nsnps_total=M, allreduce=hook)
    if emu:
        op.runEM()   # Halko.cpp:290-319: includes flip_UV across the SNP shards after every computeUSV
    else:
        op.setFlags(False, True)
        op.computeUSV(maxp, 0.0)
    return op, idx


def main():
    rank, world, local = pdist.init_process_group_from_env("nccl")
    torch.cuda.set_device(local)
    N, M, k = 900, 12800, 6
    packed = np.concatenate([synth.pack_codes(c) for _, c in synth.balding_nichols_codes(N, M, k_pop=8, seed=9)])
    hook = pdist.make_allreduce_hook()
    packed_m = np.concatenate([synth.pack_codes(c)
                               for _, c in synth.balding_nichols_codes(N, M, k_pop=8, seed=10, miss=0.05)])
    for svd, bands, maxp, emu in ((1, 64, 4, False), (2, 16, 6, False), (1, 64, 3, True)):
        if emu:
            packed = packed_m
        op, idx = run(svd, packed, N, k, bands, maxp, rank, world, local, hook, emu)
        Vfull = torch.zeros((M, k), dtype=torch.float64, device=f"cuda:{local}")
        Vfull[torch.from_numpy(idx).to(Vfull.device)] = torch.from_numpy(np.ascontiguousarray(op.V)).to(Vfull.device)
        dist.all_reduce(Vfull)
        if rank == 0:
            ref_op, _ = run(svd, packed, N, k, bands, maxp, 0, 1, local, None, emu)
            assert_usv_close(op.U, op.S, Vfull.cpu().numpy(), ref_op.U, ref_op.S, ref_op.V, eig_rtol=1e-9,
                             min_corr=1 - 1e-9)
            if emu:  # flip_UV fixed the signs: compare without sign alignment
                assert np.abs(op.U - ref_op.U).max() < 1e-8
            print(f"svd={svd} emu={emu} world={world}: sharded == single GPU; S rel err",
                  float(np.max(np.abs(op.S - ref_op.S) / ref_op.S)), flush=True)
            ref_op.close()
        op.close()
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK", flush=True)


if __name__ == "__main__":
    main()
