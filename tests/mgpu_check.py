"""Multi-rank parity script, launched by torchrun (see tests/test_gpu_multi.py): a job sharded over
WORLD_SIZE ranks must reproduce the single-GPU result on the same bed and the same Omega.

  SNP-sharded      every rank owns 1/world of every window; H (N x l) summed at every Omega update,
                   the l x l Gram of G once per epoch                               
  sample-sharded   every rank owns 1/world of the samples of ALL SNPs; the exact int64 partial sums of
                   G_b = X_b^T Omega are summed per window, H / Omega stay row-sharded, the
                   orthonormalisation exchanges l x l Gram matrices

  --transport nccl   one GPU per rank, collectives inside the library (pcaone_comm_init, NCCL)
  --transport gloo   all ranks time-share GPU 0, collectives through the typed host hook over gloo:
                     the same schedules on a ONE-GPU box (NCCL refuses two ranks per device)
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from conftest import assert_usv_close  # noqa: E402
from pcaone_b200 import _lib  # noqa: E402
from pcaone_b200 import dist as pdist  # noqa: E402
from pcaone_b200 import halko, synth  # noqa: E402


def comm_kw(args):
    if args.transport == "nccl":
        return {"library_comm": True}
    return {"allreduce2": pdist.make_allreduce2_hook()}


def run_snp(svd, packed, N, k, bands, maxp, rank, world, local, kw, emu=False, prec=_lib.PREC_FP64):
    M = packed.shape[0]
    p = halko.Param(k=k, svd=svd, bands=bands, maxp=maxp, tol=0.0, no_shuffle=True, device=local, emu=emu, maxiter=3,
                    precision=prec)
    if svd == 2:
        idx, start, stop = pdist.shard_windows(M, bands, rank, world)
    else:
        s, e = pdist.shard_range(M, rank, world)
        idx, start, stop = np.arange(s, e), None, None
    d = halko.FileBed(p, packed=np.ascontiguousarray(packed[idx]), nsamples=N)
    d.start, d.stop = start, stop
    cls = halko.FancyRsvdOpData if svd == 2 else halko.NormalRsvdOpData
    op = cls(d, p.k, p.oversamples, rank=rank, world=world, nsnps_total=M, **(kw if world > 1 else {}))
    if emu:
        op.runEM()   # Halko.cpp:290-319: includes flip_UV across the SNP shards after every computeUSV
    else:
        op.setFlags(False, True)
        op.computeUSV(maxp, 0.0)
    return op, idx


def run_samples(svd, packed, N, k, bands, maxp, rank, world, local, kw, ooc):
    """Sample-sharded: this rank's byte columns of every SNP row. `ooc`: streamed from host memory
    with the reference's block plan (and the HBM tile cache), else resident."""
    M = packed.shape[0]
    s0, s1 = pdist.shard_samples_range(N, rank, world) if world > 1 else (0, N)
    local_packed = np.ascontiguousarray(packed[:, s0 // 4:s0 // 4 + (s1 - s0 + 3) // 4])
    p = halko.Param(k=k, svd=svd, bands=bands, maxp=maxp, tol=0.0, no_shuffle=True, device=local,
                    memory=0.001 if ooc else 0.0, precision=_lib.PREC_INT8X3)
    d = halko.FileBed(p, packed=local_packed, nsamples=s1 - s0)
    if ooc:  # the plan is the JOB's (it depends on N, M, l of the whole matrix)
        d.blocksize, d.nblocks, d.bandFactor, d.start, d.stop = halko.ooc_block_plan(N, M, p.l, 0.002, svd == 2, bands)
    cls = halko.FancyRsvdOpData if svd == 2 else halko.NormalRsvdOpData
    op = cls(d, p.k, p.oversamples, rank=rank, world=world, shard_samples=world > 1, nsamples_total=N,
             sample_offset=s0, **(kw if world > 1 else {}))
    op.setFlags(False, True)
    op.computeUSV(maxp, 0.0)
    return op, (s0, s1)


def gather_rows(x, world):
    if world == 1:
        return x
    out = [None] * world
    dist.all_gather_object(out, np.ascontiguousarray(x))
    return np.concatenate(out, axis=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--transport", default="nccl", choices=["nccl", "gloo"])
    args = ap.parse_args()
    rank, world, local = pdist.init_process_group_from_env(args.transport)
    if args.transport == "gloo":
        local = 0  # every rank on the one GPU
    torch.cuda.set_device(local)
    kw = comm_kw(args)
    N, M, k = 900, 12800, 6
    packed = np.concatenate([synth.pack_codes(c) for _, c in synth.balding_nichols_codes(N, M, k_pop=8, seed=9)])
    packed_m = np.concatenate([synth.pack_codes(c)
                               for _, c in synth.balding_nichols_codes(N, M, k_pop=8, seed=10, miss=0.05)])
    # ---- SNP-sharded: sSVD, winSVD (FP64 and int8 routes), EMU
    for svd, bands, maxp, emu, prec in ((1, 64, 4, False, _lib.PREC_FP64), (2, 16, 6, False, _lib.PREC_FP64),
                                        (2, 16, 6, False, _lib.PREC_INT8X3), (1, 64, 3, True, _lib.PREC_FP64),
                                        (1, 64, 3, True, _lib.PREC_INT8X3)):
        src = packed_m if emu else packed
        op, idx = run_snp(svd, src, N, k, bands, maxp, rank, world, local, kw, emu, prec)
        Vfull = np.zeros((M, k))
        Vfull[idx] = op.V
        Vt = torch.from_numpy(Vfull)
        if args.transport == "nccl":
            Vt = Vt.cuda()
        dist.all_reduce(Vt)
        Vfull = Vt.cpu().numpy()
        if rank == 0:
            ref_op, _ = run_snp(svd, src, N, k, bands, maxp, 0, 1, local, {}, emu, prec)
            assert_usv_close(op.U, op.S, Vfull, ref_op.U, ref_op.S, ref_op.V, eig_rtol=1e-9, min_corr=1 - 1e-9)
            if emu:  # flip_UV fixed the signs: compare without sign alignment
                assert np.abs(op.U - ref_op.U).max() < 1e-8
            print(f"snp-shard svd={svd} emu={emu} prec={prec} world={world}: == single GPU; S rel err",
                  float(np.max(np.abs(op.S - ref_op.S) / ref_op.S)), flush=True)
            ref_op.close()
        op.close()
        dist.barrier()
    # ---- sample-sharded (int8 route): winSVD resident, winSVD / sSVD streamed with the tile cache, missing calls
    for svd, bands, maxp, ooc, src in ((2, 16, 6, False, packed), (2, 16, 6, True, packed), (1, 64, 4, True, packed),
                                       (2, 16, 6, True, packed_m)):
        op, (s0, s1) = run_samples(svd, src, N, k, bands, maxp, rank, world, local, kw, ooc)
        U = gather_rows(op.U, world)
        F = op.F()
        hits = int(op.timers().cache_hits)
        if rank == 0:
            ref_op, _ = run_samples(svd, src, N, k, bands, maxp, 0, 1, local, {}, ooc)
            assert np.array_equal(F, ref_op.F()), "allele frequencies of the sample shards are not bit-identical"
            assert_usv_close(U, op.S, op.V, ref_op.U, ref_op.S, ref_op.V, eig_rtol=1e-9, min_corr=1 - 1e-9)
            if ooc:
                assert hits > 0, "streamed blocks never came from the HBM tile cache"
            print(f"sample-shard svd={svd} ooc={ooc} miss={src is packed_m} world={world}: == single GPU; S rel err",
                  float(np.max(np.abs(op.S - ref_op.S) / ref_op.S)), "cache hits", hits, flush=True)
            ref_op.close()
        op.close()
        dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK", flush=True)


if __name__ == "__main__":
    main()
