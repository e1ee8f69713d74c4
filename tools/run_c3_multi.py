"""configs[2] (UK-Biobank-scale synthetic bed, N=500k x M=500k, k=40, winSVD) SNP-sharded over the GPUs
of one box: STRONG scaling of one PCA. Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29541 tools/run_c3_multi.py [--scale 1.0] [--out gpurun_out/c3_multi.jsonl]

Every rank owns M / world SNPs of every window. With >= 2 GPUs the shard (62.5 GB /
world packed bytes, x3 with the two re-tiled copies of the int8 route) is RESIDENT in HBM, so the
host->device link that binds the one-GPU run (tools/run_configs.py c3: 8.4 s, 53 GB/s per pass)
drops out; what remains is tensor work plus one NCCL allreduce of the N x l partial H per Omega
update (320 MB at l = 80) and of the l x l Gram of G per epoch. Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from pcaone_b200 import dist as pdist  # noqa: E402
from pcaone_b200 import halko, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="multiply M by this (debug)")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = pdist.init_process_group_from_env("nccl")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    N, k = 500_000, 40
    M = int(500_000 * args.scale)
    s, e = pdist.shard_range(M, rank, world)
    m_local = e - s
    bpr = synth.bytes_per_snp(N)
    free_b, _ = torch.cuda.mem_get_info()
    need = 3.2 * m_local * bpr + 12e9
    if need > free_b:
        raise SystemExit(f"rank {rank}: resident shard needs {need / 1e9:.0f} GB, {free_b / 1e9:.0f} GB free: use more GPUs")
    packed = torch.empty((m_local, bpr), dtype=torch.uint8, device=dev)
    chunk = 8_000
    for c0 in range(0, m_local, chunk):
        m = min(chunk, m_local - c0)
        packed[c0:c0 + m].copy_(synth.torch_packed(N, m, k_pop=k + 4, seed=100 + s + c0, device=dev, chunk=2048))
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    hook = pdist.make_allreduce_hook() if world > 1 else None
    p = halko.Param(k=k, svd=2, bands=64, maxp=20, tol=1e-4, no_shuffle=True, device=local, precision=3)
    d = halko.FileBed(p, packed=packed, nsamples=N)
    t0 = time.perf_counter()
    op = halko.FancyRsvdOpData(d, p.k, p.oversamples, rank=rank, world=world, nsnps_total=M, allreduce=hook)
    op.setFlags(False, True)
    op.sync()
    setup_s = time.perf_counter() - t0

    def timed_run():
        dist.barrier()
        torch.cuda.synchronize()
        t = time.perf_counter()
        op.computeUSV(p.maxp, p.tol)
        op.sync()
        dist.barrier()
        return time.perf_counter() - t

    first = timed_run()          # includes building the tiled operand copies
    secs = timed_run()
    ep = op.epochs
    tm = op.timers(reset=True)
    # late passes (pi >= 6: one Omega update per pass)
    dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for pi in (6, 7, 8):
        op._chk(op.L.pcaone_compute_gandh(op.h, pi))
    op.sync()
    dist.barrier()
    pt = (time.perf_counter() - t) / 3
    mx = torch.tensor([first, secs, pt], dtype=torch.float64, device=dev)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    first, secs, pt = (float(x) for x in mx.tolist())
    U = op.U
    if rank == 0:
        rec = {"config": "C3-multi", "n_gpus": world,
               "workload": f"winSVD in-memory SNP-sharded N={N} M={M} k={k} l={2 * k} int8x3, {m_local} SNPs per GPU",
               "bytes_per_pass": M * bpr, "time_to_pcs_s": secs, "first_run_s": first, "upload_af_s": setup_s, "epochs": ep,
               "late_pass_ms": 1e3 * pt, "gbs_per_late_pass": M * bpr / pt / 1e9,
               "allreduce_ms_total": tm.allreduce_ms, "omega_updates": int(tm.omega_updates),
               "tc_ranges": int(tm.tc_ranges), "fp64_ranges": int(tm.fp64_ranges),
               "U_orthonormality_err": float(np.abs(U.T @ U - np.eye(k)).max()),
               "eigvals_top5": (op.S[:5] ** 2 / M).tolist(),
               "one_gpu_streamed_reference": "tools/run_configs.py c3: 8.41 s to PCs, 1172.5 ms per late pass (profiles/r01_configs_session8.jsonl)"}
        line = json.dumps(rec)
        print(line, flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(line + "\n")
    op.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
