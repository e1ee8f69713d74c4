"""Run BASELINE.json's five configs on ONE B200 at their named sizes (or a stated scale) and
print one JSON line per config: time-to-top-k PCs (s), packed GB/s per power-iteration pass,
and an accuracy / property check. Run under gpurun; results are copied to profiles/.

    python tools/run_configs.py [c1 c2 c3 c4 c5] [--scale S] [--out gpurun_out/configs.jsonl]

The CPU reference (oracle/_ref) is only the checker of config 1 here (the one config the
reference runs in-core on CPU today); it is never on the measured path.
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

from pcaone_b200 import _lib, halko, ld, synth  # noqa: E402


def _emit(out, rec):
    line = json.dumps(rec)
    print(line, flush=True)
    if out:
        with open(out, "a") as f:
            f.write(line + "\n")


def _corr_cols(A, B):
    A = A / np.linalg.norm(A, axis=0)
    B = B / np.linalg.norm(B, axis=0)
    return np.abs((A * B).sum(0))


def _run(op, p):
    """computeUSV to convergence; returns (seconds, epochs) device-synchronised."""
    op.sync()
    t0 = time.perf_counter()
    op.computeUSV(p.maxp, p.tol)  # includes the D2H of U, S, V
    return time.perf_counter() - t0, op.epochs


def _pass_time(op, pi_list):
    """seconds per computeGandH pass (averaged over the given epochs), device-synchronised."""
    op.sync()
    t0 = time.perf_counter()
    for pi in pi_list:
        op._chk(op.L.pcaone_compute_gandh(op.h, pi))
    op.sync()
    return (time.perf_counter() - t0) / len(pi_list)


def c1(args, out):
    N, M, k = 2504, int(100_000 * args.scale), 10
    packed = synth.torch_packed(N, M, k_pop=k + 4, seed=1, device="cuda:0", chunk=16384)
    res = {}
    for name, prec in (("fp64", 0), ("int8x3", 3)):
        p = halko.Param(k=k, svd=1, maxp=20, tol=1e-4, precision=prec)
        t0 = time.perf_counter()
        d = halko.FileBed(p, packed=packed, nsamples=N)
        op = halko.NormalRsvdOpData(d, p.k, p.oversamples)
        op.setFlags(False, True)
        secs, ep = _run(op, p)
        total = time.perf_counter() - t0
        pt = _pass_time(op, [1, 2, 3])
        res[name] = dict(U=op.U.copy(), S=op.S.copy(), secs=secs, total=total, epochs=ep, pass_s=pt)
        op.close()
    rec = {"config": "C1", "workload": f"sSVD in-memory N={N} M={M} k={k}", "bytes_per_pass": M * packed.shape[1]}
    for name in res:
        r = res[name]
        rec[name] = {"time_to_pcs_s": r["secs"], "incl_upload_af_s": r["total"], "epochs": r["epochs"],
                     "pass_ms": 1e3 * r["pass_s"], "gbs_per_pass": rec["bytes_per_pass"] / r["pass_s"] / 1e9}
    rec["int8x3_vs_fp64"] = {"eig_rel": float(np.max(np.abs(res["int8x3"]["S"] ** 2 - res["fp64"]["S"] ** 2) / res["fp64"]["S"] ** 2)),
                             "min_abs_corr": float(_corr_cols(res["int8x3"]["U"], res["fp64"]["U"]).min())}
    # the reference itself on the same bed (CPU, all cores): time and parity
    try:
        from oracle import ref
        if ref.available() and not args.no_ref:
            import tempfile
            tmp = tempfile.mkdtemp(prefix="c1_")
            synth.write_bed_from_packed(os.path.join(tmp, "s"), packed.cpu().numpy(), N)
            th = os.cpu_count() or 1
            t0 = time.perf_counter()
            r = ref.Ref(f"PCAone -b {tmp}/s -k {k} -d 1 -o {tmp}/o -n {th}", threads=th)
            r.new_op()
            r.set_flags(False, True)
            U, S, V = r.compute_usv(20, 1e-4)
            cpu_s = time.perf_counter() - t0
            r.close()
            rec["reference_cpu"] = {"time_to_pcs_s": cpu_s, "cores": th,
                                    "eig_rel_fp64": float(np.max(np.abs(res["fp64"]["S"] ** 2 - S ** 2) / S ** 2)),
                                    "min_abs_corr_fp64": float(_corr_cols(res["fp64"]["U"], U).min()),
                                    "eig_rel_int8x3": float(np.max(np.abs(res["int8x3"]["S"] ** 2 - S ** 2) / S ** 2)),
                                    "min_abs_corr_int8x3": float(_corr_cols(res["int8x3"]["U"], U).min())}
    except Exception as e:  # the checker is optional on the box
        rec["reference_cpu"] = {"error": str(e)[:200]}
    _emit(out, rec)


def c2(args, out):
    N, M, k = 10_000, int(1_000_000 * args.scale), 20
    packed = synth.torch_packed(N, M, k_pop=k + 4, seed=1, device="cuda:0", chunk=16384)
    rec = {"config": "C2", "workload": f"winSVD in-memory N={N} M={M} k={k} 64 windows", "bytes_per_pass": M * packed.shape[1]}
    res = {}
    for name, prec in (("int8x3", 3), ("fp64", 0)):
        p = halko.Param(k=k, svd=2, bands=64, maxp=20, tol=1e-4, no_shuffle=True, precision=prec)
        t0 = time.perf_counter()
        d = halko.FileBed(p, packed=packed, nsamples=N)
        op = halko.FancyRsvdOpData(d, p.k, p.oversamples)
        op.setFlags(False, True)
        secs, ep = _run(op, p)
        total = time.perf_counter() - t0
        secs2, _ = _run(op, p)  # second run: tiles / allocations warm
        pt = _pass_time(op, [6, 7, 8])
        res[name] = (op.U.copy(), op.S.copy())
        rec[name] = {"time_to_pcs_s": secs2, "first_run_s": secs, "incl_upload_af_s": total, "epochs": ep,
                     "late_pass_ms": 1e3 * pt, "gbs_per_late_pass": rec["bytes_per_pass"] / pt / 1e9}
        op.close()
    rec["int8x3_vs_fp64"] = {"eig_rel": float(np.max(np.abs(res["int8x3"][1] ** 2 - res["fp64"][1] ** 2) / res["fp64"][1] ** 2)),
                             "min_abs_corr": float(_corr_cols(res["int8x3"][0], res["fp64"][0]).min())}
    _emit(out, rec)


def c2m(args, out):
    """configs[1] with 1 % missing calls everywhere: every range takes the (count, mask) product pair
    on the int8 kernels; the FP64 DMMA route (LUT decode of code 01) is the checker."""
    N, M, k = 10_000, int(1_000_000 * args.scale), 20
    packed = synth.torch_packed(N, M, k_pop=k + 4, seed=1, miss=0.01, device="cuda:0", chunk=16384)
    rec = {"config": "C2m", "workload": f"winSVD in-memory N={N} M={M} k={k} 64 windows, 1% missing calls",
           "bytes_per_pass": M * packed.shape[1]}
    res = {}
    for name, prec in (("int8x3", 3), ("fp64", 0)):
        p = halko.Param(k=k, svd=2, bands=64, maxp=20, tol=1e-4, no_shuffle=True, precision=prec)
        d = halko.FileBed(p, packed=packed, nsamples=N)
        op = halko.FancyRsvdOpData(d, p.k, p.oversamples)
        op.setFlags(False, True)
        secs, ep = _run(op, p)
        secs2, _ = _run(op, p)
        pt = _pass_time(op, [6, 7, 8])
        tm = op.timers(reset=True)
        res[name] = (op.U.copy(), op.S.copy())
        rec[name] = {"time_to_pcs_s": secs2, "first_run_s": secs, "epochs": ep, "late_pass_ms": 1e3 * pt,
                     "gbs_per_late_pass": rec["bytes_per_pass"] / pt / 1e9, "tc_ranges": int(tm.tc_ranges),
                     "tc_miss_ranges": int(tm.tc_miss_ranges), "fp64_ranges": int(tm.fp64_ranges)}
        if prec == 3:
            rec["missing_fraction"] = op.missing_count() / (N * M)
        op.close()
    rec["int8x3_vs_fp64"] = {"eig_rel": float(np.max(np.abs(res["int8x3"][1] ** 2 - res["fp64"][1] ** 2) / res["fp64"][1] ** 2)),
                             "min_abs_corr": float(_corr_cols(res["int8x3"][0], res["fp64"][0]).min())}
    _emit(out, rec)


def _host_avail_gb():
    for ln in open("/proc/meminfo"):
        if ln.startswith("MemAvailable"):
            return int(ln.split()[1]) / 1e6
    return 0.0


def c3(args, out):
    N, k = 500_000, 40
    M = int(500_000 * args.scale)
    bpr = synth.bytes_per_snp(N)
    need_gb = M * bpr / 1e9
    avail = _host_avail_gb()
    while need_gb * 1.6 > avail and M > 50_000:  # stay well inside host RAM (a dead box is a strike)
        M //= 2
        need_gb = M * bpr / 1e9
    host = torch.empty((M, bpr), dtype=torch.uint8, pin_memory=True)
    chunk = 20_000
    for s in range(0, M, chunk):
        m = min(chunk, M - s)
        host[s:s + m].copy_(synth.torch_packed(N, m, k_pop=k + 4, seed=100 + s, device="cuda:0", chunk=2048))
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    p = halko.Param(k=k, svd=2, bands=64, maxp=20, tol=1e-4, no_shuffle=True, memory=1.0, precision=3)
    d = halko.FileBed(p, packed=host, nsamples=N)
    bs = -(-M // 64)
    d.start = np.arange(64, dtype=np.uint64) * np.uint64(bs)
    d.stop = np.minimum(d.start + np.uint64(bs - 1), np.uint64(M - 1))
    d.nblocks, d.blocksize, d.bandFactor = 64, bs, 1
    op = halko.FancyRsvdOpData(d, p.k, p.oversamples)
    op.setFlags(False, True)
    secs, ep = _run(op, p)
    tm = op.timers(reset=True)
    pt = _pass_time(op, [6, 7])
    U = op.U
    orth = float(np.abs(U.T @ U - np.eye(k)).max())
    rec = {"config": "C3", "workload": f"winSVD out-of-core (64 host-streamed blocks) N={N} M={M} k={k} l={2 * k} int8x3, 1 GPU",
           "host_mem_avail_gb": avail, "bytes_per_pass": M * bpr, "time_to_pcs_s": secs, "epochs": ep,
           "late_pass_ms": 1e3 * pt, "gbs_per_late_pass": M * bpr / pt / 1e9,
           "h2d_gb_total": tm.h2d_bytes / 1e9, "tc_ranges": int(tm.tc_ranges), "fp64_ranges": int(tm.fp64_ranges),
           "U_orthonormality_err": orth, "eigvals_top5": (op.S[:5] ** 2 / M).tolist()}
    op.close()
    _emit(out, rec)


def c4(args, out):
    """configs[3]: EMU on N = 50k x M = 500k, 10 % missing calls, k = 10 (l = 20), winSVD in memory. Legs: the int8 route
    with the FP64 correction over the missing calls (default), the same with PCAONE_EMU_TC=0 (update passes on the
    FP64 DMMA kernels, what round 1 measured), and for the first a late update pass next to a late pass without the
    fill (mean imputation only) with the per-kernel shares."""
    N, M, k = 50_000, int(500_000 * args.scale), 10
    l = 2 * k
    packed = synth.torch_packed(N, M, k_pop=k + 4, miss=0.10, seed=4, device="cuda:0", chunk=4096)
    legs = [("int8x3 + FP64 correction over the missing calls", 3, "1"), ("update passes on the FP64 DMMA kernels", 3, "0")]
    if args.c4_prec == 0:
        legs = [("FP64 DMMA kernels throughout", 0, "1")]
    legs = legs[:args.c4_legs]
    for name, prec, sw in legs:
        os.environ["PCAONE_EMU_TC"] = sw
        p = halko.Param(k=k, svd=2, bands=64, maxp=20, tol=1e-4, no_shuffle=True, emu=True, precision=prec)
        d = halko.FileBed(p, packed=packed, nsamples=N)
        op = halko.FancyRsvdOpData(d, p.k, p.oversamples)
        op.runEM()                       # warm-up: module load, tile build, buffers
        op.sync()
        op.timers(reset=True)
        t0 = time.perf_counter()
        iters = op.runEM()
        secs = time.perf_counter() - t0
        tm = op.timers(reset=True)
        U = op.U
        nmiss = op.missing_count()
        rec = {"config": "C4", "workload": f"EMU winSVD in-memory N={N} M={M} k={k} 10% missing: {name}",
               "bytes_per_pass": M * packed.shape[1], "time_to_pcs_s": secs, "genotypes_per_s": N * M / secs, "em_iterations": iters,
               "missing_fraction": nmiss / (N * M), "tc_ranges": int(tm.tc_ranges), "tc_emu_ranges": int(tm.tc_emu_ranges),
               "fp64_ranges": int(tm.fp64_ranges), "kernel_launches": int(tm.kernel_launches),
               "U_orthonormality_err": float(np.abs(U.T @ U - np.eye(k)).max()), "eigvals_top5": (op.S[:5] ** 2 / M).tolist()}
        # one late epoch (pi >= 6: two half-range products, one Omega update) with and without the fill
        op.enable_timing(True)
        for label, upd, pi in (("late_update_pass", True, 20), ("late_plain_pass", False, 20), ("first_update_pass", True, 0),
                               ("first_plain_pass", False, 0)):
            # pi = 0: the first epoch of a computeUSV, 64 windows with an Omega update after each
            op.setFlags(upd, False)
            op.computeGandH(19 if pi else 0, want=False)       # warm
            op.sync()
            op.timers(reset=True)
            t0 = time.perf_counter()
            op.computeGandH(pi, want=False)
            op.sync()
            ms = 1e3 * (time.perf_counter() - t0)
            t = op.timers(reset=True)
            rec[label] = {"ms": ms, "gemm_g_ms": t.gemm_g_ms, "gemm_h_ms": t.gemm_h_ms, "tc_g_ms": t.tc_g_ms, "tc_h_ms": t.tc_h_ms,
                          "emu_fix_ms": t.emu_fix_ms, "orth_ms": t.orth_ms, "tc_ranges": int(t.tc_ranges), "kernel_launches": int(t.kernel_launches),
                          "fp64_ranges": int(t.fp64_ranges), "gbs": M * packed.shape[1] / ms / 1e6}
            if upd and t.emu_fix_ms > 0:
                # FP64 work of the correction: per missing call and product a k-term dot and an l-term axpy
                rec[label]["emu_fix_gflops"] = 2.0 * 2.0 * nmiss * (k + l) / t.emu_fix_ms / 1e6
            if upd and t.fp64_ranges:
                rec[label]["dmma_tflops"] = 2.0 * 2.0 * N * M * l / (t.gemm_g_ms + t.gemm_h_ms) / 1e9
        op.close()
        _emit(out, rec)
    os.environ.pop("PCAONE_EMU_TC", None)
    # CPU baseline: the unmodified reference (oracle/_ref) on a bounded sample of the same workload — an EMU update pass
    # (read_block_update + the two products, Halko.cpp:188-222 in core) on 2,048 samples x 32,768 SNPs with 10 % missing
    try:
        from oracle import ref
        if ref.available() and not args.no_cpu:
            import tempfile
            ns, ms = 2048, 32768
            pk = synth.torch_packed(ns, ms, k_pop=k + 4, miss=0.10, seed=4, device="cuda:0", chunk=4096).cpu().numpy()
            tmp = tempfile.mkdtemp(prefix="c4_cpu_")
            synth.write_bed_from_packed(os.path.join(tmp, "s"), pk, ns)
            thr = os.cpu_count() or 1
            r = ref.Ref(f"PCAone -b {tmp}/s -k {k} -d 2 -S --emu -o {tmp}/o -n {thr}", threads=thr)
            r.new_op()
            t0 = time.perf_counter()
            r.run_em()
            em_s = time.perf_counter() - t0
            r.set_flags(True, False)
            r.time_gandh(19)
            ts = [r.time_gandh(20) for _ in range(3)]
            r.close()
            _emit(out, {"config": "C4-cpu", "workload": f"reference EMU (oracle/_ref, {thr} threads) on {ns} x {ms}, 10 % missing, k = {k}",
                        "em_run_s": em_s, "genotypes_per_s": ns * ms / em_s,
                        "late_pass_s": min(ts), "gbs_per_late_pass": ms * pk.shape[1] / min(ts) / 1e9,
                        "note": "in core the reference fills the missing calls once per computeUSV (fit_with_pi at pi = 0), so its late pass is two dense GEMMs; compare whole EM runs per genotype",
                        "cores": thr, "sample": f"{ns} samples x {ms} SNPs of the same population model"})
    except Exception as e:  # the baseline is a side note: never fail the config run on it
        print("# c4 cpu baseline skipped:", e)


def c5(args, out):
    """configs[4]: LD r2 over bp windows, N = 20k x M = 200k. Two legs: (1) the r2 tiles on the centred genotypes of the
    resident bed; (2) the whole ancestry-adjusted path in one go — PCA (k = 5, --ld: centred, unscaled), G -= U S V^T,
    column-centre, float32 round trip, r2 tiles — with no residual file and no 32 GB host matrix."""
    N, M = 20_000, int(200_000 * args.scale)
    packed = synth.torch_packed(N, M, k_pop=6, seed=5, device="cuda:0", chunk=8192)
    p = halko.Param(k=5, svd=1, ld=True, precision=3)
    d = halko.FileBed(p, packed=packed, nsamples=N)
    op = halko.NormalRsvdOpData(d, p.k, p.oversamples)
    # .bim of the synthetic bed: 22 chromosomes, 100 bp apart; --ld-bp 100000 => ~1000 SNPs per window
    per = -(-M // 22)
    chrom = np.minimum(np.arange(M) // per, 21)
    pos = (np.arange(M) % per + 1) * 100
    t0 = time.perf_counter()
    ws, we = ld.divide_pos_by_window(chrom, pos, 100_000)
    plan_s = time.perf_counter() - t0
    op.enable_timing(True)
    op.sync()
    t0 = time.perf_counter()
    r2 = ld.ld_r2_big(op, None, ws, we)
    secs = time.perf_counter() - t0
    tm = op.timers(reset=True)
    rec = {"config": "C5", "workload": f"LD r2 from the resident bed N={N} M={M}, --ld-bp 100000 (mean window {float(we.mean()):.0f} SNPs)",
           "pairs": int(r2.size), "host_window_plan_s": plan_s, "r2_total_s": secs, "kernel_ms": tm.ld_ms,
           "tiles": int(tm.ld_tiles), "gram_tflops_fp64": 2.0 * N * tm.ld_tiles * 128 * 128 / (tm.ld_ms * 1e-3) / 1e12 if tm.ld_ms else None,
           "r2_range": [float(r2.min()), float(r2.max())], "pairs_per_s": r2.size / secs}
    _emit(out, rec)
    # ---- ancestry-adjusted LD from the bed in one command
    op.enable_timing(False)
    op.setFlags(False, False)
    t0 = time.perf_counter()
    op.computeUSV(p.maxp, p.tol)
    pca_s = time.perf_counter() - t0
    op.enable_timing(True)
    op.timers(reset=True)
    t0 = time.perf_counter()
    r2a = ld.ld_adjusted_from_bed(op, ws, we, ld_stats=0)
    adj_s = time.perf_counter() - t0
    tm = op.timers(reset=True)
    rec = {"config": "C5-adjusted", "workload": f"--ld (k=5) + ancestry-adjusted r2 straight from the bed, N={N} M={M}, no .residuals file",
           "pairs": int(r2a.size), "pca_s": pca_s, "epochs": op.epochs, "residualise_plus_r2_s": adj_s, "k_ld_tiles_ms": tm.ld_ms,
           "tiles": int(tm.ld_tiles), "gram_tflops_fp64": 2.0 * N * tm.ld_tiles * 128 * 128 / (tm.ld_ms * 1e-3) / 1e12 if tm.ld_ms else None,
           "d2h_gb": tm.d2h_bytes / 1e9, "mean_abs_r2_shift_vs_unadjusted": float(np.abs(r2a - r2).mean()),
           "total_s": pca_s + adj_s}
    op.close()
    _emit(out, rec)


def f_bgen(args, out):
    """§8 f-1: BGEN-style float dosages (4 B per genotype resident in HBM), sSVD, FP64 DMMA with the
    decode fused into the operand load."""
    N, M, k = 10_000, int(200_000 * args.scale), 20
    g = torch.Generator(device="cuda:0")
    g.manual_seed(3)
    dos = torch.empty((M, N), dtype=torch.float32, device="cuda:0")
    for s0 in range(0, M, 8192):
        m = min(8192, M - s0)
        pfreq = torch.rand(m, 1, generator=g, device="cuda:0") * 0.9 + 0.05
        blk = (torch.rand(m, N, generator=g, device="cuda:0") < pfreq).float() + (torch.rand(m, N, generator=g, device="cuda:0") < pfreq).float()
        blk = (blk + 0.05 * torch.randn(m, N, generator=g, device="cuda:0")).clamp_(0, 2)
        blk[torch.rand(m, N, generator=g, device="cuda:0") < 0.01] = float("nan")
        dos[s0:s0 + m] = blk
    p = halko.Param(k=k, svd=1, maxp=20, tol=1e-4, precision=_lib.PREC_FP64)
    d = halko.FileBgen(p, dos)      # selection skipped for the synthetic matrix (no all-zero variant)
    op = halko.NormalRsvdOpData(d, p.k, p.oversamples)
    op.setFlags(False, True)
    secs, ep = _run(op, p)
    secs2, _ = _run(op, p)
    pt = _pass_time(op, [1, 2, 3])
    l = op.size()
    rec = {"config": "F1-bgen", "workload": f"sSVD on float32 dosages N={N} M={M} k={k}, 1% NaN, FP64 DMMA fused decode",
           "dosage_bytes": 4 * N * M, "time_to_pcs_s": secs2, "first_run_s": secs, "epochs": ep, "pass_ms": 1e3 * pt,
           "dosage_gbs_per_pass": 2 * 4.0 * N * M / pt / 1e9, "fp64_tflops": 4.0 * N * M * l / pt / 1e12,
           "note": "each pass reads the dosage matrix twice (G product, H product); FP64 tensor nominal ~37 TFLOP/s"}
    op.close()
    _emit(out, rec)


def f_beagle(args, out):
    """§8 f-1: Beagle genotype likelihoods, PCAngsd EM (allele-frequency EM, expected genotypes from the
    individual allele frequencies once per EM iteration, dense FP64 DMMA products)."""
    N, M, k = 2_000, int(200_000 * args.scale), 4
    rng = np.random.default_rng(5)
    pop = rng.integers(0, 5, N)
    P = np.zeros((2 * N, M), order="F")
    for s0 in range(0, M, 20_000):
        m = min(20_000, M - s0)
        pf = np.clip(rng.uniform(0.1, 0.9, (m, 1)) + 0.12 * rng.standard_normal((m, 5)), 0.05, 0.95)[:, pop]
        gt = rng.binomial(2, pf)
        d = rng.poisson(2.0, gt.shape)
        alt = rng.binomial(d, np.clip(gt / 2.0, 0.01, 0.99))
        lik = [np.clip(q / 2.0, 0.01, 0.99) ** alt * (1 - np.clip(q / 2.0, 0.01, 0.99)) ** (d - alt) for q in (0, 1, 2)]
        tot = lik[0] + lik[1] + lik[2]
        P[0::2, s0:s0 + m] = (lik[0] / tot).T
        P[1::2, s0:s0 + m] = (lik[1] / tot).T
    p = halko.Param(k=k, svd=1, maxp=20, tol=1e-4, maxiter=100, precision=_lib.PREC_FP64)
    d = halko.FileBeagle(p, P)
    t0 = time.perf_counter()
    d.prepare()
    maf_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    op = halko.run_pca_with_halko(d, p)
    op.sync()
    em_s = time.perf_counter() - t0
    rec = {"config": "F1-beagle", "workload": f"PCAngsd EM on genotype likelihoods N={N} M={M} k={k} (depth ~2x), sSVD, FP64",
           "gl_bytes": 16 * N * M, "maf_em_s_incl_upload": maf_s, "maf_em_iterations": d.maf_iters,
           "pcangsd_s_incl_upload": em_s, "em_iterations": op.em_iters, "eigvals": (op.S ** 2 / M).tolist()}
    op.close()
    _emit(out, rec)


def f_prune(args, out):
    """§8 f-2: greedy LD pruning on the device at config-5 size (r2 tiles never leave HBM)."""
    N, M = 20_000, int(200_000 * args.scale)
    packed = synth.torch_packed(N, M, k_pop=6, seed=5, device="cuda:0", chunk=8192)
    p = halko.Param(k=2, svd=1, ld=True)
    d = halko.FileBed(p, packed=packed, nsamples=N)
    op = halko.NormalRsvdOpData(d, p.k, p.oversamples)
    per = -(-M // 22)
    chrom = np.minimum(np.arange(M) // per, 21)
    pos = (np.arange(M) % per + 1) * 100
    ws, we = ld.divide_pos_by_window(chrom, pos, 100_000)
    af = op.F()
    op.sync()
    res = {}
    for tol in (0.2, 0.02):
        t0 = time.perf_counter()
        keep = ld.ld_prune_big(op, None, ws, we, tol, af)
        res[str(tol)] = {"seconds": time.perf_counter() - t0, "kept": int(keep.sum())}
    pairs = int((we.astype(np.int64) - 1).sum())
    rec = {"config": "F2-ldprune", "workload": f"LD pruning from the resident bed N={N} M={M}, --ld-bp 100000, MAF rule",
           "pairs": pairs, "by_r2_tol": res, "pairs_per_s": pairs / res["0.2"]["seconds"]}
    op.close()
    _emit(out, rec)


def f_iram(args, out):
    """§8 f-3: one ArnoldiOpData::perform_op at config-2 size (GEMV-shaped pass, l = 1)."""
    N, M = 10_000, int(1_000_000 * args.scale)
    packed = synth.torch_packed(N, M, k_pop=24, seed=1, device="cuda:0", chunk=16384)
    rec = {"config": "F3-iram-op", "workload": f"y = X X^T x, N={N} M={M}, resident bed", "bytes_per_op": M * packed.shape[1]}
    x = np.random.default_rng(0).standard_normal(N)
    ys = {}
    for name, prec in (("int8x4", 4), ("fp64", 0)):
        p = halko.Param(k=20, svd=1, precision=prec)
        d = halko.FileBed(p, packed=packed, nsamples=N)
        op = halko.ArnoldiOpData(d)
        op.setFlags(False, True)
        ys[name] = op.perform_op(x)
        op.sync()
        t0 = time.perf_counter()
        for _ in range(5):
            op.perform_op(x)
        dt = (time.perf_counter() - t0) / 5
        rec[name] = {"ms_per_op": 1e3 * dt, "packed_gbs": rec["bytes_per_op"] / dt / 1e9}
        op.close()
    rec["int8x4_vs_fp64_rel"] = float(np.abs(ys["int8x4"] - ys["fp64"]).max() / np.abs(ys["fp64"]).max())
    rec["note"] = "HBM-bound shape: the packed matrix is read twice per op (G product, H product)"
    _emit(out, rec)


def f_dense(args, out):
    """§8 a15: RsvdOne on a dense FP64 matrix (PCAoneR's entry point)."""
    from pcaone_b200.rsvd import RsvdOne
    rows, cols, k = int(100_000 * args.scale), 2_000, 20
    rng = np.random.default_rng(1)
    A = (rng.standard_normal((rows, 30)) * np.linspace(50, 5, 30)) @ rng.standard_normal((30, cols)) \
        + 0.1 * rng.standard_normal((rows, cols))
    rec = {"config": "A15-dense", "workload": f"RsvdOne {rows} x {cols} doubles, k={k}, os=20", "bytes": 8 * rows * cols}
    s = np.sqrt(np.linalg.eigvalsh(A.T @ A)[::-1][:k])     # exact singular values (checker)
    for name, (pp, w) in (("plain_p7", (7, 0)), ("windows64_p7", (7, 64))):
        r = RsvdOne(A, k, 20, 1)
        t0 = time.perf_counter()
        r.compute(pp, w)
        rec[name] = {"seconds_incl_upload": time.perf_counter() - t0}
        rec[name]["sv_rel_err_vs_lapack"] = float(np.max(np.abs(r.singularValues() - s) / s))
    _emit(out, rec)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--scale", type=float, default=1.0, help="multiply every M by this (debug)")
    ap.add_argument("--out", default="")
    ap.add_argument("--no-ref", action="store_true")
    ap.add_argument("--c4-prec", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baselines")
    ap.add_argument("--c4-legs", type=int, default=2, help="c4: 1 = the int8 route only, 2 = also the FP64 DMMA update passes")
    args = ap.parse_args()
    _lib.load()
    fns = {"c1": c1, "c2": c2, "c2m": c2m, "c3": c3, "c4": c4, "c5": c5, "bgen": f_bgen, "beagle": f_beagle, "prune": f_prune,
           "iram": f_iram, "dense": f_dense}
    for w in args.which:
        t0 = time.perf_counter()
        try:
            fns[w](args, args.out)
        except Exception as e:
            _emit(args.out, {"config": w.upper(), "error": repr(e)[:500]})
        torch.cuda.empty_cache()
        print(f"# {w} done in {time.perf_counter() - t0:.1f} s wall", file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
