#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on the config it is quoted on.

metric  : time-to-top-k PCs (s) at the 500k x 500k bed; genotype GB/s per power-iteration pass
workload: configs[2] — UK-Biobank-scale synthetic bed, N = 500,000 samples x M = 500,000 SNPs, k = 40
          (l = 80), PCAone's window-based RSVD (winSVD, 64 windows), out-of-core block plan (64 blocks =
          the 64 windows, what `-m` gives at this size), int8x3 tensor-core route.
step    : ONE complete PCA = RsvdOpData::computeUSV with the reference's defaults (--maxp 20,
          --tol-rsvd 1e-4, winSVD minimum of log2(64)+1 = 7 epochs): every epoch is computeGandH over the
          64 blocks (all Omega updates of that epoch) plus the dense stage. ms_per_step = time to PCs.
value   : packed GB per pass = epochs x (N/4 x M bytes) / time, whole job, inputs resident in HBM (the
          streamed blocks' re-tiled operands are in the library's HBM tile cache when the timed region
          starts), device-timed, max over ranks.
e2e     : the same PCA through the C-ABI from HOST buffers: the packed bed sits in pinned host memory,
          every step invalidates the tile cache (pcaone_set_host_source), streams all 62.5 GB
          host->device inside the timed region (double-buffered, overlapped with epoch 0) and reads
          U, S, V back.
N > 1   : STRONG scaling of that one PCA, sample-sharded (include/pcaone_b200.h: shard_samples): rank r
          owns samples [r N/W, (r+1) N/W) of every SNP; per window the exact int64 partial sums of
          G_b = X_b^T Omega (7.8k x 80) are summed over NCCL, H / Omega stay row-sharded, the
          orthonormalisation exchanges l x l Gram matrices. (SNP-sharding would exchange the 320 MB
          N x l partial H at each of the 133 Omega updates: 64x the bytes at this shape.)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2] [--scale f]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "time-to-top-k PCs (s) at 500k x 500k bed; genotype GB/s per power-iteration pass"
UNIT = "GB/s"
BANDS = 64
SAMPLE_BLOCKS = 8           # fixed sample grid of the synthetic bed (world sizes 1, 2, 4, 8 cut along it)
WORKLOADS = {
    # name: (N samples, M SNPs, k, out-of-core plan, description)
    "c3": (500_000, 500_000, 40, True, "configs[2]: UK-Biobank-scale bed, winSVD, out-of-core block plan"),
    "c2": (10_000, 1_000_000, 20, False, "configs[1]: winSVD in-memory"),
}
CPU_SAMPLES_BASELINE = 1024   # cpu_baseline leg: this many samples x ALL SNPs, one full PCA (44.6 s on the box's 16 cores)
CPU_SAMPLES_REF_ARM = 256     # --impl reference: K + W full PCAs must end within minutes, so the sample is 256 samples x
CPU_SNP_FRACTION_REF_ARM = 8  # every 8th of the SNP axis (a PCA on 256 x 500k took 37 s; per-pass time scales with N x M)


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region, sampled in-process through NVML
    (nvidia_ml_py) every 5 ms (PCAONE_BENCH_CLOCK_MS; 5 vs 20 ms: no change of `value`, 420.5-420.7 vs 420.5-422.9 GB/s); falls back to one `nvidia-smi` query when NVML is unavailable.
    (A polling `nvidia-smi -lms` child process contends for the driver lock and slows down a
    launch-bound timed region several-fold, so it is not used.)"""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.period_s = float(os.environ.get("PCAONE_BENCH_CLOCK_MS", "5")) * 1e-3
        self.samples = []
        self.stop_flag = threading.Event()
        self.t = None
        self.nv = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical GPUs: map the CUDA ordinal through CUDA_VISIBLE_DEVICES
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.idx
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.idx < len(ids) and ids[self.idx].isdigit():
                    phys = int(ids[self.idx])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((sm, mx, rs))
            except Exception:
                pass
            self.stop_flag.wait(self.period_s)

    def stop(self):
        if self.nv is None:
            return self._smi_once()
        self.stop_flag.set()
        self.t.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = sorted({n for _, _, rs in self.samples for n, bit in names.items() if rs & bit})
        sm = [x[0] for x in self.samples]
        mx = [x[1] for x in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": f"nvml in-process, {self.period_s * 1e3:.0f} ms"}

    def _smi_once(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                 capture_output=True, text=True, timeout=20).stdout.strip().split(",")
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]),
                    "reasons": [n for n, v in zip(names, out[2:6]) if v.strip().lower().startswith("active")],
                    "samples": 1, "source": "nvidia-smi, one query after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}



def workload_dims(args):
    n, m, k, ooc, desc = WORKLOADS[args.workload]
    if args.scale != 1.0:
        n = max(SAMPLE_BLOCKS * 64, int(n * args.scale) // (SAMPLE_BLOCKS * 4) * (SAMPLE_BLOCKS * 4))
        m = max(BANDS * 64, int(m * args.scale))
    return n, m, k, ooc, desc


def cpu_reference_pca(n_s, m, k, pcas, warm, threads, seed=1, m_s=None):
    """The reference's own CPU computeUSV (oracle/_ref = unmodified PCAone, winSVD in-core, -S, defaults
    --maxp 20 --tol-rsvd 1e-4) on a bounded sample: n_s samples of the same population model x ALL m
    SNPs. Sampling the SAMPLE axis keeps both the per-pass GEMM work and the Omega-update work
    proportional to n_s, so packed GB/s per pass and (time to PCs) / n_s carry over to the full N."""
    import torch
    from oracle import ref
    from pcaone_b200 import synth

    if not ref.available():
        return None
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    packed = synth.torch_packed(n_s, m_s or m, k_pop=k + 4, seed=seed, device=dev, chunk=8192).cpu().numpy()
    tmp = tempfile.mkdtemp(prefix="pcaone_cpu_")
    prefix = os.path.join(tmp, "s")
    synth.write_bed_from_packed(prefix, packed, n_s)
    ref.lib().ref_set_threads(threads)
    t0 = time.perf_counter()
    r = ref.Ref(f"PCAone -b {prefix} -k {k} -d 2 -S -o {tmp}/o -n {threads}", threads=threads)
    r.new_op()
    load_s = time.perf_counter() - t0
    times, epochs = [], []
    for i in range(warm + pcas):
        t = time.perf_counter()
        r.compute_usv(20, 1e-4, want=False)
        dt = time.perf_counter() - t
        if i >= warm:
            times.append(dt)
            epochs.append(r.last_epochs())
    r.close()
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)
    return {"times": times, "epochs": epochs, "load_s": load_s, "bytes_per_pass": packed.shape[0] * packed.shape[1]}


def run_reference_arm(args, rank):
    """--impl reference: PCAone's CPU implementation of the same PCA on the box's host cores."""
    if rank != 0:
        return
    n, m, k, ooc, desc = workload_dims(args)
    threads = os.cpu_count() or 1
    n_s = min(n, CPU_SAMPLES_REF_ARM)
    m_s = max(BANDS * 64, m // CPU_SNP_FRACTION_REF_ARM)
    res = cpu_reference_pca(n_s, m, k, args.steps, min(args.warmup, 1), threads, m_s=m_s)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpcaone_ref.so was not built"}))
        return
    tot = sum(res["times"])
    passes = sum(res["epochs"])
    val = res["bytes_per_pass"] * passes / tot / 1e9
    scale_up = (n / n_s) * (m / m_s)
    sample = (f"{n_s} of {n} samples (same population model) x {m_s} of {m} SNPs, unmodified PCAone winSVD in-core -S, "
              f"defaults --maxp 20 --tol-rsvd 1e-4: {res['epochs'][0]} epochs per PCA, {tot / len(res['times']):.2f} s per PCA "
              f"on the sample; GEMM and Omega-update work per pass both scale with samples x SNPs, so packed GB/s per pass "
              f"carries over and time to PCs scales by {scale_up:.0f}; Eigen built-in GEMM, no MKL; "
              f"{min(args.warmup, 1)} warm-up PCA")
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(res["times"]), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "time_to_pcs_s": tot / len(res["times"]) * scale_up,
            "time_to_pcs_note": f"measured on {n_s} samples x {m_s} SNPs, scaled by {scale_up:.0f}",
            "config": {"workload": f"{desc}, N={n} x M={m}, k={k} (bounded CPU sample)", "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="scale N and M (debug; not a bench number)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--shard", default="auto", choices=["auto", "samples", "snps"])
    ap.add_argument("--precision", default="int8x3", choices=["int8x2", "int8x3", "int8x4"],
                    help="slices of the exact int8 tensor-core route")
    args = ap.parse_args()
    if args.warmup < 3 and args.scale == 1.0:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from pcaone_b200 import dist as pdist
    from pcaone_b200 import halko, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pcaone_b200 has no CPU fallback")
    rank, world, local = pdist.init_process_group_from_env("nccl")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    n, m, K, ooc, desc = workload_dims(args)
    if SAMPLE_BLOCKS % world != 0:
        raise SystemExit(f"--gpus must divide {SAMPLE_BLOCKS}")
    peaks, peak_src = _peaks()
    prec = {"int8x2": 2, "int8x3": 3, "int8x4": 4}[args.precision]
    l = 2 * K
    # exchange volume per Omega update decides the split (include/pcaone_b200.h, shard_samples)
    by_samples = world > 1 and (args.shard == "samples" or (args.shard == "auto" and -(-m // BANDS) < n))
    by_snps = world > 1 and not by_samples

    # ---- this rank's part of the ONE synthetic bed, generated on the GPU tile by tile into pinned host memory
    nblk = (-(-n // SAMPLE_BLOCKS) + 3) // 4 * 4          # block boundaries on whole bed bytes; the last block may be short

    def blk(sb):
        lo = min(n, sb * nblk)
        return lo, min(n, lo + nblk) - lo

    if by_samples or world == 1:
        sb0, sb1 = (rank * SAMPLE_BLOCKS // world, (rank + 1) * SAMPLE_BLOCKS // world) if by_samples else (0, SAMPLE_BLOCKS)
        samp0 = blk(sb0)[0]
        n_loc = blk(sb1 - 1)[0] + blk(sb1 - 1)[1] - samp0
        snp_idx = None
        m_loc = m
    else:
        sb0, sb1, samp0, n_loc = 0, SAMPLE_BLOCKS, 0, n
        snp_idx, w_start, w_stop = pdist.shard_windows(m, BANDS, rank, world)
        m_loc = len(snp_idx)
    bpr = synth.bytes_per_snp(n_loc)
    t_gen = time.perf_counter()
    host = torch.empty((m_loc, bpr), dtype=torch.uint8, pin_memory=True)
    chunk = max(64, min(4096, (1 << 28) // max(nblk, 1)))   # ~256M genotypes per tile
    if snp_idx is None:
        stage = torch.empty((chunk, bpr), dtype=torch.uint8, device=dev)   # one contiguous D2H per SNP chunk
        for s0 in range(0, m, chunk):
            mm = min(chunk, m - s0)
            for sb in range(sb0, sb1):
                lo, cnt = blk(sb)
                tile = synth.torch_packed_tile(n, lo, cnt, s0, mm, k_pop=K + 4, seed=1, device=dev)
                c0 = (lo - samp0) // 4
                stage[:mm, c0:c0 + tile.shape[1]] = tile
            host[s0:s0 + mm].copy_(stage[:mm], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        del stage
    else:
        # SNP shard: this rank's SNPs of every window (rows of the same tiles)
        pos = 0
        sel = torch.from_numpy(snp_idx).to(dev)
        for s0 in range(0, m, chunk):
            mm = min(chunk, m - s0)
            mine = sel[(sel >= s0) & (sel < s0 + mm)] - s0
            if mine.numel() == 0:
                continue
            for sb in range(SAMPLE_BLOCKS):
                lo, cnt = blk(sb)
                tile = synth.torch_packed_tile(n, lo, cnt, s0, mm, k_pop=K + 4, seed=1, device=dev)
                c0 = lo // 4
                host[pos:pos + mine.numel(), c0:c0 + tile.shape[1]].copy_(tile[mine])
            pos += int(mine.numel())
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    gen_s = time.perf_counter() - t_gen
    bytes_per_pass = m * synth.bytes_per_snp(n)           # the whole job's packed bed

    p = halko.Param(k=K, svd=2, bands=BANDS, maxp=20, tol=1e-4, no_shuffle=True, device=local,
                    memory=64.0 if ooc else 0.0, precision=prec)
    d = halko.FileBed(p, packed=host if ooc else host.to(dev), nsamples=n_loc)
    if by_snps:
        d.start, d.stop, d.nblocks, d.bandFactor = w_start, w_stop, len(w_start), 1
    elif ooc:
        bs = -(-m // BANDS)   # Data.cpp:66-69: nblocks < bands -> blocksize = ceil(M / bands)
        d.start = np.arange(BANDS, dtype=np.uint64) * np.uint64(bs)
        d.stop = np.minimum(d.start + np.uint64(bs - 1), np.uint64(m - 1))
        d.nblocks, d.blocksize, d.bandFactor = BANDS, bs, 1
    op = halko.FancyRsvdOpData(d, p.k, p.oversamples, rank=rank, world=world, nsnps_total=m,
                               library_comm=world > 1, shard_samples=by_samples, nsamples_total=n,
                               sample_offset=samp0)
    op.setFlags(False, True)
    stream = torch.cuda.ExternalStream(op.L.pcaone_stream(op.h))

    def barrier():
        if world > 1:
            dist.barrier()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    import ctypes as C
    ep = C.c_int(0)
    df = C.c_double(0)

    def pca():
        op._chk(op.L.pcaone_compute_usv(op.h, p.maxp, C.c_double(p.tol), C.byref(df), C.byref(ep)))
        return ep.value

    def new_data():
        """the host hands the bed over again: the library forgets its cached tiles and allele frequencies"""
        if ooc:
            op._chk(op.L.pcaone_set_host_source(op.h, halko._vp(host), m_loc))

    # ---- value leg: operands resident in HBM (tile cache warm)
    sampler = ClockSampler(local)
    sampler.start()
    op.enable_timing(False)
    for _ in range(args.warmup):
        epochs = pca()
    op.sync()
    op.timers(reset=True)
    sampler.samples.clear()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    passes = 0
    for _ in range(args.steps):
        passes += pca()
    e1.record(stream)
    op.sync()
    torch.cuda.synchronize()
    barrier()
    dev_ms = maxr(e0.elapsed_time(e1))
    clocks = sampler.stop()
    tm0 = op.timers(reset=True)
    gpu_launches = int(tm0.kernel_launches)
    value = bytes_per_pass * passes / (dev_ms * 1e-3) / 1e9
    op._fetch_usv()
    eig = op.S ** 2 / m
    u_orth = float(np.abs(op.U.T @ op.U - (np.eye(K) if (world == 1 or not by_samples) else 0)).max()) if world == 1 else None

    # ---- instrumented replica (per-kernel CUDA events) for the roofline object
    op.enable_timing(True)
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record(stream)
    ep_i = pca()
    i1.record(stream)
    op.sync()
    torch.cuda.synchronize()
    instr_ms = i0.elapsed_time(i1)
    tm = op.timers(reset=True)
    flops_gemm_rank = 2.0 * n_loc * m_loc * l * ep_i      # one of the two GEMMs, all epochs of the PCA, this rank
    tg, th = tm.tc_g_ms, tm.tc_h_ms
    dom, dom_ms = ("k_tc_gemm (H pass)", th) if th >= tg else ("k_tc_gemm (G pass)", tg)
    peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    achieved_tf = flops_gemm_rank / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else None
    int8_peak = None
    try:
        int8_peak = json.load(open(os.path.join(ROOT, "profiles", "r02_int8_peak.json")))
    except Exception:
        pass
    nl = int(tm.gemm_h_launches if "H pass" in dom else tm.gemm_g_launches)
    packed_per_launch = bytes_per_pass / world * ep_i / max(nl, 1)   # algorithmic packed bytes of an average launch
    traffic, traffic_note = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_tc_gemm_traffic.json")))
        leg = tj["h_pass" if "H pass" in dom else "g_pass"]
        traffic = leg["ratio"] * packed_per_launch
        traffic_note = (f"per launch: {leg['ratio']:.3f} x the {packed_per_launch / 1e9:.2f} GB of packed operand an average launch reads; "
                        f"the ratio (dram__bytes_read.sum + dram__bytes_write.sum over algorithmic packed bytes) is from ONE ncu --set "
                        f"full capture at 1/4 linear scale ({tj['source']}): kernel replay cannot save / restore the 130 GB of the "
                        f"full-size run; tensor pipe active {leg['tensor_pipe_active_pct']} % in that capture")
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf if achieved_tf else None, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": f"{peak_src} bf16 sustained (MEASURED_PEAKS.json has no int8 entry)",
                "note": (f"int8 Ozaki route, {prec} slices: achieved = algorithmic flops 2*N*M*l per GEMM pass x {ep_i} epochs / "
                         f"the kernel's CUDA-event time over those {nl} launches ({dom_ms / max(nl, 1):.3f} ms per launch of one "
                         f"{-(-m // BANDS)}-SNP block); the tensor cores execute {prec}x that as exact int8 MACs"),
                "launches": nl, "ms_per_launch": dom_ms / max(nl, 1),
                "algorithmic_flops_per_launch": flops_gemm_rank / max(nl, 1),
                "tc_g_ms_per_pca": tg, "tc_h_ms_per_pca": th, "gemm_g_ms_per_pca": tm.gemm_g_ms, "gemm_h_ms_per_pca": tm.gemm_h_ms,
                "orth_ms_per_pca": tm.orth_ms, "small_stage_ms_per_pca": tm.small_ms, "allreduce_ms_per_pca": tm.allreduce_ms,
                "omega_updates_per_pca": int(tm.omega_updates), "epochs": ep_i, "cache_hits": int(tm.cache_hits),
                "tc_ranges": int(tm.tc_ranges), "fp64_ranges": int(tm.fp64_ranges),
                "hbm_algorithmic_gbs": bytes_per_pass / world * ep_i / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else None,
                "timing_note": f"instrumented replica of one PCA: {instr_ms:.1f} ms with the per-scope events on"}
    if int8_peak:
        ip = int8_peak.get("int8_dense_tops")
        roofline["int8_peak_tops"] = ip
        # the tensor cores execute prec x the algorithmic flops as int8 ops: fraction of the measured kind::i8 rate
        roofline["frac_of_int8_peak"] = achieved_tf * prec / ip if (achieved_tf and ip) else None
        roofline["int8_peak_source"] = int8_peak.get("source")
    # one late pass on its own (pi >= 6: plain power iteration, one Omega update)
    op.timers(reset=True)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record(stream)
    op._chk(op.L.pcaone_compute_gandh(op.h, 7))
    op._chk(op.L.pcaone_small_stage(op.h))
    l1.record(stream)
    op.sync()
    torch.cuda.synchronize()
    tl = op.timers(reset=True)
    late_ms = maxr(l0.elapsed_time(l1))
    op.enable_timing(False)
    tcl = max(tl.tc_g_ms, tl.tc_h_ms)
    fl_late = 2.0 * n_loc * m_loc * l
    roofline["late_pass"] = {
        "what": "one epoch with pi >= 6 (one Omega update) + dense stage, untimed leg", "ms": late_ms,
        "gbs": bytes_per_pass / (late_ms * 1e-3) / 1e9, "tc_g_ms": tl.tc_g_ms, "tc_h_ms": tl.tc_h_ms,
        "gemm_g_ms": tl.gemm_g_ms, "gemm_h_ms": tl.gemm_h_ms, "orth_ms": tl.orth_ms, "small_stage_ms": tl.small_ms,
        "allreduce_ms": tl.allreduce_ms,
        "achieved": fl_late / (tcl * 1e-3) / 1e12 if tcl > 0 else None,
        "frac": fl_late / (tcl * 1e-3) / 1e12 / peak_tf if tcl > 0 else None}

    # ---- end-to-end leg: host buffers -> PCs, every step streams the bed again
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.steps
        for _ in range(1):
            new_data()
            if not ooc:
                op._chk(op.L.pcaone_upload_bed(op.h, halko._vp(host), m_loc, 0))
                op._chk(op.L.pcaone_allele_freq(op.h))
            pca()
            op._fetch_usv()
        op.sync()
        op.timers(reset=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            new_data()
            if not ooc:
                op._chk(op.L.pcaone_upload_bed(op.h, halko._vp(host), m_loc, 0))
                op._chk(op.L.pcaone_allele_freq(op.h))
            pe = pca()
            op._fetch_usv()
        op.sync()
        barrier()
        e2e_s = maxr(time.perf_counter() - t0)
        tm2 = op.timers(reset=True)
        h2d = torch.tensor([float(tm2.h2d_bytes)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(h2d)
        e2e = {"value": bytes_per_pass * pe * e2e_steps / e2e_s / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d.item() // e2e_steps), "d2h_bytes_per_step": int(tm2.d2h_bytes // e2e_steps),
               "ms_per_step": 1e3 * e2e_s / e2e_steps, "time_to_pcs_s": e2e_s / e2e_steps, "steps": e2e_steps,
               "cache_hits_per_step": int(tm2.cache_hits // e2e_steps),
               "timing": "host wall clock between stream syncs, max over ranks; pinned host bed -> U, S, V on the host"}
    op.close()
    del host

    # ---- CPU baseline (rank 0, N = 1 only): the compiled reference on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        n_s = min(n, CPU_SAMPLES_BASELINE)
        res = cpu_reference_pca(n_s, m, K, 1, 0, threads)
        if res is not None:
            tot = sum(res["times"])
            cpu = {"value": res["bytes_per_pass"] * sum(res["epochs"]) / tot / 1e9, "unit": UNIT, "cores": threads,
                   "kind": "reference", "time_to_pcs_s_extrapolated": tot * n / n_s,
                   "sample": (f"{n_s} of {n} samples (same population model) x all {m} SNPs, unmodified PCAone winSVD "
                              f"in-core -S, one full PCA: {res['epochs'][0]} epochs in {tot:.1f} s (+ {res['load_s']:.1f} s "
                              f"read/decode); per-pass work and Omega-update work both scale with the sample count, so "
                              f"GB/s carries over and time to PCs scales by N / {n_s}; Eigen GEMM, no MKL")}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": "oracle/_ref/libpcaone_ref.so missing"}

    if rank == 0:
        shard = "one GPU" if world == 1 else (f"sample-shard x{world}" if by_samples else f"snp-shard x{world}")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": f"int8x{prec} (exact integer products, FP64 epilogue)",
                "data": "synthetic", "time_to_pcs_s": dev_ms * 1e-3 / args.steps, "epochs_per_pca": passes / args.steps,
                "config": {"workload": f"{desc}: N={n} x M={m}, k={K}, l={l}, {BANDS} windows, no-shuffle; step = one complete "
                                       f"PCA (computeUSV, --maxp 20 --tol-rsvd 1e-4)",
                           "l2": f"inputs ({2 * bytes_per_pass / world / 1e9:.1f} GB of tiled operands per GPU) are larger than L2; no flush needed",
                           "parallelism": shard, "synthetic_bed_s": gen_s,
                           "top_eigenvalues": eig[:3].tolist(), "U_orthonormality_err": u_orth},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches,
                "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
