#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

metric  : genotype GB/s per power-iteration pass (packed bed bytes / pass time)
workload: configs[1] — PCAone window-based RSVD (winSVD, 64 windows), N=10,000 samples x
          M=1,000,000 SNPs per GPU, k=20 (l=40), in-memory on B200. One "step" = one epoch of
          RsvdOpData::computeUSV: computeGandH (decode + X^T Omega + X G for every window and all
          Omega updates of that epoch) followed by the dense stage (QR(G) x2, B, SVD). Steps walk
          epochs pi = 0,1,2,... of one winSVD run (7 epochs = a complete default run).
value   : whole-job packed GB/s with the packed shard resident in HBM (device-timed, max over ranks)
e2e     : same metric through the C-ABI with the packed matrix in pinned HOST memory: every
          step streams all blocks host->device (double-buffered) and reads U,S back.
N > 1   : SNP-sharded, weak scaling (every rank owns its own 1M SNPs of a world*1M-SNP job);
          H (N x l) is all-reduced over NCCL at every Omega update, the l x l Gram of G once per epoch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
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

METRIC = "genotype GB/s per power-iteration pass"
UNIT = "GB/s"
N_SAMPLES, M_SNPS, K, BANDS = 10_000, 1_000_000, 20, 64
CPU_SAMPLE_SNPS = 32_768


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


def cpu_reference_pass(packed_host, n_samples, k, steps, warmup, threads):
    """Time the reference's own CPU computeGandH (oracle/_ref, unmodified PCAone) on a bounded
    sample of the workload: the first CPU_SAMPLE_SNPS SNPs, all samples, winSVD in-core, -S."""
    from oracle import ref
    from pcaone_b200 import synth

    if not ref.available():
        return None
    tmp = tempfile.mkdtemp(prefix="pcaone_cpu_")
    prefix = os.path.join(tmp, "s")
    synth.write_bed_from_packed(prefix, packed_host, n_samples)
    ref.lib().ref_set_threads(threads)
    t0 = time.perf_counter()
    r = ref.Ref(f"PCAone -b {prefix} -k {k} -d 2 -S -o {tmp}/o -n {threads}", threads=threads)
    r.new_op()
    load_s = time.perf_counter() - t0
    times = []
    for i in range(warmup + steps):
        t = r.time_gandh(i)  # epochs 0,1,2,... of one winSVD run, like the GPU arm
        if i >= warmup:
            times.append(t)
    r.close()
    nbytes = packed_host.shape[0] * packed_host.shape[1]
    return {"times": times, "load_s": load_s, "bytes": nbytes}


def run_reference_arm(args, rank):
    """--impl reference: PCAone's CPU implementation of the pass on the box's host cores."""
    if rank != 0:
        return
    import torch
    from pcaone_b200 import synth

    threads = os.cpu_count() or 1
    m = CPU_SAMPLE_SNPS if not args.small else 4096
    n = N_SAMPLES if not args.small else 1000
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    packed = synth.torch_packed(n, m, k_pop=K + 4, seed=1, device=dev, chunk=8192).cpu().numpy()
    res = cpu_reference_pass(packed, n, K, args.steps, args.warmup, threads)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpcaone_ref.so was not built"}))
        return
    tot = sum(res["times"])
    val = res["bytes"] * len(res["times"]) / tot / 1e9
    sample = (f"first {m} of {M_SNPS} SNPs x {n} samples (1/{M_SNPS // m} of the workload), winSVD in-core -S, "
              f"epochs {args.warmup}..{args.warmup + args.steps - 1}; Eigen built-in GEMM, no MKL")
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(res["times"]), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "configs[1]: winSVD N=10000 x M=1000000 k=20 (bounded CPU sample)",
                       "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=7)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--small", action="store_true", help="debug-sized workload (not a bench number)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--precision", default="int8x3", choices=["fp64", "int8x2", "int8x3", "int8x4"],
                    help="GEMM arithmetic: FP64 DMMA, or the exact int8 tensor-core path with 2/3/4 slices")
    args = ap.parse_args()
    if args.warmup < 3 and not args.small:
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
    n, m = (N_SAMPLES, M_SNPS) if not args.small else (1000, 65536)
    peaks, peak_src = _peaks()
    prec = {"fp64": 0, "int8x2": 2, "int8x3": 3, "int8x4": 4}[args.precision]

    # ---- synthetic packed shard, generated in HBM (seeded per rank)
    packed = synth.torch_packed(n, m, k_pop=K + 4, seed=1 + rank, device=f"cuda:{local}", chunk=16384)
    bpr = packed.shape[1]
    shard_bytes = m * bpr

    hook = pdist.make_allreduce_hook() if world > 1 else None

    def make_op(src, ooc):
        p = halko.Param(k=K, svd=2, bands=BANDS, maxp=20, tol=1e-4, no_shuffle=True, device=local,
                        memory=1.0 if ooc else 0.0, precision=prec)
        d = halko.FileBed(p, packed=src, nsamples=n)
        if ooc:  # 64 streamed blocks == the 64 windows (what -m gives when nblocks < bands)
            bs = -(-m // BANDS)
            d.start = np.arange(BANDS, dtype=np.uint64) * np.uint64(bs)
            d.stop = np.minimum(d.start + np.uint64(bs - 1), np.uint64(m - 1))
            d.nblocks, d.blocksize, d.bandFactor = BANDS, bs, 1
        op = halko.FancyRsvdOpData(d, p.k, p.oversamples, rank=rank, world=world, nsnps_total=m * world,
                                   allreduce=hook)
        op.setFlags(False, True)
        return op

    def barrier():
        if world > 1:
            dist.barrier()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg
    op = make_op(packed, ooc=False)
    stream = torch.cuda.ExternalStream(op.L.pcaone_stream(op.h))

    def step(o, i):
        o._chk(o.L.pcaone_compute_gandh(o.h, i))
        o._chk(o.L.pcaone_small_stage(o.h))

    # NVML is initialised and the sampler thread started BEFORE the warm-up (nvmlInit costs ~30 ms
    # of driver lock, which must not land in the timed region); its samples are cleared below.
    sampler = ClockSampler(local)
    sampler.start()
    # The library's per-scope CUDA events (the kernel breakdown behind `roofline`) cost ~7 % of this
    # launch-bound region when they are on (48.6 vs 52.2 ms for the 7 epochs, tools/step_times.py), so
    # the timed region runs WITHOUT them and an instrumented replica of the same K steps follows it.
    op.enable_timing(False)
    for i in range(args.warmup):
        step(op, i)
    # one more untimed replica of the timed loop so that every (pi -> schedule) code path, cuda
    # function attribute and workspace of the timed steps has been touched once
    for i in range(args.steps):
        step(op, i)
    op.sync()
    op.timers(reset=True)
    sampler.samples.clear()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(op, i)
    e1.record(stream)
    op.sync()
    torch.cuda.synchronize()
    barrier()
    dev_ms = maxr(e0.elapsed_time(e1))
    clocks = sampler.stop()
    gpu_launches = int(op.timers(reset=True).kernel_launches)
    # instrumented replica of the timed steps: per-kernel CUDA-event times for the roofline object
    op.enable_timing(True)
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record(stream)
    for i in range(args.steps):
        step(op, i)
    i1.record(stream)
    op.sync()
    torch.cuda.synchronize()
    instr_ms = i0.elapsed_time(i1)
    tm = op.timers(reset=True)
    op.enable_timing(False)
    value = world * shard_bytes * args.steps / (dev_ms * 1e-3) / 1e9
    l = op.size()
    flops_per_gemm_total = 2.0 * n * m * l  # per pass, each of the two GEMMs
    g_ms, h_ms = tm.gemm_g_ms / args.steps, tm.gemm_h_ms / args.steps
    peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
    if prec == 0:
        # dominant kernel: the fused decode->DMMA GEMM pair; report the slower of the two
        dom, dom_ms, dom_launches = (("k_gemm_h", h_ms, tm.gemm_h_launches) if h_ms >= g_ms else
                                     ("k_gemm_g", g_ms, tm.gemm_g_launches))
        note = ("FP64 DMMA path: algorithmic flops 2*N*M*l per GEMM per pass; peak is the measured bf16 "
                "tensor figure because MEASURED_PEAKS has no FP64 entry (B200 FP64 tensor nominal ~37 TF)")
    else:
        # dominant kernel: k_tc_gemm (tcgen05 kind::i8, both passes use the same kernel); its
        # own CUDA-event time (tc_g_ms / tc_h_ms), without the slice / finish kernels around it
        tg, th = tm.tc_g_ms / args.steps, tm.tc_h_ms / args.steps
        dom, dom_ms, dom_launches = (("k_tc_gemm (H pass)", th, tm.gemm_h_launches) if th >= tg else
                                     ("k_tc_gemm (G pass)", tg, tm.gemm_g_launches))
        note = (f"int8 Ozaki path, {prec} slices: algorithmic flops 2*N*M*l per GEMM per pass (the tensor cores "
                f"execute {prec}x that as exact int8 MACs, so frac <= {1.0 / prec * 2:.2f} of the bf16 peak at the "
                "int8 rate of 2x bf16); peak = measured bf16 sustained")
    achieved_tf = flops_per_gemm_total / (dom_ms * 1e-3) / 1e12
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (per pass =
    # the two full-size launches of a late epoch), to hold against the algorithmic packed bytes
    traffic, traffic_note = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_tc_gemm_traffic.json")))
        leg = tj["h_pass" if "H pass" in dom else "g_pass"]
        if prec == 3:
            traffic = tj["launches_per_pass_full_size"] * (leg["dram_read_bytes"] + leg["dram_write_bytes"])
            traffic_note = (f"bytes per pass (2 full-size launches), ncu dram__bytes_read.sum + dram__bytes_write.sum, "
                            f"{tj['source']}; algorithmic: {shard_bytes} packed bytes read + {n if 'H pass' in dom else m}"
                            f" x {l} int64 accumulators written")
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": f"{peak_src} bf16 sustained",
                "note": note,
                "gemm_g_ms_per_pass": g_ms, "gemm_h_ms_per_pass": h_ms, "orth_ms_per_pass": tm.orth_ms / args.steps,
                "small_stage_ms_per_pass": tm.small_ms / args.steps,
                "tc_g_ms_per_pass": tm.tc_g_ms / args.steps, "tc_h_ms_per_pass": tm.tc_h_ms / args.steps,
                "tc_ranges": int(tm.tc_ranges), "fp64_ranges": int(tm.fp64_ranges),
                "launches_per_pass": dom_launches / args.steps,
                "hbm_algorithmic_gbs": shard_bytes / (dom_ms * 1e-3) / 1e9}
    roofline["timing_note"] = (f"kernel times: CUDA events of an instrumented replica of the {args.steps} timed steps "
                               f"({instr_ms:.2f} ms with the per-scope events on, {dev_ms:.2f} ms timed without them)")
    # the same kernel at its full-size launches only: one more epoch with pi >= log2(bands) (one Omega
    # update per pass, every window merged into two half-shard launches), outside the timed region
    if prec != 0 and args.steps >= 1:
        op.enable_timing(True)
        pi_late = max(args.steps, 6)
        step(op, pi_late)
        op.sync()
        op.timers(reset=True)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record(stream)
        step(op, pi_late + 1)
        l1.record(stream)
        op.sync()
        torch.cuda.synchronize()
        tl = op.timers(reset=True)
        op.enable_timing(False)
        late_ms = l0.elapsed_time(l1)
        tcl = max(tl.tc_g_ms, tl.tc_h_ms)
        roofline["late_pass"] = {
            "what": "one epoch with pi >= 6 (plain power iteration, 2 half-shard launches per GEMM), untimed leg",
            "ms": late_ms, "gbs": shard_bytes / (late_ms * 1e-3) / 1e9, "tc_g_ms": tl.tc_g_ms, "tc_h_ms": tl.tc_h_ms,
            "gemm_g_ms": tl.gemm_g_ms, "gemm_h_ms": tl.gemm_h_ms, "orth_ms": tl.orth_ms, "small_stage_ms": tl.small_ms,
            "achieved": flops_per_gemm_total / (tcl * 1e-3) / 1e12 if tcl > 0 else None,
            "frac": flops_per_gemm_total / (tcl * 1e-3) / 1e12 / peak_tf if tcl > 0 else None}
    op.close()

    # ---- end-to-end leg: packed matrix in pinned host memory, streamed every pass
    e2e = None
    if not args.no_e2e:
        host = torch.empty((m, bpr), dtype=torch.uint8, pin_memory=True)
        host.copy_(packed)
        torch.cuda.synchronize()
        del packed
        torch.cuda.empty_cache()
        op2 = make_op(host, ooc=True)
        def step2(i):
            step(op2, i)
            # result of the step back on the host: sigma (l) and the current PCs are device state;
            # read the N x l H (the pass output the reference hands to computeUSV)
            op2.getH(Hh)

        Hh = np.zeros((n, l), order="F")
        for i in range(min(args.warmup, 3)):
            step2(i)
        op2.sync()
        op2.timers(reset=True)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step2(i)
        op2.sync()
        barrier()
        e2e_s = maxr(time.perf_counter() - t0)
        tm2 = op2.timers(reset=True)
        e2e = {"value": world * shard_bytes * args.steps / e2e_s / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(tm2.h2d_bytes // args.steps), "d2h_bytes_per_step": int(tm2.d2h_bytes // args.steps),
               "ms_per_step": 1e3 * e2e_s / args.steps, "timing": "host wall clock between stream syncs, max over ranks"}
        op2.close()
        cpu_src = host.numpy()[: (CPU_SAMPLE_SNPS if not args.small else 4096)]
    else:
        cpu_src = packed[: (CPU_SAMPLE_SNPS if not args.small else 4096)].cpu().numpy()

    # ---- CPU baseline (rank 0, N=1 only): the compiled reference on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        res = cpu_reference_pass(np.ascontiguousarray(cpu_src), n, K, 2, 1, threads)
        if res is not None:
            tot = sum(res["times"])
            cpu = {"value": res["bytes"] * len(res["times"]) / tot / 1e9, "unit": UNIT, "cores": threads,
                   "kind": "reference",
                   "sample": (f"first {cpu_src.shape[0]} of {m} SNPs x {n} samples, winSVD in-core -S, epochs 1-2 "
                              f"({tot:.1f} s of CPU passes + {res['load_s']:.1f} s read/decode); Eigen GEMM, no MKL")}
        else:
            cpu = {"value": None, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": "oracle/_ref/libpcaone_ref.so missing"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64" if prec == 0 else f"int8x{prec} (exact, FP64 epilogue)",
                "data": "synthetic",
                "config": {"workload": f"configs[1]: winSVD in-memory, N={n} x M={m} SNPs per GPU, k={K}, l={l}, "
                                       f"{BANDS} windows, no-shuffle; step = one computeUSV epoch (pi = step index)",
                           "l2": "inputs (2.5 GB packed per GPU) are larger than L2; no flush needed",
                           "warmup_note": f"{args.warmup} warm-up steps + one untimed replica of the {args.steps} timed steps",
                           "parallelism": f"snp-shard x{world}"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": gpu_launches,
                "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
