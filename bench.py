#!/usr/bin/env python
"""bench.py — CDAN forward throughput on B200 (the metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c3|c2] [--scaling weak|strong]

A "step" is one eval-mode CDAN forward over one synthetic batch.
  --config c3 (default): BASELINE config 3, the configuration the metric is quoted on: 32 x 3 x 1080 x 1920 images.
      --scaling weak (default): 32 images PER GPU at every N (batch-sharded, no collective).
      --scaling strong: SURVEY 8(e)'s split of C3: 32 images in total, 32/16/8/4 per GPU.
  --config c2: BASELINE config 2, 64 x 3 x 256 x 256 on one GPU (per GPU when N > 1).
`value` = megapixels/s with inputs resident in HBM, device-timed (CUDA events, max over ranks).
`e2e` = the same through the host-buffer C-ABI call cdan_forward_host (pinned fp32 NCHW host input -> H2D -> forward ->
D2H fp32 NCHW, the reference's tensor contract at the module boundary); `e2e_u8` = the reference's image data path
(uint8 HWC -> /255 -> forward -> x255 -> uint8 HWC) through cdan_forward_host_u8, a quarter of the PCIe bytes.
`roofline` is for the tcgen05 convolutions (all conv launches of one step) against the measured sustained bf16 peak;
`roofline_dense` for the dense-block kernels against the measured HBM bandwidth; `roofline_cbam` for the CBAM group.
`cpu_baseline` (N = 1 only, computed before any process group exists) and `--impl reference` time the reference's own
CPU forward (oracle/_ref = the unmodified /root/reference files vendored at build time; the oracle port only when that
directory is missing) on this box's host cores on a bounded sample.
Inputs and activations are far larger than the 126 MB L2 at C3; at C2 a 256 MB scratch buffer is written between
timed forwards to flush L2 (stated in `config.l2`).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "multi-degradation-image-enhancement_b200")
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "megapixels/sec CDAN fwd (1080p bf16)"
UNIT = "MP/s"
FLOP_PER_PIXEL = 252770  # algorithmic conv FLOPs (2*MAC, unpadded), SURVEY 8(d) / BASELINE.md 3
# Dense-block 3x3 layers: logical input channels and resolution divisor; algorithmic HBM bytes per layer
# and output pixel in bf16 = 2 * (Cin + 16): read the concat prefix once, write the 16 new channels once.
DENSE3X3 = {f"{blk}.layers.{l}": (c0 + 16 * l, div)
            for blk, c0, div in (("encoder.dense1", 64, 2), ("encoder.dense2", 128, 4), ("encoder.dense3", 256, 8),
                                 ("decoder.final_dense", 3, 1)) for l in range(4)}
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per 32 x 1080p step from the `ncu --set full` capture of ALL
# launches of a step summarised in profiles/r02_step_ncu.md: the 12 dense-block 3x3 launches of dense blocks 1-3 (the final
# dense block is one fused kernel), all 24 convolution launches, and the CBAM kernels.
DENSE3X3_NCU_TRAFFIC_BYTES = 22317280000
ALLCONV_NCU_TRAFFIC_BYTES = 50553688000
CBAM_NCU_TRAFFIC_BYTES = 19545652000


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"tensor": float(d["bf16_tflops_sustained"]), "hbm": float(d["hbm_gbs"]), "source": "measured"}
    return {"tensor": 1400.0, "hbm": 6650.0, "source": "fallback"}  # B200_PROFILING.md fallback figures


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None, t_load0=None):
        """Summarise the samples received inside [t0, t1] (perf_counter; the timed region).  nvidia-smi needs ~0.2 s to
        start, so it is launched before the warm-up; if the timed region is shorter than one sampling period the samples
        taken under the same load during warm-up are used and the window says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for ts, r in self.rows if t0 is None or (t0 <= ts <= t1 + 0.05)]
        window = "timed region"
        if not rows:
            rows = [r for ts, r in self.rows if t_load0 is None or (t_load0 <= ts <= t1 + 0.05)]
            window = "warm-up + timed region (same load)"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def synthetic_batch(n, h, w, seed=42):
    g = torch.Generator().manual_seed(seed)
    return torch.rand((n, 3, h, w), generator=g)


def default_weights(seed=42):
    """`torch.manual_seed(42); CDAN()` — PyTorch's default initialisation of the module tree."""
    from models.cdan import CDAN
    torch.manual_seed(seed)
    net = CDAN()
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def reference_weights(seed=42):
    """Same seeded default init drawn through the REFERENCE's own module (oracle/_ref) when present, so that the reference
    arm imports none of this repository's model code; bit-identical to default_weights (same registration order)."""
    from oracle.build_ref import load_ref, ref_available
    if not ref_available():
        return default_weights(seed)
    cdan_mod, _, _ = load_ref()
    torch.manual_seed(seed)
    net = cdan_mod.CDAN()
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def cpu_reference_rate(sd, h, w, budget_s, threads):
    """MP/s of the reference's CPU forward on a bounded sample: one image, as many rows (multiple of 8) as fit the
    budget.  Returns (MP/s, sample description, seconds, kind) with kind = "reference" (oracle/_ref) or "port"."""
    from oracle.build_ref import reference_forward_fn
    torch.set_num_threads(threads)
    fwd, kind = reference_forward_fn(sd)
    probe = synthetic_batch(1, 64, min(w, 256), seed=1)
    fwd(probe)  # warm-up (thread pools, allocator)
    t0 = time.perf_counter()
    fwd(probe)
    rate = probe.shape[2] * probe.shape[3] / (time.perf_counter() - t0)  # px/s, pessimistic for larger images
    rows = int(max(8, min(h, (budget_s * rate / w) // 8 * 8)))
    x = synthetic_batch(1, rows, w, seed=42)
    t0 = time.perf_counter()
    fwd(x)
    dt = time.perf_counter() - t0
    what = "unmodified reference models/cdan.py (oracle/_ref)" if kind == "reference" else "torch CPU oracle port"
    return rows * w / 1e6 / dt, f"1 x 3 x {rows} x {w} fp32 image (rows sized to ~{budget_s:.0f} s), {what}", dt, kind


def run_reference(args, rank, world, emit):
    """Reference arm: the reference's own CPU forward on host cores, bounded sample per step; rank 0 only."""
    if rank != 0:
        return
    from oracle.build_ref import reference_forward_fn
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = reference_weights(42)
    total = args.steps + args.warmup
    budget = max(1.0, 150.0 / max(1, total))
    mp_s, sample, _, kind = cpu_reference_rate(sd, args.height, args.width, budget, threads)
    rows = int(sample.split(" x ")[2])
    fwd, kind = reference_forward_fn(sd)
    x = synthetic_batch(1, rows, args.width, seed=42)
    for _ in range(args.warmup):
        fwd(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fwd(x)
    dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = rows * args.width / 1e6 / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, args.batch), "note": "reference arm times a bounded sample per step",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_name(args, n):
    if args.config == "c5":
        return f"CDAN forward {n}x3x{args.height}x{args.width}, ONE image batch split into row bands over the GPUs (BASELINE C5)"
    if args.config == "c4":
        return (f"multi-degradation routing: MultiHeadClassifier (ResNet18, torch) -> thresholds -> 5 per-degradation CDAN weight sets "
                f"on a mixed synthetic batch {n}x3x{args.height}x{args.width} per GPU (BASELINE C4)")
    tag = {"c3": "BASELINE C3", "c2": "BASELINE C2"}[args.config]
    return f"CDAN forward {n}x3x{args.height}x{args.width} per GPU ({tag}), batch-sharded, no collective"


def run_c4(args, rank, world, dev, dist, numa, emit):
    """BASELINE config 4: a mixed synthetic batch per GPU (reference degradation functions, 0-3 degradations per image) is
    classified by the multi-label classifier (ResNet18 on torch), thresholded per class and routed through the CDAN weight
    set of every flagged degradation (routing.MultiDegradationPipeline).  Weights are seeded random (no checkpoints
    offline), so the classifier's thresholds are set to the per-class median probability of the batch: every enhancer
    receives about half of the images, ~2.5 CDAN passes per image.  A step = classify + route + enhance one batch."""
    import cdan_b200_native as native
    from classification.multilabel_classifier import DEGRADATIONS, MultiHeadClassifier, predict_probs
    from models.cdan import CDAN
    from routing import ENHANCER_CLASSES, MultiDegradationPipeline, synthetic_mixed_batch
    n, h, w = args.batch, args.height, args.width
    torch.manual_seed(7)
    clf = MultiHeadClassifier().to(dev).eval()
    enhancers = {}
    for k, name in enumerate(ENHANCER_CLASSES):
        net = CDAN().set_compute_dtype(args.dtype)
        net.load_state_dict(default_weights(100 + k))
        enhancers[name] = net.to(dev).eval()
    xu_host, _ = synthetic_mixed_batch(n, h, w, seed=1000 + rank)
    xu_host = xu_host.pin_memory()
    yu_host = torch.empty_like(xu_host).pin_memory()
    x = (xu_host.to(dev).permute(0, 3, 1, 2).float() / 255).contiguous()
    with torch.no_grad():
        probs, _ = predict_probs(clf, x)
    th = [float(probs[:, i].median()) for i in range(len(DEGRADATIONS))]
    pipe = MultiDegradationPipeline(clf, enhancers, thresholds=th)
    sampler = ClockSampler(dev.index)
    sampler.start()

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)
    t_warm0 = time.perf_counter()
    for _ in range(args.warmup):
        y = pipe(x)
    barrier()
    launches = sum(e.native_plan(dev).last_launch_count for e in enhancers.values())
    t_region0 = time.perf_counter()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        flush.fill_(1)
        a.record()
        y = pipe(x)
        b.record()
    barrier()
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.stop(t_region0, time.perf_counter(), t_warm0)
    passes = sum(pipe.router.last_bucket_sizes.values())

    def e2e_once():  # uint8 host image batch -> H2D -> /255 -> classify + route + enhance -> x255 uint8 -> D2H
        xd = xu_host.to(dev, non_blocking=True).permute(0, 3, 1, 2).float().mul_(1.0 / 255).contiguous()
        yu_host.copy_(native.quantize_u8(pipe(xd)), non_blocking=True)
        torch.cuda.synchronize(dev)

    e2e_once()
    barrier()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_once()
    e2e_s = (time.perf_counter() - t0) / steps
    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, e2e_s = float(t[0]) / args.steps, float(t[1])
    if rank == 0:
        mp = world * n * h * w / 1e6
        peaks = measured_peaks()
        flops = FLOP_PER_PIXEL * passes * h * w  # CDAN passes only (the classifier adds ~1.8 GFLOP per image on cuDNN)
        emit({
            "metric": "images/sec multi-degradation routing (classifier + per-degradation CDAN, bf16)", "value": world * n / (ms_step * 1e-3),
            "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "megapixels_per_s": mp / (ms_step * 1e-3),
            "config": {"workload": workload_name(args, n), "name": "c4", "batch_per_gpu": n, "height": h, "width": w,
                       "classes": DEGRADATIONS, "routed_classes": ENHANCER_CLASSES, "bucket_sizes_rank0": pipe.router.last_bucket_sizes,
                       "cdan_passes_per_image": passes / n, "weights": "seeded random (classifier and 5 CDAN sets); thresholds = per-class median",
                       "l2": "256 MB buffer written between timed steps (L2 flush); per-step CUDA events summed",
                       "parallelism": f"dp{world}", "host_numa": numa},
            "clocks": clocks,
            "e2e": {"value": world * n / e2e_s, "unit": "img/s", "h2d_bytes_per_step": n * h * w * 3, "d2h_bytes_per_step": n * h * w * 3,
                    "ms_per_step": e2e_s * 1e3, "api": "uint8 host batch -> H2D -> MultiDegradationPipeline -> cdan_quantize_u8 -> D2H"},
            "gpu_launches": launches * args.steps,
            "roofline": {"bound": "tensor", "kernel": "CDAN convolution FLOPs of the routed passes over the whole step (classifier and routing included in the time)",
                         "achieved": flops / (ms_step * 1e-3) / 1e12, "peak": peaks["tensor"], "unit": "TFLOP/s",
                         "frac": flops / (ms_step * 1e-3) / 1e12 / peaks["tensor"], "traffic": None,
                         "peak_source": peaks["source"] + " (sustained)"},
        })


def run_c5(args, rank, world, dev, dist, cpu_baseline, numa, emit):
    """BASELINE config 5: ONE large image (default 1 x 3 x 2160 x 3840) split into row bands over the GPUs
    (spatial_tiling.NcclBandedCDAN: recompute halos, halo refreshes by ncclSend/ncclRecv, ChannelGate all-reduce).
    At N = 1 the same image runs through the untiled forward.  Strong scaling: value = image megapixels / step time."""
    import cdan_b200_native as native
    import spatial_tiling as st
    n, h, w = args.batch, args.height, args.width
    sd = default_weights(42)
    sampler = ClockSampler(dev.index)
    sampler.start()
    x_host = synthetic_batch(n, h, w, seed=42).pin_memory()  # the same image on every rank
    if world > 1:
        runner = st.NcclBandedCDAN(sd, args.dtype, dev, halo=args.halo)
        plan = runner.plan
        r0, r1, e0, e1 = runner.rows(h)
    else:
        plan = native.Plan(dev, args.dtype)
        plan.load_state_dict(sd)
        r0, r1, e0, e1 = 0, h, 0, h
    x_ext_host = x_host[:, :, e0:e1].contiguous().pin_memory()
    y_own_host = torch.empty((n, 3, r1 - r0, w), dtype=torch.float32).pin_memory()
    x_ext = x_ext_host.to(dev)
    y_ext = torch.empty_like(x_ext)

    def fwd():
        if world > 1:
            plan.forward_band(x_ext, h, args.halo, out=y_ext)
        else:
            plan.forward(x_ext, out=y_ext)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev)
    t_warm0 = time.perf_counter()
    for _ in range(args.warmup):
        fwd()
    barrier()
    # device-timed: per-forward CUDA events (an L2 flush between forwards, untimed), summed; max over ranks
    t_region0 = time.perf_counter()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for e_a, e_b in evs:
        flush.fill_(1)
        if dist is not None:
            dist.barrier()  # bands start a forward together (a collective schedule)
        e_a.record()
        fwd()
        e_b.record()
    barrier()
    ms_total = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.stop(t_region0, time.perf_counter(), t_warm0)
    launches = plan.last_launch_count
    stats = plan.band_stats() if world > 1 else {"halo_exchanges": 0, "halo_bytes_received": 0, "allreduces": 0}
    # where the time goes (rank 0, outside the timed region): per-group CUDA-event spans of a few more forwards
    plan.set_option("profile", 1)
    plan.profile_read()
    for _ in range(3):
        if dist is not None:
            dist.barrier()
        fwd()
    torch.cuda.synchronize(dev)
    spans = plan.profile_read()
    plan.set_option("profile", 0)
    breakdown = {}
    for k, (ms, _) in spans.items():
        g = k.split("|")[0]
        g = {"conv": "convolutions", "cbam": "cbam (+ pooled-statistics all-reduce)", "glue": "upsample+add", "band": "halo refreshes"}.get(g, g)
        breakdown[g] = breakdown.get(g, 0.0) + ms / 3
    breakdown = {k: round(v, 4) for k, v in breakdown.items()}

    # end to end: pinned host rows of the band (halo included) -> H2D -> banded forward -> D2H of the OWNED rows
    def e2e_once():
        x_ext.copy_(x_ext_host, non_blocking=True)
        fwd()
        y_own_host.copy_(y_ext[:, :, r0 - e0:r1 - e0], non_blocking=True)
        torch.cuda.synchronize(dev)

    e2e_once()
    barrier()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        if dist is not None:
            dist.barrier()
        e2e_once()
    e2e_s = (time.perf_counter() - t0) / steps
    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, e2e_s = float(t[0]) / args.steps, float(t[1])
    mp = n * h * w / 1e6
    if rank == 0:
        peaks = measured_peaks()
        flops = FLOP_PER_PIXEL * n * h * w
        line = {
            "metric": "megapixels/sec CDAN fwd (one 4K image in row bands, bf16)" if args.dtype == "bf16" else "megapixels/sec CDAN fwd (row bands)",
            "value": mp / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic", "images_per_s": n / (ms_step * 1e-3),
            "config": {"workload": workload_name(args, n), "name": "c5", "batch": n, "height": h, "width": w,
                       "bands": world, "halo_rows": args.halo if world > 1 else 0,
                       "rows_rank0": {"owned": [r0, r1], "extended": [e0, e1]},
                       "halo_refreshes_per_forward": stats["halo_exchanges"], "allreduces_per_forward": stats["allreduces"],
                       "halo_bytes_received_rank0": stats["halo_bytes_received"],
                       "kernel_ms_rank0": breakdown, "launches_per_forward": launches,
                       "weights": "torch.manual_seed(42); CDAN() default init (random)",
                       "l2": "256 MB buffer written between timed forwards (L2 flush); per-forward CUDA events summed",
                       "parallelism": f"rows{world}" if world > 1 else "untiled", "host_numa": numa},
            "clocks": clocks,
            "e2e": {"value": mp / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(x_ext_host.numel() * 4),
                    "d2h_bytes_per_step": int(y_own_host.numel() * 4), "ms_per_step": e2e_s * 1e3,
                    "api": "pinned host rows of the band -> H2D -> cdan_forward_band (NCCL halo refreshes) -> D2H of the owned rows"},
            "gpu_launches": launches * args.steps,
            "roofline": {"bound": "tensor", "kernel": "whole banded forward (latency-bound: launches + halo refreshes; see DESIGN.md 6)",
                         "achieved": flops / (ms_step * 1e-3) / 1e12, "peak": peaks["tensor"] * world, "unit": "TFLOP/s",
                         "frac": flops / (ms_step * 1e-3) / 1e12 / (peaks["tensor"] * world), "traffic": None,
                         "peak_source": peaks["source"] + " (sustained) x n_gpus"},
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c2", "c4", "c5"])
    ap.add_argument("--halo", type=int, default=24, help="c5: recompute halo rows per interior band side")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU (default: the config's batch)")
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", action="store_true", help="print the per-launch timing table to stderr")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    cfg_batch, cfg_h, cfg_w = {"c3": (32, 1080, 1920), "c2": (64, 256, 256), "c4": (64, 256, 384), "c5": (1, 2160, 3840)}[args.config]
    args.height = args.height or cfg_h
    args.width = args.width or cfg_w
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.config == "c5":
        args.scaling = "strong"  # one image, split by rows: total work is fixed
    if args.batch is None:
        args.batch = cfg_batch if args.config == "c5" else (max(1, cfg_batch // world_env) if args.scaling == "strong" else cfg_batch)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # Rank 0 prints ONE JSON line on stdout.  Libraries write banners there from native code (e.g. "NCCL version ..." at
    # process-group creation), so file descriptor 1 points at stderr until the line is printed.
    sys.stdout.flush()
    _stdout_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(_stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA kernels are the product (no CPU fallback)")
    n, h, w = args.batch, args.height, args.width

    # ---- CPU baseline: rank 0, N = 1 only, BEFORE any process group / GPU work exists (at N > 1 the other ranks would
    #      spin in a barrier on the same cores and the figure is meaningless; the N = 1 line carries it)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        mp_s, sample, _, kind = cpu_reference_rate(default_weights(42), h, w, args.cpu_budget, threads)
        cpu_baseline = {"value": mp_s, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    from host_affinity import bind_to_gpu_node
    numa = bind_to_gpu_node(local_rank) if world > 1 else {"bound": False, "note": "single rank: not bound"}
    torch.set_num_threads(max(1, min(8, len(os.sched_getaffinity(0)))))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    if args.config == "c4":
        run_c4(args, rank, world, dev, dist, numa, emit)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    if args.config == "c5":
        run_c5(args, rank, world, dev, dist, cpu_baseline, numa, emit)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    from models.cdan import CDAN

    net = CDAN().set_compute_dtype(args.dtype)
    net.load_state_dict(default_weights(42))
    net = net.to(dev).eval()
    sampler = ClockSampler(local_rank)  # started early: nvidia-smi needs a few hundred ms before its first sample
    sampler.start()
    x_host = synthetic_batch(n, h, w, seed=42 + rank).pin_memory()
    y_host = torch.empty_like(x_host).pin_memory()
    xu_host = (x_host * 255.0).to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
    yu_host = torch.empty_like(xu_host).pin_memory()
    x = x_host.to(dev)
    y = torch.empty_like(x)
    plan = net.native_plan(dev)
    # L2 flush between timed forwards when the working set could stay resident (C2: 50 MB of input): write 256 MB
    small = n * h * w * 3 * 4 < 512 * 2 ** 20
    flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device=dev) if small else None

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- warm-up (also sizes the workspace); the clock sampler is already streaming
    t_warm0 = time.perf_counter()
    for _ in range(args.warmup):
        plan.forward(x, out=y)
    barrier()

    # ---- timed region: K forwards, inputs resident in HBM; per-launch CUDA-event spans recorded alongside
    plan.set_option("profile", 1)
    plan.profile_read()
    barrier()
    t_region0 = time.perf_counter()
    if flush is None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            plan.forward(x, out=y)
        ev1.record()
        barrier()
        ms_total = ev0.elapsed_time(ev1)
    else:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for e0, e1 in evs:
            flush.fill_(1)  # untimed: evict the previous forward's tensors from L2
            e0.record()
            plan.forward(x, out=y)
            e1.record()
        barrier()
        ms_total = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    clocks = sampler.stop(t_region0, time.perf_counter(), t_warm0)
    spans = plan.profile_read()
    plan.set_option("profile", 0)
    launches_per_step = plan.last_launch_count

    # ---- end-to-end through the host-buffer entry points (H2D + forward + D2H inside the timed region)
    def time_host(fn, xin, yout):
        steps = max(1, min(args.steps, 5))
        fn(xin, yout)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn(xin, yout)
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / steps

    e2e_s = time_host(plan.forward_host, x_host, y_host)
    e2e_u8_s = time_host(plan.forward_host_u8, xu_host, yu_host)

    # what the host side can deliver at best: the fp32 batch H2D and D2H at the same time, no compute (all ranks at once)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        with torch.cuda.stream(s_in):
            x.copy_(x_host, non_blocking=True)
        with torch.cuda.stream(s_out):
            y_host.copy_(y, non_blocking=True)
    torch.cuda.synchronize(dev)
    copy_floor_s = (time.perf_counter() - t0) / 2

    t = torch.tensor([ms_total, e2e_s, e2e_u8_s, copy_floor_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t[0]) / args.steps
    e2e_s, e2e_u8_s, copy_floor_s = float(t[1]), float(t[2]), float(t[3])
    mp_per_step_total = world * n * h * w / 1e6

    if rank == 0:
        peaks = measured_peaks()
        conv_ms = sum(v[0] for k, v in spans.items() if k.startswith("conv|")) / args.steps
        conv_launches = sum(v[1] for k, v in spans.items() if k.startswith("conv|")) // args.steps
        cbam_ms = sum(v[0] for k, v in spans.items() if k.startswith("cbam|")) / args.steps
        glue_ms = sum(v[0] for k, v in spans.items() if k.startswith("glue|")) / args.steps
        flops_step = FLOP_PER_PIXEL * n * h * w
        achieved = flops_step / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        cbam_bytes = 144.0 * n * h * w  # compulsory CBAM traffic, 1R+1W of 36 elem/px bf16
        if args.layers:
            for k, v in sorted(spans.items(), key=lambda kv: -kv[1][0]):
                print(f"  {k:60s} {v[0] / args.steps:9.3f} ms/step  x{v[1] // args.steps}", file=sys.stderr)
        c3_full = (n, h, w) == (32, 1080, 1920)
        line = {
            "metric": METRIC, "value": mp_per_step_total / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "images_per_s": world * n / (ms_step * 1e-3),
            "config": {"workload": workload_name(args, n), "name": args.config,
                       "batch_per_gpu": n, "global_batch": n * world, "height": h, "width": w,
                       "weights": "torch.manual_seed(42); CDAN() default init (random)",
                       "l2": ("256 MB buffer written between timed forwards (L2 flush); per-forward CUDA events summed"
                              if flush is not None else "inputs+activations >> 126 MB L2 (no flush needed)"),
                       "parallelism": f"dp{world}", "host_numa": numa},
            "clocks": clocks,
            "e2e": {"value": mp_per_step_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * 3 * h * w * 4,
                    "d2h_bytes_per_step": n * 3 * h * w * 4, "ms_per_step": e2e_s * 1e3,
                    "api": "Plan.forward_host -> cdan_forward_host (pinned fp32 NCHW host buffers)",
                    "host_copy_floor_ms": copy_floor_s * 1e3,
                    "host_copy_floor_note": "the same H2D + D2H bytes moved concurrently with NO compute, all ranks at once, max over ranks: "
                                            f"{n * 3 * h * w * 4 / copy_floor_s / 1e9:.1f} GB/s per direction and GPU"},
            "e2e_u8": {"value": mp_per_step_total / e2e_u8_s, "unit": UNIT, "h2d_bytes_per_step": n * 3 * h * w,
                       "d2h_bytes_per_step": n * 3 * h * w, "ms_per_step": e2e_u8_s * 1e3,
                       "api": "Plan.forward_host_u8 -> cdan_forward_host_u8 (pinned uint8 NHWC host buffers; /255 and "
                              "x255 quantisation on the device, the reference's image data path)"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "tensor", "kernel": "tcgen05 convolutions: dense_fused_kernel + conv_stream2_kernel + conv_stream_kernel + conv_umma_kernel (all conv launches of a step)",
                         "achieved": achieved, "peak": peaks["tensor"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tensor"],
                         "traffic": ALLCONV_NCU_TRAFFIC_BYTES if c3_full else None,
                         "peak_source": peaks["source"] + " (sustained)",
                         "launches_per_step": conv_launches, "kernel_ms_per_step": conv_ms,
                         "kernel_share_of_step": conv_ms / ms_step if ms_step else None,
                         "whole_step_frac": flops_step / (ms_step * 1e-3) / 1e12 / peaks["tensor"]},
            "roofline_cbam": {"bound": "hbm", "achieved": cbam_bytes / (cbam_ms * 1e-3) / 1e9 if cbam_ms else 0.0,
                              "peak": peaks["hbm"], "unit": "GB/s",
                              "frac": (cbam_bytes / (cbam_ms * 1e-3) / 1e9 / peaks["hbm"]) if cbam_ms else 0.0,
                              "traffic": CBAM_NCU_TRAFFIC_BYTES if c3_full else None,
                              "kernel_ms_per_step": cbam_ms, "note": "achieved = compulsory 144 B/px over the 4 CBAM sites"},
            "glue_ms_per_step": glue_ms,
        }
        # only the layers that ran as launches of their own (the fused final dense block has no per-layer spans)
        dense_run = {k.split("|")[1] for k in spans if k.split("|")[1] in DENSE3X3}
        dense_ms = sum(v[0] for k, v in spans.items() if k.split("|")[1] in dense_run) / args.steps
        dense_bytes = sum(2.0 * (cin + 16) * n * (h // div) * (w // div) for name, (cin, div) in DENSE3X3.items() if name in dense_run)
        if dense_ms > 0:
            gbs = dense_bytes / (dense_ms * 1e-3) / 1e9
            line["roofline_dense"] = {
                "bound": "hbm", "kernel": "conv_stream2_kernel<4,0,GP> (dense-block 3x3 launches of a step that are not fused)",
                "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                "traffic": DENSE3X3_NCU_TRAFFIC_BYTES if c3_full and len(dense_run) == 12 else None,
                "algorithmic_bytes_per_step": dense_bytes, "kernel_ms_per_step": dense_ms,
                "kernel_share_of_step": dense_ms / ms_step if ms_step else None}
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
