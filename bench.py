#!/usr/bin/env python
"""bench.py — CDAN forward throughput on B200 (the metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--height H] [--width W]

A "step" is one eval-mode CDAN forward over one synthetic batch.  Workload at every N: 32 x 3 x 1080 x 1920 images
PER GPU (BASELINE config C3's batch, the configuration the metric is quoted on; batch-sharded, no collective ->
weak scaling).  `value` = megapixels/s with inputs resident in HBM, device-timed (CUDA events, max over ranks);
`e2e` = the same through the host-buffer C-ABI call (pinned host input -> H2D -> forward -> D2H) per step.
`roofline` is for the tcgen05 convolutions (conv_stream_kernel + conv_umma_kernel, all conv launches of one step) against
the measured bf16 peak; `roofline_dense` is for the single heaviest kernel, conv_stream2_kernel<4,0,GP> (the 16 dense-block
3x3 layers, ~33 % of the step), against the measured HBM bandwidth; `cpu_baseline` is the CPU oracle port timed on this
box's host cores on a bounded sample.
Inputs (796 MB per step) and activations are far larger than the 126 MB L2, so no explicit L2 flush is needed.
`--impl reference` times the reference algorithm's CPU restatement (oracle/, the reference itself is a Python tree that
does not travel to the GPU box) on all host threads.
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
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per 32 x 1080p step from the `ncu --set full` capture
# summarised in profiles/r01_allconv_ncu.md: the 16 dense-block 3x3 launches, and all 29 convolution launches.
DENSE3X3_NCU_TRAFFIC_BYTES = 52039405000
ALLCONV_NCU_TRAFFIC_BYTES = 89809026000


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


def cpu_oracle_rate(h, w, budget_s, threads):
    """MP/s of the CPU oracle port on a bounded sample: one image, as many rows (multiple of 8) as fit the budget."""
    from oracle.cdan_oracle import cdan_forward
    from oracle.stress_init import default_state_dict
    torch.set_num_threads(threads)
    sd = default_state_dict(42)
    probe = synthetic_batch(1, 64, min(w, 256), seed=1)
    with torch.no_grad():
        cdan_forward(sd, probe)  # warm-up (thread pools, allocator)
        t0 = time.perf_counter()
        cdan_forward(sd, probe)
        rate = probe.shape[2] * probe.shape[3] / (time.perf_counter() - t0)  # px/s, pessimistic for larger images
    rows = int(max(8, min(h, (budget_s * rate / w) // 8 * 8)))
    x = synthetic_batch(1, rows, w, seed=42)
    with torch.no_grad():
        t0 = time.perf_counter()
        cdan_forward(sd, x)
        dt = time.perf_counter() - t0
    return rows * w / 1e6 / dt, f"1 x 3 x {rows} x {w} fp32 image (rows sized to ~{budget_s:.0f} s), torch CPU oracle port", dt


def run_reference(args, rank, world, emit):
    """Reference arm: the reference's algorithm on host cores (oracle port), bounded sample per step."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle.cdan_oracle import cdan_forward
    from oracle.stress_init import default_state_dict
    sd = default_state_dict(42)
    total = args.steps + args.warmup
    budget = max(1.0, 150.0 / max(1, total))
    mp_s, sample, _ = cpu_oracle_rate(args.height, args.width, budget, threads)
    rows = int(sample.split(" x ")[2])
    x = synthetic_batch(1, rows, args.width, seed=42)
    with torch.no_grad():
        for _ in range(args.warmup):
            cdan_forward(sd, x)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cdan_forward(sd, x)
        dt = (time.perf_counter() - t0) / max(1, args.steps)
    value = rows * args.width / 1e6 / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CDAN forward {args.batch}x3x{args.height}x{args.width} per GPU (C3); reference arm "
                               f"times a bounded sample per step", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", action="store_true", help="print the per-launch timing table to stderr")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

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
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from models.cdan import CDAN
    from oracle.stress_init import default_state_dict  # weights only (seeded init from the key schema)

    n, h, w = args.batch, args.height, args.width
    net = CDAN().set_compute_dtype(args.dtype)
    net.load_state_dict(default_state_dict(42))
    net = net.to(dev).eval()
    sampler = ClockSampler(local_rank)  # started early: nvidia-smi needs a few hundred ms before its first sample
    sampler.start()
    x_host = synthetic_batch(n, h, w, seed=42 + rank).pin_memory()
    y_host = torch.empty_like(x_host).pin_memory()
    x = x_host.to(dev)
    y = torch.empty_like(x)
    plan = net.native_plan(dev)

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
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_region0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        plan.forward(x, out=y)
    ev1.record()
    barrier()
    clocks = sampler.stop(t_region0, time.perf_counter(), t_warm0)
    ms_total = ev0.elapsed_time(ev1)
    spans = plan.profile_read()
    plan.set_option("profile", 0)
    launches_per_step = plan.last_launch_count

    # ---- end-to-end through the host-buffer entry point (H2D + forward + D2H inside the timed region)
    e2e_steps = max(1, min(args.steps, 5))
    plan.forward_host(x_host, y_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.forward_host(x_host, y_host)
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / e2e_steps

    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t[0]) / args.steps
    e2e_s = float(t[1])
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
        line = {
            "metric": METRIC, "value": mp_per_step_total / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "images_per_s": world * n / (ms_step * 1e-3),
            "config": {"workload": f"CDAN forward {n}x3x{h}x{w} per GPU (BASELINE C3 batch), batch-sharded, no collective",
                       "batch_per_gpu": n, "height": h, "width": w, "weights": "seeded default-like init (random)",
                       "l2": "inputs+activations >> 126 MB L2 (no flush needed)", "parallelism": f"dp{world}"},
            "clocks": clocks,
            "e2e": {"value": mp_per_step_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": n * 3 * h * w * 4,
                    "d2h_bytes_per_step": n * 3 * h * w * 4, "ms_per_step": e2e_s * 1e3,
                    "api": "Plan.forward_host -> cdan_forward_host (pinned fp32 NCHW host buffers)"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "tensor", "kernel": "tcgen05 convolutions: conv_stream2_kernel + conv_stream_kernel + conv_umma_kernel (all conv launches of a step)",
                         "achieved": achieved, "peak": peaks["tensor"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tensor"],
                         "traffic": ALLCONV_NCU_TRAFFIC_BYTES if (n, h, w) == (32, 1080, 1920) else None,
                         "peak_source": peaks["source"] + " (sustained)",
                         "launches_per_step": conv_launches, "kernel_ms_per_step": conv_ms,
                         "kernel_share_of_step": conv_ms / ms_step if ms_step else None,
                         "whole_step_frac": flops_step / (ms_step * 1e-3) / 1e12 / peaks["tensor"]},
            "roofline_cbam": {"bound": "hbm", "achieved": cbam_bytes / (cbam_ms * 1e-3) / 1e9 if cbam_ms else 0.0,
                              "peak": peaks["hbm"], "unit": "GB/s",
                              "frac": (cbam_bytes / (cbam_ms * 1e-3) / 1e9 / peaks["hbm"]) if cbam_ms else 0.0,
                              "kernel_ms_per_step": cbam_ms, "note": "achieved = compulsory 144 B/px over the 4 CBAM sites"},
            "glue_ms_per_step": glue_ms,
        }
        dense_ms = sum(v[0] for k, v in spans.items() if k.split("|")[1] in DENSE3X3) / args.steps
        dense_bytes = sum(2.0 * (cin + 16) * n * (h // div) * (w // div) for cin, div in DENSE3X3.values())
        if dense_ms > 0:
            gbs = dense_bytes / (dense_ms * 1e-3) / 1e9
            line["roofline_dense"] = {
                "bound": "hbm", "kernel": "conv_stream2_kernel<4,0,GP> (16 dense-block 3x3 launches of a step; GP=1 group-planar final dense block, GP=2 hybrid buffers of dense blocks 1-3)",
                "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                "traffic": DENSE3X3_NCU_TRAFFIC_BYTES if (n, h, w) == (32, 1080, 1920) else None,
                "algorithmic_bytes_per_step": dense_bytes, "kernel_ms_per_step": dense_ms,
                "kernel_share_of_step": dense_ms / ms_step if ms_step else None, "launches_per_step": len(DENSE3X3)}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            mp_s, sample, _ = cpu_oracle_rate(h, w, args.cpu_budget, threads)
            line["cpu_baseline"] = {"value": mp_s, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
