"""Forward time by batch size and end-to-end time by host_chunk (tuning of cdan_forward_host's sub-batch pipeline)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-degradation-image-enhancement_b200")); sys.path.insert(0, ROOT)
import torch
import cdan_b200_native as native
from bench import default_weights, synthetic_batch
dev = torch.device("cuda", 0)
plan = native.Plan(dev, "bf16"); plan.load_state_dict(default_weights(42))
h, w = 1080, 1920
for n in (1, 2, 4, 8, 16, 32):
    x = torch.rand(n, 3, h, w, device=dev); y = torch.empty_like(x)
    for _ in range(3): plan.forward(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): plan.forward(x, out=y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"batch {n:2d}: {ms:7.3f} ms  {ms / n:6.3f} ms/img  {n * h * w / 1e3 / ms:7.1f} MP/s", flush=True)
n = 32
xh = synthetic_batch(n, h, w).pin_memory(); yh = torch.empty_like(xh).pin_memory()
xu = (xh * 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory(); yu = torch.empty_like(xu).pin_memory()
for chunk in (0, 2, 4, 8, 16):
    plan.set_option("host_chunk", chunk)
    for name, fn, a, b in (("f32", plan.forward_host, xh, yh), ("u8", plan.forward_host_u8, xu, yu)):
        fn(a, b); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): fn(a, b)
        dt = (time.perf_counter() - t0) / 3
        print(f"host_chunk {chunk:2d} {name}: {dt * 1e3:7.2f} ms  {n * h * w / 1e6 / dt:7.1f} MP/s", flush=True)
