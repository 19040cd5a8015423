"""Debug helper (GPU box): find the first stage whose batched result differs from the single-image result at 1080p."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-degradation-image-enhancement_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_forward import make_net, stress_state_dict, ramp_input  # noqa: E402

dev = torch.device("cuda", 0)
sd = stress_state_dict(1234)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
x = ramp_input(n, 1080, 1920, seed=21).to(dev)
net = make_net(sd, "bf16", dev)
plan = net.native_plan()
names = ["enc.out1", "enc.dense1", "enc.out2", "enc.dense2", "enc.out3", "enc.dense3", "enc.conv4", "bottleneck", "dec.bn1",
         "dec.gated1", "dec.bn2", "dec.gated2", "dec.bn3", "dec.gated3", "dec.bn4", "dec.final_in"]
with torch.no_grad():
    y = net(x)
    full = {k: plan.stage(k)[1:2].clone() for k in names}
    y1 = net(x[1:2].contiguous())
    single = {k: plan.stage(k).clone() for k in names}
    y_again = net(x)
for k in names:
    d = (full[k].float() - single[k].float()).abs()
    print(f"{k:14s} equal={bool(torch.equal(full[k], single[k]))} maxdiff={float(d.max()):.3e} ndiff={int((d > 0).sum())}")
d = (y[1:2] - y1).abs()
print("output equal", bool(torch.equal(y[1:2], y1)), float(d.max()), int((d > 0).sum()), "repeat equal", bool(torch.equal(y, y_again)))
if int((d > 0).sum()):
    idx = (d > 0).nonzero()
    print("first diffs (n,c,h,w):", idx[:8].tolist(), "rows:", sorted(set(idx[:, 2].tolist()))[:40], "cols min/max", int(idx[:, 3].min()), int(idx[:, 3].max()))
