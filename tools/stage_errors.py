"""Per-stage relative error of the bf16 plan vs the CPU oracle for a given shape (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-degradation-image-enhancement_b200")); sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import cdan_oracle as O
from oracle.stress_init import ramp_input, stress_state_dict
from models.cdan import CDAN
from test_gpu_forward import expected_stages
n, h, w, seed = [int(v) for v in sys.argv[1:5]]
sd = stress_state_dict(1234)
x = ramp_input(n, h, w, seed=seed)
y_ref, st = O.cdan_forward(sd, x, return_stages=True)
for dtype in ("fp32", "bf16"):
    net = CDAN().set_compute_dtype(dtype); net.load_state_dict(sd); net = net.to("cuda:0").eval()
    with torch.no_grad():
        y = net(x.cuda()).cpu()
    plan = net.native_plan()
    print(dtype, "output max abs", float((y - y_ref).abs().max()), "mean", float((y - y_ref).abs().mean()))
    for name, ref in expected_stages(st).items():
        got = plan.stage(name).cpu()
        err = float((got - ref).abs().max()) / max(1.0, float(ref.abs().max()))
        rms = float((got - ref).double().pow(2).mean().sqrt() / ref.double().pow(2).mean().sqrt())
        print(f"  {name:12s} rel max {err:.3e} rel rms {rms:.3e}  (|ref| max {float(ref.abs().max()):.2f})")
