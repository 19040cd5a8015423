"""Timeline trace of the fused final dense block (CDAN_FUSED_TRACE=1): one 1080p-wide forward, CTA 0, first item."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-degradation-image-enhancement_b200")); sys.path.insert(0, ROOT)
os.environ["CDAN_FUSED_TRACE"] = "1"
import torch
from models.cdan import CDAN
torch.manual_seed(0)
net = CDAN().set_compute_dtype("bf16").to("cuda:0").eval()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
x = torch.rand(n, 3, 1080, 1920, device="cuda:0")
with torch.no_grad():
    net(x); torch.cuda.synchronize()
