"""`utils.loss_factory` — `build_loss_pipeline` / `LossPipeline` (reference utils/loss_factory.py:17-235).

Training losses are OUTSIDE the accelerated hot path (out of scope); this module exists because
`models.model.Model` builds the pipeline even in the test phase and `test_step` reports the terms.  mse / l1 /
charbonnier / gradient_l1 are plain torch (autograd-capable, used by `train_step`).  `ssim` (= 1 - SSIM) uses the
native fused metric kernel when no gradient is required (test phase).  `vgg_perceptual` and `lpips` need pretrained
VGG19 / AlexNet weights that cannot be downloaded here: they are accepted in configs and skipped with a notice."""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch
import torch.nn.functional as F


@dataclass
class LossTerm:
    name: str
    weight: float
    mode: str  # "paired" or "unpaired"
    fn: Any


class LossPipeline:
    """Weighted sum of terms; returns every component plus 'total' (reference :25-56)."""

    def __init__(self, terms: List[LossTerm]):
        self.terms = terms

    def __call__(self, outputs, targets=None, inputs=None, is_paired: bool = True) -> Dict[str, torch.Tensor]:
        comps: Dict[str, torch.Tensor] = {}
        total = torch.zeros((), device=outputs.device)
        for term in self.terms:
            if (term.mode == "paired") != bool(is_paired):
                continue
            val = torch.as_tensor(term.fn(outputs=outputs, targets=targets, inputs=inputs), device=outputs.device)
            val = val.mean() if val.ndim else val
            comps[term.name] = val
            total = total + term.weight * val
        comps["total"] = total
        return comps


def _sobel_gradients(x: torch.Tensor) -> torch.Tensor:
    """Per-channel Sobel x/y responses, zero padding -> [B,C,2,H,W] (reference :71-103)."""
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], device=x.device, dtype=x.dtype)
    k = torch.stack((kx, kx.t())).unsqueeze(1)  # [2,1,3,3]
    b, c, h, w = x.shape
    g = F.conv2d(x.reshape(b * c, 1, h, w), k, padding=1)
    return g.reshape(b, c, 2, h, w)


def _gray(x: torch.Tensor) -> torch.Tensor:
    if x.shape[1] != 3:
        return x.mean(dim=1, keepdim=True)
    return 0.2989 * x[:, 0:1] + 0.5870 * x[:, 1:2] + 0.1140 * x[:, 2:3]


def _ssim_torch(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Differentiable SSIM with torchmetrics' default settings restated (11x11 Gaussian sigma 1.5, k1 0.01, k2 0.03,
    data range from the batch, reflect padding then crop = statistics over the valid region; parity unpinned, see
    oracle/metrics_oracle.py).  Used by the 'ssim' loss term when a gradient is required (train_step)."""
    c = x.shape[1]
    d = torch.arange(11, dtype=x.dtype, device=x.device) - 5
    g = torch.exp(-(d * d) / (2 * 1.5 * 1.5))
    g = g / g.sum()
    k = (g[:, None] * g[None, :]).expand(c, 1, 11, 11).contiguous()
    rng = torch.maximum(x.detach().max() - x.detach().min(), y.max() - y.min())
    c1, c2 = (0.01 * rng) ** 2, (0.03 * rng) ** 2
    mu_x, mu_y = F.conv2d(x, k, groups=c), F.conv2d(y, k, groups=c)
    sxx = F.conv2d(x * x, k, groups=c) - mu_x * mu_x
    syy = F.conv2d(y * y, k, groups=c) - mu_y * mu_y
    sxy = F.conv2d(x * y, k, groups=c) - mu_x * mu_y
    ssim = ((2 * mu_x * mu_y + c1) * (2 * sxy + c2)) / ((mu_x * mu_x + mu_y * mu_y + c1) * (sxx + syy + c2))
    return ssim.flatten(1).mean(1).mean()


def _need(targets, name):
    if targets is None:
        raise ValueError(f"{name} loss requires targets (paired dataset).")


def build_loss_pipeline(loss_cfg: Optional[Dict[str, Any]], device: str) -> LossPipeline:
    if not loss_cfg or not loss_cfg.get("enabled", True):  # reference :117-123: default fallback = plain MSE
        loss_cfg = {"terms": [{"name": "mse", "weight": 1.0, "args": {}}]}
    terms_cfg = loss_cfg.get("terms", []) or [{"name": "mse", "weight": 1.0, "args": {}}]
    terms: List[LossTerm] = []
    for t in terms_cfg:
        name, weight = t["name"], float(t.get("weight", 1.0))
        args, mode = t.get("args", {}) or {}, t.get("mode", "paired")
        if name == "mse":
            def fn(outputs, targets, inputs=None):
                _need(targets, "mse")
                return F.mse_loss(outputs, targets)
        elif name == "l1":
            def fn(outputs, targets, inputs=None):
                _need(targets, "l1")
                return F.l1_loss(outputs, targets)
        elif name == "charbonnier":
            eps = float(args.get("eps", 1e-3))

            def fn(outputs, targets, inputs=None, eps=eps):
                _need(targets, "charbonnier")
                d = outputs - targets
                return torch.sqrt(d * d + eps * eps).mean()
        elif name == "gradient_l1":
            to_gray = bool(args.get("to_gray", False))

            def fn(outputs, targets, inputs=None, to_gray=to_gray):
                _need(targets, "gradient_l1")
                x, y = (_gray(outputs), _gray(targets)) if to_gray else (outputs, targets)
                return (_sobel_gradients(x) - _sobel_gradients(y)).abs().mean()
        elif name == "ssim":
            def fn(outputs, targets, inputs=None):
                _need(targets, "ssim")
                if outputs.requires_grad or not outputs.is_cuda:
                    return 1.0 - _ssim_torch(outputs, targets)  # training / CPU: differentiable torch composition
                import cdan_b200_native as native
                return 1.0 - native.psnr_ssim(outputs, targets)[1]
        elif name in ("vgg_perceptual", "lpips"):
            warnings.warn(f"cdan_b200: loss term '{name}' needs pretrained weights (unavailable offline); skipped")
            continue
        else:
            raise ValueError(f"Unknown loss term: {name}")
        terms.append(LossTerm(name=name, weight=weight, mode=mode, fn=fn))
    return LossPipeline(terms)
