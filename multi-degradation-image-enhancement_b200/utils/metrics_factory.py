"""`utils.metrics_factory` — `build_metrics_pipeline` / `MetricsPipeline` (reference utils/metrics_factory.py:14-111).

psnr and ssim are computed by ONE fused native reduction per (outputs, targets) pair (csrc/postproc.cu,
torchmetrics-default semantics restated: parity unpinned, see oracle/metrics_oracle.py) and shared between the two
items, so a batch costs one kernel sequence and one D2H instead of one `.item()` sync per metric.
lpips needs the pretrained AlexNet of the `lpips`/torchmetrics packages (absent offline, out of scope): the item is
accepted in configs and skipped with a one-time notice."""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch

import cdan_b200_native as _native


@dataclass
class MetricItem:
    name: str
    mode: str  # "paired" or "unpaired"
    fn: Any


class _PsnrSsim:
    """Caches the fused (psnr, ssim) result for the most recent (outputs, targets) pair."""

    def __init__(self):
        self._key, self._val = None, (float("nan"), float("nan"))

    def __call__(self, outputs, targets):
        if targets is None:
            raise ValueError("psnr/ssim metrics require targets (paired dataset).")
        key = (outputs.data_ptr(), targets.data_ptr(), outputs._version, tuple(outputs.shape))
        if key != self._key:
            self._val = _native.psnr_ssim(outputs, targets)
            self._key = key
        return self._val


class MetricsPipeline:
    def __init__(self, metrics: Dict[str, MetricItem]):
        self.metrics = metrics

    def __call__(self, outputs, targets=None, inputs=None, is_paired: bool = True) -> Dict[str, torch.Tensor]:
        out: Dict[str, torch.Tensor] = {}
        for name, item in self.metrics.items():
            if (item.mode == "paired") != bool(is_paired):
                continue
            val = item.fn(outputs=outputs, targets=targets, inputs=inputs)
            val = torch.as_tensor(val)
            out[name] = val.mean() if val.ndim else val
        return out


def build_metrics_pipeline(metrics_cfg: Optional[Dict[str, Any]], device: str) -> MetricsPipeline:
    if not metrics_cfg or not metrics_cfg.get("enabled", True):
        return MetricsPipeline({})
    fused = _PsnrSsim()
    metrics: Dict[str, MetricItem] = {}
    for it in metrics_cfg.get("items", []) or []:
        name, mode = it["name"], it.get("mode", "paired")
        if name == "psnr":
            metrics[name] = MetricItem(name, mode, lambda outputs, targets, inputs=None: fused(outputs, targets)[0])
        elif name == "ssim":
            metrics[name] = MetricItem(name, mode, lambda outputs, targets, inputs=None: fused(outputs, targets)[1])
        elif name == "lpips":
            warnings.warn("cdan_b200: metric 'lpips' needs pretrained AlexNet weights (unavailable offline); skipped")
        else:
            raise ValueError(f"Unknown metric: {name}")
    return MetricsPipeline(metrics)
