"""PSNR / SSIM oracle.  TEST INFRASTRUCTURE ONLY.   *** PARITY UNPINNED ***

The reference computes these through ``torchmetrics`` with all-default constructor arguments
(reference utils/metrics_factory.py:74-94: ``PeakSignalNoiseRatio()``, ``StructuralSimilarityIndexMeasure()``).
torchmetrics is listed WITHOUT a version in the reference's requirements.txt:9, its source is not part of the
reference tree and it is not installed in the build image, so no output of the real dependency can be generated
here.  This file restates the published default algorithm (torchmetrics 1.x functional implementations):

PSNR   : data_range=None -> the metric tracks min/max of the TARGET against 0-initialised states, so
         range = max(target.max(), 0) - min(target.min(), 0); value = 10*log10(range^2 / mean((p-t)^2)) over
         all elements of the batch.
SSIM   : gaussian 11x11, sigma 1.5, k1 0.01, k2 0.03, data_range=None -> max(p.max()-p.min(), t.max()-t.min());
         inputs reflect-padded by 5, filtered, and the result cropped by 5 on each side again, i.e. only the
         (H-10)x(W-10) interior whose window lies fully inside the image contributes; variances clamped at 0;
         per-image mean, then batch mean.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def psnr(pred: torch.Tensor, target: torch.Tensor) -> float:
    p, t = pred.double().cpu(), target.double().cpu()
    rng = max(float(t.max()), 0.0) - min(float(t.min()), 0.0)
    mse = float(((p - t) ** 2).mean())
    return float(10.0 * torch.log10(torch.tensor(rng * rng / mse, dtype=torch.float64)))


def ssim(pred: torch.Tensor, target: torch.Tensor) -> float:
    p, t = pred.double().cpu(), target.double().cpu()
    c = p.shape[1]
    rng = max(float(p.max() - p.min()), float(t.max() - t.min()))
    c1, c2 = (0.01 * rng) ** 2, (0.03 * rng) ** 2
    d = torch.arange(11, dtype=torch.float64) - 5
    g = torch.exp(-(d ** 2) / (2 * 1.5 ** 2))
    g = g / g.sum()
    k = (g[:, None] * g[None, :]).expand(c, 1, 11, 11).contiguous()

    def filt(z):  # valid filtering == reflect-pad + filter + crop(5) of the torchmetrics implementation
        return F.conv2d(z, k, groups=c)

    mp, mt = filt(p), filt(t)
    vp = (filt(p * p) - mp * mp).clamp(min=0)
    vt = (filt(t * t) - mt * mt).clamp(min=0)
    cov = filt(p * t) - mp * mt
    m = ((2 * mp * mt + c1) * (2 * cov + c2)) / ((mp * mp + mt * mt + c1) * (vp + vt + c2))
    return float(m.reshape(m.shape[0], -1).mean(-1).mean())
