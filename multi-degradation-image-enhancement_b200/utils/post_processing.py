"""`utils.post_processing` — the four image-space ops of the reference (utils/post_processing.py:5-77) executed by
fused sm_100a kernels (csrc/postproc.cu).  Same names, arguments and semantics, including the
`images.max() > 1 -> /255` rule (evaluated on the device: no host sync per op) and `sharpen`'s use of the 3x3
identity MATRIX (+1 on the three diagonal taps).  CUDA tensors only: there is no CPU fallback."""
from __future__ import annotations

import torch

import cdan_b200_native as _native


def _run(images: torch.Tensor, op: str, arg: float) -> torch.Tensor:
    if not images.is_cuda:
        raise RuntimeError("cdan_b200: post-processing kernels run on CUDA tensors only (no CPU fallback)")
    return _native.postprocess(images, op, float(arg))


def enhance_contrast(images, contrast_factor=1.1):
    return _run(images, "enhance_contrast", contrast_factor)


def enhance_color(images, saturation_factor=1.1):
    return _run(images, "enhance_color", saturation_factor)


def sharpen(images, strength=0.5):
    return _run(images, "sharpen", strength)


def soft_denoise(images, sigma=0.2):
    return _run(images, "soft_denoise", sigma)
