"""Multi-degradation routing (BASELINE config C4): run every image of a mixed batch through the CDAN
weight set(s) of the degradations a classifier flagged for it.

The reference trains a multi-label classifier (classification/train_multilabel_classifier.py) whose per-class sigmoid
probabilities are thresholded per class (`apply_thresholds`, :251-253: `probs >= thresholds`, default 0.5 at :35) and it
trains one CDAN per degradation (config/{noise,blur,jpeg,low_contrast,pixelation,...}.json), but it ships NO code that
connects the two.  This module defines that missing step; the semantics are this build's own and deliberately minimal:

* an image with several active classes is enhanced SEQUENTIALLY in the fixed order of `class_order`;
* an image with no active class is returned unchanged (identity);
* images are bucketed per class, each bucket runs as ONE sub-batch through that class's enhancer and is scattered back;
  the forward is batch-independent (bitwise), so bucketing never changes a result.

The enhancers are callables `[n,3,H,W] -> [n,3,H,W]` — normally eval-mode `models.cdan.CDAN` instances that share one
architecture and differ only in weights (each owns a native plan on its device).  The classifier's own convolution stack
is not on the accelerated path and stays on torch.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Mapping, Sequence

import torch

Enhancer = Callable[[torch.Tensor], torch.Tensor]


def active_classes(probs: torch.Tensor, thresholds) -> torch.Tensor:
    """`probs >= thresholds` per class (reference classification/train_multilabel_classifier.py:251-253)."""
    th = torch.as_tensor(thresholds, dtype=probs.dtype, device=probs.device).reshape(1, -1)
    if th.shape[1] == 1:
        th = th.expand(1, probs.shape[1])
    if th.shape[1] != probs.shape[1]:
        raise ValueError("one threshold per class (or a single scalar) expected")
    return probs >= th


class DegradationRouter:
    def __init__(self, enhancers: Mapping[str, Enhancer], class_order: Sequence[str], thresholds=0.5):
        missing = [c for c in class_order if c not in enhancers]
        if missing:
            raise KeyError(f"no enhancer for classes {missing}")
        if len(set(class_order)) != len(class_order):
            raise ValueError("class_order must not repeat a class")
        self.enhancers: Dict[str, Enhancer] = dict(enhancers)
        self.class_order: List[str] = list(class_order)
        self.thresholds = thresholds
        self.last_bucket_sizes: Dict[str, int] = {}

    @torch.no_grad()
    def __call__(self, images: torch.Tensor, probs: torch.Tensor) -> torch.Tensor:
        """images [N,3,H,W]; probs [N,K] with K == len(class_order) (classifier sigmoid outputs in class_order)."""
        if images.dim() != 4 or probs.dim() != 2 or probs.shape[0] != images.shape[0]:
            raise ValueError("expected images [N,C,H,W] and probs [N,K]")
        if probs.shape[1] != len(self.class_order):
            raise ValueError(f"probs has {probs.shape[1]} classes, router has {len(self.class_order)}")
        active = active_classes(probs, self.thresholds).to("cpu")
        out = images.clone()
        self.last_bucket_sizes = {}
        for k, name in enumerate(self.class_order):
            idx = torch.nonzero(active[:, k], as_tuple=False).flatten()
            self.last_bucket_sizes[name] = int(idx.numel())
            if idx.numel() == 0:
                continue
            idx_dev = idx.to(images.device)
            enhanced = self.enhancers[name](out.index_select(0, idx_dev).contiguous())
            if enhanced.shape != (idx.numel(),) + tuple(images.shape[1:]):
                raise RuntimeError(f"enhancer '{name}' changed the image shape")
            out.index_copy_(0, idx_dev, enhanced.to(out.dtype))
        return out
