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

from typing import Callable, Dict, List, Mapping, Optional, Sequence

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


# ------------------------------------------------------------------------------------------------ C4 end to end
ENHANCER_CLASSES = ["noise", "blur", "jpeg", "low_contrast", "pixelation"]
"""BASELINE config C4's five per-degradation CDAN weight sets (reference config/{noise,blur,jpeg,low_contrast,pixelation}.json)."""


def degrade(img, name: str, sev: float, rng):
    """uint8 RGB [H,W,3] -> uint8 RGB: restatement of the reference's degradation functions for the five routed classes
    (datasets_generation/generate_classifier_dataset.py:212-262,300-307), same parameter ranges; rng = numpy Generator."""
    import cv2
    import numpy as np
    if name == "blur":  # :212-219, Gaussian kernel 3..9
        k = [3, 5, 7, 9][max(0, min(int(round(sev * 3)), 3))]
        return cv2.GaussianBlur(img, (k, k), 0)
    if name == "noise":  # :222-228, std 5..50
        out = img.astype(np.float32) + rng.normal(0.0, 5.0 + sev * 45.0, img.shape).astype(np.float32)
        return np.clip(out, 0, 255).astype(np.uint8)
    if name == "jpeg":  # :239-249, quality 80..10
        ok, enc = cv2.imencode(".jpg", cv2.cvtColor(img, cv2.COLOR_RGB2BGR), [int(cv2.IMWRITE_JPEG_QUALITY), int(round(80 - sev * 70))])
        return cv2.cvtColor(cv2.imdecode(enc, 1), cv2.COLOR_BGR2RGB) if ok else img
    if name == "pixelation":  # :252-262, factor 4..16
        h, w = img.shape[:2]
        f = max(2, min(int(round(4 + sev * 12)), min(h, w) // 2))
        small = cv2.resize(img, (max(1, w // f), max(1, h // f)), interpolation=cv2.INTER_LINEAR)
        return cv2.resize(small, (w, h), interpolation=cv2.INTER_NEAREST)
    if name == "low_contrast":  # :300-307, alpha 0.8..0.2
        alpha = 0.8 - sev * 0.6
        mean = img.mean(axis=(0, 1), keepdims=True).astype(np.float32)
        return np.clip(alpha * img.astype(np.float32) + (1 - alpha) * mean, 0, 255).astype(np.uint8)
    raise KeyError(name)


def synthetic_mixed_batch(n: int, h: int, w: int, seed: int = 0, classes: Sequence[str] = tuple(ENHANCER_CLASSES)):
    """A seeded mixed batch: smooth synthetic scenes, each degraded by 0-3 of `classes` (1: 60 %, 2: 30 %, 3: 10 % as in
    generate_classifier_dataset.py:42-43, plus ~10 % clean images).  Returns (uint8 [N,H,W,3] tensor, bool labels [N,K])."""
    import cv2
    import numpy as np
    rng = np.random.default_rng(seed)
    imgs, labels = [], np.zeros((n, len(classes)), dtype=bool)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for i in range(n):
        base = np.stack([127 + 120 * np.sin(xx / rng.uniform(9, 60) + rng.uniform(0, 6)) * np.cos(yy / rng.uniform(9, 60) + rng.uniform(0, 6))
                         for _ in range(3)], axis=-1)
        base += cv2.GaussianBlur(rng.normal(0, 25, (h, w, 3)).astype(np.float32), (5, 5), 0)
        img = np.clip(base, 0, 255).astype(np.uint8)
        k = 0 if rng.random() < 0.1 else int(rng.choice([1, 2, 3], p=[0.6, 0.3, 0.1]))
        for c in rng.choice(len(classes), size=k, replace=False):
            img = degrade(img, classes[int(c)], float(rng.beta(2.0, 2.0)), rng)
            labels[i, int(c)] = True
        imgs.append(img)
    return torch.from_numpy(np.stack(imgs)), torch.from_numpy(labels)


class MultiDegradationPipeline:
    """classifier -> per-class thresholds -> DegradationRouter over per-degradation CDAN enhancers (BASELINE config C4).

    `classifier` is a classification.multilabel_classifier.MultiHeadClassifier over `classifier_classes` (the reference's nine);
    only the classes that have an enhancer are routed, in `class_order`."""

    def __init__(self, classifier, enhancers: Mapping[str, Enhancer], class_order: Sequence[str] = tuple(ENHANCER_CLASSES),
                 classifier_classes: Optional[Sequence[str]] = None, thresholds=None, normalize: bool = True):
        from classification.multilabel_classifier import DEFAULT_THRESH, DEGRADATIONS
        self.classifier = classifier
        self.classes = list(classifier_classes or DEGRADATIONS)
        self.normalize = normalize
        th = list(thresholds) if thresholds is not None else [DEFAULT_THRESH] * len(self.classes)
        self.columns = [self.classes.index(c) for c in class_order]
        self.router = DegradationRouter(enhancers, class_order, [th[k] for k in self.columns])
        self.last_probs: Optional[torch.Tensor] = None

    @torch.no_grad()
    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        """images: float [N,3,H,W] in [0,1] on the enhancers' device.  Returns the enhanced batch."""
        from classification.multilabel_classifier import predict_probs
        probs, _ = predict_probs(self.classifier, images, self.normalize)
        self.last_probs = probs
        return self.router(images, probs[:, self.columns])
