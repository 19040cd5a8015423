"""Inference side of the reference's multi-label degradation classifier (BASELINE config C4).

Mirrors classification/train_multilabel_classifier.py of the reference: `MultiHeadClassifier` (:117-131, ResNet18 backbone with
`fc = Identity`, heads `head_cls` / `head_sev`; same module tree, so a reference checkpoint's `model_state` loads with
strict=True), the evaluation transform (:773-777: Resize((256, 384)), ToTensor, ImageNet Normalize :47-48), the class order of
datasets_generation/generate_classifier_dataset.py:47-57 (written to meta/classes.json), `apply_thresholds` (:251-253) and the
`thresholds_val.json` report written at :923 (`{"thresholds": {class: value}}`).

The classifier's convolution stack runs on stock torch / torchvision (cuDNN on a GPU): it is ~1 % of the routed pipeline's
time and outside SURVEY 8(a)'s hot path; the CDAN enhancers behind it run on this library's kernels.  The reference builds the
backbone from ImageNet weights (a download); without network access the constructor takes `pretrained=False` — the weights
that matter are the trained checkpoint's.
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

DEGRADATIONS: List[str] = ["blur", "noise", "low_light", "jpeg", "pixelation", "motion_blur", "high_light", "low_contrast",
                           "color_distortion"]  # generate_classifier_dataset.py:47-57
DEFAULT_THRESH = 0.5  # train_multilabel_classifier.py:35
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
EVAL_SIZE = (256, 384)  # transforms.Resize((256, 384)), :774


class MultiHeadClassifier(nn.Module):
    def __init__(self, num_classes: int = len(DEGRADATIONS), pretrained: bool = False):
        super().__init__()
        from torchvision import models
        backbone = models.resnet18(weights=models.ResNet18_Weights.IMAGENET1K_V1 if pretrained else None)
        in_features = backbone.fc.in_features
        backbone.fc = nn.Identity()
        self.backbone = backbone
        self.head_cls = nn.Linear(in_features, num_classes)  # logits
        self.head_sev = nn.Linear(in_features, num_classes)  # logits -> sigmoid -> [0,1]

    def forward(self, x):
        feat = self.backbone(x)
        return self.head_cls(feat), self.head_sev(feat)


def load_checkpoint(model: MultiHeadClassifier, path: str, map_location="cpu") -> MultiHeadClassifier:
    """`torch.save({"model_state": ...})` of the reference's training loop (loaded at :898-899)."""
    ckpt = torch.load(path, map_location=map_location)
    model.load_state_dict(ckpt["model_state"] if "model_state" in ckpt else ckpt)
    return model.eval()


def load_thresholds(path: Optional[str], classes: Sequence[str] = DEGRADATIONS) -> List[float]:
    """Per-class thresholds from the reference's thresholds_val.json (:298,:923); DEFAULT_THRESH for every class without a file."""
    if not path:
        return [DEFAULT_THRESH] * len(classes)
    with open(path, encoding="utf-8") as f:
        rep = json.load(f)
    th = rep.get("thresholds", rep)
    if isinstance(th, dict):
        missing = [c for c in classes if c not in th]
        if missing:
            raise KeyError(f"thresholds file lacks classes {missing}")
        return [float(th[c]) for c in classes]
    if len(th) != len(classes):
        raise ValueError("one threshold per class expected")
    return [float(v) for v in th]


def apply_thresholds(probs: torch.Tensor, thresholds: Sequence[float]) -> torch.Tensor:
    """`(probs >= th)` per class (:251-253)."""
    return probs >= torch.as_tensor(list(thresholds), dtype=probs.dtype, device=probs.device).reshape(1, -1)


def preprocess(images: torch.Tensor, normalize: bool = True, size: Tuple[int, int] = EVAL_SIZE) -> torch.Tensor:
    """The evaluation transform on a [N,3,H,W] float batch in [0,1] (what ToTensor yields): resize to 256 x 384 (bilinear with
    antialiasing, torchvision's Resize on tensors) and ImageNet normalisation."""
    if tuple(images.shape[-2:]) != tuple(size):
        images = F.interpolate(images, size=size, mode="bilinear", align_corners=False, antialias=True)
    if normalize:
        mean = torch.tensor(IMAGENET_MEAN, dtype=images.dtype, device=images.device).view(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD, dtype=images.dtype, device=images.device).view(1, 3, 1, 1)
        images = (images - mean) / std
    return images


@torch.no_grad()
def predict_probs(model: MultiHeadClassifier, images: torch.Tensor, normalize: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sigmoid class probabilities and severities (the reference's evaluation loop, :224-247)."""
    cls_logits, sev_logits = model(preprocess(images, normalize))
    return torch.sigmoid(cls_logits), torch.sigmoid(sev_logits)
