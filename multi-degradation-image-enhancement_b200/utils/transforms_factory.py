"""`utils.transforms_factory` — config -> transform (reference utils/transforms_factory.py:19-127).

Input side of the pipeline ("next"): kept minimal.  The deterministic test-time ops of the shipped
configs (Resize -> Normalize(mean 0, std 1) -> ToTensorV2, i.e. bilinear resize and /255) are implemented directly
on PIL/torch for BOTH backend names, so configs load without albumentations (absent offline).  Random training
augmentations require the real albumentations package."""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
from PIL import Image

_RANDOM_AUG = {"HorizontalFlip", "VerticalFlip", "RandomRotate90", "RandomGamma", "RandomBrightnessContrast"}


class _Pipeline:
    def __init__(self, ops):
        self.resize, self.mean, self.std = None, None, None
        for op in ops or []:
            name, args = op["name"], op.get("args", {}) or {}
            if name == "Resize":
                size = args.get("size")
                self.resize = (int(args["height"]), int(args["width"])) if size is None else tuple(size)
            elif name == "Normalize":
                self.mean = torch.tensor(args.get("mean", [0.0, 0.0, 0.0])).view(3, 1, 1)
                self.std = torch.tensor(args.get("std", [1.0, 1.0, 1.0])).view(3, 1, 1)
            elif name in ("ToTensorV2", "ToTensor"):
                pass
            elif name in _RANDOM_AUG:
                try:
                    import albumentations  # noqa: F401
                except Exception as exc:
                    raise ImportError(f"transform '{name}' needs albumentations (training augmentation)") from exc
                raise NotImplementedError("random augmentations are outside the accelerated test path")
            else:
                raise ValueError(f"Unknown transform op: {name}")

    def __call__(self, img: Image.Image) -> torch.Tensor:
        if self.resize is not None:
            img = img.resize((self.resize[1], self.resize[0]), Image.BILINEAR)
        t = torch.from_numpy(np.asarray(img, dtype=np.uint8).copy()).permute(2, 0, 1).float() / 255.0
        if self.mean is not None:
            t = (t - self.mean) / self.std
        return t


def build_transforms(transform_cfg: Optional[Dict[str, Any]], is_paired: bool):
    cfg = transform_cfg or {}
    backend = cfg.get("backend", "torchvision")
    return backend, _Pipeline(cfg.get("ops", []))


def apply_paired_transform(backend: str, tf, inp_pil: Image.Image, tgt_pil: Image.Image) -> Tuple[torch.Tensor, torch.Tensor]:
    return tf(inp_pil), tf(tgt_pil)


def apply_single_transform(backend: str, tf, inp_pil: Image.Image) -> torch.Tensor:
    return tf(inp_pil)
