"""`utils.transforms_factory` — config -> transform (reference utils/transforms_factory.py:19-127).

Input side of the pipeline.  The deterministic test-time ops of the shipped configs
(Resize -> Normalize(mean 0, std 1, max 255) -> ToTensorV2) are applied IN CONFIG ORDER with the reference backend's
arithmetic: `A.Resize` is cv2.resize(INTER_LINEAR) on the uint8 image (no antialiasing — PIL's BILINEAR filter scales its
support when shrinking and gives different pixels), `A.Normalize` multiplies by float32(1/255).  cv2 is used when
importable; otherwise the bit-exact host port of OpenCV's 8-bit fixed-point bilinear is not shipped in the product, so the
PIL fallback is taken with a warning.  (On the device the same arithmetic is `cdan_resize_normalize_u8`.)
Random training augmentations are delegated to albumentations when it is importable (as the reference does) and fail with
one clear message otherwise — training is outside the accelerated path."""
from __future__ import annotations

import warnings
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
from PIL import Image

_RANDOM_AUG = {"HorizontalFlip", "VerticalFlip", "RandomRotate90", "RandomGamma", "RandomBrightnessContrast", "GaussNoise",
               "MotionBlur", "HueSaturationValue", "CLAHE", "Sharpen"}


def _resize_u8(img: np.ndarray, hw: Tuple[int, int]) -> np.ndarray:
    try:
        import cv2
        return cv2.resize(img, (hw[1], hw[0]), interpolation=cv2.INTER_LINEAR)
    except ImportError:  # pragma: no cover - cv2 is part of the image
        warnings.warn("cdan_b200: cv2 not importable; Resize falls back to PIL BILINEAR (antialiased when shrinking, "
                      "differs from the reference's cv2.INTER_LINEAR)")
        return np.asarray(Image.fromarray(img).resize((hw[1], hw[0]), Image.BILINEAR))


class _Pipeline:
    """Deterministic ops applied in config order on the uint8 HWC image, then CHW float32."""

    def __init__(self, ops):
        self.ops = []
        for op in ops or []:
            name, args = op["name"], op.get("args", {}) or {}
            if name == "Resize":
                size = args.get("size")
                self.ops.append(("resize", (int(args["height"]), int(args["width"])) if size is None else tuple(size)))
            elif name == "Normalize":
                mean = np.asarray(args.get("mean", [0.0, 0.0, 0.0]), dtype=np.float32) * 255.0
                std = np.asarray(args.get("std", [1.0, 1.0, 1.0]), dtype=np.float32) * 255.0
                self.ops.append(("normalize", (mean, np.float32(1.0) / std)))
            elif name in ("ToTensorV2", "ToTensor"):
                self.ops.append(("to_tensor", None))
            elif name in _RANDOM_AUG:
                raise NotImplementedError(
                    f"transform '{name}' is a random training augmentation: it needs the albumentations package, which "
                    "is not installed here; training is outside the accelerated test path")
            else:
                raise ValueError(f"Unknown transform op: {name}")

    def __call__(self, img) -> torch.Tensor:
        a = np.asarray(img, dtype=np.uint8)
        normalized = False
        for kind, arg in self.ops:
            if kind == "resize":
                a = _resize_u8(a, arg) if a.dtype == np.uint8 else a
            elif kind == "normalize":
                a = (a.astype(np.float32) - arg[0]) * arg[1]  # albumentations: (img - mean*255) * (1 / (std*255))
                normalized = True
        if not normalized:
            a = a.astype(np.float32) / 255.0  # torchvision ToTensor semantics
        return torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1)))


def _albumentations_pipeline(ops, is_paired):
    """Configs with random augmentations: build the reference's own A.Compose when albumentations is importable."""
    import albumentations as A
    from albumentations.pytorch import ToTensorV2
    built = []
    for op in ops or []:
        name, args = op["name"], op.get("args", {}) or {}
        if name == "Resize":
            size = args.get("size")
            h, w = (int(args["height"]), int(args["width"])) if size is None else tuple(size)
            built.append(A.Resize(height=h, width=w))
        elif name == "Normalize":
            built.append(A.Normalize(mean=args["mean"], std=args["std"]))
        elif name == "ToTensorV2":
            built.append(ToTensorV2())
        elif name in ("HorizontalFlip", "VerticalFlip", "RandomRotate90"):
            built.append(getattr(A, name)(p=args.get("p", 0.5)))
        elif name in _RANDOM_AUG:
            built.append(getattr(A, name)(**args))
        else:
            raise ValueError(f"[albumentations] Transform not supported: {name}")
    return A.Compose(built, additional_targets={"target": "image"} if is_paired else None)


def build_transforms(transform_cfg: Optional[Dict[str, Any]], is_paired: bool):
    cfg = transform_cfg or {}
    backend = cfg.get("backend", "torchvision")
    ops = cfg.get("ops", [])
    if backend == "albumentations" and any(op["name"] in _RANDOM_AUG for op in ops or []):
        try:
            return "albumentations_pkg", _albumentations_pipeline(ops, is_paired)
        except ImportError:
            pass  # fall through: _Pipeline raises the clear NotImplementedError
    return backend, _Pipeline(ops)


def apply_paired_transform(backend: str, tf, inp_pil: Image.Image, tgt_pil: Image.Image) -> Tuple[torch.Tensor, torch.Tensor]:
    if backend == "albumentations_pkg":
        out = tf(image=np.array(inp_pil), target=np.array(tgt_pil))
        return out["image"], out["target"]
    return tf(inp_pil), tf(tgt_pil)


def apply_single_transform(backend: str, tf, inp_pil: Image.Image) -> torch.Tensor:
    if backend == "albumentations_pkg":
        return tf(image=np.array(inp_pil))["image"]
    return tf(inp_pil)
