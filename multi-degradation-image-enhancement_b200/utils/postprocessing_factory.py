"""`utils.postprocessing_factory` — name -> op table and `apply_postprocessing` (reference
utils/postprocessing_factory.py:10-41): disabled/empty config returns the input tensor itself, unknown op names
raise ValueError."""
from __future__ import annotations

from typing import Any, Dict

import torch

from utils.post_processing import enhance_color, enhance_contrast, sharpen, soft_denoise

_OPS = {"enhance_contrast": enhance_contrast, "enhance_color": enhance_color, "sharpen": sharpen,
        "soft_denoise": soft_denoise}


def apply_postprocessing(images: torch.Tensor, pp_cfg: Dict[str, Any]) -> torch.Tensor:
    if not pp_cfg or not pp_cfg.get("enabled", False):
        return images
    out = images
    for op in pp_cfg.get("ops", []) or []:
        if op["name"] not in _OPS:
            raise ValueError(f"Unknown post-processing op: {op['name']}")
        out = _OPS[op["name"]](out, **(op.get("args", {}) or {}))
    return out
