"""`data.dataset` — paired / unpaired image-folder datasets (reference data/dataset.py:29-112): same constructor
arguments, pairing modes ("filename", "stem", "sorted") and error behaviour.  CPU data loading is outside the hot path."""
from __future__ import annotations

import os
from typing import Dict, List, Optional

from PIL import Image
from torch.utils.data import Dataset

from utils.transforms_factory import apply_paired_transform, apply_single_transform, build_transforms

_EXTS = (".png", ".jpg", ".jpeg", ".bmp", ".tif", ".tiff", ".webp")


def _list_images(folder: str) -> List[str]:
    return sorted(f for f in os.listdir(folder) if not f.startswith(".") and f.lower().endswith(_EXTS))


class PairedDataset(Dataset):
    def __init__(self, input_root: str, target_root: str, pairing_mode: str = "filename",
                 transform: Optional[Dict] = None, image_size: Optional[List[int]] = None):
        super().__init__()
        self.input_root, self.target_root = input_root, target_root
        inp, tgt = _list_images(input_root), _list_images(target_root)
        if pairing_mode == "sorted":
            self.pairs = [(os.path.join(input_root, a), os.path.join(target_root, b)) for a, b in zip(inp, tgt)]
        elif pairing_mode in ("filename", "stem"):
            key = (lambda f: f) if pairing_mode == "filename" else (lambda f: os.path.splitext(f)[0])
            a = {key(f): os.path.join(input_root, f) for f in inp}
            b = {key(f): os.path.join(target_root, f) for f in tgt}
            keys = sorted(set(a) & set(b))
            if not keys:
                raise RuntimeError(f"No paired files found with pairing_mode='{pairing_mode}'.\n"
                                   f"input_root={input_root}\ntarget_root={target_root}")
            self.pairs = [(a[k], b[k]) for k in keys]
        else:
            raise ValueError(f"Unknown pairing_mode: {pairing_mode}")
        self.backend, self.tf = build_transforms(transform, is_paired=True)

    def __len__(self):
        return len(self.pairs)

    def __getitem__(self, idx: int):
        a, b = self.pairs[idx]
        return apply_paired_transform(self.backend, self.tf, Image.open(a).convert("RGB"), Image.open(b).convert("RGB"))


class UnpairedDataset(Dataset):
    def __init__(self, input_root: str, transform: Optional[Dict] = None):
        super().__init__()
        self.input_root = input_root
        self.files = [os.path.join(input_root, f) for f in _list_images(input_root)]
        self.backend, self.tf = build_transforms(transform, is_paired=False)

    def __len__(self):
        return len(self.files)

    def __getitem__(self, idx: int):
        return apply_single_transform(self.backend, self.tf, Image.open(self.files[idx]).convert("RGB"))
