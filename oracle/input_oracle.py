"""Input side of the hot path: uint8 HWC image -> Resize(256, 384) -> /255 -> CHW float32.
TEST INFRASTRUCTURE ONLY, like the rest of oracle/.

Reference call sites: data/dataset.py:86-92 (PIL -> numpy uint8 HWC -> transform), utils/transforms_factory.py:50-86
(`A.Resize(height, width)` = cv2.resize(..., interpolation=cv2.INTER_LINEAR), then `A.Normalize(mean=0, std=1,
max_pixel_value=255)`, then `ToTensorV2`).  Both steps are THIRD-PARTY arithmetic:

* cv2.resize INTER_LINEAR on uint8 — OpenCV's fixed-point bilinear (11-bit coefficients, imgproc/src/resize.cpp:
  `HResizeLinear` / `VResizeLinear<uchar,int,short,FixedPtCast<...,22>>`).  OpenCV IS installed in the build container,
  so `resize_linear_u8` below is PINNED against `cv2.resize` itself (tests/test_oracle_input.py), bit-exactly.
* albumentations.Normalize — not installed here (requirements.txt:1, unpinned): restated from its published
  implementation (`img.astype(float32); img -= mean*max; img *= reciprocal(std*max, dtype=float32)`), i.e. a multiply
  by float32(1/255), not a division.  PARITY UNPINNED for this one step.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS  # INTER_RESIZE_COEF_SCALE


def linear_coeffs(src: int, dst: int, vertical: bool = False):
    """Per output index: the two source indices and the two 11-bit weights, exactly as cv::resize computes them
    (resize.cpp `resizeGeneric_` set-up: float fraction of a double coordinate, cvFloor, cvRound of weight * 2048).
    Horizontally the fraction is forced to 0 at the borders (`if (sx < 0) fx = 0, sx = 0; if (sx >= width-1) fx = 0,
    sx = width-1`); VERTICALLY it is not — the invoker only clips the two row indices (`clip(sy + k, 0, height)`), so a
    border row is blended with itself using the unclamped weights (which sum to 2048 +- 1 after rounding)."""
    scale = 1.0 / (float(dst) / float(src))  # scale_x = 1. / inv_scale_x, both double
    i0 = np.empty(dst, np.int32)
    i1 = np.empty(dst, np.int32)
    w = np.empty((dst, 2), np.int32)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if not vertical:
            if s < 0:
                f, s = np.float32(0.0), 0
            if s >= src - 1:
                f, s = np.float32(0.0), src - 1
        i0[d] = min(max(s, 0), src - 1)
        i1[d] = min(max(s + 1, 0), src - 1)
        # saturate_cast<short>(float * 2048) == cvRound (round half to even)
        w[d, 0] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(COEF_SCALE)))
        w[d, 1] = int(np.rint(f * np.float32(COEF_SCALE)))
    return i0, i1, w


def resize_linear_u8(img: np.ndarray, out_hw) -> np.ndarray:
    """cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR) for uint8 HWC images, restated."""
    assert img.dtype == np.uint8 and img.ndim == 3
    hs, ws, _ = img.shape
    hd, wd = out_hw
    x0, x1, xa = linear_coeffs(ws, wd)
    y0, y1, yb = linear_coeffs(hs, hd, vertical=True)
    src = img.astype(np.int32)
    # horizontal pass: int rows scaled by 2^11
    rows = src[:, x0, :] * xa[None, :, 0, None] + src[:, x1, :] * xa[None, :, 1, None]
    s0, s1 = rows[y0], rows[y1]
    b0, b1 = yb[:, 0][:, None, None], yb[:, 1][:, None, None]
    out = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def normalize_to_chw(img_u8: np.ndarray) -> np.ndarray:
    """A.Normalize(mean=0, std=1, max_pixel_value=255) + ToTensorV2: float32(u8) * float32(1/255), HWC -> CHW."""
    denom = np.reciprocal(np.float32(255.0), dtype=np.float32)
    return np.ascontiguousarray((img_u8.astype(np.float32) * denom).transpose(2, 0, 1))


def network_input(img_u8: np.ndarray, out_hw=(256, 384)) -> np.ndarray:
    return normalize_to_chw(resize_linear_u8(img_u8, out_hw))
