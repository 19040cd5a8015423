"""Independent cross-check of oracle/metrics_oracle.py: torchmetrics is not installed, so the restatement
is checked against a SECOND, separately written fp64 implementation built on scipy.ndimage (Wang et al.'s SSIM as skimage
computes it with gaussian_weights=True: `gaussian_filter(sigma=1.5, truncate=3.5)` = an 11-tap window, 'reflect' boundary,
the 5-pixel border cropped before the mean) and cv2.PSNR for the PSNR formula.  Parity with the real dependency stays
unpinned (stated in the oracle's header); this test only rules out a private misreading of the published algorithm."""
import numpy as np
import pytest
import torch
from scipy import ndimage

from oracle import metrics_oracle


def ssim_scipy(pred: np.ndarray, target: np.ndarray) -> float:
    p, t = pred.astype(np.float64), target.astype(np.float64)
    rng = max(p.max() - p.min(), t.max() - t.min())
    c1, c2 = (0.01 * rng) ** 2, (0.03 * rng) ** 2
    vals = []
    for n in range(p.shape[0]):
        per_channel = []
        for c in range(p.shape[1]):
            x, y = p[n, c], t[n, c]
            f = lambda z: ndimage.gaussian_filter(z, sigma=1.5, truncate=3.5, mode="reflect")
            ux, uy = f(x), f(y)
            vx, vy, vxy = f(x * x) - ux * ux, f(y * y) - uy * uy, f(x * y) - ux * uy
            s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
            per_channel.append(s[5:-5, 5:-5])
        vals.append(np.mean(per_channel))
    return float(np.mean(vals))


@pytest.mark.parametrize("seed,shape", [(0, (2, 3, 40, 56)), (1, (1, 3, 64, 64)), (2, (3, 1, 33, 47))])
def test_ssim_oracle_equals_independent_scipy_implementation(seed, shape):
    g = torch.Generator().manual_seed(seed)
    t = torch.rand(shape, generator=g)
    p = (t + 0.1 * torch.randn(shape, generator=g)).clamp(0, 1)
    assert metrics_oracle.ssim(p, t) == pytest.approx(ssim_scipy(p.numpy(), t.numpy()), abs=1e-9)


def test_psnr_oracle_equals_cv2_psnr_on_unit_range_targets():
    import cv2
    g = torch.Generator().manual_seed(3)
    t = torch.rand((2, 3, 32, 48), generator=g)
    t[0, 0, 0, 0], t[0, 0, 0, 1] = 0.0, 1.0  # data range of the target = 1
    p = (t + 0.05 * torch.randn(t.shape, generator=g)).clamp(0, 1)
    ref = cv2.PSNR(p.numpy().astype(np.float64), t.numpy().astype(np.float64), 1.0)
    assert metrics_oracle.psnr(p, t) == pytest.approx(ref, abs=1e-9)
