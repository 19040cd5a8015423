"""Per-kernel parity (GPU): each operator through the C ABI vs the CPU oracle / torch.nn.functional fp32.
Tolerances: fp32 plan <= 1e-4 abs (relative to magnitude); bf16 plan: error of bf16 storage + bf16 MMA operands,
bounded at 2e-2 relative to the tensor's max magnitude."""
import numpy as np
import os
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from oracle import cdan_oracle as O
from oracle import metrics_oracle as MO

pytestmark = pytest.mark.gpu


def _tol(dtype, ref):
    scale = max(1.0, float(ref.abs().max()))
    return (1e-4 if dtype == "fp32" else 2e-2) * scale


def _conv_ref(x, w, b, pre=None, relu=False, pool=False):
    a = x if pre is None else F.relu(x * pre[0].view(1, -1, 1, 1) + pre[1].view(1, -1, 1, 1))
    y = F.conv2d(a, w, b, padding=w.shape[-1] // 2)
    if relu:
        y = F.relu(y)
    return F.max_pool2d(y, 2, 2) if pool else y


CONV_CASES = [
    # (N, Cin, Cout, H, W, ks, pre, relu, pool)
    (2, 3, 64, 16, 24, 3, False, True, True),     # encoder.conv1 shape class
    (1, 64, 128, 16, 16, 3, False, True, True),   # encoder.conv2
    (2, 64, 16, 24, 40, 3, True, False, False),   # dense layer 0
    (1, 80, 16, 8, 8, 3, True, False, False),     # dense layer 1 (Cin not a multiple of 64)
    (1, 112, 16, 16, 8, 3, True, False, False),
    (1, 128, 64, 8, 16, 1, True, False, False),   # transition 1x1
    (1, 19, 16, 16, 16, 3, True, False, False),   # final_dense odd channel counts
    (1, 35, 16, 8, 24, 3, True, False, False),
    (1, 67, 3, 16, 16, 1, True, False, False),    # final transition
    (2, 64, 3, 8, 8, 3, False, True, False),      # decoder.conv4
    (1, 256, 512, 8, 8, 3, False, True, False),   # encoder.conv4
    (1, 512, 256, 5, 7, 3, False, True, False),   # odd spatial extent
    (3, 128, 64, 135 // 5, 16, 3, False, True, False),
    # larger extents: several column strips / row segments of the streaming kernel, ring wrap-around, N passes
    (1, 16, 16, 70, 300, 3, True, False, False),   # final_dense layer 0 class, 3 strips x 3 segments
    (2, 48, 16, 45, 130, 3, True, False, False),
    (1, 304, 16, 40, 24, 3, True, False, False),   # dense3 layer 3 (5 K-chunks)
    (2, 3, 64, 72, 260, 3, False, True, True),     # encoder.conv1: planar fp32 input, K-folded taps, fused pool
    (1, 3, 64, 38, 50, 3, False, True, False),
    (1, 128, 64, 40, 200, 1, True, False, False),  # transition, 2 strips
    (1, 320, 256, 24, 136, 1, True, False, False), # dense3 transition: two 128-channel passes
    (1, 64, 3, 40, 140, 3, False, True, False),    # decoder.conv4
    (1, 64, 128, 36, 250, 3, False, True, True),   # encoder.conv2, several tiles
    (1, 128, 64, 33, 150, 3, False, True, False),  # decoder.conv3
    # CTA-pair tile kernel (cta_group::2): odd tile counts (the last pair computes one tile twice, copy masked), fused pool,
    # two N-passes, two M-blocks per weight stage
    (1, 128, 256, 20, 36, 3, False, True, True),   # encoder.conv3 class: pooled epilogue, 5 x 2 = 10 tiles
    (3, 128, 256, 12, 30, 3, False, True, True),   # 3 images x 3 tiles = 9 tiles (odd)
    (1, 256, 512, 19, 70, 3, False, True, False),  # encoder.conv4 class, two 256-channel passes
    (3, 256, 128, 9, 50, 3, False, True, False),   # decoder.conv2 class (NT = 128: 64 + 64 rows per CTA)
    (1, 512, 256, 13, 31, 3, False, True, False),  # decoder.conv1 class, one ragged tile row
]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d(cuda_device, dtype, case):
    import cdan_b200_native as native
    n, ci, co, h, w, ks, pre, relu, pool = case
    g = torch.Generator().manual_seed(hash(case) % 2**31)
    x = torch.randn((n, ci, h, w), generator=g)
    wt = torch.randn((co, ci, ks, ks), generator=g) * (2.0 / (ci * ks * ks)) ** 0.5
    b = torch.randn((co,), generator=g) * 0.1
    p = (torch.rand((ci,), generator=g) + 0.5, torch.randn((ci,), generator=g) * 0.3) if pre else None
    ref = _conv_ref(x, wt, b, p, relu, pool)
    for impl in ([1] if dtype == "fp32" else [1, 0]):
        got = native.op_conv2d(x.to(cuda_device), wt, b, None if p is None else p[0], None if p is None else p[1],
                               relu=relu, pool=pool, dtype=dtype, impl=impl).cpu()
        assert got.shape == ref.shape
        assert (got - ref).abs().max() < _tol(dtype, ref), f"impl={impl}"


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("c,h,w,with_mul", [(64, 16, 24, True), (128, 8, 8, True), (256, 5, 7, False), (512, 4, 6, False)])
def test_cbam(cuda_device, dtype, c, h, w, with_mul):
    import cdan_b200_native as native
    g = torch.Generator().manual_seed(c + h)
    x = torch.rand((2, c, h, w), generator=g) * torch.linspace(0.1, 2.0, h).view(1, 1, h, 1)
    sd = {
        "p.ChannelGate.mlp.1.weight": torch.randn((c // 16, c), generator=g) * 3.0 / c ** 0.5,
        "p.ChannelGate.mlp.1.bias": torch.randn((c // 16,), generator=g) * 0.5,
        "p.ChannelGate.mlp.3.weight": torch.randn((c, c // 16), generator=g) * 3.0 / (c // 16) ** 0.5,
        "p.ChannelGate.mlp.3.bias": torch.randn((c,), generator=g) * 0.5,
        "p.SpatialGate.spatial.conv.weight": torch.randn((1, 2, 7, 7), generator=g) * 0.2,
        "p.SpatialGate.spatial.bn.weight": torch.tensor([1.3]), "p.SpatialGate.spatial.bn.bias": torch.tensor([-0.2]),
        "p.SpatialGate.spatial.bn.running_mean": torch.tensor([0.1]), "p.SpatialGate.spatial.bn.running_var": torch.tensor([0.7]),
    }
    mul = torch.randn((2, c, h, w), generator=g) if with_mul else None
    ref = O.cbam(sd, "p", x)
    if mul is not None:
        ref = ref * mul
    bn = [sd["p.SpatialGate.spatial.bn." + k].item() for k in ("weight", "bias", "running_mean", "running_var")]
    got = native.op_cbam(x.to(cuda_device), sd["p.ChannelGate.mlp.1.weight"], sd["p.ChannelGate.mlp.1.bias"],
                         sd["p.ChannelGate.mlp.3.weight"], sd["p.ChannelGate.mlp.3.bias"],
                         sd["p.SpatialGate.spatial.conv.weight"], bn, mul=None if mul is None else mul.to(cuda_device),
                         dtype=dtype).cpu()
    assert (got - ref).abs().max() < _tol(dtype, ref)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("up", [True, False])
def test_upsample_add(cuda_device, dtype, up):
    import cdan_b200_native as native
    g = torch.Generator().manual_seed(3)
    a = torch.randn((2, 64, 5, 9), generator=g)
    skip = torch.randn((2, 64, 10, 18) if up else (2, 64, 5, 9), generator=g)
    ref = (O.upsample2x(a) if up else a) + skip
    ref_torch = (F.interpolate(a, scale_factor=2, mode="bilinear", align_corners=False) if up else a) + skip
    assert (ref - ref_torch).abs().max() < 1e-6  # the oracle's explicit bilinear restatement equals F.interpolate
    got = native.op_upsample_add(a.to(cuda_device), skip.to(cuda_device), up=up, dtype=dtype).cpu()
    assert (got - ref).abs().max() < _tol(dtype, ref)


def test_postprocess_ops_match_reference_golden(cuda_device):
    import cdan_b200_native as native
    g = np.load(os.path.join(GOLDEN, "postproc.npz"))
    img = torch.from_numpy(g["img"]).to(cuda_device)
    img255 = torch.from_numpy(g["img255"]).to(cuda_device)
    cases = [("enhance_contrast", 1.03, img, "enhance_contrast_1.03"), ("enhance_color", 1.55, img, "enhance_color_1.55"),
             ("sharpen", 0.5, img, "sharpen_0.5"), ("soft_denoise", 0.15, img, "soft_denoise_0.15"),
             ("enhance_contrast", 1.1, img255, "enhance_contrast_255")]
    for op, arg, x, key in cases:
        got = native.postprocess(x, op, arg).cpu().numpy()
        assert np.abs(got - g[key]).max() < 2e-6, key
    chain = native.postprocess(native.postprocess(img, "enhance_contrast", 1.03), "enhance_color", 1.55).cpu().numpy()
    assert np.abs(chain - g["low_light_chain"]).max() < 2e-6


def test_postprocess_large_matches_oracle(cuda_device):
    import cdan_b200_native as native
    g = torch.Generator().manual_seed(5)
    x = torch.rand((3, 3, 72, 200), generator=g)
    for op, arg, fn in [("enhance_contrast", 1.2, O.enhance_contrast), ("enhance_color", 0.7, O.enhance_color),
                        ("sharpen", 1.5, O.sharpen), ("soft_denoise", 0.4, O.soft_denoise)]:
        got = native.postprocess(x.to(cuda_device), op, arg).cpu()
        assert (got - fn(x, arg)).abs().max() < 3e-6, op


def test_quantize_u8_bit_exact(cuda_device):
    """cdan_quantize_u8 against the reference's `(img * 255).clip(0, 255).astype(uint8)` (models/model.py:80-83): integer
    output, so the bar is bit-exact — including the k/255 grid points, values just around them and out-of-range inputs."""
    import cdan_b200_native as native
    g = torch.Generator().manual_seed(17)
    x = torch.rand((3, 3, 24, 40), generator=g) * 1.2 - 0.1          # some values below 0 and above 1
    grid = torch.arange(256, dtype=torch.float32) / 255.0            # exact grid points and their fp32 neighbours
    edge = torch.cat([grid, torch.nextafter(grid, torch.tensor(2.0)), torch.nextafter(grid, torch.tensor(-1.0))])
    x.view(-1)[:edge.numel()] = edge
    got = native.quantize_u8(x.to(cuda_device)).cpu().numpy()
    want = O.quantize_u8(x)
    assert got.dtype == np.uint8 and got.shape == (3, 24, 40, 3)
    assert np.array_equal(got, want)
    big = torch.rand((2, 3, 1080, 1920), generator=g)                # full-size frame
    assert np.array_equal(native.quantize_u8(big.to(cuda_device)).cpu().numpy(), O.quantize_u8(big))


def test_resize_normalize_u8_bit_exact(cuda_device):
    """cdan_resize_normalize_u8 against oracle/input_oracle.py (itself pinned bit-exactly to cv2.resize INTER_LINEAR):
    integer bilinear -> the float32 outputs must be equal bit for bit.  Down- and up-scaling, odd sizes, batch > 1."""
    import cdan_b200_native as native
    from oracle.input_oracle import network_input
    rng = np.random.default_rng(23)
    for (hs, ws), (hd, wd), n in [((301, 517), (256, 384), 2), ((37, 53), (256, 384), 1), ((1080, 1920), (256, 384), 1),
                                  ((8, 8), (24, 40), 3), ((256, 384), (256, 384), 1)]:
        img = rng.integers(0, 256, (n, hs, ws, 3), dtype=np.uint8)
        got = native.resize_normalize_u8(torch.from_numpy(img).to(cuda_device), (hd, wd)).cpu().numpy()
        want = np.stack([network_input(img[i], (hd, wd)) for i in range(n)])
        assert got.shape == (n, 3, hd, wd) and got.dtype == np.float32
        assert np.array_equal(got, want), ((hs, ws), (hd, wd))


def test_psnr_ssim_match_oracle(cuda_device):
    import cdan_b200_native as native
    g = torch.Generator().manual_seed(9)
    t = torch.rand((2, 3, 40, 56), generator=g)
    p = (t + 0.1 * torch.randn((2, 3, 40, 56), generator=g)).clamp(0, 1)
    psnr, ssim = native.psnr_ssim(p.to(cuda_device), t.to(cuda_device))
    assert abs(psnr - MO.psnr(p, t)) < 1e-3
    assert abs(ssim - MO.ssim(p, t)) < 1e-4
    psnr_same, ssim_same = native.psnr_ssim(t.to(cuda_device), (t * 0.5).to(cuda_device))
    assert abs(psnr_same - MO.psnr(t, t * 0.5)) < 1e-3 and abs(ssim_same - MO.ssim(t, t * 0.5)) < 1e-4
