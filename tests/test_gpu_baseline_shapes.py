"""Parity at BASELINE.json's own shapes: one full 1080p image (C3's image size) and the C2 batch
64x3x256x256, compared with the CPU reference — the unmodified reference module from oracle/_ref when the build step
vendored it (oracle/build_ref.py), else the oracle port — not with another plan of this library.

Stated tolerances:
  fp32 plan : max abs <= 1e-3 and PSNR >= 60 dB (north star); under the stress init additionally max abs <= 1e-4
  bf16 plan : SURVEY 4.2's rule — max abs <= 2x torch's own CPU bf16-autocast error of the REFERENCE on the same weights and
              input (tests/golden/bf16_autocast_floor.json) and PSNR >= the autocast PSNR - 3 dB; default init additionally
              max abs <= 5e-3 and PSNR >= 55 dB.  Measured under the stress init (ours / autocast): 1080p 7.9e-2, 53.5 dB /
              0.18, 40.2 dB; 2x64x96 8.0e-2, 47.7 dB / 0.49, 27.8 dB; 8x256x256 0.37, 37.0 dB / 0.36, 38.2 dB — the last
              case is bf16 noise amplified by the stress init's large ChannelGate MLP on small feature maps (the encoder
              stages carry the same ~0.5 % relative error as at 1080p; tools/stage_errors.py), the fp32 plan is at 4e-5.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.build_ref import reference_forward_fn
from oracle.stress_init import default_state_dict, ramp_input, stress_state_dict, uniform_input

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN, "bf16_autocast_floor.json")) as _f:
    FLOOR = json.load(_f)["cases"]


def psnr_db(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 200.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)


def make_net(sd, dtype, device):
    from models.cdan import CDAN
    net = CDAN().set_compute_dtype(dtype)
    net.load_state_dict(sd, strict=True)
    return net.to(device).eval()


def check(y, ref, dtype, init, floor_key):
    err = float((y - ref).abs().max())
    p = psnr_db(y, ref)
    if dtype == "fp32":
        assert err <= (1e-4 if init == "stress" else 1e-3), err
        assert p >= 60.0, p
    else:
        fl = FLOOR[floor_key]
        print(f"  bf16 {floor_key}: max abs {err:.3e} (autocast floor {fl['max_abs']:.3e}), PSNR {p:.1f} dB (autocast {fl['psnr_db']:.1f})")
        assert err <= 2.0 * fl["max_abs"] and (init != "default" or err <= 5e-3), (err, fl)
        assert p >= (55.0 if init == "default" else fl["psnr_db"] - 3.0), (p, fl)
    return err, p


@pytest.mark.parametrize("init", ["default", "stress"])
def test_one_1080p_image_vs_reference(cuda_device, init):
    """C3's image size: 16 column strips x several row segments per image, every kernel form at production geometry."""
    sd = default_state_dict(42) if init == "default" else stress_state_dict(1234)
    x = uniform_input(1, 1080, 1920, seed=42) if init == "default" else ramp_input(1, 1080, 1920, seed=21)
    torch.set_num_threads(os.cpu_count() or 1)
    fwd, kind = reference_forward_fn(sd)
    ref = fwd(x)
    for dtype in ("fp32", "bf16"):
        net = make_net(sd, dtype, cuda_device)
        with torch.no_grad():
            y = net(x.to(cuda_device)).cpu()
        err, p = check(y, ref, dtype, init, f"{init}_1x1080x1920")
        print(f"1080p[{init},{dtype}] vs {kind}: max abs {err:.3e}, PSNR {p:.1f} dB")
        del net
        torch.cuda.empty_cache()


@pytest.mark.parametrize("init", ["default", "stress"])
def test_c2_batch_64x256x256_vs_reference(cuda_device, init):
    """BASELINE config 2: 64x3x256x256.  The first 8 images are compared with the CPU reference; the remaining ones
    through batch independence (batched == per-sample, bitwise), which the first 8 also satisfy."""
    sd = default_state_dict(42) if init == "default" else stress_state_dict(1234)
    x8 = uniform_input(8, 256, 256, seed=42) if init == "default" else ramp_input(8, 256, 256, seed=5)
    g = torch.Generator().manual_seed(99)
    x = torch.cat([x8, torch.rand((56, 3, 256, 256), generator=g)])
    torch.set_num_threads(os.cpu_count() or 1)
    fwd, kind = reference_forward_fn(sd)
    ref = fwd(x8)
    for dtype in ("fp32", "bf16"):
        net = make_net(sd, dtype, cuda_device)
        with torch.no_grad():
            y = net(x.to(cuda_device))
            for i in (0, 7, 8, 37, 63):
                assert torch.equal(y[i:i + 1], net(x[i:i + 1].to(cuda_device))), i
        err, p = check(y[:8].cpu(), ref, dtype, init, f"{init}_8x256x256")
        print(f"C2[{init},{dtype}] vs {kind}: max abs {err:.3e}, PSNR {p:.1f} dB")


def test_small_case_bf16_vs_autocast_floor(cuda_device):
    """The shape of tests/test_gpu_forward.py::test_end_to_end_vs_oracle under the SURVEY 4.2 rule."""
    for init, sd, x in (("default", default_state_dict(42), uniform_input(2, 64, 96, seed=42)),
                        ("stress", stress_state_dict(1234), ramp_input(2, 64, 96, seed=7))):
        fwd, _ = reference_forward_fn(sd)
        ref = fwd(x)
        net = make_net(sd, "bf16", cuda_device)
        with torch.no_grad():
            y = net(x.to(cuda_device)).cpu()
        check(y, ref, "bf16", init, f"{init}_2x64x96")


def test_u8_host_path_matches_explicit_pipeline(cuda_device):
    """cdan_forward_host_u8 == normalise (u8 * float32(1/255)) -> cdan_forward -> cdan_quantize_u8, bit for bit, for
    ragged sub-batch schedules; and the normalisation equals the reference transform's arithmetic."""
    import cdan_b200_native as native
    sd = stress_state_dict(1234)
    g = torch.Generator().manual_seed(3)
    xu = torch.randint(0, 256, (5, 40, 64, 3), generator=g, dtype=torch.uint8)
    for dtype in ("fp32", "bf16"):
        net = make_net(sd, dtype, cuda_device)
        plan = net.native_plan()
        xf = (xu.permute(0, 3, 1, 2).to(torch.float32) * np.float32(1.0 / 255.0)).contiguous()
        with torch.no_grad():
            y = net(xf.to(cuda_device))
        want = native.quantize_u8(y).cpu()
        for chunk in (1, 2, 8):
            plan.set_option("host_chunk", chunk)
            got = plan.forward_host_u8(xu.pin_memory())
            assert torch.equal(got, want), (dtype, chunk)
        plan.set_option("host_chunk", 16)


def test_repeated_full_batch_forward_is_stable(cuda_device):
    """Stress loop: the 32 x 1080p bf16 forward, 50 times, must give the same bits every time
    (pipeline desynchronisations found in round 1 showed up only at batch >= 8 and as rare events)."""
    sd = stress_state_dict(1234)
    free, _ = torch.cuda.mem_get_info(cuda_device)
    n = 32 if free > 60 * 2 ** 30 else 8
    x = ramp_input(n, 1080, 1920, seed=33).to(cuda_device)
    net = make_net(sd, "bf16", cuda_device)
    y = torch.empty_like(x)
    plan = net.native_plan()
    plan.forward(x, out=y)
    first = y.clone()
    for it in range(50):
        plan.forward(x, out=y)
        assert torch.equal(y, first), it
    assert torch.isfinite(y).all()
