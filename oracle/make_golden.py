"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py [--ref /root/reference]

The reference is Python and cannot travel to the GPU box, so its outputs on seeded inputs are committed as
small fixtures.  Weights are not stored: they are regenerated bit-identically from
``oracle.stress_init.stress_state_dict(seed)`` / ``default_state_dict(seed)`` (CPU torch RNG).
Stage tensors are stored as a deterministic strided subsample (<= 4096 values) plus sum / abs-sum / max.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.stress_init import default_state_dict, ramp_input, stress_state_dict, uniform_input  # noqa: E402

STAGES = ["encoder.conv1", "encoder.dense1", "encoder.conv2", "encoder.dense2", "encoder.conv3", "encoder.dense3",
          "encoder.conv4", "bottleneck", "decoder.bn1", "decoder.cbam1", "decoder.bn2", "decoder.cbam2",
          "decoder.bn3", "decoder.cbam3", "decoder.bn4", "decoder.final_dense"]
MAX_SAMPLES = 4096


def subsample_index(numel: int) -> np.ndarray:
    if numel <= MAX_SAMPLES:
        return np.arange(numel)
    return np.linspace(0, numel - 1, MAX_SAMPLES).astype(np.int64)


def import_reference(ref_root: str):
    """Import the reference's modules from `ref_root` without polluting sys.modules for the caller."""
    from oracle.build_ref import load_ref
    return load_ref(ref_root)


def run_reference(cdan_mod, sd, x):
    net = cdan_mod.CDAN()
    net.load_state_dict(sd, strict=True)
    net.eval()
    stages = {}
    hooks = []
    for name in STAGES:
        mod = net.get_submodule(name)
        if name.startswith("decoder.bn"):
            # stage = relu(bn(.)) ; the ReLU is in-place on the BN output (models/cdan.py:129)
            hooks.append(mod.register_forward_hook(lambda m, i, o, n=name: stages.__setitem__(n, torch.relu(o).clone())))
        else:
            hooks.append(mod.register_forward_hook(lambda m, i, o, n=name: stages.__setitem__(n, o.clone())))
    with torch.no_grad():
        y = net(x)
    for h in hooks:
        h.remove()
    stages["output"] = y
    return y, stages


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cdan_mod, pp_mod, ppf_mod = import_reference(args.ref)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    cases = [
        ("cdan_stress_2x32x48", stress_state_dict(1234), ramp_input(2, 32, 48, seed=7)),
        ("cdan_default_1x24x40", default_state_dict(42), uniform_input(1, 24, 40, seed=42)),
    ]
    for name, sd, x in cases:
        y, stages = run_reference(cdan_mod, sd, x)
        blob = {"x": x.numpy(), "y": y.numpy()}
        for sname, t in stages.items():
            flat = t.detach().reshape(-1).double().numpy()
            idx = subsample_index(flat.size)
            blob["stage/" + sname + "/shape"] = np.array(t.shape, dtype=np.int64)
            blob["stage/" + sname + "/sample"] = flat[idx].astype(np.float32)
            blob["stage/" + sname + "/stats"] = np.array([flat.sum(), np.abs(flat).sum(), flat.max(), flat.min()])
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **blob)
        print(name, "out range", float(y.min()), float(y.max()), "std", float(y.std()))

    # torch's own bf16-autocast error of the REFERENCE on the same weights / inputs (the bf16 plan's bound is
    # "<= 2x this floor"); the 1080p rows take a few minutes of CPU time
    import json
    floors = {}
    floor_cases = [("stress_2x64x96", stress_state_dict(1234), ramp_input(2, 64, 96, seed=7)),
                   ("default_2x64x96", default_state_dict(42), uniform_input(2, 64, 96, seed=42)),
                   ("stress_1x1080x1920", stress_state_dict(1234), ramp_input(1, 1080, 1920, seed=21)),
                   ("default_1x1080x1920", default_state_dict(42), uniform_input(1, 1080, 1920, seed=42)),
                   ("stress_8x256x256", stress_state_dict(1234), ramp_input(8, 256, 256, seed=5)),
                   ("default_8x256x256", default_state_dict(42), uniform_input(8, 256, 256, seed=42))]
    for name, sd, x in floor_cases:
        net = cdan_mod.CDAN()
        net.load_state_dict(sd, strict=True)
        net.eval()
        with torch.no_grad():
            y32 = net(x)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                y16 = net(x).float()
        err = (y16 - y32).abs().max().item()
        mse = ((y16.double() - y32.double()) ** 2).mean().item()
        floors[name] = {"max_abs": err, "psnr_db": 10.0 * np.log10(1.0 / mse)}
        print("bf16 autocast floor", name, floors[name])
    with open(os.path.join(out_dir, "bf16_autocast_floor.json"), "w") as fh:
        json.dump({"what": "reference CDAN on CPU: torch.autocast(bfloat16) vs fp32, same weights and inputs "
                           "(oracle/make_golden.py)", "cases": floors}, fh, indent=1)

    # post-processing goldens (utils/post_processing.py through utils/postprocessing_factory.py)
    g = torch.Generator().manual_seed(11)
    img = torch.rand((2, 3, 16, 24), generator=g)
    img255 = (img * 255.0).clone()
    blob = {"img": img.numpy(), "img255": img255.numpy()}
    blob["enhance_contrast_1.03"] = pp_mod.enhance_contrast(img.clone(), 1.03).numpy()
    blob["enhance_color_1.55"] = pp_mod.enhance_color(img.clone(), 1.55).numpy()
    blob["sharpen_0.5"] = pp_mod.sharpen(img.clone(), 0.5).numpy()
    blob["soft_denoise_0.15"] = pp_mod.soft_denoise(img.clone(), 0.15).numpy()
    blob["enhance_contrast_255"] = pp_mod.enhance_contrast(img255.clone(), 1.1).numpy()
    cfg = {"enabled": True, "ops": [{"name": "enhance_contrast", "args": {"contrast_factor": 1.03}},
                                    {"name": "enhance_color", "args": {"saturation_factor": 1.55}}]}
    blob["low_light_chain"] = ppf_mod.apply_postprocessing(img.clone(), cfg).numpy()
    np.savez_compressed(os.path.join(out_dir, "postproc.npz"), **blob)
    print("postproc goldens written")


if __name__ == "__main__":
    main()
