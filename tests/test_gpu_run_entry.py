"""`run.py -p test` end to end on the GPU box: synthetic PNG dataset + seeded weights file -> Model.test_step through the
native forward, post-processing and PSNR/SSIM kernels -> PNGs and logger rows; checked against the CPU oracle."""
import argparse
import json
import os

import numpy as np
import pytest
import torch
from PIL import Image

from conftest import PKG
from oracle import cdan_oracle as O
from oracle import metrics_oracle as MO
from oracle.stress_init import stress_state_dict

pytestmark = pytest.mark.gpu


def test_run_py_test_phase(cuda_device, tmp_path, monkeypatch):
    import run
    from utils.parser import parse
    rng = np.random.RandomState(0)
    for sub in ("degraded", "clean"):
        os.makedirs(tmp_path / "data" / sub)
    for i in range(5):
        clean = rng.randint(0, 256, (32, 48, 3)).astype(np.uint8)
        Image.fromarray(clean).save(tmp_path / "data" / "clean" / f"{i}.png")
        Image.fromarray((clean * 0.3).astype(np.uint8)).save(tmp_path / "data" / "degraded" / f"{i}.png")
    os.makedirs(tmp_path / "weights")
    sd = stress_state_dict(21)
    torch.save(sd, tmp_path / "weights" / "CDAN_low_light.pt")
    cfg = parse(argparse.Namespace(config=os.path.join(PKG, "config", "low_light.json"), phase="test"))
    t = cfg["test"]
    t["dataset"]["args"]["input_root"] = str(tmp_path / "data" / "degraded")
    t["dataset"]["args"]["target_root"] = str(tmp_path / "data" / "clean")
    t["dataset"]["args"]["transform"]["ops"][0]["args"] = {"height": 32, "width": 48}
    t["dataloader"]["args"].update(batch_size=2, num_workers=0)
    t["model_path"] = str(tmp_path / "weights")
    t["output_images_path"] = cfg["save_outputs"]["output_dir"] = str(tmp_path / "out")
    cfg["logging"]["root_dir"] = str(tmp_path / "runs")
    monkeypatch.setenv("CDAN_B200_DTYPE", "fp32")
    run.main(cfg)
    outs = sorted(os.listdir(tmp_path / "out"))
    assert len([f for f in outs if f.startswith("raw_")]) == 5 and len([f for f in outs if f.startswith("pp_")]) == 5
    # first image against the oracle (uint8 quantisation -> tolerance 1/255 + fp slack)
    x = torch.from_numpy(np.asarray(Image.open(tmp_path / "data" / "degraded" / "0.png"))).permute(2, 0, 1).float()[None] / 255
    y = O.cdan_forward(sd, x)
    raw = torch.from_numpy(np.asarray(Image.open(tmp_path / "out" / "raw_1.png"))).permute(2, 0, 1).float()[None] / 255
    assert (raw - y).abs().max() < 1.5 / 255
    pp = O.apply_postprocessing(y, cfg["post_processing"])
    got_pp = torch.from_numpy(np.asarray(Image.open(tmp_path / "out" / "pp_1.png"))).permute(2, 0, 1).float()[None] / 255
    assert (got_pp - pp).abs().max() < 1.5 / 255
    run_dirs = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp_path / "runs") for f in fs if f == "test.jsonl"]
    rows = [json.loads(l) for l in open(run_dirs[0])]
    assert {r["stage"] for r in rows} == {"pre", "post"}
    assert all(np.isfinite(r["metric_psnr"]) and 0 < r["metric_ssim"] <= 1 for r in rows)
