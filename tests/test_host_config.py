"""Config / plugin seam (no GPU): the generated configs and, when present, the reference's own config files parse
through utils.parser and select models.cdan.CDAN via init_obj."""
import argparse
import glob
import os

import pytest
import torch
from PIL import Image

from conftest import PKG, REFERENCE_ROOT


def _configs():
    files = sorted(glob.glob(os.path.join(PKG, "config", "*.json")))
    if os.path.isdir(REFERENCE_ROOT):
        files += sorted(glob.glob(os.path.join(REFERENCE_ROOT, "config", "*.json")))
    return files


def test_all_configs_parse_and_name_the_plugin():
    from utils.parser import NoneDict, define_network, parse
    files = _configs()
    assert len([f for f in files if f.startswith(PKG)]) == 11
    for f in files:
        cfg = parse(argparse.Namespace(config=f, phase="test"))
        assert isinstance(cfg, NoneDict) and cfg["phase"] == "test" and cfg["no_such_key"] is None
        assert cfg["model"]["networks"][0]["name"] == ["models.cdan", "CDAN"]
        assert cfg["test"]["model_name"] == f"CDAN_{cfg['name']}.pt"
        assert cfg["train"]["n_epoch"] and cfg["train"]["lr"]  # BaseModel reads these even in test phase
    cfg = parse(argparse.Namespace(config=os.path.join(PKG, "config", "low_light.json"), phase="test"))
    assert cfg["post_processing"]["enabled"] and [o["name"] for o in cfg["post_processing"]["ops"]] == ["enhance_contrast", "enhance_color"]
    net = define_network(cfg["model"]["networks"][0])
    assert net.__name__ == "CDAN" and len(net.state_dict()) == 236


def test_init_obj_error_behaviour():
    from utils.parser import init_obj
    with pytest.raises(NotImplementedError, match="not recognized"):
        init_obj({"name": ["models.nope", "X"], "args": {}}, init_type="Network")
    fn = init_obj({"name": ["utils.reproducibility", "set_seed_and_cudnn"], "args": {"seed_value": 1}})
    assert fn.__name__ == "set_seed_and_cudnn"


def test_paired_dataset_and_transforms(tmp_path):
    from data.dataset import PairedDataset, UnpairedDataset
    for sub in ("deg", "clean"):
        os.makedirs(tmp_path / sub)
        for i in range(3):
            Image.new("RGB", (40, 30), (10 * i, 100, 200)).save(tmp_path / sub / f"im{i}.png")
    tf = {"backend": "albumentations", "ops": [{"name": "Resize", "args": {"height": 16, "width": 24}},
                                                {"name": "Normalize", "args": {"mean": [0, 0, 0], "std": [1, 1, 1]}},
                                                {"name": "ToTensorV2", "args": {}}]}
    ds = PairedDataset(str(tmp_path / "deg"), str(tmp_path / "clean"), pairing_mode="filename", transform=tf)
    a, b = ds[1]
    assert len(ds) == 3 and a.shape == (3, 16, 24) and a.dtype == torch.float32 and float(a.max()) <= 1.0
    assert torch.equal(a, b)
    assert len(UnpairedDataset(str(tmp_path / "deg"), transform=tf)) == 3
    with pytest.raises(ValueError):
        PairedDataset(str(tmp_path / "deg"), str(tmp_path / "clean"), pairing_mode="bogus")


def test_postprocessing_factory_contract():
    from utils.postprocessing_factory import apply_postprocessing
    x = torch.rand(1, 3, 8, 8)
    assert apply_postprocessing(x, {"enabled": False, "ops": [{"name": "sharpen"}]}) is x
    assert apply_postprocessing(x, None) is x
    with pytest.raises(ValueError, match="Unknown post-processing op"):
        apply_postprocessing(x, {"enabled": True, "ops": [{"name": "nope"}]})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        apply_postprocessing(x, {"enabled": True, "ops": [{"name": "sharpen", "args": {"strength": 0.5}}]})
