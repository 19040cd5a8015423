"""`models.base.BaseModel` — config plumbing behind `models.model.Model` (reference models/base.py:11-55).

Same constructor `(config, dataloader, logger=None)`, same attributes and the same required config keys:
`config[phase].device / dataloader.args.batch_size / model_path / model_name` and `config.train.n_epoch / lr`
(read even in the test phase, as in the reference :18-19)."""
from __future__ import annotations

import os
import time

import torch


class BaseModel:
    def __init__(self, config, dataloader, logger=None):
        self.config, self.dataloader, self.logger = config, dataloader, logger
        self.phase = config["phase"]
        phase_cfg = config[self.phase]
        self.device = phase_cfg["device"]
        self.batch_size = phase_cfg["dataloader"]["args"]["batch_size"]
        self.epoch, self.lr = config["train"]["n_epoch"], config["train"]["lr"]
        self.model_path, self.model_name = phase_cfg["model_path"], phase_cfg["model_name"]
        test_cfg = config.get("test", {}) or {}
        self.is_dataset_paired = bool((test_cfg.get("dataset", {}) or {}).get("is_paired", True))
        self.output_images_path = test_cfg.get("output_images_path", "outputs/")

    def train(self):
        t0 = time.time()
        self.train_step()
        dt = time.time() - t0
        print(f"Training completed in {dt // 60:.0f}m {dt % 60:.0f}s")

    def test(self):
        self.test_step()

    def train_step(self):
        raise NotImplementedError

    def val_step(self):
        raise NotImplementedError

    def save_model(self, model):
        os.makedirs(self.model_path, exist_ok=True)
        torch.save(model.state_dict(), os.path.join(self.model_path, self.model_name))
