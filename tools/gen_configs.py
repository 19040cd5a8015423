#!/usr/bin/env python
"""Generate multi-degradation-image-enhancement_b200/config/<task>.json for the 11 degradation tasks.

The files follow the reference's config schema (same keys, so `utils.parser.parse` + `run.py` consume them and the
reference's own config files work unchanged too); the per-task table below holds only what differs between tasks
(loss terms, post-processing, evaluation flags).  Everything else is shared."""
import json
import os

TASKS = {
    # task: (loss terms, post_processing (enabled, ops), evaluate post-processed)
    "blur": ([("charbonnier", 1.0, {"eps": 0.001}), ("ssim", 0.5, None)], (False, []), False),
    "color_distortion": ([("charbonnier", 1.0, {"eps": 0.001}), ("ssim", 0.4, None)], (False, []), False),
    "high_light": ([("charbonnier", 1.0, {"eps": 0.001}), ("ssim", 0.6, None), ("vgg_perceptual", 0.05, {"layers": 20}),
                    ("lpips", 0.05, {"net": "alex"})], (False, []), False),
    "jpeg": ([("l1", 1.0, None), ("vgg_perceptual", 0.25, {"layers": 20}), ("ssim", 0.5, None),
              ("lpips", 0.5, {"net": "alex"})],
             (False, [("enhance_contrast", {"contrast_factor": 1.03}), ("enhance_color", {"saturation_factor": 1.55})]), False),
    "low_contrast": ([("charbonnier", 1.0, {"eps": 0.001}), ("ssim", 0.4, None)], (False, []), False),
    "low_light": ([("mse", 1.0, None), ("vgg_perceptual", 0.25, {"layers": 20}), ("ssim", 0.5, None),
                   ("lpips", 0.5, {"net": "alex"})],
                  (True, [("enhance_contrast", {"contrast_factor": 1.03}), ("enhance_color", {"saturation_factor": 1.55})]), True),
    "motion_blur": ([("charbonnier", 1.0, {"eps": 0.001}), ("ssim", 0.6, None), ("vgg_perceptual", 0.05, {"layers": 20})],
                    (False, []), False),
    "noise": ([("charbonnier", 1.0, {"eps": 0.001}), ("ssim", 0.5, None)], (False, [("soft_denoise", {"sigma": 0.15})]), False),
    "pixelation": ([("charbonnier", 1.0, {"eps": 0.001}), ("ssim", 0.5, None), ("vgg_perceptual", 0.03, {"layers": 20}),
                    ("gradient_l1", 0.1, {"to_gray": True})], (False, []), False),
    "pixelation_easy": ([("charbonnier", 1.0, {"eps": 0.001}), ("ssim", 0.6, None), ("vgg_perceptual", 0.05, {"layers": 20}),
                         ("lpips", 0.05, {"net": "alex"}), ("gradient_l1", 0.3, {"to_gray": True})], (False, []), False),
    "pixelation_hard": ([("charbonnier", 1.0, {"eps": 0.001}), ("ssim", 0.5, None), ("vgg_perceptual", 0.05, {"layers": 20}),
                         ("lpips", 0.05, {"net": "alex"}), ("gradient_l1", 0.35, {"to_gray": True})], (False, []), False),
}

EVAL_TF = [{"name": "Resize", "args": {"height": 256, "width": 384}},
           {"name": "Normalize", "args": {"mean": [0.0, 0.0, 0.0], "std": [1.0, 1.0, 1.0]}},
           {"name": "ToTensorV2", "args": {}}]
TRAIN_AUG = [{"name": "HorizontalFlip", "args": {"p": 0.5}}, {"name": "VerticalFlip", "args": {"p": 0.15}},
             {"name": "RandomRotate90", "args": {"p": 0.1}}]


def term(name, weight, args):
    t = {"name": name, "weight": weight}
    if args:
        t["args"] = args
    return t


def dataset(task, split, ops, paired_flag):
    d = {"name": ["data.dataset", "PairedDataset"],
         "args": {"input_root": f"../{task}/{split}/degraded", "target_root": f"../{task}/{split}/clean",
                  "pairing_mode": "filename", "transform": {"backend": "albumentations", "ops": ops}}}
    if paired_flag:
        d["is_paired"] = True
    return d


def build(task):
    loss, (pp_on, pp_ops), eval_post = TASKS[task]
    return {
        "name": task, "task": task,
        "model": {"which_model": {"name": ["models.model", "Model"], "args": {}},
                  "networks": [{"name": ["models.cdan", "CDAN"], "args": {}}]},
        "loss": {"enabled": True, "terms": [term(*t) for t in loss]},
        "metrics": {"enabled": True, "items": [{"name": "psnr"}, {"name": "ssim"}, {"name": "lpips", "args": {"net": "alex"}}]},
        "evaluation": {"raw": True, "postprocessed": eval_post},
        "post_processing": {"enabled": pp_on, "ops": [{"name": n, "args": a} for n, a in pp_ops]},
        "save_outputs": {"enabled": True, "output_dir": f"outputs/{task}/", "max_images": 200, "format": "png",
                         "save_raw": True, "save_postprocessed": pp_on, "raw_prefix": "raw_", "post_prefix": "pp_"},
        "logging": {"enabled": True, "root_dir": "runs", "save_config_copy": True,
                    "train": {"log_every_n_batches": 0, "save_csv": True, "save_jsonl": True},
                    "test": {"save_csv": True, "save_jsonl": True},
                    "checkpoints": {"enabled": False, "every_n_epochs": 10}},
        "train": {"device": "cuda", "n_epoch": 80, "lr": 0.001, "dataset": dataset(task, "train", TRAIN_AUG + EVAL_TF, False),
                  "dataloader": {"args": {"batch_size": 16, "shuffle": True, "num_workers": 4}},
                  "model_path": "weights/", "model_name": f"CDAN_{task}.pt"},
        "test": {"device": "cuda", "dataset": dataset(task, "test", EVAL_TF, True),
                 "dataloader": {"args": {"batch_size": 16, "shuffle": False, "num_workers": 4}},
                 "model_path": "weights/", "model_name": f"CDAN_{task}.pt", "output_images_path": f"outputs/{task}/"},
    }


def main():
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                       "multi-degradation-image-enhancement_b200", "config")
    os.makedirs(out, exist_ok=True)
    for task in TASKS:
        with open(os.path.join(out, task + ".json"), "w") as f:
            f.write("// generated by tools/gen_configs.py (reference config schema; '//' comments are legal here)\n")
            json.dump(build(task), f, indent=1)
            f.write("\n")
    print(f"wrote {len(TASKS)} configs to {out}")


if __name__ == "__main__":
    main()
