"""`utils.parser` — JSON config loader and the dynamic-import plugin seam (reference utils/parser.py:10-104).

`init_obj` is the boundary the drop-in sits behind: configs name `["models.cdan", "CDAN"]`, this function imports
that module and instantiates the class.  Unlike the reference (:69-71), the original exception is chained instead of
being swallowed, but the raised type and message are the same."""
from __future__ import annotations

import importlib
import json
from collections import OrderedDict
from functools import partial
from types import FunctionType

from torch.utils.data import DataLoader


class NoneDict(dict):
    """dict whose missing keys read as None (reference :10-12)."""

    def __missing__(self, key):
        return None


def dict_to_nonedict(config):
    if isinstance(config, dict):
        return NoneDict(**{k: dict_to_nonedict(v) for k, v in config.items()})
    if isinstance(config, list):
        return [dict_to_nonedict(v) for v in config]
    return config


def parse(args):
    """Strip `//` comments line by line, parse JSON, record the phase (reference :28-39)."""
    with open(args.config, "r") as f:
        text = "\n".join(line.split("//")[0] for line in f.read().splitlines())
    config = json.loads(text, object_pairs_hook=OrderedDict)
    config["phase"] = args.phase
    return dict_to_nonedict(config)


def init_obj(obj_config, *args, default_file_name="default file", given_module=None, init_type="Network",
             **modify_kwargs):
    name = obj_config["name"]
    file_name, class_name = (name[0], name[1]) if isinstance(name, list) else (default_file_name, name)
    try:
        module = given_module if given_module is not None else importlib.import_module(file_name)
        attr = getattr(module, class_name)
        kwargs = obj_config.get("args", {}) or {}
        kwargs.update(modify_kwargs)
        if isinstance(attr, type):
            obj = attr(*args, **kwargs)
            obj.__name__ = obj.__class__.__name__
        elif isinstance(attr, FunctionType):
            obj = partial(attr, *args, **kwargs)
            obj.__name__ = attr.__name__
        else:
            raise TypeError(f"{class_name} is neither a class nor a function")
    except Exception as exc:
        raise NotImplementedError(f"{init_type} [{class_name}() from {file_name}] not recognized.") from exc
    return obj


def create_model(**cfg_model):
    model_config = cfg_model["config"]["model"]["which_model"]
    model_config["args"].update(cfg_model)
    return init_obj(model_config, default_file_name="models.model", init_type="Model")


def define_network(network_config):
    return init_obj(network_config, default_file_name="models.network", init_type="Network")


def define_dataset(dataset_config):
    return init_obj(dataset_config, default_file_name="data", init_type="Dataset")


def define_dataloader(dataset, dataloader_config):
    return DataLoader(dataset, batch_size=dataloader_config["batch_size"], shuffle=dataloader_config["shuffle"],
                      num_workers=dataloader_config["num_workers"])
