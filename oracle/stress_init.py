"""Seeded weight / input fixtures for parity tests.  TEST INFRASTRUCTURE ONLY (see cdan_oracle.py header).

Why: under PyTorch's default initialisation the CDAN output is numerically blind to almost the whole network
(SURVEY 4.2): sign-flipping encoder.conv3 moves the output by < 3e-6.  The *stress init* below draws every
tensor of the 236-key state_dict from distributions under which every stage matters (He-scaled convs,
non-trivial BatchNorm running statistics, large ChannelGate MLP weights), so wrong kernels move the output by
> 0.1.  The tensors are generated from the KEY SCHEMA alone (no nn.Module), so the reference, the oracle and the
CUDA build can all be loaded with bit-identical fp32 weights via ``load_state_dict``.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import torch


def cdan_schema() -> "OrderedDict[str, Tuple[str, tuple]]":
    """key -> (kind, shape) for the reference's state_dict (models/cdan.py, models/cbam.py).
    Order follows module registration order so ``list(schema) == list(CDAN().state_dict())``."""
    s: "OrderedDict[str, Tuple[str, tuple]]" = OrderedDict()

    def bn(p, c):
        s[p + ".weight"] = ("bn_w", (c,))
        s[p + ".bias"] = ("bn_b", (c,))
        s[p + ".running_mean"] = ("bn_m", (c,))
        s[p + ".running_var"] = ("bn_v", (c,))
        s[p + ".num_batches_tracked"] = ("nbt", ())

    def conv(p, co, ci, k):
        s[p + ".weight"] = ("conv_w", (co, ci, k, k))
        s[p + ".bias"] = ("conv_b", (co,))

    def dense(p, c, cout):
        for l in range(4):
            bn(f"{p}.layers.{l}.0", c + 16 * l)
            conv(f"{p}.layers.{l}.2", 16, c + 16 * l, 3)
        bn(f"{p}.transition_layer.0", c + 64)
        conv(f"{p}.transition_layer.2", cout, c + 64, 1)

    def cbam(p, c):
        s[p + ".ChannelGate.mlp.1.weight"] = ("lin_w", (c // 16, c))
        s[p + ".ChannelGate.mlp.1.bias"] = ("lin_b", (c // 16,))
        s[p + ".ChannelGate.mlp.3.weight"] = ("lin_w", (c, c // 16))
        s[p + ".ChannelGate.mlp.3.bias"] = ("lin_b", (c,))
        s[p + ".SpatialGate.spatial.conv.weight"] = ("conv_w", (1, 2, 7, 7))
        bn(p + ".SpatialGate.spatial.bn", 1)

    # Encoder.__init__ (models/cdan.py:58-65): conv1..4 then dense1..3
    for i, (ci, co) in enumerate([(3, 64), (64, 128), (128, 256), (256, 512)], start=1):
        conv(f"encoder.conv{i}.conv", co, ci, 3)
        bn(f"encoder.conv{i}.bn", co)
    for i, c in enumerate([64, 128, 256], start=1):
        dense(f"encoder.dense{i}", c, c)
    cbam("bottleneck", 512)
    # Decoder.__init__ (models/cdan.py:103-119): conv_i, cbam_i, bn_i ...
    for i, (ci, co) in enumerate([(512, 256), (256, 128), (128, 64), (64, 3)], start=1):
        s[f"decoder.conv{i}.weight"] = ("convT_w", (ci, co, 3, 3))
        s[f"decoder.conv{i}.bias"] = ("conv_b", (co,))
        if i < 4:
            cbam(f"decoder.cbam{i}", co)
        bn(f"decoder.bn{i}", co)
    dense("decoder.final_dense", 3, 3)
    return s


def stress_state_dict(seed: int = 1234) -> "OrderedDict[str, torch.Tensor]":
    """Stress initialisation (distributions of SURVEY Appendix C), drawn key by key in schema order."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, (kind, shape) in cdan_schema().items():
        if kind == "conv_w":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif kind == "convT_w":
            fan_in = shape[0] * shape[2] * shape[3]  # in_channels * k * k
            t = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif kind == "conv_b":
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_w":
            t = torch.rand(shape, generator=g) + 0.5
        elif kind == "bn_b":
            t = torch.randn(shape, generator=g) * 0.2
        elif kind == "bn_m":
            t = torch.randn(shape, generator=g) * 0.2
        elif kind == "bn_v":
            t = torch.rand(shape, generator=g) + 0.5
        elif kind == "lin_w":
            t = torch.randn(shape, generator=g) * 6.0 / shape[1] ** 0.5
        elif kind == "lin_b":
            t = torch.randn(shape, generator=g) * 0.5
        elif kind == "nbt":
            t = torch.tensor(0, dtype=torch.long)
        else:  # pragma: no cover
            raise AssertionError(kind)
        sd[key] = t
    return sd


def default_state_dict(seed: int = 42) -> "OrderedDict[str, torch.Tensor]":
    """PyTorch-default-like initialisation from the schema (kaiming-uniform(a=sqrt(5)) convs / linears,
    identity BatchNorm) — the north star's "random-init weights" case, without needing any nn.Module."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    last_fan_in = 1
    for key, (kind, shape) in cdan_schema().items():
        if kind in ("conv_w", "convT_w", "lin_w"):
            if kind == "conv_w":
                fan_in = shape[1] * shape[2] * shape[3]
            elif kind == "convT_w":
                fan_in = shape[1] * shape[2] * shape[3]  # torch uses weight.size(1)*k*k for ConvTranspose too
            else:
                fan_in = shape[1]
            last_fan_in = fan_in
            bound = (1.0 / fan_in) ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind in ("conv_b", "lin_b"):
            bound = (1.0 / last_fan_in) ** 0.5
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind in ("bn_w", "bn_v"):
            t = torch.ones(shape)
        elif kind in ("bn_b", "bn_m"):
            t = torch.zeros(shape)
        else:
            t = torch.tensor(0, dtype=torch.long)
        sd[key] = t
    return sd


def ramp_input(n: int, h: int, w: int, seed: int = 7) -> torch.Tensor:
    """Non-stationary input in [0,1]: uniform noise times a dark-top -> bright-bottom ramp, so that the global
    pooling of ChannelGate differs between row tiles."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((n, 3, h, w), generator=g)
    return x * torch.linspace(0.05, 1.0, h).view(1, 1, h, 1)


def uniform_input(n: int, h: int, w: int, seed: int = 42) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand((n, 3, h, w), generator=g)


MUTATIONS: List[Tuple[str, str]] = [
    # (name, description) — applied by tests/test_mutations via mutate_state_dict
    ("enc_conv3_sign", "encoder.conv3.conv.weight *= -1"),
    ("bneck_mlp_zero", "bottleneck.ChannelGate.mlp.1.weight = 0"),
    ("cbam2_7x7_sign", "decoder.cbam2.SpatialGate.spatial.conv.weight *= -1"),
    ("dense2_l2_zero", "encoder.dense2.layers.2.2.weight = 0"),
    ("dec_conv1_noflip", "decoder.conv1.weight spatially flipped (= forgetting the ConvT flip)"),
    ("dense3_l1_mean0", "encoder.dense3.layers.1.0.running_mean = 0"),
]


def mutate_state_dict(sd: Dict[str, torch.Tensor], name: str) -> "OrderedDict[str, torch.Tensor]":
    out = OrderedDict((k, v.clone()) for k, v in sd.items())
    if name == "enc_conv3_sign":
        out["encoder.conv3.conv.weight"] *= -1
    elif name == "bneck_mlp_zero":
        out["bottleneck.ChannelGate.mlp.1.weight"].zero_()
    elif name == "cbam2_7x7_sign":
        out["decoder.cbam2.SpatialGate.spatial.conv.weight"] *= -1
    elif name == "dense2_l2_zero":
        out["encoder.dense2.layers.2.2.weight"].zero_()
    elif name == "dec_conv1_noflip":
        out["decoder.conv1.weight"] = out["decoder.conv1.weight"].flip(2, 3).contiguous()
    elif name == "dense3_l1_mean0":
        out["encoder.dense3.layers.1.0.running_mean"].zero_()
    else:
        raise KeyError(name)
    return out
