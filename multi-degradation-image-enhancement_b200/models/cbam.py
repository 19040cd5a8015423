"""`models.cbam` — drop-in for the reference module of the same name (reference models/cbam.py:6-95).

Same classes, constructor signatures and parameter names (so ``state_dict`` keys are identical:
``ChannelGate.mlp.1/3.*``, ``SpatialGate.spatial.conv/bn.*``).  In eval mode on a CUDA device ``CBAM.forward``
runs the fused sm_100a kernels of libcdan_b200 (pool -> MLP gate -> channel compress -> 7x7 gate -> apply);
inside ``CDAN`` the blocks are executed by the network-level plan instead and these forwards are not called.
The plain PyTorch composition below exists only for autograd (training, out of scope of the hot path) and for
the never-selected 'lp' / 'lse' pool types that are part of the constructor surface.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class BasicConv(nn.Module):
    """conv (no bias by default) -> optional BatchNorm(eps 1e-5, momentum 0.01) -> optional ReLU
    (reference models/cbam.py:6-20)."""

    def __init__(self, in_planes, out_planes, kernel_size, stride=1, padding=0, dilation=1, groups=1, relu=True,
                 bn=True, bias=False):
        super().__init__()
        self.out_channels = out_planes
        self.conv = nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, groups=groups, bias=bias)
        self.bn = nn.BatchNorm2d(out_planes, eps=1e-5, momentum=0.01, affine=True) if bn else None
        self.relu = nn.ReLU() if relu else None

    def forward(self, x):
        y = self.conv(x)
        if self.bn is not None:
            y = self.bn(y)
        return y if self.relu is None else self.relu(y)


class Flatten(nn.Module):
    def forward(self, x):
        return x.reshape(x.size(0), -1)


def logsumexp_2d(tensor):
    flat = tensor.reshape(tensor.size(0), tensor.size(1), -1)
    return torch.logsumexp(flat, dim=2, keepdim=True)


class ChannelGate(nn.Module):
    def __init__(self, gate_channels, reduction_ratio=16, pool_types=['avg', 'max']):
        super().__init__()
        self.gate_channels = gate_channels
        hidden = gate_channels // reduction_ratio
        # index 0 is the parameter-free Flatten so the Linear layers keep the keys mlp.1 / mlp.3
        self.mlp = nn.Sequential(Flatten(), nn.Linear(gate_channels, hidden), nn.ReLU(),
                                 nn.Linear(hidden, gate_channels))
        self.pool_types = pool_types

    def _pooled(self, x, kind):
        if kind == 'avg':
            return x.mean(dim=(2, 3), keepdim=True)
        if kind == 'max':
            return x.amax(dim=(2, 3), keepdim=True)
        if kind == 'lp':
            return F.lp_pool2d(x, 2, (x.size(2), x.size(3)), stride=(x.size(2), x.size(3)))
        if kind == 'lse':
            return logsumexp_2d(x)
        raise ValueError(f"unknown pool type {kind!r}")

    def forward(self, x):
        att = sum(self.mlp(self._pooled(x, kind)) for kind in self.pool_types)
        return x * torch.sigmoid(att)[:, :, None, None]


class ChannelPool(nn.Module):
    def forward(self, x):
        return torch.stack((x.amax(dim=1), x.mean(dim=1)), dim=1)  # channel 0 = max, 1 = mean


class SpatialGate(nn.Module):
    def __init__(self):
        super().__init__()
        self.compress = ChannelPool()
        self.spatial = BasicConv(2, 1, 7, stride=1, padding=3, relu=False)

    def forward(self, x):
        return x * torch.sigmoid(self.spatial(self.compress(x)))


class CBAM(nn.Module):
    def __init__(self, gate_channels, reduction_ratio=16, pool_types=['avg', 'max'], no_spatial=False):
        super().__init__()
        self.ChannelGate = ChannelGate(gate_channels, reduction_ratio, pool_types)
        self.no_spatial = no_spatial
        if not no_spatial:
            self.SpatialGate = SpatialGate()

    def _native_ok(self, x) -> bool:
        c = self.ChannelGate.gate_channels
        return (not self.training and x.is_cuda and not self.no_spatial
                and list(self.ChannelGate.pool_types) == ['avg', 'max']
                and c % 64 == 0 and (c & (c - 1)) == 0 and self.ChannelGate.mlp[1].out_features == c // 16)

    def forward(self, x):
        if self._native_ok(x):
            import cdan_b200_native as native
            cg, sg = self.ChannelGate, self.SpatialGate.spatial
            bn = sg.bn
            return native.op_cbam(x, cg.mlp[1].weight, cg.mlp[1].bias, cg.mlp[3].weight, cg.mlp[3].bias,
                                  sg.conv.weight, (bn.weight.item(), bn.bias.item(), bn.running_mean.item(),
                                                   bn.running_var.item()),
                                  dtype="fp32").to(x.dtype)
        if not self.training and not x.is_cuda:
            raise RuntimeError("cdan_b200: eval-mode CBAM runs on CUDA only (no CPU fallback)")
        y = self.ChannelGate(x)
        return y if self.no_spatial else self.SpatialGate(y)
