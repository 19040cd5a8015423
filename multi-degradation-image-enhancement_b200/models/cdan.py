"""`models.cdan` — drop-in for the reference network module (reference models/cdan.py:8-176).

``CDAN()`` builds the same parameter tree as the reference (236 ``state_dict`` entries with identical keys, shapes
and registration order), so reference checkpoints load with ``strict=True`` and ``state_dict()`` round-trips.
What differs is the execution: in eval mode on a CUDA device ``CDAN.forward`` hands the input to a native plan
(libcdan_b200: hand-written sm_100a kernels, NHWC bf16 or fp32, BatchNorm folded, dense-block concats written
in place, CBAM fused) through the C ABI in include/cdan_b200.h.  There is no cuDNN/ATen/CPU fallback on that
path: a missing library or a CPU tensor in eval mode raises.

Training mode keeps a plain PyTorch (autograd) composition of the same sub-modules so that the reference's
``train_step`` still runs; that path is outside the accelerated hot path.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from models.cbam import CBAM


class ConvBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding)
        self.bn = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.relu(self.bn(self.conv(x)))


class DenseBlock(nn.Module):
    """`num_layers` x [BN -> ReLU -> conv3x3 -> growth_rate channels] on the running concat, then a
    BN -> ReLU -> conv1x1 transition (reference models/cdan.py:22-53)."""

    def __init__(self, in_channels, out_channels, growth_rate, num_layers):
        super().__init__()
        widths = [in_channels + i * growth_rate for i in range(num_layers + 1)]
        self.layers = nn.ModuleList(self._unit(c, growth_rate, 3, 1) for c in widths[:-1])
        self.transition_layer = self._unit(widths[-1], out_channels, 1, 0)

    @staticmethod
    def _unit(cin, cout, k, pad):
        return nn.Sequential(nn.BatchNorm2d(cin), nn.ReLU(inplace=True),
                             nn.Conv2d(cin, cout, kernel_size=k, stride=1, padding=pad))

    def forward(self, x):
        cat = x
        for layer in self.layers:
            cat = torch.cat((cat, layer(cat)), dim=1)
        return self.transition_layer(cat)


class Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = ConvBlock(3, 64)
        self.conv2 = ConvBlock(64, 128)
        self.conv3 = ConvBlock(128, 256)
        self.conv4 = ConvBlock(256, 512)
        self.dense1 = DenseBlock(64, 64, 16, 4)
        self.dense2 = DenseBlock(128, 128, 16, 4)
        self.dense3 = DenseBlock(256, 256, 16, 4)
        self.maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.dp = nn.Dropout(0.2)

    def forward(self, x):
        skips, denses = [], []
        out = x
        for conv, dense in ((self.conv1, self.dense1), (self.conv2, self.dense2), (self.conv3, self.dense3)):
            out = self.maxpool(conv(out))
            denses.append(dense(out))
            out = self.dp(out)
            skips.append(out)
        return self.dp(self.conv4(out)), skips, denses


class Decoder(nn.Module):
    def __init__(self):
        super().__init__()
        # registration order conv_i, cbam_i, bn_i mirrors the reference so state_dict() iterates identically
        self.conv1 = nn.ConvTranspose2d(512, 256, kernel_size=3, stride=1, padding=1)
        self.cbam1 = CBAM(256)
        self.bn1 = nn.BatchNorm2d(256)
        self.conv2 = nn.ConvTranspose2d(256, 128, kernel_size=3, stride=1, padding=1)
        self.cbam2 = CBAM(128)
        self.bn2 = nn.BatchNorm2d(128)
        self.conv3 = nn.ConvTranspose2d(128, 64, kernel_size=3, stride=1, padding=1)
        self.cbam3 = CBAM(64)
        self.bn3 = nn.BatchNorm2d(64)
        self.conv4 = nn.ConvTranspose2d(64, 3, kernel_size=3, stride=1, padding=1)
        self.bn4 = nn.BatchNorm2d(3)
        self.final_dense = DenseBlock(3, 3, 16, 4)
        self.maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        self.dp = nn.Dropout(0.2)
        self.relu = nn.ReLU(inplace=True)
        self.sigmoid = nn.Sigmoid()

    @staticmethod
    def _up(t):
        return F.interpolate(t, scale_factor=2, mode='bilinear', align_corners=False)

    def forward(self, x, out, skip_connections, denses):
        out = self.relu(self.bn1(self.conv1(out))) + skip_connections[2]
        out = self.cbam1(out) * denses[2]
        out = self._up(self.relu(self.bn2(self.conv2(out)))) + skip_connections[1]
        out = self.cbam2(out) * denses[1]
        out = self._up(self.relu(self.bn3(self.conv3(out)))) + skip_connections[0]
        out = self.cbam3(out) * denses[0]
        out = self._up(self.relu(self.bn4(self.conv4(out)))) + x
        return self.sigmoid(self.final_dense(out))


class CDAN(nn.Module):
    """encoder -> CBAM(512) bottleneck -> decoder (reference models/cdan.py:164-176).

    ``compute_dtype``: 'bf16' (default; tcgen05 tensor-core path) or 'fp32' (CUDA-core path, matches the
    reference's fp32 forward to ~1e-5).  Override per process with CDAN_B200_DTYPE, or per instance with
    ``set_compute_dtype``."""

    def __init__(self):
        super().__init__()
        self.encoder = Encoder()
        self.bottleneck = CBAM(512)
        self.decoder = Decoder()
        self._compute_dtype = os.environ.get("CDAN_B200_DTYPE", "bf16")
        self._plan = None
        self._plan_key = None
        self._weights_dirty = True

    # ---- native-plan bookkeeping ---------------------------------------------------------------------------
    def set_compute_dtype(self, dtype: str) -> "CDAN":
        if dtype not in ("bf16", "fp32"):
            raise ValueError("compute dtype must be 'bf16' or 'fp32'")
        if dtype != self._compute_dtype:
            self._compute_dtype = dtype
            self._drop_plan()
        return self

    def refresh_weights(self) -> None:
        """Call after editing parameters/buffers in place so the packed native copy is rebuilt."""
        self._weights_dirty = True

    def _drop_plan(self):
        if self._plan is not None:
            self._plan.close()
        self._plan, self._plan_key, self._weights_dirty = None, None, True

    def train(self, mode: bool = True):
        if mode:
            self._weights_dirty = True  # parameters may change while training
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._weights_dirty = True
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._weights_dirty = True
        return out

    def native_plan(self, device: Optional[torch.device] = None):
        import cdan_b200_native as native
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("cdan_b200: the CDAN forward is implemented for CUDA (sm_100a) only; "
                               "move the module and input to a CUDA device (no CPU fallback)")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, self._compute_dtype)
        if self._plan is None or self._plan_key != key:
            self._drop_plan()
            self._plan = native.Plan(torch.device("cuda", idx), self._compute_dtype)
            self._plan_key = key
        if self._weights_dirty:
            self._plan.load_state_dict(self.state_dict())
            self._weights_dirty = False
        return self._plan

    # ---- forward ---------------------------------------------------------------------------------------------
    def forward(self, x):
        if self.training:
            return self._forward_autograd(x)
        if not x.is_cuda:
            raise RuntimeError("cdan_b200: eval-mode CDAN.forward needs a CUDA tensor (no CPU fallback); "
                               "the CPU oracle lives in oracle/ and is test infrastructure only")
        return self.native_plan(x.device).forward(x)

    def _forward_autograd(self, x):
        out, skips, denses = self.encoder(x)
        out = self.bottleneck(out)
        return self.decoder(x, out, skips, denses)
