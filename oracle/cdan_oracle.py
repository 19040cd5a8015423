"""CPU oracle for the CDAN forward hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  The product (``multi-degradation-image-enhancement_b200/``) never does: on a
CUDA device it calls the sm_100a kernels through the C-ABI library and fails loudly if that is missing.

What this is: a from-scratch *functional* restatement (torch.nn.functional on CPU, fp32 or fp64) of the
reference's eval-mode forward, driven directly by a ``state_dict`` (the 236-key weight-file format), with no
``nn.Module`` of the reference involved.  Each function cites the reference lines it follows
(paths relative to the upstream repository root).

Parity pinning: the upstream repository ships no tests, golden vectors or weights.  The oracle is therefore
pinned against OUTPUTS OF THE REFERENCE ITSELF: ``oracle/make_golden.py`` imports the unmodified reference
(`models/cdan.py`, `utils/post_processing.py`) in the build container, runs it on seeded inputs under the
seeded stress initialisation of ``oracle/stress_init.py`` and stores per-stage tensors in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this oracle against them (and against the live
reference when it is importable).  PSNR/SSIM are third-party arithmetic (torchmetrics, unpinned, absent) ->
see ``oracle/metrics_oracle.py`` ("parity unpinned").
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default (models/cdan.py:12,43,50) and explicit in models/cbam.py:11


def _bn(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """Eval-mode BatchNorm2d: running statistics, affine (models/cdan.py:12)."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    m, v = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    s = w / torch.sqrt(v + BN_EPS)
    return x * s.view(1, -1, 1, 1) + (b - m * s).view(1, -1, 1, 1)


def conv_block(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """ConvBlock.forward: relu(bn(conv3x3(x)))  (models/cdan.py:15-19)."""
    y = F.conv2d(x, sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"], stride=1, padding=1)
    return F.relu(_bn(sd, prefix + ".bn", y))


def dense_block(sd, prefix: str, x: torch.Tensor, num_layers: int = 4) -> torch.Tensor:
    """DenseBlock.forward (models/cdan.py:32-39): layers are BN -> ReLU -> conv3x3 on the running concat
    (models/cdan.py:41-46); transition is BN -> ReLU -> conv1x1 (models/cdan.py:48-53)."""
    feats = [x]
    for l in range(num_layers):
        cat = torch.cat(feats, dim=1)
        a = F.relu(_bn(sd, f"{prefix}.layers.{l}.0", cat))
        feats.append(F.conv2d(a, sd[f"{prefix}.layers.{l}.2.weight"], sd[f"{prefix}.layers.{l}.2.bias"], padding=1))
    cat = torch.cat(feats, dim=1)
    a = F.relu(_bn(sd, f"{prefix}.transition_layer.0", cat))
    return F.conv2d(a, sd[f"{prefix}.transition_layer.2.weight"], sd[f"{prefix}.transition_layer.2.bias"])


def channel_gate(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """ChannelGate.forward with pool_types ['avg','max'] (models/cbam.py:37-60): the shared MLP is applied to
    the global average and the global max; the two results are summed (so the 2nd Linear's bias counts twice)."""
    w1, b1 = sd[prefix + ".mlp.1.weight"], sd[prefix + ".mlp.1.bias"]
    w2, b2 = sd[prefix + ".mlp.3.weight"], sd[prefix + ".mlp.3.bias"]

    def mlp(v):
        return F.linear(F.relu(F.linear(v, w1, b1)), w2, b2)

    avg = x.mean(dim=(2, 3))
    mx = x.amax(dim=(2, 3))
    att = mlp(avg) + mlp(mx)
    return x * torch.sigmoid(att)[:, :, None, None]


def spatial_gate(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """SpatialGate.forward (models/cbam.py:78-82): ChannelPool = cat(max_c, mean_c) (models/cbam.py:70),
    conv7x7 pad 3 no bias -> BN(1) -> no ReLU (models/cbam.py:77), sigmoid gate."""
    comp = torch.cat([x.amax(dim=1, keepdim=True), x.mean(dim=1, keepdim=True)], dim=1)
    s = F.conv2d(comp, sd[prefix + ".spatial.conv.weight"], None, padding=3)
    s = _bn(sd, prefix + ".spatial.bn", s)
    return x * torch.sigmoid(s)


def cbam(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """CBAM.forward (models/cbam.py:91-95)."""
    return spatial_gate(sd, prefix + ".SpatialGate", channel_gate(sd, prefix + ".ChannelGate", x))


def conv_transpose3x3(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """ConvTranspose2d(k=3, s=1, p=1) (models/cdan.py:103,107,111,115).  Restated as the equivalent direct
    correlation: y[o] = sum_i sum_{u,v} x[i, p+1-u, q+1-v] W[i,o,u,v] + b[o]  <=>  conv2d with
    W'[o,i,u,v] = W[i,o,2-u,2-v], padding 1."""
    w = sd[prefix + ".weight"]  # [Cin, Cout, 3, 3]
    w2 = w.flip(2, 3).permute(1, 0, 2, 3).contiguous()
    return F.conv2d(x, w2, sd[prefix + ".bias"], padding=1)


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    """F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) (models/cdan.py:137,145,153),
    restated explicitly: even outputs 0.25*in[i-1] + 0.75*in[i], odd 0.75*in[i] + 0.25*in[i+1], indices clamped."""

    def up1(t, dim):
        n = t.shape[dim]
        idx = torch.arange(n)
        prev = t.index_select(dim, (idx - 1).clamp(min=0))
        nxt = t.index_select(dim, (idx + 1).clamp(max=n - 1))
        even = 0.25 * prev + 0.75 * t
        odd = 0.75 * t + 0.25 * nxt
        out = torch.stack([even, odd], dim=dim + 1)
        shape = list(t.shape)
        shape[dim] = 2 * n
        return out.reshape(shape)

    return up1(up1(x, 2), 3)


def cdan_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, dtype: torch.dtype = torch.float32,
                 return_stages: bool = False):
    """CDAN.forward in eval mode (models/cdan.py:171-176; Encoder :70-98; Decoder :126-159).
    Dropout is the identity in eval.  Returns the output and, optionally, an OrderedDict of stage tensors
    named like the reference submodules (the per-stage parity tests hook the same names)."""
    sd = {k: v.detach().to("cpu", dtype) for k, v in sd.items() if v.is_floating_point()}
    x = x.detach().to("cpu", dtype)
    if x.shape[2] % 8 or x.shape[3] % 8:
        raise RuntimeError("CDAN needs H and W divisible by 8 (three 2x2 max-pools, models/cdan.py:75-89)")
    st: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    # ---- Encoder (models/cdan.py:70-98)
    c1 = conv_block(sd, "encoder.conv1", x)
    st["encoder.conv1"] = c1
    out1 = F.max_pool2d(c1, 2, 2)
    d1 = dense_block(sd, "encoder.dense1", out1)
    st["encoder.dense1"] = d1
    c2 = conv_block(sd, "encoder.conv2", out1)
    st["encoder.conv2"] = c2
    out2 = F.max_pool2d(c2, 2, 2)
    d2 = dense_block(sd, "encoder.dense2", out2)
    st["encoder.dense2"] = d2
    c3 = conv_block(sd, "encoder.conv3", out2)
    st["encoder.conv3"] = c3
    out3 = F.max_pool2d(c3, 2, 2)
    d3 = dense_block(sd, "encoder.dense3", out3)
    st["encoder.dense3"] = d3
    c4 = conv_block(sd, "encoder.conv4", out3)
    st["encoder.conv4"] = c4

    # ---- bottleneck (models/cdan.py:173)
    b = cbam(sd, "bottleneck", c4)
    st["bottleneck"] = b

    # ---- Decoder (models/cdan.py:126-159)
    o = F.relu(_bn(sd, "decoder.bn1", conv_transpose3x3(sd, "decoder.conv1", b)))
    st["decoder.bn1"] = o
    o = cbam(sd, "decoder.cbam1", o + out3)
    st["decoder.cbam1"] = o
    o = o * d3
    o = F.relu(_bn(sd, "decoder.bn2", conv_transpose3x3(sd, "decoder.conv2", o)))
    st["decoder.bn2"] = o
    o = cbam(sd, "decoder.cbam2", upsample2x(o) + out2)
    st["decoder.cbam2"] = o
    o = o * d2
    o = F.relu(_bn(sd, "decoder.bn3", conv_transpose3x3(sd, "decoder.conv3", o)))
    st["decoder.bn3"] = o
    o = cbam(sd, "decoder.cbam3", upsample2x(o) + out1)
    st["decoder.cbam3"] = o
    o = o * d1
    o = F.relu(_bn(sd, "decoder.bn4", conv_transpose3x3(sd, "decoder.conv4", o)))
    st["decoder.bn4"] = o
    o = upsample2x(o) + x
    fd = dense_block(sd, "decoder.final_dense", o)
    st["decoder.final_dense"] = fd
    y = torch.sigmoid(fd)
    st["output"] = y
    return (y, st) if return_stages else y


# --------------------------------------------------------------------------------------------------------------
# Post-processing tail (utils/post_processing.py), restated.  Each op first rescales /255 when max > 1.
def _maybe_unit(images: torch.Tensor) -> torch.Tensor:
    return images / 255.0 if float(images.max()) > 1.0 else images  # utils/post_processing.py:9,22,37,62


def enhance_contrast(images: torch.Tensor, contrast_factor: float = 1.1) -> torch.Tensor:
    """utils/post_processing.py:5-15."""
    images = _maybe_unit(images)
    mean = images.mean(dim=(2, 3), keepdim=True)
    return ((images - mean) * contrast_factor + mean).clamp(0.0, 1.0)


def enhance_color(images: torch.Tensor, saturation_factor: float = 1.1) -> torch.Tensor:
    """utils/post_processing.py:18-30."""
    images = _maybe_unit(images)
    g = (0.2989 * images[:, 0] + 0.5870 * images[:, 1] + 0.1140 * images[:, 2]).unsqueeze(1)
    return (g + saturation_factor * (images - g)).clamp(0.0, 1.0)


def sharpen(images: torch.Tensor, strength: float = 0.5) -> torch.Tensor:
    """utils/post_processing.py:33-54.  Note the reference adds torch.eye(3) (the identity MATRIX, :47), i.e.
    +1 on the three diagonal taps, then normalises by the kernel sum (= strength + 3)."""
    images = _maybe_unit(images)
    k = torch.tensor([[0.0, -1.0, 0.0], [-1.0, 5.0, -1.0], [0.0, -1.0, 0.0]], dtype=images.dtype) * strength
    k = k + torch.eye(3, dtype=images.dtype)
    k = k / k.sum()
    c = images.shape[1]
    return F.conv2d(images, k.view(1, 1, 3, 3).repeat(c, 1, 1, 1), padding=1, groups=c).clamp(0.0, 1.0)


def soft_denoise(images: torch.Tensor, sigma: float = 0.2) -> torch.Tensor:
    """utils/post_processing.py:57-77."""
    images = _maybe_unit(images)
    k = torch.tensor([[1.0, 2.0, 1.0], [2.0, 4.0, 2.0], [1.0, 2.0, 1.0]], dtype=images.dtype) / 16.0
    c = images.shape[1]
    blurred = F.conv2d(images, k.view(1, 1, 3, 3).repeat(c, 1, 1, 1), padding=1, groups=c)
    return ((1 - sigma) * images + sigma * blurred).clamp(0.0, 1.0)


POSTPROC_OPS = {
    "enhance_contrast": enhance_contrast,
    "enhance_color": enhance_color,
    "sharpen": sharpen,
    "soft_denoise": soft_denoise,
}


def apply_postprocessing(images: torch.Tensor, pp_cfg: Optional[dict]) -> torch.Tensor:
    """utils/postprocessing_factory.py:19-41."""
    if not pp_cfg or not pp_cfg.get("enabled", False):
        return images
    out = images
    for op in pp_cfg.get("ops", []):
        if op["name"] not in POSTPROC_OPS:
            raise ValueError(f"Unknown post-processing op: {op['name']}")
        out = POSTPROC_OPS[op["name"]](out, **op.get("args", {}))
    return out


def quantize_u8(images: torch.Tensor):
    """Output quantisation of Model._save_batch_outputs (reference models/model.py:80-83): per image
    `img = outputs[i].permute(1, 2, 0).numpy(); img = (img * 255).clip(0, 255).astype(np.uint8)` -> uint8 [N,H,W,3]
    (fp32 product, truncation toward zero)."""
    import numpy as np
    x = images.detach().cpu().float()
    return np.stack([(x[i].permute(1, 2, 0).numpy() * 255).clip(0, 255).astype(np.uint8) for i in range(x.shape[0])])


def conv_flops_per_pixel() -> int:
    """Algorithmic conv FLOPs (2*MAC, unpadded) per input pixel; SURVEY A.1 = 252 770."""
    def c3(ci, co):
        return 2 * 9 * ci * co

    def dense(c, cout):
        return sum(c3(c + 16 * l, 16) for l in range(4)) + 2 * (c + 64) * cout

    total = 0.0
    total += c3(3, 64) + (dense(64, 64) + c3(64, 128)) / 4 + (dense(128, 128) + c3(128, 256)) / 16
    total += (dense(256, 256) + c3(256, 512) + c3(512, 256) + c3(256, 128)) / 64
    total += c3(128, 64) / 16 + c3(64, 3) / 4 + dense(3, 3)
    return int(round(total))
