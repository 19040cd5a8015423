"""Row-tiled CDAN forward (SURVEY 8(e) "spatial rows", config C5).  TEST INFRASTRUCTURE ONLY — like the rest of oracle/.

One rank owns a band of image rows (band boundaries at multiples of 8, so the three 2x2 max-pools never straddle a
boundary) and runs the whole network on it.  Everything that looks across the boundary is an explicit exchange:

* every 3x3 convolution / 3x3 transposed convolution : 1 halo row of its INPUT at that layer's resolution; at the image
  border the halo is zero (the convolution's own padding).  Inside a dense block every feature map travels once, RAW:
  the block input, then the 16 new channels of layers 0-2; each layer's pre-activation is applied to the received rows
  locally and the out-of-image rows are zeroed after it (the reference pads after the activation, models/cdan.py:41-46);
* every bilinear x2 upsampling (align_corners=False, models/cdan.py:137,145,153): 1 halo row of its input, at the image
  border the edge row is replicated (index clamping);
* every SpatialGate 7x7 convolution (models/cbam.py:72-82): 3 halo rows of the 2-channel pooled map, zero at the border;
* every ChannelGate (models/cbam.py:37-60): the per-(n,c) sum and max over H x W are all-reduced (SUM, MAX) over the
  bands; the mean divides by the FULL image's pixel count.

This file states that schedule executably and counts its communication (`TileComm.stats`); tests/test_tiled_gloo.py runs
it on 2 and 3 gloo ranks and checks the stitched result against the untiled oracle.  It is the parity target for a
future NCCL / peer-memory implementation of C5; the CUDA product does not implement spatial tiling yet.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from oracle.cdan_oracle import _bn


class TileComm:
    """Communication of one band with its vertical neighbours (torch.distributed, any backend with send/recv)."""

    def __init__(self, rank: int, world: int, dist=None):
        self.rank, self.world, self.dist = rank, world, dist
        self.stats = {"halo_exchanges": 0, "halo_bytes": 0, "allreduces": 0}

    # -- halo: returns x extended by h rows on both sides; border rows are zero ('zero') or the edge row ('replicate')
    def halo(self, x: torch.Tensor, h: int, border: str) -> torch.Tensor:
        n, c, rows, w = x.shape
        if rows < h:
            raise RuntimeError(f"band of {rows} rows is thinner than the {h}-row halo")
        top_send, bot_send = x[:, :, :h].contiguous(), x[:, :, rows - h:].contiguous()
        up, down = self.rank - 1, self.rank + 1
        top_recv = torch.empty_like(top_send) if up >= 0 else None
        bot_recv = torch.empty_like(bot_send) if down < self.world else None
        if self.world > 1:
            ops = []
            if up >= 0:
                ops += [self.dist.P2POp(self.dist.isend, top_send, up), self.dist.P2POp(self.dist.irecv, top_recv, up)]
            if down < self.world:
                ops += [self.dist.P2POp(self.dist.isend, bot_send, down), self.dist.P2POp(self.dist.irecv, bot_recv, down)]
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()
            self.stats["halo_exchanges"] += 1
            self.stats["halo_bytes"] += sum(t.numel() * t.element_size() for t in (top_recv, bot_recv) if t is not None)

        def edge(t):
            return torch.zeros_like(t) if border == "zero" else t
        top = top_recv if top_recv is not None else edge(x[:, :, :1].expand(n, c, h, w).contiguous())
        bot = bot_recv if bot_recv is not None else edge(x[:, :, rows - 1:].expand(n, c, h, w).contiguous())
        return torch.cat([top, x, bot], dim=2)

    def zero_border_rows(self, x_ext: torch.Tensor, h: int) -> torch.Tensor:
        """Zero the h halo rows that lie outside the image (first / last band only): padding applied AFTER an activation."""
        if self.rank == 0 or self.rank == self.world - 1:
            x_ext = x_ext.clone()
            if self.rank == 0:
                x_ext[:, :, :h] = 0
            if self.rank == self.world - 1:
                x_ext[:, :, x_ext.shape[2] - h:] = 0
        return x_ext

    def allreduce(self, t: torch.Tensor, op: str) -> torch.Tensor:
        if self.world > 1:
            t = t.clone()
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM if op == "sum" else self.dist.ReduceOp.MAX)
            self.stats["allreduces"] += 1
        return t


def band_rows(height: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous bands of rows, boundaries at multiples of 8 (2160 rows on 8 GPUs -> 6 x 272 + 2 x 264)."""
    if height % 8:
        raise ValueError("H must be a multiple of 8")
    units = height // 8
    if units < 3 * world:  # a band must hold the 3-row halo of the 7x7 SpatialGate at 1/8 resolution
        raise ValueError("bands must be at least 24 rows tall (3 rows at 1/8 resolution)")
    base, extra = divmod(units, world)
    out, r = [], 0
    for k in range(world):
        n = 8 * (base + (1 if k < extra else 0))
        out.append((r, r + n))
        r += n
    return out


def _conv3x3(comm, x, w, b):
    return F.conv2d(comm.halo(x, 1, "zero"), w, b, padding=(0, 1))


def _conv_block(comm, sd, prefix, x):
    y = _conv3x3(comm, x, sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"])
    return F.relu(_bn(sd, prefix + ".bn", y))


def _dense_block(comm, sd, prefix, x, num_layers=4):
    """Every feature map of the concat travels ONCE, raw: the block input when the block starts, each layer's 16 new
    channels when they are produced (the last layer's only feed the 1x1 transition and do not travel).  The receiving
    band applies each layer's own pre-activation to the halo rows itself and zeroes the rows outside the image after it
    (the reference pads after the activation, models/cdan.py:41-46)."""
    ext = [comm.halo(x, 1, "zero")]  # feature maps with their halo rows
    for l in range(num_layers):
        a = F.relu(_bn(sd, f"{prefix}.layers.{l}.0", torch.cat(ext, dim=1)))
        a = comm.zero_border_rows(a, 1)
        y = F.conv2d(a, sd[f"{prefix}.layers.{l}.2.weight"], sd[f"{prefix}.layers.{l}.2.bias"], padding=(0, 1))
        ext.append(comm.halo(y, 1, "zero") if l + 1 < num_layers else F.pad(y, (0, 0, 1, 1)))
    a = F.relu(_bn(sd, f"{prefix}.transition_layer.0", torch.cat([e[:, :, 1:-1] for e in ext], dim=1)))
    return F.conv2d(a, sd[f"{prefix}.transition_layer.2.weight"], sd[f"{prefix}.transition_layer.2.bias"])


def _channel_gate(comm, sd, prefix, x, full_rows):
    w1, b1 = sd[prefix + ".mlp.1.weight"], sd[prefix + ".mlp.1.bias"]
    w2, b2 = sd[prefix + ".mlp.3.weight"], sd[prefix + ".mlp.3.bias"]

    def mlp(v):
        return F.linear(F.relu(F.linear(v, w1, b1)), w2, b2)

    avg = comm.allreduce(x.sum(dim=(2, 3)), "sum") / float(full_rows * x.shape[3])  # owned rows only, full-image mean
    mx = comm.allreduce(x.amax(dim=(2, 3)), "max")
    return x * torch.sigmoid(mlp(avg) + mlp(mx))[:, :, None, None]


def _spatial_gate(comm, sd, prefix, x):
    comp = torch.cat([x.amax(dim=1, keepdim=True), x.mean(dim=1, keepdim=True)], dim=1)
    s = F.conv2d(comm.halo(comp, 3, "zero"), sd[prefix + ".spatial.conv.weight"], None, padding=(0, 3))
    return x * torch.sigmoid(_bn(sd, prefix + ".spatial.bn", s))


def _cbam(comm, sd, prefix, x, full_rows):
    return _spatial_gate(comm, sd, prefix + ".SpatialGate", _channel_gate(comm, sd, prefix + ".ChannelGate", x, full_rows))


def _conv_transpose3x3(comm, sd, prefix, x):
    w2 = sd[prefix + ".weight"].flip(2, 3).permute(1, 0, 2, 3).contiguous()
    return _conv3x3(comm, x, w2, sd[prefix + ".bias"])


def _upsample2x(comm, x):
    """Rows: even 0.25*in[i-1] + 0.75*in[i], odd 0.75*in[i] + 0.25*in[i+1] with the neighbour band's row (or the clamped
    edge row) as in[-1] / in[rows]; columns exactly as oracle.cdan_oracle.upsample2x."""
    e = comm.halo(x, 1, "replicate")
    mid, prev, nxt = e[:, :, 1:-1], e[:, :, :-2], e[:, :, 2:]
    rows = torch.stack([0.25 * prev + 0.75 * mid, 0.75 * mid + 0.25 * nxt], dim=3)
    r = rows.reshape(x.shape[0], x.shape[1], 2 * x.shape[2], x.shape[3])
    n = r.shape[3]
    idx = torch.arange(n)
    pc, nc = r.index_select(3, (idx - 1).clamp(min=0)), r.index_select(3, (idx + 1).clamp(max=n - 1))
    cols = torch.stack([0.25 * pc + 0.75 * r, 0.75 * r + 0.25 * nc], dim=4)
    return cols.reshape(r.shape[0], r.shape[1], r.shape[2], 2 * n)


def cdan_forward_band(sd: Dict[str, torch.Tensor], x_band: torch.Tensor, full_height: int, comm: TileComm,
                      dtype: torch.dtype = torch.float64) -> torch.Tensor:
    """The CDAN eval forward (reference models/cdan.py:171-176) on one band of rows of an image of `full_height` rows."""
    sd = {k: v.detach().to("cpu", dtype) for k, v in sd.items() if v.is_floating_point()}
    x = x_band.detach().to("cpu", dtype)
    if x.shape[2] % 8 or x.shape[3] % 8 or full_height % 8:
        raise RuntimeError("band height, image height and width must be multiples of 8")
    H = full_height
    c1 = _conv_block(comm, sd, "encoder.conv1", x)
    out1 = F.max_pool2d(c1, 2, 2)
    d1 = _dense_block(comm, sd, "encoder.dense1", out1)
    out2 = F.max_pool2d(_conv_block(comm, sd, "encoder.conv2", out1), 2, 2)
    d2 = _dense_block(comm, sd, "encoder.dense2", out2)
    out3 = F.max_pool2d(_conv_block(comm, sd, "encoder.conv3", out2), 2, 2)
    d3 = _dense_block(comm, sd, "encoder.dense3", out3)
    c4 = _conv_block(comm, sd, "encoder.conv4", out3)
    b = _cbam(comm, sd, "bottleneck", c4, H // 8)
    o = F.relu(_bn(sd, "decoder.bn1", _conv_transpose3x3(comm, sd, "decoder.conv1", b)))
    o = _cbam(comm, sd, "decoder.cbam1", o + out3, H // 8) * d3
    o = F.relu(_bn(sd, "decoder.bn2", _conv_transpose3x3(comm, sd, "decoder.conv2", o)))
    o = _cbam(comm, sd, "decoder.cbam2", _upsample2x(comm, o) + out2, H // 4) * d2
    o = F.relu(_bn(sd, "decoder.bn3", _conv_transpose3x3(comm, sd, "decoder.conv3", o)))
    o = _cbam(comm, sd, "decoder.cbam3", _upsample2x(comm, o) + out1, H // 2) * d1
    o = F.relu(_bn(sd, "decoder.bn4", _conv_transpose3x3(comm, sd, "decoder.conv4", o)))
    o = _upsample2x(comm, o) + x
    return torch.sigmoid(_dense_block(comm, sd, "decoder.final_dense", o))
