"""Spatial row tiling of very large images (SURVEY 8(e) "spatial rows", config C5: one 4K image on 8 GPUs).

The CDAN forward is not separable by rows: every 3x3 convolution, bilinear x2 and SpatialGate 7x7 looks one to three rows
across a band boundary and every ChannelGate pools over the whole image (reference models/cdan.py:70-159,
models/cbam.py:37-82).  `halo_schedule()` lists those 31 cross-row operators (the schedule oracle/tiled_oracle.py executes on
CPU, one exchange per operator).  The GPU product (csrc/band.cu, csrc/plan.cu `forward_impl`) does not exchange before every
operator: a band carries `halo` extra rows (default 24) on each interior side, runs the unchanged kernels on the extended
band and lets the wrong rows next to the artificial border ("dirt") grow inward — 3x3: +1 row, 7x7: +3, bilinear x2: 2d+1,
2x2 max-pool: ceil(d/2) — refreshing a tensor's halo rows from the neighbours only when the next operator would push the dirt
into the owned rows.  `refresh_schedule()` restates that bookkeeping on the host (7 refreshes per forward at halo 24) and the
GPU tests check the library's counters against it.  ChannelGate statistics are pooled over the owned rows and all-reduced.

Host drivers: `LocalBandedCDAN` (all bands in this process, one thread per band, in-process transport; runs on one GPU) and
`NcclBandedCDAN` (one process per GPU under torchrun; halo rows travel by ncclSend/ncclRecv over NVLink).
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

DEFAULT_HALO = 24

def band_rows(height: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous bands, boundaries at multiples of 8 rows (the three 2x2 max-pools never straddle a boundary), every
    band at least 24 rows (the SpatialGate's 3-row halo at 1/8 resolution must fit).  2160 rows on 8 GPUs ->
    6 x 272 + 2 x 264."""
    if height % 8:
        raise ValueError("H must be a multiple of 8")
    units = height // 8
    if units < 3 * world:
        raise ValueError("bands must be at least 24 rows tall (3 rows at 1/8 resolution)")
    base, extra = divmod(units, world)
    out, r = [], 0
    for k in range(world):
        n = 8 * (base + (1 if k < extra else 0))
        out.append((r, r + n))
        r += n
    return out


@dataclass(frozen=True)
class Exchange:
    name: str        # the consumer that needs the rows
    div: int         # resolution divisor of the exchanged tensor (rows / width = H / div, W / div)
    rows: int        # halo rows per neighbour
    channels: int
    border: str      # what stands in for the rows outside the image: "zero" or "replicate"


def halo_schedule() -> List[Exchange]:
    """The 31 halo exchanges of one forward, in execution order.  Inside a dense block every feature map travels once,
    raw (block input, then the 16 new channels of layers 0-2); the receiver applies each layer's pre-activation itself."""
    s: List[Exchange] = []

    def dense(prefix, div, cin):
        s.append(Exchange(f"{prefix}.input", div, 1, cin, "zero"))
        s.extend(Exchange(f"{prefix}.layers.{l}.out", div, 1, 16, "zero") for l in range(3))

    s.append(Exchange("encoder.conv1", 1, 1, 3, "zero"))
    dense("encoder.dense1", 2, 64)
    s.append(Exchange("encoder.conv2", 2, 1, 64, "zero"))
    dense("encoder.dense2", 4, 128)
    s.append(Exchange("encoder.conv3", 4, 1, 128, "zero"))
    dense("encoder.dense3", 8, 256)
    s.append(Exchange("encoder.conv4", 8, 1, 256, "zero"))
    s.append(Exchange("bottleneck.SpatialGate", 8, 3, 2, "zero"))
    s.append(Exchange("decoder.conv1", 8, 1, 512, "zero"))
    s.append(Exchange("decoder.cbam1.SpatialGate", 8, 3, 2, "zero"))
    s.append(Exchange("decoder.conv2", 8, 1, 256, "zero"))
    s.append(Exchange("decoder.up2", 8, 1, 128, "replicate"))
    s.append(Exchange("decoder.cbam2.SpatialGate", 4, 3, 2, "zero"))
    s.append(Exchange("decoder.conv3", 4, 1, 128, "zero"))
    s.append(Exchange("decoder.up3", 4, 1, 64, "replicate"))
    s.append(Exchange("decoder.cbam3.SpatialGate", 2, 3, 2, "zero"))
    s.append(Exchange("decoder.conv4", 2, 1, 64, "zero"))
    s.append(Exchange("decoder.up4", 2, 1, 3, "replicate"))
    dense("decoder.final_dense", 1, 3)
    return s


CHANNEL_GATE_ALLREDUCES = [("bottleneck", 512), ("decoder.cbam1", 256), ("decoder.cbam2", 128), ("decoder.cbam3", 64)]
"""Per ChannelGate one SUM and one MAX all-reduce of an fp32 [N, C] tensor (pooled statistics of the OWNED rows only)."""


def halo_bytes_received(batch: int, width: int, rank: int, world: int, elem_size: int = 2) -> int:
    """Bytes one band receives per forward (bf16 activations by default)."""
    neighbours = (1 if rank > 0 else 0) + (1 if rank < world - 1 else 0)
    return sum(neighbours * e.rows * e.channels * (width // e.div) * batch * elem_size for e in halo_schedule())


# ---------------------------------------------------------------------------------------------- the product's schedule
def extended_rows(height: int, world: int, rank: int, halo: int = DEFAULT_HALO) -> Tuple[int, int, int, int]:
    """(owned begin, owned end, extended begin, extended end): band_rows plus `halo` rows on each interior side."""
    if halo % 8 or halo < 24:
        raise ValueError("halo must be a multiple of 8 and at least 24 rows")
    r0, r1 = band_rows(height, world)[rank]
    if world > 1 and min(b - a for a, b in band_rows(height, world)) < halo:
        raise ValueError("bands are thinner than the halo")
    return r0, r1, (r0 - halo if rank > 0 else r0), (r1 + halo if rank < world - 1 else r1)


@dataclass(frozen=True)
class Refresh:
    tensor: str    # which tensor's halo rows are refreshed (before the operator that needed it)
    div: int       # resolution divisor
    rows: int      # halo rows per neighbour = halo // div
    channels: int  # channel stride of the rows that travel (padding included)


def refresh_schedule(halo: int = DEFAULT_HALO, hybrid: bool = True, fused_fd: bool = True) -> List[Refresh]:
    """Host restatement of the dirt bookkeeping in csrc/plan.cu forward_impl: which halo refreshes one forward performs.
    hybrid / fused_fd select the bf16 tensor-core plan's layouts (compact heads + 16-channel planes, fused final dense
    block); False/False is the fp32 plan (NHWC concat buffers, whose rows travel whole)."""
    out: List[Refresh] = []
    D = lambda lvl: halo >> lvl

    class Ten:
        def __init__(self, name, ld, lvl, dirt=0):
            self.name, self.ld, self.lvl, self.dirt = name, ld, lvl, dirt

    def refresh(t):
        out.append(Refresh(t.name, 1 << t.lvl, D(t.lvl), t.ld))
        t.dirt = 0

    def need(t, e):
        if t.dirt + e > D(t.lvl):
            refresh(t)

    def need_up(t):
        if 2 * t.dirt + 1 > D(t.lvl - 1):
            refresh(t)

    def dense(name, cpre, ld, lvl, planar, head_dirt):
        head = Ten(f"{name}.concat" if not planar else f"{name}.input", ld, lvl, head_dirt)
        g = [head] + [Ten(f"{name}.layers.{l}.out", 16, lvl) if planar else head for l in range(4)]
        for l in range(4):
            for t in g[:l + 1]:
                need(t, 1)
            worst = max(t.dirt for t in g[:l + 1])
            g[l + 1].dirt = max(g[l + 1].dirt if g[l + 1] is head else 0, worst + 1)
        return head, max(t.dirt for t in g)

    ld = (64, 128, 256) if hybrid else (128, 192, 320)
    h1, dn1 = dense("encoder.dense1", 64, ld[0], 1, hybrid, 1)
    need(h1, 1)
    h2, dn2 = dense("encoder.dense2", 128, ld[1], 2, hybrid, (h1.dirt + 2) // 2)
    need(h2, 1)
    h3, dn3 = dense("encoder.dense3", 256, ld[2], 3, hybrid, (h2.dirt + 2) // 2)
    need(h3, 1)
    e4 = Ten("encoder.conv4", 512, 3, h3.dirt + 1)
    need(e4, 3)
    b0 = Ten("bottleneck", 512, 3, e4.dirt + 3)
    need(b0, 1)
    a1 = Ten("decoder.add1", 256, 3, max(b0.dirt + 1, h3.dirt))
    need(a1, 3)
    c1 = Ten("decoder.gated1", 256, 3, max(a1.dirt + 3, dn3))
    need(c1, 1)
    t2 = Ten("decoder.bn2", 128, 3, c1.dirt + 1)
    need_up(t2)
    u2 = Ten("decoder.add2", 128, 2, max(2 * t2.dirt + 1, h2.dirt))
    need(u2, 3)
    c2 = Ten("decoder.gated2", 128, 2, max(u2.dirt + 3, dn2))
    need(c2, 1)
    t3 = Ten("decoder.bn3", 64, 2, c2.dirt + 1)
    need_up(t3)
    u3 = Ten("decoder.add3", 64, 1, max(2 * t3.dirt + 1, h1.dirt))
    need(u3, 3)
    c3 = Ten("decoder.gated3", 64, 1, max(u3.dirt + 3, dn1))
    need(c3, 1)
    t4 = Ten("decoder.bn4", 8, 1, c3.dirt + 1)
    if fused_fd:
        if 2 * t4.dirt + 5 > D(0):
            refresh(t4)
        final = 2 * t4.dirt + 5
    else:
        need_up(t4)
        _, final = dense("decoder.final_dense", 16, 16 if hybrid else 128, 0, hybrid, 2 * t4.dirt + 1)
    assert final <= D(0), "the output's dirty zone would reach the owned rows"
    return out


def refresh_bytes_received(batch: int, width: int, rank: int, world: int, halo: int = DEFAULT_HALO, elem_size: int = 2,
                           hybrid: bool = True, fused_fd: bool = True) -> int:
    """Bytes one band receives per forward under refresh_schedule (what cdan_band_stats reports)."""
    neighbours = (1 if rank > 0 else 0) + (1 if rank < world - 1 else 0)
    return sum(neighbours * r.rows * r.channels * (width // r.div) * batch * elem_size
               for r in refresh_schedule(halo, hybrid, fused_fd))


# ---------------------------------------------------------------------------------------------- host drivers
class LocalBandedCDAN:
    """All bands of the row-tiled forward inside this process: one native plan, one CUDA stream and one host thread per
    band, halo rows copied device to device by the library's in-process transport.  Exists so that the tiling schedule is
    exercised on a single GPU; with `devices` it also spreads the bands over several GPUs of one process."""

    def __init__(self, state_dict, nbands: int, dtype: str = "bf16", devices: Optional[Sequence] = None, halo: int = DEFAULT_HALO,
                 options: Optional[Dict[str, int]] = None):
        import torch
        import cdan_b200_native as native
        self.nbands, self.halo = int(nbands), int(halo)
        devices = list(devices) if devices else [torch.device("cuda", torch.cuda.current_device())]
        self.devices = [torch.device(devices[r % len(devices)]) for r in range(self.nbands)]
        self.group = native.BandGroup(self.nbands)
        self.plans = []
        for r, dev in enumerate(self.devices):
            plan = native.Plan(dev, dtype)
            for k, v in (options or {}).items():
                plan.set_option(k, v)
            plan.load_state_dict(state_dict)
            plan.band_attach_local(self.group, r)
            self.plans.append(plan)
        self.streams = [torch.cuda.Stream(device=d) for d in self.devices]

    def forward(self, x):
        """x: fp32 [N,3,H,W] (any device).  Returns the stitched [N,3,H,W] result on x's device."""
        import torch
        import cdan_b200_native as native
        n, _, h, w = x.shape
        rows = [native.band_rows(h, self.nbands, r, self.halo) for r in range(self.nbands)]
        xs = [x[:, :, e0:e1].to(self.devices[r], torch.float32).contiguous() for r, (_, _, e0, e1) in enumerate(rows)]
        ys: List[Optional[torch.Tensor]] = [None] * self.nbands
        errs: List[Optional[BaseException]] = [None] * self.nbands
        for d in set(self.devices):
            torch.cuda.synchronize(d)

        def run(r):
            try:
                with torch.cuda.device(self.devices[r]):
                    ys[r] = self.plans[r].forward_band(xs[r], h, self.halo, stream=self.streams[r].cuda_stream)
                    self.streams[r].synchronize()
            except BaseException as e:  # noqa: BLE001 - reported to the caller below
                errs[r] = e

        threads = [threading.Thread(target=run, args=(r,)) for r in range(self.nbands)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in errs:
            if e is not None:
                raise e
        out = torch.empty_like(x, dtype=torch.float32)
        for r, (r0, r1, e0, _) in enumerate(rows):
            out[:, :, r0:r1] = ys[r][:, :, r0 - e0:r1 - e0].to(x.device)
        return out

    def stats(self) -> List[Dict[str, int]]:
        return [p.band_stats() for p in self.plans]

    def close(self):
        for p in self.plans:
            p.close()
        self.group.close()


class NcclBandedCDAN:
    """One band per process / GPU (torchrun).  torch.distributed is only used to hand the NCCL unique id to every rank;
    the halo rows and ChannelGate statistics travel through the library's own NCCL communicator on the caller's stream."""

    def __init__(self, state_dict, dtype: str = "bf16", device=None, halo: int = DEFAULT_HALO, options: Optional[Dict[str, int]] = None):
        import torch
        import torch.distributed as dist
        import cdan_b200_native as native
        if not dist.is_initialized():
            raise RuntimeError("NcclBandedCDAN needs torch.distributed (one process per GPU, e.g. torchrun)")
        self.rank, self.world, self.halo = dist.get_rank(), dist.get_world_size(), int(halo)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        ident = [native.nccl_unique_id() if self.rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        self.plan = native.Plan(self.device, dtype)
        for k, v in (options or {}).items():
            self.plan.set_option(k, v)
        self.plan.load_state_dict(state_dict)
        self.plan.band_attach_nccl(self.rank, self.world, ident[0])

    def rows(self, height: int) -> Tuple[int, int, int, int]:
        import cdan_b200_native as native
        return native.band_rows(height, self.world, self.rank, self.halo)

    def forward_band(self, x_ext, height: int, out=None):
        """x_ext: this rank's extended rows on the GPU.  Returns y_ext (owned rows valid).  Collective."""
        return self.plan.forward_band(x_ext, height, self.halo, out=out)

    def forward(self, x):
        """x: the full fp32 [N,3,H,W] image (same on every rank, any device).  Returns this rank's owned rows and (r0, r1)."""
        import torch
        r0, r1, e0, e1 = self.rows(x.shape[2])
        y = self.forward_band(x[:, :, e0:e1].to(self.device, torch.float32).contiguous(), x.shape[2])
        return y[:, :, r0 - e0:r1 - e0], (r0, r1)

    def stats(self) -> Dict[str, int]:
        return self.plan.band_stats()

    def close(self):
        self.plan.close()
