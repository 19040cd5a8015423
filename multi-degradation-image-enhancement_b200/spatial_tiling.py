"""Host-side plan for spatial row tiling of very large images (SURVEY 8(e) "spatial rows", config C5: one 4K image on
8 GPUs).  The CDAN forward is not separable by rows: every 3x3 convolution, bilinear x2 and SpatialGate 7x7 looks one to
three rows across a band boundary and every ChannelGate pools over the whole image (reference models/cdan.py:70-159,
models/cbam.py:37-82).  This module states WHO owns which rows and WHAT crosses each boundary per forward; the executable
statement of the same schedule (and its parity proof against the untiled forward) is oracle/tiled_oracle.py with
tests/test_tiled_gloo.py.  The CUDA kernels do not implement the exchange yet — this is the contract they
will be built against.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple


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
