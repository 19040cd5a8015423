"""Spatial row tiling on CPU: 2 and 3 gloo ranks each run one band of rows through
oracle/tiled_oracle.py (halo exchanges + ChannelGate all-reduces) — the stitched result must equal the untiled oracle,
and the communication schedule is counted (31 halo exchanges, 8 tiny all-reduces per forward)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_band_rows_follow_the_survey_split():
    from oracle.tiled_oracle import band_rows
    bands = band_rows(2160, 8)  # C5: 270 rows at 1/8 resolution -> 6 bands of 34 units + 2 of 33
    sizes = [b - a for a, b in bands]
    assert bands[0][0] == 0 and bands[-1][1] == 2160 and all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
    assert sorted(sizes) == [264, 264] + [272] * 6 and all(s % 8 == 0 for s in sizes)
    import spatial_tiling
    assert spatial_tiling.band_rows(2160, 8) == bands  # the product-side plan and the oracle agree
    # C5: an interior band of the 4K image receives ~7 MB per forward in bf16, every message <= 0.5 MB
    assert max(e.rows * e.channels * (3840 // e.div) * 2 for e in spatial_tiling.halo_schedule()) == 512 * 480 * 2
    assert 6.5e6 < spatial_tiling.halo_bytes_received(1, 3840, rank=3, world=8) < 7.5e6
    with pytest.raises(ValueError):
        band_rows(36, 2)
    with pytest.raises(ValueError):
        band_rows(56, 3)  # 7 units: a band would be thinner than the SpatialGate's halo at 1/8 resolution


def _worker(rank, world, port, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "multi-degradation-image-enhancement_b200"), root]
    from oracle.cdan_oracle import cdan_forward
    from oracle.stress_init import ramp_input, stress_state_dict
    from oracle.tiled_oracle import TileComm, band_rows, cdan_forward_band
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sd = stress_state_dict(11)
    x = ramp_input(2, 80, 40, seed=3)  # 10 eight-row units: 5+5 for 2 ranks, 4+3+3 for 3 (a band needs >= 3 units)
    bands = band_rows(x.shape[2], world)
    r0, r1 = bands[rank]
    comm = TileComm(rank, world, dist)
    y_band = cdan_forward_band(sd, x[:, :, r0:r1], x.shape[2], comm)
    hmax = max(b - a for a, b in bands)
    padded = torch.zeros((x.shape[0], 3, hmax, x.shape[3]), dtype=y_band.dtype)
    padded[:, :, : r1 - r0] = y_band
    gathered = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded)
    if rank == 0:
        y = torch.cat([g[:, :, : b - a] for g, (a, b) in zip(gathered, bands)], dim=2)
        torch.save({"tiled": y, "full": cdan_forward(sd, x, dtype=torch.float64), "stats": comm.stats}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_tiled_forward_matches_untiled(tmp_path, world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "out.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r = torch.load(out)
    assert r["tiled"].shape == r["full"].shape
    assert (r["tiled"] - r["full"]).abs().max() < 1e-10  # fp64: the schedule is exact, only summation order differs
    # 20 3x3 convolutions (4 ConvBlocks + 16 dense layers; the 1x1 transitions need none) + 4 transposed + 3 bilinear
    # + 4 SpatialGate 7x7 = 31 halo exchanges; 4 ChannelGates x (SUM, MAX) = 8 all-reduces
    assert r["stats"]["halo_exchanges"] == 31 and r["stats"]["allreduces"] == 8
    # the product-side plan (spatial_tiling.halo_schedule) predicts exactly what the executable schedule moved
    from spatial_tiling import CHANNEL_GATE_ALLREDUCES, halo_bytes_received, halo_schedule
    assert len(halo_schedule()) == 31 and 2 * len(CHANNEL_GATE_ALLREDUCES) == 8
    assert r["stats"]["halo_bytes"] == halo_bytes_received(batch=2, width=40, rank=0, world=world, elem_size=8)
