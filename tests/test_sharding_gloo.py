"""Multi-GPU path without a cluster: world_size-2 gloo processes shard a batch, run their shard through the CPU
oracle (standing in for the per-GPU plan) and gather — the result must equal the unsharded forward bitwise-per-image."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from batch_sharding import all_shards, shard_range


def test_shard_ranges_cover_batch_exactly():
    for n in (0, 1, 5, 32, 33):
        for world in (1, 2, 3, 8):
            shards = all_shards(n, world)
            assert shards[0][0] == 0 and shards[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
            sizes = [e - s for s, e in shards]
            assert max(sizes) - min(sizes) <= 1
    assert all_shards(32, 8) == [(4 * r, 4 * r + 4) for r in range(8)]  # BASELINE C3: 32 images on 8 GPUs
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "multi-degradation-image-enhancement_b200"), root]
    from oracle.cdan_oracle import cdan_forward
    from oracle.stress_init import ramp_input, stress_state_dict
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    x = ramp_input(3, 16, 24, seed=5)  # odd batch -> uneven shards
    s, e = shard_range(x.shape[0], rank, world)
    y_local = cdan_forward(stress_state_dict(7), x[s:e])
    sizes = [b - a for a, b in all_shards(x.shape[0], world)]
    padded = torch.zeros((max(sizes),) + tuple(y_local.shape[1:]))
    padded[: e - s] = y_local
    gathered = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded)
    if rank == 0:
        y = torch.cat([g[:n] for g, n in zip(gathered, sizes)])
        torch.save({"sharded": y, "full": cdan_forward(stress_state_dict(7), x)}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_batch_sharding_matches_unsharded(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "out.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["sharded"].shape == r["full"].shape
    assert (r["sharded"] - r["full"]).abs().max() < 1e-6  # per-sample independence (reference: 1.2e-7)
