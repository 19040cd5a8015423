"""Worker of tests/test_gpu_band.py::test_banded_nccl_two_processes (launched by torchrun, one process per GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-degradation-image-enhancement_b200"))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    import cdan_b200_native as native
    import spatial_tiling as st
    from oracle.stress_init import ramp_input, stress_state_dict
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    sd, x = stress_state_dict(77), ramp_input(1, 192, 256, seed=3)
    msgs = []
    for dtype, tol in (("fp32", 1e-5), ("bf16", 2e-2)):
        plan = native.Plan(dev, dtype)
        plan.load_state_dict(sd)
        y0 = plan.forward(x.to(dev))
        plan.close()
        runner = st.NcclBandedCDAN(sd, dtype, dev)
        y, (r0, r1) = runner.forward(x)
        err = float((y - y0[:, :, r0:r1]).abs().max())
        stats = runner.stats()
        sched = st.refresh_schedule(st.DEFAULT_HALO, hybrid=dtype == "bf16", fused_fd=dtype == "bf16")
        ok = err <= tol and stats["halo_exchanges"] == len(sched)
        errs = [None] * world
        dist.all_gather_object(errs, (ok, err, stats))
        msgs.append((dtype, errs))
        runner.close()
    if rank == 0:
        good = all(e[0] for _, errs in msgs for e in errs)
        with open(sys.argv[1], "w") as f:
            f.write(("ok " if good else "FAIL ") + repr(msgs))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
