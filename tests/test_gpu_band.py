"""Row-tiled forward on the GPU (SURVEY 8(e) "spatial rows", BASELINE config C5): the banded forward of the library
(csrc/band.cu, csrc/plan.cu) against the untiled forward of the same plan type and against the CPU oracle.

Stated tolerances: fp32 plan, banded vs untiled <= 1e-5 (the only difference is the association of the ChannelGate sums:
per band, then over bands); banded fp32 vs CPU oracle <= 1e-4 (the untiled plan's own bound under the stress init);
bf16 plan, banded vs untiled <= 2e-2 (a ChannelGate sum that moves by an ulp can flip bf16 roundings downstream) and vs the
oracle within the untiled bf16 plan's bound.  The in-process transport runs all bands on ONE GPU (one thread per band); the
NCCL transport is tested with one process per GPU when the box has at least two.
"""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT
from oracle.cdan_oracle import cdan_forward
from oracle.stress_init import ramp_input, stress_state_dict

pytestmark = pytest.mark.gpu


def _untiled(sd, dtype, x, device, options=None):
    import cdan_b200_native as native
    plan = native.Plan(device, dtype)
    for k, v in (options or {}).items():
        plan.set_option(k, v)
    plan.load_state_dict(sd)
    y = plan.forward(x.to(device)).cpu()
    plan.close()
    return y


@pytest.mark.parametrize("nbands,halo,h,w", [(2, 24, 96, 64), (3, 24, 200, 136), (4, 32, 256, 72), (2, 40, 160, 64)])
def test_banded_fp32_equals_untiled_and_oracle(cuda_device, nbands, halo, h, w):
    import spatial_tiling as st
    sd, x = stress_state_dict(77), ramp_input(2, h, w, seed=3)
    ref = cdan_forward(sd, x)
    y0 = _untiled(sd, "fp32", x, cuda_device)
    runner = st.LocalBandedCDAN(sd, nbands, "fp32", [cuda_device], halo=halo)
    y = runner.forward(x)
    stats = runner.stats()
    runner.close()
    assert float((y - y0).abs().max()) <= 1e-5
    assert float((y - ref).abs().max()) <= 1e-4
    # the library's counters equal the host restatement of its schedule
    sched = st.refresh_schedule(halo, hybrid=False, fused_fd=False)
    for r, s in enumerate(stats):
        assert s["halo_exchanges"] == len(sched)
        assert s["allreduces"] == 2 * len(st.CHANNEL_GATE_ALLREDUCES)
        assert s["halo_bytes_received"] == st.refresh_bytes_received(2, w, r, nbands, halo, 4, hybrid=False, fused_fd=False)


@pytest.mark.parametrize("fused", [1, 0])
def test_banded_bf16_tensor_core_plan(cuda_device, fused):
    """The throughput plan (tcgen05 kernels, hybrid concat buffers, fused or layer-by-layer final dense block) in bands."""
    import spatial_tiling as st
    nbands, halo, h, w = 3, 24, 264, 256
    sd, x = stress_state_dict(1234), ramp_input(1, h, w, seed=9)
    ref = cdan_forward(sd, x)
    y0 = _untiled(sd, "bf16", x, cuda_device, {"fd_fused": fused})
    runner = st.LocalBandedCDAN(sd, nbands, "bf16", [cuda_device], halo=halo, options={"fd_fused": fused})
    y = runner.forward(x)
    y_again = runner.forward(x)
    stats = runner.stats()
    runner.close()
    assert torch.equal(y, y_again)  # repeatable bit for bit
    e_tiled, e_untiled = float((y - ref).abs().max()), float((y0 - ref).abs().max())
    print(f"bf16 banded vs untiled {float((y - y0).abs().max()):.3e}; vs oracle {e_tiled:.3e} (untiled {e_untiled:.3e})")
    assert float((y - y0).abs().max()) <= 2e-2
    assert e_tiled <= max(1.5 * e_untiled, 5e-2)
    sched = st.refresh_schedule(halo, hybrid=True, fused_fd=bool(fused))
    assert [s["halo_exchanges"] for s in stats] == [len(sched)] * nbands
    assert stats[1]["halo_bytes_received"] == st.refresh_bytes_received(1, w, 1, nbands, halo, 2, True, bool(fused))


def test_single_band_is_the_untiled_forward(cuda_device):
    import spatial_tiling as st
    sd, x = stress_state_dict(5), ramp_input(1, 64, 64, seed=1)
    y0 = _untiled(sd, "fp32", x, cuda_device)
    runner = st.LocalBandedCDAN(sd, 1, "fp32", [cuda_device])
    y = runner.forward(x)
    runner.close()
    assert float((y - y0).abs().max()) <= 1e-6


def test_band_errors(cuda_device):
    import cdan_b200_native as native
    import spatial_tiling as st
    sd = stress_state_dict(5)
    plan = native.Plan(cuda_device, "fp32")
    plan.load_state_dict(sd)
    with pytest.raises(RuntimeError, match="no band transport"):
        plan.forward_band(torch.zeros(1, 3, 64, 64, device=cuda_device), 64, 24)
    plan.close()
    with pytest.raises(RuntimeError, match="thinner than"):
        st.LocalBandedCDAN(sd, 4, "fp32", [cuda_device], halo=24).forward(ramp_input(1, 64, 64, seed=1))


def test_banded_nccl_two_processes(cuda_device, tmp_path):
    """One process per GPU, halo rows by ncclSend/ncclRecv: needs two devices (skipped on a one-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    out = tmp_path / "res.txt"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "band_nccl_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert out.read_text().startswith("ok"), out.read_text()
