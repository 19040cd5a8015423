"""Pipeline-desynchronisation stress: the warp-specialised kernels hand rows between roles through
mbarrier phases; a phase slip shows up as a wrong or different result only occasionally and only at production geometry.
The forward is therefore looped at BASELINE C3's image size and every output compared bit for bit with the first one
(ring slots and segmentations are functions of the absolute image row, so the result must not depend on timing)."""
import pytest
import torch

from oracle.stress_init import stress_state_dict

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch,loops", [(8, 50), (32, 12), (3, 30)])
def test_repeated_1080p_forward_is_bitwise_stable(cuda_device, batch, loops):
    import cdan_b200_native as native
    plan = native.Plan(cuda_device, "bf16")
    plan.load_state_dict(stress_state_dict(99))
    g = torch.Generator().manual_seed(batch)
    x = torch.rand((batch, 3, 1080, 1920), generator=g).to(cuda_device)
    first = plan.forward(x).clone()
    assert bool(torch.isfinite(first).all())
    y = torch.empty_like(first)
    bad = 0
    for _ in range(loops):
        plan.forward(x, out=y)
        bad += int(not torch.equal(y, first))
    plan.close()
    assert bad == 0, f"{bad} of {loops} repeated forwards differ from the first"
