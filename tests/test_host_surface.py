"""Host-side drop-in surface (no GPU): state_dict schema, loud failures, C-ABI symbols."""
import ctypes
import os
import re

import pytest
import torch

from conftest import PKG, ROOT
from oracle.stress_init import cdan_schema, stress_state_dict


def test_state_dict_schema_matches_reference_layout():
    from models.cdan import CDAN
    net = CDAN()
    sd = net.state_dict()
    schema = cdan_schema()
    assert list(sd.keys()) == list(schema.keys())
    for k, (_, shape) in schema.items():
        assert tuple(sd[k].shape) == shape, k
    assert sum(p.numel() for p in net.parameters()) == 3585663
    net.load_state_dict(stress_state_dict(5), strict=True)
    back = net.state_dict()
    for k, v in stress_state_dict(5).items():
        assert torch.equal(back[k], v), k


def test_eval_forward_on_cpu_fails_loudly():
    from models.cdan import CDAN
    from models.cbam import CBAM
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CDAN().eval()(torch.rand(1, 3, 16, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CBAM(64).eval()(torch.rand(1, 64, 8, 8))


def test_cbam_constructor_surface():
    from models.cbam import CBAM
    m = CBAM(64, reduction_ratio=16, pool_types=['avg', 'max'], no_spatial=False)
    keys = list(m.state_dict().keys())
    assert keys[:4] == ["ChannelGate.mlp.1.weight", "ChannelGate.mlp.1.bias", "ChannelGate.mlp.3.weight",
                        "ChannelGate.mlp.3.bias"]
    assert "SpatialGate.spatial.conv.weight" in keys and "SpatialGate.spatial.conv.bias" not in keys
    assert not hasattr(CBAM(64, no_spatial=True), "SpatialGate")
    CBAM(64, pool_types=['avg', 'max', 'lp', 'lse']).train()(torch.rand(2, 64, 8, 8))  # dead-code pool types still run


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cdan_b200.h")).read()
    declared = re.findall(r"CDAN_API[^;(]*?\b(cdan_\w+)\s*\(", header)
    assert len(declared) >= 16
    import cdan_b200_native as native
    assert set(declared) == set(native.EXPORTED_SYMBOLS)
    path = native.library_path()
    assert os.path.exists(path), "build csrc/build.sh first (the driver's build() does)"
    handle = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(handle, name), name
    handle.cdan_version.restype = ctypes.c_char_p
    assert b"sm_100a" in handle.cdan_version()


def test_plan_create_without_gpu_reports_error():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cdan_b200_native as native
    h = ctypes.c_void_p()
    rc = native.lib().cdan_plan_create(0, 1, ctypes.byref(h))
    assert rc != 0
    assert b"no CPU fallback" in native.lib().cdan_last_error()
