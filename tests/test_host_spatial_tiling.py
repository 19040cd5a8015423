"""Host logic of the row-tiled forward (no GPU): the band split of the library's C ABI equals the Python plan, and the dirt
bookkeeping (spatial_tiling.refresh_schedule, a restatement of csrc/plan.cu forward_impl) is sound — checked against a
brute-force receptive-field simulation on a 1-D column of rows."""
import numpy as np
import pytest

import spatial_tiling as st


def test_band_rows_of_the_library_match_the_python_plan():
    import cdan_b200_native as native
    for h, world in [(2160, 8), (1080, 4), (264, 3), (96, 2), (64, 1), (2160, 7)]:
        for halo in (24, 32):
            if world > 1 and min(b - a for a, b in st.band_rows(h, world)) < halo:
                continue
            for r in range(world):
                assert native.band_rows(h, world, r, halo) == st.extended_rows(h, world, r, halo)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        native.band_rows(100, 2, 0, 24)
    with pytest.raises(RuntimeError, match="halo"):
        native.band_rows(96, 2, 0, 16)
    with pytest.raises(RuntimeError, match="thinner"):
        native.band_rows(64, 4, 0, 24)


def test_refresh_schedule_counts():
    s24 = st.refresh_schedule(24)
    assert len(s24) == 7 and [r.rows for r in s24] == [3, 3, 3, 3, 3, 3, 6]
    assert len(st.refresh_schedule(32)) == 5
    assert len(st.halo_schedule()) == 31  # the per-operator schedule of oracle/tiled_oracle.py
    # an interior band of the 4K image receives far less than under the per-operator schedule's 31 latency-bound messages
    assert st.refresh_bytes_received(1, 3840, 1, 8) == sum(2 * r.rows * r.channels * (3840 // r.div) * 2 for r in s24)


def _simulate(halo, hybrid, fused_fd):
    """Brute force: propagate a per-row 'wrong' flag through the operator chain on a band with an artificial border at row
    0 (rows [0, halo) are the halo, row halo.. owned), refreshing exactly where refresh_schedule says; assert that no
    owned row is ever wrong.  Operators act on boolean columns: 3x3 / 7x7 dilate by 1 / 3, pool ORs row pairs, bilinear
    x2 makes out row r wrong if either source row is (or it is clamped at the artificial border)."""
    sched = list(st.refresh_schedule(halo, hybrid, fused_fd))
    taken = []
    rows0 = 8 * halo  # tall enough that the far side never matters

    def fresh(lvl):
        return np.zeros(rows0 >> lvl, bool)

    def dil(v, e):  # the artificial border at row 0 contributes wrong data (zero padding instead of the neighbour's rows)
        out = v.copy()
        for k in range(1, e + 1):
            out[:-k] |= v[k:]
            out[k:] |= v[:-k]
        out[:e] = True
        return out

    def pool(v):
        return v[0::2] | v[1::2]

    def up(v):
        n = len(v)
        out = np.zeros(2 * n, bool)
        for r in range(2 * n):
            src = (r + 0.5) / 2 - 0.5
            i0 = int(np.floor(src))
            out[r] = (i0 < 0) or v[max(i0, 0)] or v[min(i0 + 1, n - 1)]
        return out

    class T:
        def __init__(self, name, lvl, v):
            self.name, self.lvl, self.v = name, lvl, v

    def maybe_refresh(t, names):
        if sched and sched[0].tensor in names:
            r = sched.pop(0)
            taken.append(r.tensor)
            assert not t.v[halo >> t.lvl:].any(), f"{t.name}: owned rows wrong before refresh"
            t.v[:halo >> t.lvl] = False

    def owned_ok(t):
        assert not t.v[halo >> t.lvl:].any(), f"{t.name}: dirt reached the owned rows"

    def dense(name, lvl, head):
        groups = [head]
        for l in range(4):
            for i, g in enumerate(groups):
                maybe_refresh(g, {f"{name}.input" if i == 0 else f"{name}.layers.{i - 1}.out"} if hybrid else {f"{name}.concat"})
                if not hybrid and taken and taken[-1] == f"{name}.concat":
                    for gg in groups:
                        gg.v[:halo >> lvl] = False
            out = T(f"{name}.g{l + 1}", lvl, dil(np.logical_or.reduce([g.v for g in groups]), 1))
            owned_ok(out)
            groups.append(out)
        return head, T(f"{name}.out", lvl, np.logical_or.reduce([g.v for g in groups]))

    x = T("x", 0, fresh(0))
    h1 = T("h1", 1, pool(dil(x.v, 1)))
    h1, dn1 = dense("encoder.dense1", 1, h1)
    maybe_refresh(h1, {"encoder.dense1.input", "encoder.dense1.concat"})
    h2 = T("h2", 2, pool(dil(h1.v, 1)))
    h2, dn2 = dense("encoder.dense2", 2, h2)
    maybe_refresh(h2, {"encoder.dense2.input", "encoder.dense2.concat"})
    h3 = T("h3", 3, pool(dil(h2.v, 1)))
    h3, dn3 = dense("encoder.dense3", 3, h3)
    maybe_refresh(h3, {"encoder.dense3.input", "encoder.dense3.concat"})
    e4 = T("e4", 3, dil(h3.v, 1)); maybe_refresh(e4, {"encoder.conv4"})
    b0 = T("b0", 3, dil(e4.v, 3)); owned_ok(b0); maybe_refresh(b0, {"bottleneck"})
    a1 = T("a1", 3, dil(b0.v, 1) | h3.v); maybe_refresh(a1, {"decoder.add1"})
    c1 = T("c1", 3, dil(a1.v, 3) | dn3.v); owned_ok(c1); maybe_refresh(c1, {"decoder.gated1"})
    t2 = T("t2", 3, dil(c1.v, 1)); maybe_refresh(t2, {"decoder.bn2"})
    u2 = T("u2", 2, up(t2.v) | h2.v); maybe_refresh(u2, {"decoder.add2"})
    c2 = T("c2", 2, dil(u2.v, 3) | dn2.v); owned_ok(c2); maybe_refresh(c2, {"decoder.gated2"})
    t3 = T("t3", 2, dil(c2.v, 1)); maybe_refresh(t3, {"decoder.bn3"})
    u3 = T("u3", 1, up(t3.v) | h1.v); maybe_refresh(u3, {"decoder.add3"})
    c3 = T("c3", 1, dil(u3.v, 3) | dn1.v); owned_ok(c3); maybe_refresh(c3, {"decoder.gated3"})
    t4 = T("t4", 1, dil(c3.v, 1)); maybe_refresh(t4, {"decoder.bn4"})
    f0 = T("f0", 0, up(t4.v))
    if fused_fd:
        v = f0.v
        for _ in range(4):  # every later layer reads all earlier groups: the union grows by one row per layer
            v = v | dil(v, 1)
        out = T("out", 0, v)
    else:
        _, out = dense("decoder.final_dense", 0, f0)
    owned_ok(out)
    assert not sched, f"refreshes never taken: {sched}"


@pytest.mark.parametrize("halo", [24, 32, 40, 48, 64])
@pytest.mark.parametrize("hybrid,fused_fd", [(True, True), (True, False), (False, False)])
def test_refresh_schedule_keeps_owned_rows_clean(halo, hybrid, fused_fd):
    _simulate(halo, hybrid, fused_fd)
