"""Property tests (hypothesis): for arbitrary small scenes — clustered, coincident, out of the
box, zero-mass, any θ — the engine's algorithm core (executed on the CPU by tests/emul/) makes
exactly the reference's decisions and builds exactly the reference's tree."""
import ctypes as C

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import make_engine
from test_core_emulation import _dp, emul_acc, emul_tree, I64

POINT = st.tuples(st.floats(-50.0, 2450.0, allow_nan=False, width=64), st.floats(-850.0, 1650.0, allow_nan=False, width=64))


@st.composite
def scenes_strategy(draw):
    pts = draw(st.lists(POINT, min_size=0, max_size=60))
    # clusters around a few anchors: bodies a few 1e-5 apart (jitter regime) or exactly coincident
    for ax, ay in draw(st.lists(POINT, min_size=0, max_size=3)):
        k = draw(st.integers(2, 5))
        for _ in range(k):
            dx = draw(st.sampled_from([0.0, 1e-5, -2e-5, 3e-4, 1e-3]))
            dy = draw(st.sampled_from([0.0, -1e-5, 2e-5, 2e-4]))
            pts.append((ax + dx, ay + dy))
    n = len(pts)
    masses = draw(st.lists(st.sampled_from([0.5, 1.0, 5000.0, 0.0, 3.25]), min_size=n, max_size=n))
    x = np.array([p[0] for p in pts], float)
    y = np.array([p[1] for p in pts], float)
    theta = draw(st.sampled_from([0.0, 0.2, 0.5, 1.0, 1.6]))
    return (x, y, np.zeros(n), np.zeros(n), np.array(masses, float)), theta


@settings(max_examples=150, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(case=scenes_strategy())
def test_core_equals_oracle_on_arbitrary_scenes(oracle_lib, emul_lib, case):
    scene, theta = case
    n = len(scene[0])
    o = make_engine(oracle_lib, scene, 2400, 800, flags=1, theta=theta)
    p = o.params
    ax, ay = o.compute_accelerations()
    ox, oy, *_ = o.get_bodies()
    ci, co = o.body_counts()
    g = emul_acc(emul_lib, scene, p, theta)
    if n:
        assert (g["ci"] == ci).all() and (g["co"] == co).all()
        x, y, _, _, m = (np.ascontiguousarray(a, np.float64) for a in scene)
        gx, gy, stt = np.empty(n), np.empty(n), np.zeros(2, np.int64)
        emul_lib.bh_emul_positions_after_build(n, _dp(x), _dp(y), _dp(m), C.c_double(p.root_cx), C.c_double(p.root_cy),
                                               C.c_double(p.root_half), _dp(gx), _dp(gy), stt.ctypes.data_as(I64))
        assert stt[1] == 0 and (gx == ox).all() and (gy == oy).all()      # jitter mutations, bit-exact
        ok = np.isfinite(ax) & np.isfinite(g["ax"])
        assert (np.isnan(ax) == np.isnan(g["ax"])).all()
        scale = max(1e-300, float(np.hypot(ax[ok], ay[ok]).max())) if ok.any() else 1.0
        assert np.hypot(g["ax"][ok] - ax[ok], g["ay"][ok] - ay[ok]).max(initial=0.0) <= 2e-5 * scale
    to = o.tree()
    k, tg = emul_tree(emul_lib, scene, p, len(to["cx"]))
    assert k == len(to["cx"])
    for q in to:
        assert (tg[q] == to[q]).all(), q
