"""Known-answer tests that pin the oracle (oracle/bh_ref.cpp) to BarnesHutAlg.kt.

The reference ships no tests or golden vectors, so every case here is
hand-derivable from the cited Kotlin lines.  CPU only."""
import math

import numpy as np
import pytest

import os

from conftest import ROOT, leaf_paths, make_engine

G = 80.0


def _scene(rows):
    a = np.array(rows, dtype=np.float64).reshape(-1, 5)
    return tuple(a[:, k].copy() for k in range(5))


def test_default_params(oracle_lib):
    e = make_engine(oracle_lib, _scene([]))
    p = e.default_params(2400, 800)
    # Config.kt:11,14,23,20 and BarnesHutAlg.kt:360-361,315,321
    assert (p.G, p.dt, p.theta, p.soft2) == (80.0, 0.005, 0.30, 1.0)
    assert (p.root_cx, p.root_cy, p.root_half) == (1200.0, 400.0, 1202.0)
    assert (p.merge_max_mass, p.merge_min_dist) == (4000.0, 8.0)
    p = e.default_params(801, 1000)          # Int/2.0, max(W,H)/2.0 + 2.0
    assert (p.root_cx, p.root_cy, p.root_half) == (400.5, 500.0, 502.0)


def test_two_body_force_is_pointForceAcc(oracle_lib):
    """BH.kt:250-259 + :390-391 on two leaves (no criterion for leaves, BH.kt:217-221)."""
    e = make_engine(oracle_lib, _scene([[100, 100, 0, 0, 3.0], [103, 104, 0, 0, 5.0]]), theta=0.5)
    ax, ay = e.compute_accelerations()
    r2 = 9.0 + 16.0 + 1.0
    inv_r, inv_r2 = 1.0 / math.sqrt(r2), 1.0 / r2
    f = G * 3.0 * 5.0 * inv_r2
    assert ax[0] == (f * 3.0 * inv_r) / 3.0 and ay[0] == (f * 4.0 * inv_r) / 3.0
    assert ax[1] == (f * -3.0 * inv_r) / 5.0 and ay[1] == (f * -4.0 * inv_r) / 5.0
    c = e.counters()
    assert c["interactions"] == 2


def test_single_body_and_empty(oracle_lib):
    e = make_engine(oracle_lib, _scene([[10, 10, 1, 2, 7.0]]))
    ax, ay = e.compute_accelerations()
    assert ax[0] == 0.0 and ay[0] == 0.0          # root is its own leaf: `single === b`, BH.kt:219
    t = e.tree()
    assert len(t["cx"]) == 1 and t["body"][0] == 0 and t["mass"][0] == 7.0
    e.step(1)                                     # v unchanged, x drifts: BH.kt:419-422
    x, y, vx, vy, m = e.get_bodies()
    assert (x[0], y[0], vx[0], vy[0]) == (10 + 1 * 0.005, 10 + 2 * 0.005, 1.0, 2.0)
    e = make_engine(oracle_lib, _scene([]))
    e.step(2)
    assert e.n == 0
    t = e.tree()                                  # a lone empty root: mass 0, COM = centre, BH.kt:179-183
    assert len(t["cx"]) == 1 and t["body"][0] == -1 and (t["comx"][0], t["comy"][0]) == (1200.0, 400.0)


def test_quad_children_and_half_open_cells(oracle_lib):
    """Quad.child (BH.kt:73-80), contains (BH.kt:61-62) and the digit rule (BH.kt:153-155)."""
    # window 796x796 -> root centre (398,398), half 400: x,y in [-2, 798)
    bodies = [[397.0, 397.0, 0, 0, 1.0],    # NW  (x<cx, y<cy)      digit 0
              [398.0, 397.0, 0, 0, 1.0],    # NE  (x>=cx: boundary) digit 1
              [397.0, 398.0, 0, 0, 1.0],    # SW                    digit 2
              [398.0, 398.0, 0, 0, 1.0]]    # SE                    digit 3
    e = make_engine(oracle_lib, _scene(bodies), width=796, height=796)
    depth, path = leaf_paths(oracle_lib, e)
    assert depth.tolist() == [1, 1, 1, 1] and path.tolist() == [0, 1, 2, 3]
    t = e.tree()
    assert t["h"].tolist() == [400.0, 200.0, 200.0, 200.0, 200.0]
    assert t["cx"].tolist() == [398.0, 198.0, 598.0, 198.0, 598.0]     # NW, NE, SW, SE
    assert t["cy"].tolist() == [398.0, 198.0, 198.0, 598.0, 598.0]
    assert t["body"].tolist() == [-2, 0, 1, 2, 3]
    # half-open box: x = cx-h is inside, x = cx+h is not (BH.kt:61-62, :126)
    e = make_engine(oracle_lib, _scene([[-2.0, 0, 0, 0, 1.0], [798.0, 0, 0, 0, 1.0], [5, 5, 0, 0, 1.0]]), width=796, height=796)
    depth, _ = leaf_paths(oracle_lib, e)
    assert depth[0] >= 0 and depth[1] == -1 and depth[2] >= 0


def test_out_of_box_body_is_target_not_source(oracle_lib):
    """BH.kt:126 drops it from the tree; BH.kt:381-392 still evaluates and integrates it."""
    e = make_engine(oracle_lib, _scene([[100, 100, 0, 0, 2.0], [5000, 100, 0, 0, 9.0]]), theta=0.5)
    ax, ay = e.compute_accelerations()
    assert ax[0] == 0.0 and ay[0] == 0.0                      # the only source is the body itself
    dx = 100.0 - 5000.0
    r2 = dx * dx + 1.0
    want = (G * 9.0 * 2.0 * (1.0 / r2) * dx * (1.0 / math.sqrt(r2))) / 9.0
    assert ax[1] == want and ay[1] == 0.0


def test_com_and_criterion_with_softening_inside(oracle_lib):
    """computeMass order (BH.kt:185-196) and the s^2 < theta^2 (d^2 + eps^2) test (BH.kt:223-228)."""
    # two bodies in the NW child of a 796 window, one far body in SE
    b = [[10.0, 10.0, 0, 0, 1.0], [20.0, 10.0, 0, 0, 3.0], [700.0, 700.0, 0, 0, 2.0]]
    e = make_engine(oracle_lib, _scene(b), width=796, height=796, theta=1.0)
    t = e.tree()
    assert t["mass"][0] == 6.0
    assert t["comx"][0] == ((10.0 * 1.0 + 20.0 * 3.0) / 4.0 * 4.0 + 700.0 * 2.0) / 6.0   # child COM first, then root
    ax, ay = e.compute_accelerations()
    # for body 2 the NW subtree is one cell chain; count interactions: body2 sees ONE accepted cell
    e2 = make_engine(oracle_lib, _scene(b), width=796, height=796, flags=1, theta=1.0)
    e2.compute_accelerations()
    inter, opened = e2.body_counts()
    assert inter[2] == 1
    # theta = 0: s2 < 0 never holds -> every internal cell is opened, BH == direct sum
    e0 = make_engine(oracle_lib, _scene(b), width=796, height=796, theta=0.0)
    a0 = e0.compute_accelerations()
    d0 = e0.direct_sum()
    assert np.allclose(a0[0], d0[0], rtol=1e-14, atol=0) and np.allclose(a0[1], d0[1], rtol=1e-14, atol=0)


def test_accepted_cell_may_contain_the_target(oracle_lib):
    """Softening inside the criterion: a cell with side < theta*eps is always accepted, even
    by a body inside it (self-mass included) — BH.kt:223-230, SURVEY.md §0 fact 4."""
    # two bodies 0.01 apart around (100.3,100.3): their common cells are tiny
    b = [[100.300, 100.300, 0, 0, 1.0], [100.310, 100.300, 0, 0, 1.0], [900.0, 500.0, 0, 0, 1.0]]
    e = make_engine(oracle_lib, _scene(b), flags=1, theta=1.6)
    ax, ay = e.compute_accelerations()
    inter, opened = e.body_counts()
    # body 0 never reaches its sibling leaf: it accepts an ancestor cell holding both
    t = e.tree()
    d = e.direct_sum()
    assert inter[0] == 2 and abs(ax[0] - d[0][0]) > 1e-6


def test_zero_mass_target_is_nan_and_zero_mass_source_is_pruned(oracle_lib):
    b = [[100, 100, 0, 0, 0.0], [200, 100, 0, 0, 2.0], [300, 300, 0, 0, 1.0]]
    e = make_engine(oracle_lib, _scene(b), flags=1, theta=0.5)
    ax, ay = e.compute_accelerations()
    assert math.isnan(ax[0]) and math.isnan(ay[0])           # 0/0, BH.kt:390-391
    inter, _ = e.body_counts()
    assert inter[1] == 1                                     # the zero-mass leaf returns at BH.kt:216


def test_step_is_kick_drift_kick_with_two_evaluations(oracle_lib):
    b = [[100, 100, 0, 0, 3.0], [103, 104, 0, 0, 5.0]]
    e = make_engine(oracle_lib, _scene(b), theta=0.5)
    dt = 0.005
    x, y, vx, vy, m = (np.array(v, float) for v in zip(*b))

    def acc(x, y):
        ax, ay = np.zeros(2), np.zeros(2)
        for i in range(2):
            j = 1 - i
            dx, dy = x[j] - x[i], y[j] - y[i]
            r2 = dx * dx + dy * dy + 1.0
            f = G * m[i] * m[j] * (1.0 / r2)
            ax[i] = (f * dx * (1.0 / math.sqrt(r2))) / m[i]
            ay[i] = (f * dy * (1.0 / math.sqrt(r2))) / m[i]
        return ax, ay

    ax, ay = acc(x, y)
    vx = vx + ax * (dt * 0.5); vy = vy + ay * (dt * 0.5)
    x = x + vx * dt; y = y + vy * dt
    ax, ay = acc(x, y)
    vx = vx + ax * (dt * 0.5); vy = vy + ay * (dt * 0.5)
    e.step(1)
    gx, gy, gvx, gvy, gm = e.get_bodies()
    assert (gx == x).all() and (gy == y).all() and (gvx == vx).all() and (gvy == vy).all()
    assert e.counters()["total_evaluations"] == 2


def test_jitter_mutates_and_can_drop_bodies(oracle_lib):
    """BH.kt:146-151: two bodies in one cell with h < 1e-3 get +-1e-3 shifts by mantissa LSB."""
    xa = 100.25
    b = [[xa, 100.25, 0, 0, 1.0], [xa + 1e-5, 100.25, 0, 0, 1.0], [900.0, 500.0, 0, 0, 1.0]]
    e = make_engine(oracle_lib, _scene(b), theta=0.5)
    e.build_tree()
    x, y, *_ = e.get_bodies()
    assert abs(abs(x[0] - xa) - 1e-3) < 1e-9 or abs(abs(x[0] - xa) - 2e-3) < 1e-9 or x[0] != xa
    assert x[2] == 900.0


def test_merge_rule(oracle_lib):
    """BH.kt:463-532: heavy (m > 4000) absorbs MASS ONLY of bodies closer than 8; victims removed."""
    b = [[100, 100, 1, 2, 5000.0], [103, 100, 9, 9, 2.0], [100, 107.9, 0, 0, 3.0], [100, 108.0, 0, 0, 4.0], [500, 500, 0, 0, 1.0]]
    e = make_engine(oracle_lib, _scene(b), theta=0.5, merge_min_dist=8.0, merge_max_mass=4000.0, dt=0.0)
    e.step(1)
    x, y, vx, vy, m = e.get_bodies()
    assert e.n == 3
    assert m[0] == 5000.0 + 3.0 + 2.0              # descending index order: += m[2] then += m[1]
    assert e.get_origin().tolist() == [0, 3, 4]    # d == 8.0 is not < 8.0 (strict)
    assert e.counters()["total_merged"] == 2
    e2 = make_engine(oracle_lib, _scene(b), theta=0.5, merge_min_dist=0.0, dt=0.0)
    e2.step(1)
    assert e2.n == 5


def test_tree_potential_converges_to_the_pair_sum(oracle_lib):
    """bh_energy_tree (walk of BH.kt:215-239 with 1/sqrt(d^2+soft2)): theta -> 0 reproduces the
    all-pairs potential; at theta = 0.5 it is within the usual Barnes-Hut error; kinetic energy and
    momentum are the same sums."""
    from bh_b200 import scenes
    scene = scenes.snap_f32(scenes.default_two_disks(n1=1500, n2=500, seed=9))
    e = make_engine(oracle_lib, scene, theta=0.5)
    d = e.energy()
    t0 = e.energy_tree(1e-6)
    t5 = e.energy_tree(0.5)
    assert abs(t0["potential"] - d["potential"]) <= 1e-12 * abs(d["potential"])
    assert abs(t5["potential"] - d["potential"]) <= 5e-3 * abs(d["potential"])      # monopole-only cells
    assert t5["kinetic"] == d["kinetic"] and t5["px"] == d["px"] and t5["py"] == d["py"]
    # theta <= 0: the engine's theta (the port sums per-thread partial potentials, whose split varies between calls)
    assert abs(e.energy_tree()["potential"] - t5["potential"]) <= 1e-13 * abs(t5["potential"])


def test_jvm_roundtrip_format(tmp_path, oracle_lib):
    """tests/golden/jvm_roundtrip.py (the tool that pins the oracle against a real JVM run of the
    reference): case / dump formats round-trip and `check` accepts the oracle's own dump."""
    import subprocess
    import sys
    tool = os.path.join(ROOT, "tests", "golden", "jvm_roundtrip.py")
    case, dump = str(tmp_path / "case.bin"), str(tmp_path / "case.out")
    assert subprocess.call([sys.executable, tool, "make", case, "--steps", "2"]) == 0
    assert subprocess.call([sys.executable, tool, "selftest", case, dump]) == 0
    assert subprocess.call([sys.executable, tool, "check", case, dump]) == 0
