"""GPU parity: the CUDA engine (through the C ABI) against the oracle on identical inputs.

Gates (BASELINE.json north_star):
  * Morton order and cell assignment: bit-exact (keys, leaf depths, sort order, every cell
    of visitQuads incl. the f64 centres of mass);
  * per-body accept/open decisions: interaction and opened-cell counts equal as integers;
  * per-body acceleration: |da| <= 1e-5 * max(|a_i|, 0.05 rms|a|), p99 of the unfloored
    relative error <= 1e-5, norm-wise error <= 1e-6   (conftest.assert_acc_parity);
  * trajectories: 10 steps within 1e-5 relative; 1000-step energy drift bounded by the
    oracle's own drift."""
import math

import numpy as np
import pytest

import bh_b200
from bh_b200 import scenes
from conftest import ACC_TOL, acc_errors, assert_acc_parity, key_levels, leaf_paths, make_engine

pytestmark = pytest.mark.gpu


def _oob_scene():
    s = scenes.make_uniform_random(3000, 0.5, seed=5)
    s[0][:50] += 3000.0
    s[1][50:60] -= 5000.0
    return s


CASES = [
    ("C1 two-disk 12.5k θ0.5", lambda: scenes.snap_f32(scenes.default_two_disks()), 2400, 800, 0.5),
    ("C1 two-disk 12.5k θ0.30 (code default)", lambda: scenes.snap_f32(scenes.default_two_disks(seed=2)), 2400, 800, 0.30),
    ("two-disk 2x10k θ0.5", lambda: scenes.snap_f32(scenes.default_two_disks(n1=10000, n2=10000, seed=3)), 2400, 800, 0.5),
    ("uniform 20k unsnapped θ0.2", lambda: scenes.make_uniform_random(20000, 0.5, seed=4), 2400, 800, 0.2),
    ("uniform 20k θ1.6", lambda: scenes.make_uniform_random(20000, 0.5, seed=5), 2400, 800, 1.6),
    ("out-of-box targets", _oob_scene, 2400, 800, 1.0),
    ("big box 32768²", lambda: scenes.default_two_disks(32768, 32768, 40000, 10000, scale=2.0 * math.sqrt(5.0), seed=4), 32768, 32768, 0.5),
    ("mixed-mass 131072²", lambda: scenes.mixed_mass_stress(8, 4000, 16, 131072, 131072, seed=8), 131072, 131072, 0.5),
    ("uniform 200k θ0.5", lambda: scenes.make_uniform_random(200000, 0.5, seed=9), 2400, 800, 0.5),
]


@pytest.mark.parametrize("name,gen,W,H,theta", CASES, ids=[c[0] for c in CASES])
def test_tree_decisions_and_accelerations_match_oracle(oracle_lib, cuda_lib, name, gen, W, H, theta):
    scene = gen()
    o = make_engine(oracle_lib, scene, W, H, flags=1, theta=theta)
    g = make_engine(cuda_lib, scene, W, H, flags=1, theta=theta)
    ax, ay = o.compute_accelerations()
    gx, gy = g.compute_accelerations()
    oc, gc = o.counters(), g.counters()
    assert gc["n_jitter_bodies"] == 0, "parity inputs must stay out of the jitter regime"
    # decisions
    assert (oc["interactions"], oc["opened"]) == (gc["interactions"], gc["opened"])
    oi, oo = o.body_counts()
    gi, go = g.body_counts()
    assert (oi == gi).all() and (oo == go).all()
    # Morton order / cell assignment
    depth, path = leaf_paths(oracle_lib, o)
    key, gdepth, order = g.morton()
    L = key_levels(g.params.root_half)
    assert gc["key_levels"] == L
    assert (gdepth == depth).all()
    inb = depth >= 0
    assert ((key[inb] >> (2 * (L - depth[inb])).astype(np.uint64)) == path[inb]).all()
    assert (key[~inb] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    assert (order == np.argsort(key, kind="stable")).all()          # sort: stable, exact
    assert gc["n_in_tree"] == int(inb.sum()) and gc["n_out_of_box"] == int((~inb).sum())
    # every cell, bit-exact
    to, tg = o.tree(), g.tree()
    assert len(to["cx"]) == len(tg["cx"])
    for k in to:
        assert (to[k] == tg[k]).all(), k
    # accelerations
    s = assert_acc_parity(ax, ay, gx, gy, name)
    print(name, s)


def test_full_size_1m_cloud(oracle_lib, cuda_lib):
    """BASELINE config[1] at full size: 1M-body uniform cloud, θ = 0.5, vs the oracle."""
    scene = scenes.make_uniform_random(1_000_000, 0.5, seed=3)
    o = make_engine(oracle_lib, scene, theta=0.5)
    g = make_engine(cuda_lib, scene, flags=1, theta=0.5)
    ax, ay = o.compute_accelerations()
    gx, gy = g.compute_accelerations()
    oc, gc = o.counters(), g.counters()
    assert (oc["interactions"], oc["opened"]) == (gc["interactions"], gc["opened"])
    s = assert_acc_parity(ax, ay, gx, gy, "1M cloud")
    print("1M cloud", s, "retests", gc["exact_retests"])
    # size-independent properties: the sort is a sorted permutation, the scan closes the tree
    key, depth, order = g.morton()
    assert (np.sort(order) == np.arange(len(order))).all()
    assert (np.diff(key[order].astype(np.uint64)) >= 0).all() if False else (key[order][1:] >= key[order][:-1]).all()
    assert gc["n_cells"] == gc["n_in_tree"] + gc["n_internal"]
    gi, go = g.body_counts()
    assert int(gi.sum()) == gc["interactions"] and int(go.sum()) == gc["opened"]


def test_theta_sweep_error_vs_device_direct_sum(cuda_lib):
    """C2: BH force error against the on-device all-pairs kernel grows monotonically with θ
    over the reference's adjustable range 0.2-1.6 (NBodyPanel.kt:247-248)."""
    scene = scenes.make_uniform_random(100_000, 0.5, seed=3)
    g = make_engine(cuda_lib, scene)
    dx, dy = g.direct_sum()
    errs = []
    for theta in (0.2, 0.3, 0.5, 0.8, 1.0, 1.3, 1.6):
        g.set_params(theta=theta)
        bx, by = g.compute_accelerations()
        s = acc_errors(dx, dy, bx, by)
        errs.append(s["normwise"])
        print("theta", theta, "normwise", s["normwise"], "median", s["median"], "interactions/body", g.counters()["interactions"] / g.n)
    assert all(a < b for a, b in zip(errs, errs[1:])), errs
    assert errs[0] < 5e-3 and errs[2] < 5e-2


def test_theta_zero_equals_direct_sum(oracle_lib, cuda_lib):
    """θ = 0: s² < 0 never holds, every cell is opened, BH == direct sum."""
    scene = scenes.snap_f32(scenes.make_uniform_random(2000, 0.5, seed=12))
    g = make_engine(cuda_lib, scene, flags=1, theta=0.0)
    bx, by = g.compute_accelerations()
    gi, _ = g.body_counts()
    assert (gi == 1999).all()
    dx, dy = g.direct_sum()
    o = make_engine(oracle_lib, scene, theta=0.0)
    ox, oy = o.direct_sum()
    assert_acc_parity(ox, oy, dx, dy, "device direct sum vs f64 direct sum")
    assert_acc_parity(ox, oy, bx, by, "θ=0 walk vs f64 direct sum")


def test_trajectory_10_steps(oracle_lib, cuda_lib):
    scene = scenes.snap_f32(scenes.default_two_disks())
    o = make_engine(oracle_lib, scene, theta=0.5)
    g = make_engine(cuda_lib, scene, theta=0.5)
    o.step(10)
    g.step(10)
    so, sg = o.get_bodies(), g.get_bodies()
    pos_err = np.hypot(so[0] - sg[0], so[1] - sg[1]).max()
    vel_err = np.hypot(so[2] - sg[2], so[3] - sg[3])
    v = np.hypot(so[2], so[3])
    assert pos_err <= 1e-5 * 2400 * 1e-2, pos_err                    # 1e-7 of the box
    assert (vel_err / np.maximum(v, 0.05 * np.sqrt(np.mean(v ** 2)))).max() <= 1e-5
    assert (so[4] == sg[4]).all()


def test_energy_drift_1000_steps(oracle_lib, cuda_lib):
    """H9 protocol: stable setting (Δt = 0.001, merge off, θ = 0.5), 1000 steps;
    |ΔE/E0| on the GPU <= max(2 x the oracle's own BH drift, 1e-4)."""
    scene = scenes.snap_f32(scenes.default_two_disks(n1=2400, n2=600, seed=31))
    o = make_engine(oracle_lib, scene, theta=0.5, dt=0.001)
    g = make_engine(cuda_lib, scene, theta=0.5, dt=0.001)
    e0o, e0g = o.energy()["total"], g.energy()["total"]
    assert abs(e0o - e0g) <= 1e-9 * abs(e0o)
    o.step(1000)
    g.step(1000)
    do = abs(o.energy()["total"] - e0o) / abs(e0o)
    dg = abs(g.energy()["total"] - e0g) / abs(e0g)
    print("energy drift oracle", do, "gpu", dg)
    assert dg <= max(2.0 * do, 1e-4)
    # tree forces are not pairwise symmetric, so momentum drifts slightly and (the system being
    # chaotic) differently for FP32 and f64 interactions; bound the difference, not the value
    assert abs(g.energy()["px"] - o.energy()["px"]) <= 5e-3 * max(1.0, abs(o.energy()["px"]))


def test_edge_cases(oracle_lib, cuda_lib):
    z = np.zeros(0)
    g = make_engine(cuda_lib, (z, z, z, z, z))
    g.step(2)                                                   # resetBodies(emptyList), NBodyPanel.kt:144
    assert g.n == 0 and len(g.tree()["cx"]) == 1 and g.tree()["body"][0] == -1
    one = (np.array([10.0]), np.array([10.0]), np.array([1.0]), np.array([2.0]), np.array([7.0]))
    g.set_bodies(*one)
    ax, ay = g.compute_accelerations()
    assert ax[0] == 0.0 and ay[0] == 0.0
    g.step(1)
    x, y, vx, vy, m = g.get_bodies()
    assert (x[0], y[0], vx[0], vy[0]) == (10 + 1 * 0.005, 10 + 2 * 0.005, 1.0, 2.0)
    # every body outside the root box: nobody is a source, everybody is integrated
    far = (np.array([5000.0, 6000.0]), np.array([0.0, 0.0]), np.array([1.0, 0.0]), np.array([0.0, 1.0]), np.array([1.0, 1.0]))
    g.set_bodies(*far)
    ax, ay = g.compute_accelerations()
    assert (ax == 0).all() and (ay == 0).all() and g.counters()["n_out_of_box"] == 2
    # zero-mass target -> NaN like 0/0 in BH.kt:390-391; zero-mass source pruned (BH.kt:216)
    b = np.array([[100, 100, 0, 0, 0.0], [200, 100, 0, 0, 2.0], [300, 300, 0, 0, 1.0]], float)
    sc = tuple(b[:, k].copy() for k in range(5))
    g2 = make_engine(cuda_lib, sc, flags=1, theta=0.5)
    o2 = make_engine(oracle_lib, sc, flags=1, theta=0.5)
    gx, gy = g2.compute_accelerations()
    ox, oy = o2.compute_accelerations()
    assert math.isnan(gx[0]) and math.isnan(gy[0])
    assert (g2.body_counts()[0] == o2.body_counts()[0]).all()
    assert np.allclose(gx[1:], ox[1:], rtol=1e-6) and np.allclose(gy[1:], oy[1:], rtol=1e-6)
    # NaN position: fails contains() like any comparison with NaN (BH.kt:61-62)
    b[0] = [float("nan"), 5.0, 0, 0, 1.0]
    g3 = make_engine(cuda_lib, tuple(b[:, k].copy() for k in range(5)))
    g3.compute_accelerations()
    assert g3.counters()["n_out_of_box"] == 1
    # coincident bodies (jitter regime, BH.kt:146-151): detected, replayed like the reference
    s = scenes.make_uniform_random(1000, 0.5, seed=1)
    s[0][1], s[1][1] = s[0][0], s[1][0]
    g4 = make_engine(cuda_lib, s)
    o4 = make_engine(oracle_lib, s)
    ax, ay = g4.compute_accelerations()
    o4.compute_accelerations()
    assert g4.counters()["n_jitter_bodies"] == 2 and np.isfinite(ax).all()
    assert (g4.get_bodies()[0] == o4.get_bodies()[0]).all() and (g4.get_bodies()[1] == o4.get_bodies()[1]).all()


def test_params_are_reread_every_call(oracle_lib, cuda_lib):
    """Z/X, O/P, K/L keys change Config between steps (NBodyPanel.kt:247-260)."""
    scene = scenes.snap_f32(scenes.default_two_disks(n1=1500, n2=500, seed=6))
    o = make_engine(oracle_lib, scene, flags=1)
    g = make_engine(cuda_lib, scene, flags=1)
    for theta, G, dt in ((0.2, 80.0, 0.005), (1.6, 100.0, -0.003), (0.75, 0.0, 0.01)):
        for e in (o, g):
            e.set_params(theta=theta, G=G, dt=dt)
            e.step(1)
            e.compute_accelerations()
        assert (o.body_counts()[0] == g.body_counts()[0]).all()
    so, sg = o.get_bodies(), g.get_bodies()
    assert np.hypot(so[0] - sg[0], so[1] - sg[1]).max() < 1e-6


def test_kotlin_facade_drives_the_engine(cuda_lib):
    """PhysicsEngine / Body / Config with the reference's names (NBodyPanel.kt call sites)."""
    import bh_b200
    from bh_b200 import Body, Config, PhysicsEngine
    Config.reset()
    Config.theta = 0.5
    sc = scenes.default_two_disks(n1=600, n2=200, seed=5)
    bodies = [Body(*[float(v[i]) for v in sc]) for i in range(len(sc[0]))]
    eng = PhysicsEngine(bodies)
    eng.mergeMinDist = 0.0
    first = bodies[0]
    x0 = first.x
    eng.step()
    assert eng.getBodies() is bodies and bodies[0] is first and first.x != x0 or first.vx == 0
    quads = []
    eng.getTreeForDebug().visitQuads(lambda q: quads.append(q))
    assert quads[0] == bh_b200.Quad(1200.0, 400.0, 1202.0) and len(quads) > len(bodies)
    eng.resetBodies([])
    eng.step()
    assert eng.getBodies() == []
    Config.reset()


def _merge_scene(seed=41, n1=3000, n2=800, k=60):
    """Two-disk scene whose heavy centres (m = 50,000 and 5,000 > mergeMaxMass) have satellites
    inside the 8 px merge radius, plus a third heavy body close enough to be eaten itself."""
    s = list(scenes.snap_f32(scenes.default_two_disks(n1=n1, n2=n2, seed=seed)))
    rng = np.random.default_rng(seed)
    ang, rad = rng.uniform(0, 2 * np.pi, k), rng.uniform(1.0, 7.0, k)
    s[0][5:5 + k] = s[0][0] + rad * np.cos(ang)
    s[1][5:5 + k] = s[1][0] + rad * np.sin(ang)
    j = n1 + 3                                       # a satellite of the second disk near its centre n1
    s[0][j], s[1][j] = s[0][n1] + 2.0, s[1][n1] - 3.0
    s[0][n1 + 7], s[1][n1 + 7], s[4][n1 + 7] = s[0][n1] + 5.0, s[1][n1] + 1.0, 4500.0   # heavy, inside the radius of heavy n1
    return tuple(np.ascontiguousarray(a) for a in s)


def test_merge_rule_matches_oracle(oracle_lib, cuda_lib):
    """mergeCloseBodiesIfNeeded (BH.kt:463-532) with the reference defaults (4000 / 8 px): the
    same bodies disappear, masses are summed in the same order (bit-identical), list order is kept."""
    scene = _merge_scene()
    o = make_engine(oracle_lib, scene, theta=0.5, merge_min_dist=8.0)
    g = make_engine(cuda_lib, scene, theta=0.5, merge_min_dist=8.0)
    for step in range(6):
        o.step(1)
        g.step(1)
        assert g.n == o.n, step
        assert (g.get_origin() == o.get_origin()).all(), step
        so, sg = o.get_bodies(), g.get_bodies()
        assert (so[4] == sg[4]).all(), step                        # masses: same f64 sums, same order
        # satellites 1-7 px from a 50,000-mass body feel a ~ 1e5-1e6: FP32 interaction rounding
        # (1e-7 relative) moves them by ~1e-6 px per step; 1e-4 px is 4e-8 of the box
        assert np.hypot(so[0] - sg[0], so[1] - sg[1]).max() < 1e-4
    assert o.counters()["total_merged"] == g.counters()["total_merged"] > 40
    # the engine keeps working on the shrunk list: accelerations still match
    ax, ay = o.compute_accelerations()
    gx, gy = g.compute_accelerations()
    assert_acc_parity(ax, ay, gx, gy, "after merges")


def test_merge_default_scene_long_run(oracle_lib, cuda_lib):
    """The shipped app's configuration (merge on, θ = 0.30, Δt = 0.005): 60 steps, identical
    survivor lists at every step."""
    scene = scenes.snap_f32(scenes.default_two_disks(n1=2000, n2=500, seed=43))
    o = make_engine(oracle_lib, scene, theta=0.30, merge_min_dist=8.0)
    g = make_engine(cuda_lib, scene, theta=0.30, merge_min_dist=8.0)
    for step in range(60):
        o.step(1)
        g.step(1)
        assert g.n == o.n and (g.get_origin() == o.get_origin()).all(), step
    assert (o.get_bodies()[4] == g.get_bodies()[4]).all()


def test_rehoming_is_invisible_through_the_abi(cuda_lib):
    """Device state lives in Morton ('home') order; the ABI speaks list order only.  Results must
    not depend on how often the engine re-homes (bit-identical trajectories)."""
    import bh_b200
    scene = scenes.snap_f32(scenes.default_two_disks(n1=3000, n2=1000, seed=44))
    outs = []
    for interval in (1, 3, 1000):
        e = bh_b200.NativeEngine(lib=cuda_lib, rehome_interval=interval, flags=1)
        e.set_params(theta=0.5, merge_min_dist=0.0)
        e.set_bodies(*scene)
        e.step(7)
        e.set_bodies(*e.get_bodies())          # same-length list: stored through the existing permutation
        e.step(2)
        outs.append(e.get_bodies() + (e.body_counts()[0],) + e.get_positions_f32())
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert (a == b).all()


def test_jitter_regime_matches_oracle(oracle_lib, cuda_lib):
    """Bodies sharing a cell with h < 1e-3 (BarnesHutAlg.kt:145-156): buildTree() MUTATES them by
    +-1e-3 and may drop them from the tree.  Same mutated coordinates (bit-exact), same dropped
    bodies, same cells (incl. the empty shells of collided children), same decisions."""
    from test_core_emulation import _jitter_scenes
    scene = _jitter_scenes()[0][1]
    for theta in (0.5, 0.0):
        o = make_engine(oracle_lib, scene, flags=1, theta=theta)
        g = make_engine(cuda_lib, scene, flags=1, theta=theta)
        ax, ay = o.compute_accelerations()
        gx, gy = g.compute_accelerations()
        so, sg = o.get_bodies(), g.get_bodies()
        assert (so[0] == sg[0]).all() and (so[1] == sg[1]).all()
        assert (so[0] != scene[0]).sum() > 50
        depth, path = leaf_paths(oracle_lib, o)
        key, gdepth, order = g.morton()
        assert (gdepth == depth).all() and (depth < 0).sum() > 10
        oc, gc = o.counters(), g.counters()
        assert gc["n_in_tree"] == int((depth >= 0).sum())
        assert (oc["interactions"], oc["opened"]) == (gc["interactions"], gc["opened"])
        assert (o.body_counts()[0] == g.body_counts()[0]).all() and (o.body_counts()[1] == g.body_counts()[1]).all()
        to, tg = o.tree(), g.tree()
        assert len(to["cx"]) == len(tg["cx"])
        for k in to:
            assert (to[k] == tg[k]).all(), k
        ok = np.isfinite(ax)
        assert_acc_parity(ax[ok], ay[ok], gx[ok], gy[ok], "jitter regime")
    # Through whole steps every build re-jitters whatever still shares a cell, and the direction of
    # each +-1e-3 shift is the LAST MANTISSA BIT of the coordinate (BH.kt:149-150): once FP32
    # interaction rounding (1e-9 px after one step) flips that bit the shifts differ, so beyond the
    # first build only the size of the effect is comparable, not the coordinates.
    o = make_engine(oracle_lib, scene, theta=0.5)
    g = make_engine(cuda_lib, scene, theta=0.5)
    o.step(3)
    g.step(3)
    so, sg = o.get_bodies(), g.get_bodies()
    ok = np.isfinite(so[0])
    d = np.hypot(so[0] - sg[0], so[1] - sg[1])[ok]
    assert np.isfinite(sg[0][ok]).all() and d.max() < 2e-2 and np.median(d) < 1e-6


def test_acceleration_reuse_is_result_identical(cuda_lib):
    """BH_FLAG_REUSE_ACC: step n+1 starts from the a(t+dt) step n ended with (the reference
    recomputes it from unchanged positions, BH.kt:407-408).  Bit-identical state, also across
    Config changes, merges and resetBodies, with about half the evaluations."""
    import bh_b200
    scene = _merge_scene(seed=51, n1=2500, n2=700)
    engines = []
    for flags in (0, bh_b200.BH_FLAG_REUSE_ACC):
        e = bh_b200.NativeEngine(lib=cuda_lib, flags=flags)
        e.set_params(theta=0.5, merge_min_dist=8.0)
        e.set_bodies(*scene)
        e.step(5)
        e.set_params(theta=0.35)                # Z key between frames
        e.step(3)
        e.set_params(dt=-0.002, G=75.0)         # O / K keys
        e.step(3)
        e.set_bodies(*e.get_bodies())           # resetBodies
        e.step(2)
        engines.append(e)
    a, b = engines[0].get_bodies(), engines[1].get_bodies()
    for u, v in zip(a, b):
        assert u.shape == v.shape and (u == v).all()
    assert engines[0].counters()["total_merged"] == engines[1].counters()["total_merged"] > 0
    assert engines[1].counters()["total_evaluations"] < engines[0].counters()["total_evaluations"]


def test_full_size_10m_two_disk_merger(oracle_lib, cuda_lib):
    """BASELINE config[2] at full size: 8M + 2M two-disk merger with central black holes in a
    32768^2 window, θ = 0.5 — one evaluation against the oracle (decisions as integers,
    accelerations within 1e-5), then size-independent properties of the build."""
    scene = scenes.default_two_disks(32768, 32768, 8_000_000, 2_000_000, scale=math.sqrt(800.0), seed=4)
    o = make_engine(oracle_lib, scene, 32768, 32768, theta=0.5)
    g = make_engine(cuda_lib, scene, 32768, 32768, flags=1, theta=0.5)
    ax, ay = o.compute_accelerations()
    gx, gy = g.compute_accelerations()
    oc, gc = o.counters(), g.counters()
    assert (oc["interactions"], oc["opened"]) == (gc["interactions"], gc["opened"])
    s = assert_acc_parity(ax, ay, gx, gy, "10M two-disk")
    print("10M two-disk", s, "retests", gc["exact_retests"], "cells", gc["n_cells"], "depth", gc["max_depth"])
    del o, ax, ay
    key, depth, order = g.morton()
    assert (np.sort(order) == np.arange(len(order))).all()              # the sort is a permutation
    ks = key[order]
    # ... and sorted; the cusped disk centres put a few bodies into the jitter regime, whose replay
    # moves them AFTER the sort (BH.kt:146-151), so their re-derived keys may be out of place
    assert int((ks[1:] < ks[:-1]).sum()) <= 2 * gc["n_jitter_bodies"]
    print("jitter bodies", gc["n_jitter_bodies"])
    assert gc["n_cells"] >= gc["n_in_tree"] + gc["n_internal"]          # ghosts of dropped bodies keep their leaf
    gi, go = g.body_counts()
    assert int(gi.sum(dtype=np.int64)) == gc["interactions"] and int(go.sum(dtype=np.int64)) == gc["opened"]
    # root centre of mass == mass-weighted mean of the in-box bodies (f64, tree-ordered sum vs numpy)
    t0 = g.tree_root()
    inb = depth >= 0
    msum = scene[4][inb].sum()
    assert abs(t0["mass"] - msum) <= 1e-9 * msum
    assert abs(t0["comx"] - (scene[4][inb] * scene[0][inb]).sum() / msum) <= 1e-6
    assert abs(t0["comy"] - (scene[4][inb] * scene[1][inb]).sum() / msum) <= 1e-6


def test_large_32m_cloud_properties(cuda_lib):
    """32M-body uniform cloud in a density-scaled window (no oracle at this size): the step runs,
    the sort is a sorted permutation, counters are consistent, momentum stays ~0 for a cloud at rest
    and re-homing / read-back return the list order."""
    n = 32_000_000
    W, H = 13576, 4525                       # 2400x800 scaled by sqrt(32)
    scene = scenes.make_uniform_random(n, 0.5, W, H, seed=6)
    g = make_engine(cuda_lib, scene, W, H, theta=0.5)
    g.step(1)
    c = g.counters()
    # bodies dropped by the jitter replay (a handful of close pairs at this size) keep a ghost leaf
    assert c["n_cells"] == c["n_internal"] + (n - c["n_out_of_box"]) and c["n_in_tree"] + c["n_out_of_box"] <= n
    assert n - c["n_out_of_box"] - c["n_in_tree"] <= c["n_jitter_bodies"] < 1000
    assert 200 < c["interactions"] / n < 400
    key, depth, order = g.morton()
    assert (np.sort(order) == np.arange(n)).all()
    ks = key[order]
    assert int((ks[1:] < ks[:-1]).sum()) <= 2 * c["n_jitter_bodies"]    # jittered bodies moved after the sort
    x, y, vx, vy, m = g.get_bodies()
    assert (m == scene[4]).all() and np.isfinite(vx).all()
    d = np.hypot(x - scene[0], y - scene[1])
    assert d.max() < 1.0                      # dt = 0.005: nobody moved more than a pixel in one step
    px, py = (m * vx).sum(), (m * vy).sum()
    assert abs(px) + abs(py) < 1e-3 * (m * np.hypot(vx, vy)).sum()


def test_tree_energy_matches_oracle_and_scales(oracle_lib, cuda_lib):
    """bh_energy_tree: same decisions as the force walk, FP32 terms in f64 sums — potential within
    1e-6 of the f64 oracle; O(N log N), so energy drift can be tracked at sizes where the pair sum
    (bh_energy) is out of reach."""
    scene = scenes.snap_f32(scenes.default_two_disks(n1=8000, n2=2000, seed=13))
    o = make_engine(oracle_lib, scene, theta=0.5)
    g = make_engine(cuda_lib, scene, theta=0.5)
    for th in (0.5, 0.2, 1.0):
        eo, eg = o.energy_tree(th), g.energy_tree(th)
        assert abs(eg["potential"] - eo["potential"]) <= 1e-6 * abs(eo["potential"]), th
        assert abs(eg["kinetic"] - eo["kinetic"]) <= 1e-12 * abs(eo["kinetic"])
    d = g.energy()
    assert abs(g.energy_tree(0.2)["potential"] - d["potential"]) <= 5e-4 * abs(d["potential"])
    # 4M bodies: tree energy before / after 5 steps at dt = 0.001 (stable setting, H9)
    big = scenes.make_uniform_random(4_000_000, 0.5, 4800, 1600, seed=15)
    gb = make_engine(cuda_lib, big, 4800, 1600, theta=0.5, dt=0.001)
    e0 = gb.energy_tree(0.3)
    gb.step(5)
    e1 = gb.energy_tree(0.3)
    assert abs(e1["total"] - e0["total"]) <= 1e-4 * abs(e0["total"])


def test_async_render_readback_overlaps_the_next_step(oracle_lib, cuda_lib):
    """bh_request_positions_f32 / bh_wait_positions_f32 (what NBodyPanel.paintComponent needs,
    NBodyPanel.kt:302-306): the snapshot is the state at request time even though a step ran
    before it was collected; list order survives re-homing and merges."""
    scene = _merge_scene(seed=61, n1=4000, n2=1000)
    for lib in (oracle_lib, cuda_lib):
        e = make_engine(lib, scene, theta=0.5, merge_min_dist=8.0)
        e.step(3)
        x, y, vx, vy, m = e.get_bodies()
        e.request_positions_f32()
        e.step(2)                              # overlaps the copy
        xy, mf = e.wait_positions_f32()
        assert xy.shape == (len(x), 2)
        assert (xy[:, 0] == x.astype(np.float32)).all() and (xy[:, 1] == y.astype(np.float32)).all()
        assert (mf == m.astype(np.float32)).all()
        xy2, _ = e.get_positions_f32()         # the synchronous read-back now shows the later state
        assert (xy2[: min(len(xy2), len(xy))] != xy[: min(len(xy2), len(xy))]).any()


def test_device_scene_generators_follow_bodyfactory(cuda_lib):
    """bh_append_disk / bh_append_uniform_random (BodyFactory.kt:63-150, :160-177) on the device.
    Kotlin's RNG stream is not reproducible, so: (i) the deterministic parts are checked exactly —
    list order old + new, centre body, radius range, enclosed-mass circular speeds recomputed with
    numpy from the generated positions; (ii) the sampling laws are compared with the numpy
    generators of scenes.py (two-sample Kolmogorov-Smirnov on radius, angle, speed)."""
    import bh_b200
    from scipy import stats
    e = bh_b200.NativeEngine(lib=cuda_lib)
    e.set_params(theta=0.5, merge_min_dist=0.0)
    old = scenes.make_uniform_random(1000, 0.5, seed=2)
    e.set_bodies(*old)
    n_total = 200_001
    p = e.disk_params(2400, 800, r=300.0, x=1200.0, y=400.0)
    e.append_disk(n_total, p, seed=7)
    e.append_uniform_random(5000, 0.5, 2400, 800, seed=9)      # key C (NBodyPanel.kt:282-286)
    e.append_disk(0, e.disk_params(2400, 800, x=100.0, y=700.0), seed=1)   # RMB black hole (:171)
    x, y, vx, vy, m = e.get_bodies()
    assert len(x) == 1000 + n_total + 5000 + 1
    assert all((a[:1000] == b).all() for a, b in zip((x, y, vx, vy, m), old))          # old bodies first, untouched
    d = slice(1000, 1000 + n_total)
    dx, dy = x[d] - 1200.0, y[d] - 400.0
    assert (x[d][0], y[d][0], m[d][0]) == (1200.0, 400.0, 50000.0) and vx[d][0] == 0.0
    assert np.allclose(m[d][1:], 5000.0 / (n_total - 1), rtol=0, atol=0)
    R = np.hypot(dx, dy)[1:]
    assert R.min() >= 8.0 * 0.97 and R.max() <= 300.0 * 1.03
    # exact enclosed mass -> circular speed within the +-1 % speed jitter, purely tangential, clockwise
    order = np.argsort(np.hypot(dx, dy), kind="stable")
    menc = np.empty(n_total); menc[order] = np.cumsum(m[d][order])
    vc = np.sqrt(80.0 * menc[1:] / R)
    v = np.hypot(vx[d][1:], vy[d][1:])
    assert np.abs(v / vc - 1.0).max() <= 0.0100001
    assert np.abs((vx[d][1:] * dx[1:] + vy[d][1:] * dy[1:]) / (v * R)).max() < 1e-12      # no radial velocity
    assert ((dx[1:] * vy[d][1:] - dy[1:] * vx[d][1:]) < 0).all()                          # clockwise (BF.kt:138)
    u = slice(1000 + n_total, 1000 + n_total + 5000)
    assert (m[u] == 0.5).all() and (vx[u] == 0).all() and x[u].min() >= 0 and x[u].max() < 2400 and y[u].max() < 800
    assert (x[-1], y[-1], m[-1]) == (100.0, 700.0, 50000.0)
    # sampling laws vs the numpy generators
    ref = scenes.make_galaxy_disk(n_total, x=1200.0, y=400.0, r=300.0, seed=11)
    Rr = np.hypot(ref[0][1:] - 1200.0, ref[1][1:] - 400.0)
    assert stats.ks_2samp(R, Rr).pvalue > 1e-3
    assert stats.ks_2samp(np.arctan2(dy[1:], dx[1:]), np.arctan2(ref[1][1:] - 400.0, ref[0][1:] - 1200.0)).pvalue > 1e-3
    assert stats.ks_2samp(v, np.hypot(ref[2][1:], ref[3][1:])).pvalue > 1e-3
    assert stats.ks_2samp(x[u], scenes.make_uniform_random(5000, 0.5, seed=4)[0]).pvalue > 1e-3
    # the generated scene is usable right away, and a Kepler disk can be appended too
    e.append_disk(50_000, e.disk_params(2400, 800, kepler=1, radial_jitter=0.03, r=304.0), seed=3)
    e.step(2)
    assert e.n == len(x) + 50_000 and np.isfinite(e.get_bodies()[0]).all()


def test_step_io_equals_set_step_get(oracle_lib, cuda_lib):
    """bh_step_io = resetBodies + step + getBodies with overlapped transfers: bit-identical to the
    three separate calls — first call (list order), later calls (through the home permutation,
    incl. a re-homing step), several steps per call, merge enabled (fallback), and the oracle's."""
    import bh_b200
    scene = scenes.snap_f32(scenes.default_two_disks(n1=6000, n2=1500, seed=71))
    a = bh_b200.NativeEngine(lib=cuda_lib, rehome_interval=3)
    b = bh_b200.NativeEngine(lib=cuda_lib, rehome_interval=3)
    for e in (a, b):
        e.set_params(theta=0.5, merge_min_dist=0.0)
    state = scene
    for it in range(8):
        k = 1 + (it % 2)
        out = a.step_io(k, inputs=state)
        b.set_bodies(*state)
        b.step(k)
        ref = b.get_bodies()
        for u, v in zip(out, ref):
            assert (u == v).all(), it
        state = tuple(np.ascontiguousarray(v) for v in ref)
    assert a.step_io(2) is None and a.n == b.n                 # no inputs, no outputs: just steps
    b.step(2)
    for u, v in zip(a.get_bodies(), b.get_bodies()):
        assert (u == v).all()
    # merge enabled: plain sequence inside, same answer as the oracle's bh_step_io
    ms = _merge_scene(seed=72, n1=2000, n2=500)
    g = make_engine(cuda_lib, ms, theta=0.5, merge_min_dist=8.0)
    o = make_engine(oracle_lib, ms, theta=0.5, merge_min_dist=8.0)
    og, oo = g.step_io(3, inputs=ms), o.step_io(3, inputs=ms)
    assert len(og[0]) == len(oo[0]) < len(ms[0]) and (og[4] == oo[4]).all()


WALK_VARIANTS = [(1, 0), (1, 1), (2, 1)]


@pytest.mark.gpu
def test_walk_variants_group_size_and_summation(oracle_lib, cuda_lib, monkeypatch):
    """k_walk<G, ACC>: one body per lane with FP32 partial sums folded into f64 (what small target counts use), one
    body per lane with every term added to an f64 sum at once, and PAIRS of bodies per thread sharing one preorder
    position (bh_walk_multi; what large target counts use).  Every variant makes the reference's per-body decisions
    (interaction and opened counts equal to the oracle's as integers), and the two f64 variants are BIT-IDENTICAL:
    a body's result does not depend on its partner, which is what keeps any multi-GPU partition of the targets
    bit-identical to one GPU.  Scenes: cloud, two disks with out-of-box and zero-mass bodies, theta 0.3 .. 1.0."""
    scenes_ = [
        ("cloud 60k", scenes.snap_f32(scenes.make_uniform_random(60_000, 0.5, seed=71)), 0.5),
        ("two disks 30k + specials", None, 0.3),
        ("cloud 20k theta 1.0", scenes.snap_f32(scenes.make_uniform_random(20_000, 0.5, seed=72)), 1.0),
    ]
    s = [a.copy() for a in scenes.snap_f32(scenes.default_two_disks(n1=24_000, n2=6_000, seed=73))]
    s[0][100:110] = np.linspace(-500.0, -50.0, 10)      # outside the root box: targets only (BH.kt:126)
    s[4][200:220] = 0.0                                  # zero mass: pruned as sources, NaN as targets (BH.kt:216,390)
    scenes_[1] = (scenes_[1][0], tuple(s), 0.3)
    for name, scene, theta in scenes_:
        o = make_engine(oracle_lib, scene, theta=theta, flags=bh_b200.BH_FLAG_BODY_COUNTS)
        ox, oy = o.compute_accelerations()
        oi, oo = o.body_counts()
        res = {}
        for g, acc in WALK_VARIANTS:
            monkeypatch.setenv("BH_WALK_G", str(g))
            monkeypatch.setenv("BH_WALK_ACC", str(acc))
            e = make_engine(cuda_lib, scene, theta=theta, flags=bh_b200.BH_FLAG_BODY_COUNTS)
            gx, gy = e.compute_accelerations()
            gi, go = e.body_counts()
            assert (gi == oi).all() and (go == oo).all(), (name, g, acc)
            ok = np.isfinite(ox)
            assert (np.isnan(gx) == ~ok).all(), (name, g, acc)
            assert_acc_parity(ox[ok], oy[ok], gx[ok], gy[ok], what=f"{name} G={g} acc={acc}")
            res[(g, acc)] = (gx, gy)
            e.close()
        assert np.array_equal(res[(2, 1)][0], res[(1, 1)][0], equal_nan=True), name
        assert np.array_equal(res[(2, 1)][1], res[(1, 1)][1], equal_nan=True), name
        o.close()
    monkeypatch.delenv("BH_WALK_G")
    monkeypatch.delenv("BH_WALK_ACC")


def test_bounding_box_reduction_matches_oracle(oracle_lib, cuda_lib):
    """north_star's bounding-box reduction, fused into k_keygen (warp shuffles, one atomic per block and extremum):
    exact min/max of all bodies, in the root box or not; the reference's root stays the window (BH.kt:360-361)."""
    s = [a.copy() for a in scenes.snap_f32(scenes.default_two_disks(n1=20_000, n2=5_000, seed=77))]
    s[0][7], s[1][7] = -812.25, 5000.5          # far outside the root box [-2, 2402) x [-802, 1602)
    s[0][9], s[1][9] = 9000.0, -3000.125
    g = make_engine(cuda_lib, tuple(s), theta=0.5)
    o = make_engine(oracle_lib, tuple(s), theta=0.5)
    g.build_tree()
    o.build_tree()
    gc, oc = g.counters(), o.counters()
    for k in ("bbox_min_x", "bbox_max_x", "bbox_min_y", "bbox_max_y"):
        assert gc[k] == oc[k], k
    assert gc["bbox_min_x"] == -812.25 and gc["bbox_max_x"] == 9000.0 and gc["bbox_min_y"] == -3000.125 and gc["bbox_max_y"] == 5000.5
    assert gc["n_out_of_box"] == 2
    e = make_engine(cuda_lib, tuple(a[:0] for a in s), theta=0.5)
    e.build_tree()
    assert math.isnan(e.counters()["bbox_min_x"])


def test_sharded_slice_io_equals_full_state_io(cuda_lib):
    """bh_step_io_slice moves only the bodies of this rank's slice (home order; bh_get_slice_index names them, the epoch
    tells when a re-homing re-cut the slices).  With one rank the slice is the whole list, so the calls must reproduce
    set_bodies + step + get_bodies bit for bit — over re-homings (interval 3) and with a changed slice every step."""
    scene = scenes.snap_f32(scenes.default_two_disks(n1=6000, n2=2000, seed=81))
    n = len(scene[0])
    ref = bh_b200.NativeEngine(lib=cuda_lib, rehome_interval=3)
    e = bh_b200.NativeEngine(lib=cuda_lib, rehome_interval=3)
    for g in (ref, e):
        g.set_params(theta=0.5, merge_min_dist=0.0)
        g.set_bodies(*scene)
    state = [a.copy() for a in scene]
    out = [np.empty(n) for _ in range(5)]
    epochs = set()
    for s in range(7):
        idx = e.slice_index()                       # the slice as the device holds it NOW
        assert sorted(idx.tolist()) == list(range(n))
        epochs.add(e.slice_epoch())
        rng = np.random.default_rng(s)
        state[2] = state[2] + rng.normal(0, 0.01, n)        # the caller edits its bodies between steps
        k = e.step_io_slice(1, inputs=[a[idx] for a in state], out=out)
        assert k == n
        idx2 = e.slice_index()
        new = [np.empty(n) for _ in range(5)]
        for dst, src in zip(new, out):
            dst[idx2] = src
        ref.set_bodies(*state)
        ref.step(1)
        want = ref.get_bodies()
        for a, b in zip(new, want):
            assert (a == b).all(), s
        state = [a.copy() for a in new]
    assert len(epochs) >= 2                         # re-homings happened in between
    with pytest.raises(bh_b200.BhError):
        e.step_io_slice(1, inputs=[a[:10] for a in state], out=out)     # not the slice length


def test_acceleration_reuse_survives_builds_between_steps(cuda_lib):
    """BH_FLAG_REUSE_ACC and a build BETWEEN steps (round-1 advisor finding): when the 8th step since the last re-homing
    has set `rehome_due`, a `bh_build_tree` — or the overlay's rebuild after `bh_direct_sum` dropped the tree — re-homes
    the state; the accelerations on file are then in the OLD home order and must not be reused by the next step."""
    import bh_b200
    scene = scenes.snap_f32(scenes.default_two_disks(n1=2500, n2=700, seed=52))
    engines = []
    for flags in (0, bh_b200.BH_FLAG_REUSE_ACC):
        e = bh_b200.NativeEngine(lib=cuda_lib, flags=flags)
        e.set_params(theta=0.5, merge_min_dist=0.0)
        e.set_bodies(*scene)
        e.step(8)                               # the 8th step since the first build's re-homing: a re-homing is due
        e.build_tree()                          # ... and happens HERE, between two steps
        e.step(7)
        e.step(1)                               # due again
        e.direct_sum()                          # invalidates the tree; the overlay then rebuilds it (and re-homes)
        e.tree()
        e.step(9)
        engines.append(e)
    a, b = engines[0].get_bodies(), engines[1].get_bodies()
    for u, v in zip(a, b):
        assert (u == v).all()
    assert engines[1].counters()["total_evaluations"] < engines[0].counters()["total_evaluations"]
