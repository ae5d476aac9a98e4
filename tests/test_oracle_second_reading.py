"""Double-entry check of the oracle: oracle/bh_ref.cpp (arrays, indices, threads) against oracle/bh_ref_second.py
(a second reading of BarnesHutAlg.kt as Python objects and recursion, sharing nothing with the first) — every f64 of
the state, every acceleration, every visitQuads cell and the interaction counts BIT FOR BIT, on small scenes that
reach the awkward branches.  Not a pin against the JVM (none can run here); it pins the two readings to each other."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, make_engine

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bh_ref_second as second  # noqa: E402

import bh_b200  # noqa: E402
from bh_b200 import scenes  # noqa: E402


def _configure(W, H, theta, G=80.0, dt=0.005):
    second.Config.WIDTH_PX, second.Config.HEIGHT_PX = W, H
    second.Config.theta, second.Config.G, second.Config.DT = theta, G, dt


def _bodies(scene):
    return [second.Body(*[a[i] for a in scene]) for i in range(len(scene[0]))]


def _same(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and bool((a.view(np.uint64) == b.view(np.uint64)).all() or np.array_equal(a, b, equal_nan=True))


def _state(bodies):
    return [np.array([getattr(b, k) for b in bodies], np.float64) for k in ("x", "y", "vx", "vy", "m")]


def _cells(root, bodies):
    index = {id(b): i for i, b in enumerate(bodies)}
    out = {k: [] for k in ("cx", "cy", "h", "mass", "comx", "comy", "body")}

    def visit(t):
        out["cx"].append(t.quad.cx); out["cy"].append(t.quad.cy); out["h"].append(t.quad.h)
        out["mass"].append(t.mass); out["comx"].append(t.comX); out["comy"].append(t.comY)
        out["body"].append(-2 if t.children is not None else (index[id(t.body)] if t.body is not None else -1))

    root.visit_quads(visit)
    return out


def _two_disks(n1, n2, seed):
    return scenes.snap_f32(scenes.default_two_disks(n1=n1, n2=n2, seed=seed))


def _specials():
    """out-of-box targets, zero-mass bodies, a coincident pair and a near-coincident triple (jitter regime)"""
    s = [a.copy() for a in _two_disks(260, 90, 11)]
    s[0][5:9] = [-40.0, 2500.0, 100.0, 2401.999]          # x outside [-2, 2402) (the last one just inside)
    s[1][9:12] = [-900.0, 1602.0, 1601.5]                 # y outside / on / inside the half-open edge
    s[4][20:24] = 0.0                                     # zero mass: pruned as sources, NaN as targets
    s[0][31], s[1][31] = s[0][30], s[1][30]               # coincident pair
    s[0][41], s[1][41] = s[0][40] + 2.0e-4, s[1][40]
    s[0][42], s[1][42] = s[0][40], s[1][40] + 3.0e-4
    return tuple(s)


def _merging():
    """heavy bodies with satellites inside the 8 px merge radius, placed so that removals shift the heavy's index"""
    s = [a.copy() for a in _two_disks(300, 120, 12)]
    rng = np.random.default_rng(13)
    for heavy, first in ((0, 2), (300, 305)):
        k = 25
        ang, rad = rng.uniform(0, 2 * np.pi, k), rng.uniform(1.0, 7.5, k)
        s[0][first:first + k] = s[0][heavy] + rad * np.cos(ang)
        s[1][first:first + k] = s[1][heavy] + rad * np.sin(ang)
    # a third heavy late in the list whose victims sit BEFORE it (its index moves left when they go)
    s[4][410] = 6000.0
    s[0][100:104] = s[0][410] + np.array([1.0, -2.0, 3.0, 0.5])
    s[1][100:104] = s[1][410] + np.array([0.5, 1.0, -1.0, -3.0])
    return tuple(np.ascontiguousarray(a) for a in s)


CASES = [
    ("two disks, theta 0.5", lambda: _two_disks(300, 100, 10), 2400, 800, 0.5, 3, 0.0),
    ("two disks, code defaults (theta 0.30, merge 8 px)", lambda: _two_disks(300, 100, 14), 2400, 800, 0.30, 3, 8.0),
    ("out-of-box, zero mass, jitter", _specials, 2400, 800, 0.7, 2, 0.0),
    ("merges that shift indices", _merging, 2400, 800, 0.5, 4, 8.0),
    ("uniform cloud, unsnapped f64, theta 1.2, tall window", lambda: scenes.make_uniform_random(350, 0.5, 600, 1900, seed=15), 600, 1900, 1.2, 2, 0.0),
]


@pytest.mark.parametrize("name,gen,W,H,theta,steps,merge", CASES, ids=[c[0] for c in CASES])
def test_both_readings_of_the_reference_agree_bit_for_bit(oracle_lib, name, gen, W, H, theta, steps, merge):
    scene = gen()
    _configure(W, H, theta)
    # --- one evaluation: accelerations, counts, every cell
    o = make_engine(oracle_lib, scene, W, H, theta=theta, merge_min_dist=merge)
    ax, ay = o.compute_accelerations()
    oc = o.counters()
    cells_o = o.tree()
    state_o = o.get_bodies()                            # (a jittering build has moved bodies)
    bodies = _bodies(scene)
    p = second.PhysicsEngine(bodies)
    p.mergeMinDist = merge
    root = p.build_tree()
    p.compute_accelerations(root)
    assert _same(p.ax, ax) and _same(p.ay, ay)
    assert (p.interactions, p.opened) == (oc["interactions"], oc["opened"])
    if "jitter" in name:                                # the scene does reach those branches
        moved = int(((state_o[0] != scene[0]) | (state_o[1] != scene[1])).sum())      # BH.kt:146-151 mutated them
        dropped = len(scene[0]) - int((cells_o["body"] >= 0).sum())                    # BH.kt:126 rejected them
        assert moved >= 4 and dropped >= 5 and int(np.isnan(ax).sum()) == 4, (moved, dropped)
    cells_p = _cells(root, bodies)
    assert len(cells_p["cx"]) == len(cells_o["cx"])
    for k in ("cx", "cy", "h", "mass", "comx", "comy"):
        assert _same(cells_p[k], cells_o[k]), k
    assert (np.array(cells_p["body"], np.int32) == cells_o["body"]).all()
    for a, b in zip(_state(bodies), state_o):
        assert _same(a, b)
    o.close()
    # --- whole steps (merge rule included where enabled), from the original scene
    o = make_engine(oracle_lib, scene, W, H, theta=theta, merge_min_dist=merge)
    bodies = _bodies(scene)
    p = second.PhysicsEngine(bodies)
    p.mergeMinDist = merge
    originals = list(bodies)
    for _ in range(steps):
        o.step(1)
        p.step()
        so = o.get_bodies()
        assert len(p.bodies) == o.n
        for a, b in zip(_state(p.bodies), so):
            assert _same(a, b)
    survivors = np.array([originals.index(b) for b in p.bodies], np.int32)     # identity, as the UI holds references
    assert (survivors == o.get_origin()).all()
    if "merges" in name:
        assert o.n < len(scene[0]) - 40
    o.close()


def _fuzz_scene(seed):
    """tiny adversarial scenes: tight clusters (deep trees, jitter regime), exact duplicates, bodies on cell edges and
    outside the box, zero and heavy masses, fast bodies"""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 60))
    W, H = [(2400, 800), (800, 2400), (512, 512), (37, 91)][seed % 4]
    half = max(W, H) / 2.0 + 2.0
    x = rng.uniform(W / 2.0 - half, W / 2.0 + half, n)
    y = rng.uniform(H / 2.0 - half, H / 2.0 + half, n)
    k = int(rng.integers(0, n + 1))                       # a tight cluster
    if k:
        cx0, cy0 = rng.uniform(0.2 * W, 0.8 * W), rng.uniform(0.2 * H, 0.8 * H)
        r = 10.0 ** rng.uniform(-6, 1)
        x[:k], y[:k] = cx0 + rng.normal(0, r, k), cy0 + rng.normal(0, r, k)
    for _ in range(int(rng.integers(0, 4))):              # exact duplicates
        if n >= 2:
            a, b = rng.integers(0, n, 2)
            x[b], y[b] = x[a], y[a]
    if n >= 3:                                            # on the cell edges / corners of the root and its children
        x[n - 1], y[n - 1] = W / 2.0, H / 2.0
        x[n - 2] = W / 2.0 - half
        y[n - 3] = H / 2.0 + half                         # (the upper edge is outside: half-open)
    if seed % 3 == 0 and n >= 4:
        x[n - 4] = W / 2.0 + half + rng.uniform(0, 50)    # outside
    vx, vy = rng.normal(0, 300, n), rng.normal(0, 300, n)
    m = rng.choice([0.0, 0.5, 1.0, 3.0, 5000.0, 60000.0], n, p=[0.08, 0.3, 0.3, 0.2, 0.08, 0.04])
    if seed % 2:
        x, y = x.astype(np.float32).astype(np.float64), y.astype(np.float32).astype(np.float64)
    return (x, y, vx, vy, m), W, H


@pytest.mark.parametrize("seed", range(48))
def test_fuzzed_tiny_scenes_agree_bit_for_bit(oracle_lib, seed):
    scene, W, H = _fuzz_scene(seed)
    theta = [0.0, 0.3, 0.5, 1.0, 1.6][seed % 5]
    merge = [0.0, 8.0, 40.0][seed % 3]
    _configure(W, H, theta, G=[80.0, 1.0, -5.0][seed % 3], dt=[0.005, 0.02, -0.005][seed % 3])
    o = make_engine(oracle_lib, scene, W, H, theta=theta, merge_min_dist=merge, G=second.Config.G, dt=second.Config.DT)
    bodies = _bodies(scene)
    p = second.PhysicsEngine(bodies)
    p.mergeMinDist = merge
    originals = list(bodies)
    # evaluation + cells (a jittering build moves bodies in both readings alike)
    ax, ay = o.compute_accelerations()
    root = p.build_tree()
    p.compute_accelerations(root)
    assert _same(p.ax[:len(bodies)], ax) and _same(p.ay[:len(bodies)], ay)
    cells_o, cells_p = o.tree(), _cells(root, bodies)
    for k in ("cx", "cy", "h", "mass", "comx", "comy"):
        assert _same(cells_p[k], cells_o[k]), k
    assert (np.array(cells_p["body"], np.int32) == cells_o["body"]).all()
    for _ in range(3):
        o.step(1)
        p.step()
        assert len(p.bodies) == o.n
        for a, b in zip(_state(p.bodies), o.get_bodies()):
            assert _same(a, b)
    assert (np.array([originals.index(b) for b in p.bodies], np.int32) == o.get_origin()).all()
    o.close()
