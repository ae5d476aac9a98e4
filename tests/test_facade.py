"""The Kotlin-API façade (integration/kotlin/BarnesHutAlg.kt; Python twin engine.py::PhysicsEngine)
through a MERGING run — the path that hid the round-1 bug (state re-uploaded before the write-back).

`-m "not gpu"`: the twin drives the oracle library (same C ABI) and is compared with the oracle driven
directly.  `-m gpu`: the twin drives the CUDA library and is compared with the oracle after every step."""
import os
import re

import numpy as np
import pytest

import bh_b200
from bh_b200 import Acc, Body, Config, PhysicsEngine, Quad, scenes
from conftest import ROOT, make_engine


def _merge_scene(seed=41, n1=1500, n2=400, k=40):
    """Heavy centres (m = 50,000 / 5,000 > mergeMaxMass) with satellites inside the 8 px radius and a
    third heavy body that is eaten itself (BH.kt:463-532)."""
    s = list(scenes.snap_f32(scenes.default_two_disks(n1=n1, n2=n2, seed=seed)))
    rng = np.random.default_rng(seed)
    ang, rad = rng.uniform(0, 2 * np.pi, k), rng.uniform(1.0, 7.0, k)
    s[0][5:5 + k] = s[0][0] + rad * np.cos(ang)
    s[1][5:5 + k] = s[1][0] + rad * np.sin(ang)
    s[0][n1 + 7], s[1][n1 + 7], s[4][n1 + 7] = s[0][n1] + 5.0, s[1][n1] + 1.0, 4500.0
    return tuple(np.ascontiguousarray(a) for a in s)


def _run_facade_against_oracle(lib, oracle_lib, steps, exact):
    Config.reset()
    Config.theta = 0.5
    scene = _merge_scene()
    bodies = [Body(*[float(v[i]) for v in scene]) for i in range(len(scene[0]))]
    tag = {id(b): i for i, b in enumerate(bodies)}          # the UI holds these objects
    eng = PhysicsEngine(bodies, lib=lib)
    assert eng.mergeMinDist == 8.0 and eng.mergeMaxMass == 4000.0    # BH.kt:315,321: merge on by default
    o = make_engine(oracle_lib, scene, theta=0.5, merge_min_dist=8.0)
    alive = np.arange(len(bodies))
    merged_steps = 0
    for s in range(steps):
        n0 = len(bodies)
        eng.step()
        o.step(1)
        ox, oy, ovx, ovy, om = o.get_bodies()
        alive = alive[o.get_origin()] if len(ox) != n0 else alive
        o.rebase_origin()
        merged_steps += len(ox) != n0
        assert eng.getBodies() is bodies and len(bodies) == len(ox), s
        assert [tag[id(b)] for b in bodies] == alive.tolist(), s           # the SAME objects survive, in list order
        bm = np.array([b.m for b in bodies])
        assert (bm == om).all(), s                                         # grown masses, bit-identical f64 sums
        st = np.array([[b.x, b.y, b.vx, b.vy] for b in bodies]).T
        ref = np.stack([ox, oy, ovx, ovy])
        if exact:
            assert (st == ref).all(), s
        else:
            # FP32 interactions: 1e-5 of the acceleration, i.e. of the velocity change per step (dt = 0.005, up to
            # |a| ~ 1e5 beside the 50,000-mass centre); positions follow
            assert np.abs(st[2:] - ref[2:]).max() < 2e-5 * np.abs(ref[2:]).max() * (s + 1), s
            assert np.abs(st[:2] - ref[:2]).max() < 1e-4 * (s + 1), s
        # the device state IS the host state after every step (the round-1 bug re-seeded the device with t, not t+dt)
        dx, dy, dvx, dvy, dm = eng.native.get_bodies()
        assert (dx == st[0]).all() and (dy == st[1]).all() and (dvx == st[2]).all() and (dvy == st[3]).all() and (dm == bm).all(), s
        assert (eng.native.get_origin() == np.arange(len(bodies))).all(), s  # origin re-based onto the shrunk list
    assert merged_steps >= 1 and len(bodies) < len(scene[0])
    assert abs(sum(b.m for b in bodies) - scene[4].sum()) < 1e-6 * scene[4].sum()   # mass is only moved, never lost
    Config.reset()
    return eng


def test_facade_merging_run_on_the_oracle_library(oracle_lib):
    _run_facade_against_oracle(oracle_lib, oracle_lib, 20, exact=True)


@pytest.mark.gpu
def test_facade_merging_run_on_the_cuda_library(oracle_lib, cuda_lib):
    _run_facade_against_oracle(cuda_lib, oracle_lib, 20, exact=False)


def _host_walk_equals_oracle(lib, oracle_lib):
    """BHTree.accumulateForce of the façade (f64 over the exported cells, BH.kt:215-239) reproduces the
    oracle's accelerations bit for bit: the exported cells ARE the reference's tree."""
    Config.reset()
    Config.theta = 0.5
    scene = scenes.snap_f32(scenes.default_two_disks(n1=300, n2=100, seed=9))
    bodies = [Body(*[float(v[i]) for v in scene]) for i in range(len(scene[0]))]
    eng = PhysicsEngine(bodies, lib=lib)
    eng.mergeMinDist = 0.0
    tree = eng.getTreeForDebug()
    o = make_engine(oracle_lib, scene, theta=0.5)
    ax, ay = o.compute_accelerations()
    acc = Acc()
    for i in range(0, len(bodies), 7):
        acc.reset()
        tree.accumulateForce(bodies[i], Config.theta * Config.theta, acc)
        assert acc.fx / bodies[i].m == ax[i] and acc.fy / bodies[i].m == ay[i], i      # BH.kt:390-391
    with pytest.raises(NotImplementedError):
        tree.insert(bodies[0])
    Config.reset()


def test_host_walk_over_exported_tree_on_the_oracle_library(oracle_lib):
    _host_walk_equals_oracle(oracle_lib, oracle_lib)


@pytest.mark.gpu
def test_host_walk_over_exported_tree_on_the_cuda_library(oracle_lib, cuda_lib):
    _host_walk_equals_oracle(cuda_lib, oracle_lib)


def test_quad_contains_and_child_follow_the_reference():
    q = Quad(1200.0, 400.0, 1202.0)                        # BH.kt:360-361 for the 2400x800 window
    assert q.contains(Body(-2.0, -802.0, 0, 0, 1)) and not q.contains(Body(2402.0, 0.0, 0, 0, 1))   # half-open, BH.kt:61-62
    assert [q.child(k) for k in range(4)] == [Quad(599.0, -201.0, 601.0), Quad(1801.0, -201.0, 601.0),
                                              Quad(599.0, 1001.0, 601.0), Quad(1801.0, 1001.0, 601.0)]   # NW NE SW SE, BH.kt:73-81


def test_kotlin_facade_source_is_in_sync_and_ordered():
    """INTEGRATION.md embeds the .kt file (one source); the write-back precedes the re-base."""
    import subprocess
    import sys
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_integration.py"), "--check"]) == 0
    kt = open(os.path.join(ROOT, "integration", "kotlin", "BarnesHutAlg.kt")).read()
    dl = kt[kt.index("private fun download()"):kt.index("fun getBodies()")]
    assert dl.index("b.m = m[i]") < dl.index("bh_rebase_origin") and "upload()" not in dl
    for name in ("fun contains(", "fun child(", "fun accumulateForce(", "fun insert(", "fun computeMass("):
        assert name in kt, name
    # every native the façade binds is declared by the header
    hdr = open(os.path.join(ROOT, "include", "bh_engine.h")).read()
    for fn in set(re.findall(r"external fun (bh_\w+)", kt)):
        assert re.search(r"\b%s\(" % fn, hdr), fn
