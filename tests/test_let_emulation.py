"""Locally essential trees (csrc/bh_let_core.h), emulated for P ranks on the CPU: every rank builds the
tree of its own Morton code range, the level summaries are merged, and each rank assembles its LET (top
tree + own blocks + imported blocks + far roots).  A body's walk over its rank's LET must visit the
same cells in the same order as its walk over the global tree: accelerations BIT-IDENTICAL, opened
counts equal, interaction counts equal (a stray additionally meets its own leaf, which the CUDA engine
resolves by looking the leaf up in the imported block)."""
import ctypes as C

import numpy as np
import pytest

from bh_b200 import scenes

D = C.POINTER(C.c_double)
I32 = C.POINTER(C.c_int32)
I64 = C.POINTER(C.c_int64)


def _dp(a):
    return a.ctypes.data_as(D) if a is not None else None


def _global(emul, scene, W, H, theta):
    x, y, vx, vy, m = (np.ascontiguousarray(a, np.float64) for a in scene)
    n = len(x)
    ax, ay, ci, co, st = np.empty(n), np.empty(n), np.empty(n, np.int32), np.empty(n, np.int32), np.zeros(8, np.int64)
    emul.bh_emul_accelerations(n, _dp(x), _dp(y), _dp(m), C.c_double(W / 2), C.c_double(H / 2), C.c_double(max(W, H) / 2 + 2),
                               C.c_double(theta), C.c_double(1.0), C.c_double(80.0), _dp(ax), _dp(ay), ci.ctypes.data_as(I32),
                               co.ctypes.data_as(I32), None, None, None, st.ctypes.data_as(I64))
    return ax, ay, ci, co, st


def _let(emul, scene, W, H, theta, P, ell, x0=None, y0=None):
    x, y, vx, vy, m = (np.ascontiguousarray(a, np.float64) for a in scene)
    n = len(x)
    ax, ay, ci, co, st = np.empty(n), np.empty(n), np.empty(n, np.int32), np.empty(n, np.int32), np.zeros(8, np.int64)
    rc = emul.bh_emul_let(n, _dp(x0), _dp(y0), _dp(x), _dp(y), _dp(m), C.c_double(W / 2), C.c_double(H / 2),
                          C.c_double(max(W, H) / 2 + 2), C.c_double(theta), C.c_double(1.0), C.c_double(80.0), P, ell, _dp(ax), _dp(ay),
                          ci.ctypes.data_as(I32), co.ctypes.data_as(I32), st.ctypes.data_as(I64))
    assert rc == 0
    return ax, ay, ci, co, st


def _oob_scene():
    s = scenes.make_uniform_random(3000, 0.5, seed=5)
    s[0][:50] += 3000.0
    s[4][100:120] = 0.0
    return s


def _jitter_scene():
    """coincident bodies (equal keys: the jitter regime of BH.kt:145-156) inside the ranks' own ranges"""
    s = list(scenes.snap_f32(scenes.make_uniform_random(6000, 0.5, seed=12)))
    for k in range(0, 60, 3):           # 20 clusters of 3 coincident bodies, spread over the window
        s[0][k + 1] = s[0][k + 2] = s[0][k]
        s[1][k + 1] = s[1][k + 2] = s[1][k]
    return tuple(s)


CASES = [
    ("two-disk P2", lambda: scenes.snap_f32(scenes.default_two_disks()), 0.5, 2, 3, 0.0),
    ("two-disk P8", lambda: scenes.snap_f32(scenes.default_two_disks()), 0.5, 8, 5, 0.0),
    ("two-disk θ0.3 P3", lambda: scenes.snap_f32(scenes.default_two_disks(seed=11)), 0.3, 3, 4, 0.0),
    ("two-disk drifted (strays)", lambda: scenes.snap_f32(scenes.default_two_disks()), 0.5, 4, 4, 0.5),
    ("cloud θ1.6 P8", lambda: scenes.make_uniform_random(20000, 0.5), 1.6, 8, 5, 0.0),
    ("cloud θ0.2 drifted P4", lambda: scenes.make_uniform_random(20000, 0.5, seed=4), 0.2, 4, 3, 2.0),
    ("out-of-box + zero mass", _oob_scene, 1.0, 4, 3, 0.0),
    ("n=2", lambda: scenes.make_uniform_random(2, 0.5), 0.5, 2, 2, 0.0),
    ("n=1", lambda: scenes.make_uniform_random(1, 0.5), 0.5, 2, 2, 0.0),
    ("theta=0", lambda: scenes.make_uniform_random(300, 0.5, seed=8), 0.0, 4, 2, 0.0),
    ("jitter clusters P4", _jitter_scene, 0.5, 4, 3, 0.0),
    ("auto level", lambda: scenes.make_uniform_random(60000, 0.5, seed=9), 0.5, 8, 0, 0.0),
]


@pytest.mark.parametrize("name,gen,theta,P,ell,drift", CASES, ids=[c[0] for c in CASES])
def test_let_walk_is_bit_identical_to_the_global_tree(emul_lib, name, gen, theta, P, ell, drift):
    scene = gen()
    W, H = 2400, 800
    x0 = y0 = None
    if drift > 0:   # home ranks and code ranges from the old positions, trees from the drifted ones
        rng = np.random.default_rng(5)
        x0, y0 = np.ascontiguousarray(scene[0]), np.ascontiguousarray(scene[1])
        scene = (scene[0] + rng.normal(0, drift, len(x0)), scene[1] + rng.normal(0, drift, len(x0))) + tuple(scene[2:])
    g = _global(emul_lib, scene, W, H, theta)
    l = _let(emul_lib, scene, W, H, theta, P, ell, x0, y0)
    assert np.array_equal(g[0], l[0], equal_nan=True) and np.array_equal(g[1], l[1], equal_nan=True)
    assert (g[3] == l[3]).all()                                   # opened cells per body
    d = l[2] - g[2]
    assert ((d == 0) | (d == 1)).all() and np.count_nonzero(d) <= l[4][2]   # +1 only for strays (own leaf not resolved here)
    if drift > 0:
        assert l[4][2] > 0
    assert l[4][3] == 0                                           # no guest inside a jitter cluster
    if theta >= 0.5 and ell == 0:
        assert l[4][4] < g[4][1]                                  # a rank's LET is smaller than the global tree


@pytest.mark.parametrize("theta", [0.2, 0.5, 1.0, 1.6])
@pytest.mark.parametrize("ell", [2, 4, 6])
def test_region_test_never_misses_a_cell_some_body_opens(emul_lib, theta, ell):
    """bh_let_near_region is what decides that a remote subtree is NOT imported: it must be conservative
    with respect to the exact f64 opening test of every body of the rank (bodies outside the root box
    included), for any level of the cut and any theta."""
    rng = np.random.default_rng(100 * ell + int(theta * 10))
    W, H = 2400, 800
    half = max(W, H) / 2 + 2
    # a rank's bodies: two clumps, a sparse strip and a few bodies outside the root box
    x = np.concatenate([rng.normal(600, 60, 300), rng.normal(1900, 20, 200), rng.uniform(0, W, 40), [-500.0, 3000.0, 1200.0]])
    y = np.concatenate([rng.normal(300, 40, 300), rng.normal(700, 25, 200), rng.uniform(380, 420, 40), [400.0, 400.0, 2500.0]])
    # candidate cells everywhere, denser around the clumps (where the answer flips)
    cx = np.concatenate([rng.uniform(-2, 2402, 3000), rng.normal(600, 300, 1500), rng.normal(1900, 200, 1500)])
    cy = np.concatenate([rng.uniform(-802, 1602, 3000), rng.normal(300, 300, 1500), rng.normal(700, 200, 1500)])
    x, y, cx, cy = (np.ascontiguousarray(a, np.float64) for a in (x, y, cx, cy))
    out = np.zeros(2, np.int64)
    bad = emul_lib.bh_emul_let_near_check(len(x), _dp(x), _dp(y), C.c_double(W / 2), C.c_double(H / 2), C.c_double(half),
                                          C.c_double(theta), C.c_double(1.0), ell, len(cx), _dp(cx), _dp(cy), out.ctypes.data_as(I64))
    assert bad == 0
    assert out[0] > 0 and out[1] >= out[0]            # something is opened; "near" is a superset
    if theta >= 0.5 and ell >= 4:
        assert out[1] < len(cx)                       # ... and not everything is near
