"""The engine's per-element algorithm core (csrc/bh_core.h: key descent, preorder layout
formulas, galloping searches, bottom-up f64 COM, guarded FP32 opening test) executed
serially on the CPU by tests/emul/ and compared with the oracle.  This covers the tree
logic where no GPU exists; the kernels that run it in parallel are tested with -m gpu."""
import ctypes as C

import numpy as np
import pytest

from bh_b200 import scenes
from conftest import ACC_TOL, acc_errors, key_levels, leaf_paths, make_engine

D = C.POINTER(C.c_double)
I32 = C.POINTER(C.c_int32)
I64 = C.POINTER(C.c_int64)
U64 = C.POINTER(C.c_uint64)


def _dp(a):
    return a.ctypes.data_as(D)


def emul_acc(emul, scene, p, theta):
    x, y, vx, vy, m = (np.ascontiguousarray(a, np.float64) for a in scene)
    n = len(x)
    out = dict(ax=np.empty(n), ay=np.empty(n), ci=np.empty(n, np.int32), co=np.empty(n, np.int32),
               key=np.empty(n, np.uint64), depth=np.empty(n, np.int32), order=np.empty(n, np.int32),
               stats=np.zeros(8, np.int64))
    emul.bh_emul_accelerations(n, _dp(x), _dp(y), _dp(m), C.c_double(p.root_cx), C.c_double(p.root_cy),
                               C.c_double(p.root_half), C.c_double(theta), C.c_double(p.soft2), C.c_double(p.G),
                               _dp(out["ax"]), _dp(out["ay"]), out["ci"].ctypes.data_as(I32), out["co"].ctypes.data_as(I32),
                               out["key"].ctypes.data_as(U64), out["depth"].ctypes.data_as(I32),
                               out["order"].ctypes.data_as(I32), out["stats"].ctypes.data_as(I64))
    return out


def emul_tree(emul, scene, p, k):
    x, y, vx, vy, m = (np.ascontiguousarray(a, np.float64) for a in scene)
    t = {q: np.empty(k) for q in ("cx", "cy", "h", "mass", "comx", "comy")}
    body = np.empty(k, np.int32)
    nc = C.c_int64()
    emul.bh_emul_tree(len(x), _dp(x), _dp(y), _dp(m), C.c_double(p.root_cx), C.c_double(p.root_cy), C.c_double(p.root_half),
                      C.c_int64(k), C.byref(nc), _dp(t["cx"]), _dp(t["cy"]), _dp(t["h"]), _dp(t["mass"]), _dp(t["comx"]),
                      _dp(t["comy"]), body.ctypes.data_as(I32))
    t["body"] = body
    return nc.value, t


def _oob_scene():
    s = scenes.make_uniform_random(3000, 0.5, seed=5)
    s[0][:50] += 3000.0            # 50 bodies outside the root box
    s[4][100:120] = 0.0            # zero-mass bodies
    return s


CASES = [
    ("two-disk θ0.5", lambda: scenes.snap_f32(scenes.default_two_disks()), 2400, 800, 0.5),
    ("two-disk θ0.3", lambda: scenes.snap_f32(scenes.default_two_disks(seed=11)), 2400, 800, 0.3),
    ("uniform 20k unsnapped θ1.6", lambda: scenes.make_uniform_random(20000, 0.5), 2400, 800, 1.6),
    ("uniform 20k θ0.2", lambda: scenes.make_uniform_random(20000, 0.5, seed=4), 2400, 800, 0.2),
    ("out-of-box + zero mass", _oob_scene, 2400, 800, 1.0),
    ("big box", lambda: scenes.default_two_disks(32768, 32768, 8000, 2000, scale=28.0, seed=4), 32768, 32768, 0.5),
    ("n=2", lambda: scenes.make_uniform_random(2, 0.5), 2400, 800, 0.5),
    ("n=1", lambda: scenes.make_uniform_random(1, 0.5), 2400, 800, 0.5),
    ("theta=0", lambda: scenes.make_uniform_random(300, 0.5, seed=8), 2400, 800, 0.0),
]


@pytest.mark.parametrize("name,gen,W,H,theta", CASES, ids=[c[0] for c in CASES])
def test_core_matches_oracle(oracle_lib, emul_lib, name, gen, W, H, theta):
    scene = gen()
    n = len(scene[0])
    o = make_engine(oracle_lib, scene, W, H, flags=1, theta=theta)
    p = o.params
    ax, ay = o.compute_accelerations()
    ci, co = o.body_counts()
    depth, path = leaf_paths(oracle_lib, o)
    g = emul_acc(emul_lib, scene, p, theta)
    # Morton order and cell assignment: bit-exact
    L = key_levels(p.root_half)
    assert g["stats"][6] == 1 and g["stats"][7] == 0      # closed-form keys == literal descent
    assert (g["depth"] == depth).all()
    inb = depth >= 0
    assert ((g["key"][inb] >> (2 * (L - depth[inb])).astype(np.uint64)) == path[inb]).all()
    assert (g["key"][~inb] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    # per-body decisions identical: interaction and opened-cell counts are integers
    assert (g["ci"] == ci).all() and (g["co"] == co).all()
    # accelerations (host FP32 arithmetic here; the device numbers are gated in test_gpu_parity)
    ok = np.isfinite(ax)
    assert (np.isnan(g["ax"]) == ~ok).all()
    if ok.sum() > 2:
        s = acc_errors(ax[ok], ay[ok], g["ax"][ok], g["ay"][ok])
        assert s["normwise"] < 1e-6 and s["max_floored"] < 2 * ACC_TOL, s
    # every cell of visitQuads, bit-exact incl. the f64 centres of mass
    to = o.tree()
    k, tg = emul_tree(emul_lib, scene, p, len(to["cx"]))
    assert k == len(to["cx"])
    for q in to:
        assert (tg[q] == to[q]).all(), q


def test_insertion_order_invariance(oracle_lib):
    """Outside the jitter regime the reference tree does not depend on insertion order
    (BH.kt:125-137), which is what lets a sort-based build reproduce it."""
    s = scenes.make_uniform_random(5000, 0.5, seed=2)
    perm = np.random.default_rng(0).permutation(5000)
    a = make_engine(oracle_lib, s).tree()
    b = make_engine(oracle_lib, tuple(v[perm] for v in s)).tree()
    for q in ("cx", "cy", "h", "mass"):
        assert (a[q] == b[q]).all()


# ---- jitter regime (BarnesHutAlg.kt:145-156) ---------------------------------------------------
def _jitter_scenes():
    """Bodies that share a cell with h < 1e-3 (depth 21 of the default root, side 1.146e-3):
    coincident pairs, near-coincident pairs/triples/quads in every mantissa-LSB combination,
    clusters whose members are far apart in list order, a cluster with a zero-mass body."""
    rng = np.random.default_rng(77)
    base = scenes.make_uniform_random(1500, 0.5, seed=21)
    x, y, vx, vy, m = [a.copy() for a in base]
    def put(idx, x0, y0, offs):
        for k, (dx, dy) in zip(idx, offs):
            x[k], y[k] = x0 + dx, y0 + dy
    put([3, 4], 100.25, 100.25, [(0, 0), (0, 0)])                                   # exactly coincident
    put([10, 900], 333.333, 250.5, [(0, 0), (1e-5, -2e-5)])                         # far apart in list order
    put([20, 21, 22], 1500.1, 400.2, [(0, 0), (3e-5, 1e-5), (-2e-5, 4e-5)])         # triple
    put([30, 31, 32, 33], 700.7, 123.4, [(0, 0), (1e-4, 0), (0, 1e-4), (1e-4, 1e-4)])
    for q in range(40):                                                              # random LSB parities
        i0 = 100 + 3 * q
        cx0, cy0 = rng.uniform(50, 2350), rng.uniform(50, 750)
        put([i0, i0 + 1, i0 + 2], cx0, cy0, rng.uniform(-2e-4, 2e-4, (3, 2)))
    put([400, 401, 402, 403, 404, 405, 406], 1200.0, 400.0, rng.uniform(-1e-4, 1e-4, (7, 2)))   # centre of the box
    m[401] = 0.0
    return [("mixed clusters", (x, y, vx, vy, m))]


@pytest.mark.parametrize("name,scene", _jitter_scenes(), ids=[s[0] for s in _jitter_scenes()])
@pytest.mark.parametrize("theta", [0.5, 0.0])
def test_jitter_regime_matches_oracle(oracle_lib, emul_lib, name, scene, theta):
    o = make_engine(oracle_lib, scene, 2400, 800, flags=1, theta=theta)
    p = o.params
    ax, ay = o.compute_accelerations()        # buildTree() mutates the clustered bodies
    ci, co = o.body_counts()
    depth, path = leaf_paths(oracle_lib, o)
    ox, oy, *_ = o.get_bodies()
    assert (ox != scene[0]).sum() > 50        # the regime was really exercised
    assert (depth < 0).sum() > 10             # ... and dropped bodies exist
    x, y, vx, vy, m = (np.ascontiguousarray(a, np.float64) for a in scene)
    n = len(x)
    gx, gy, st = np.empty(n), np.empty(n), np.zeros(2, np.int64)
    emul_lib.bh_emul_positions_after_build(n, _dp(x), _dp(y), _dp(m), C.c_double(p.root_cx), C.c_double(p.root_cy),
                                           C.c_double(p.root_half), _dp(gx), _dp(gy), st.ctypes.data_as(I64))
    assert st[1] == 0
    assert (gx == ox).all() and (gy == oy).all()          # mutated coordinates, bit-exact
    assert st[0] == int((depth < 0).sum())                # the same bodies were dropped
    g = emul_acc(emul_lib, scene, p, theta)
    assert (g["depth"] == depth).all()
    assert (g["ci"] == ci).all() and (g["co"] == co).all()
    ok = np.isfinite(ax)
    s = acc_errors(ax[ok], ay[ok], g["ax"][ok], g["ay"][ok])
    assert s["normwise"] < 1e-6, s
    to = o.tree()
    k, tg = emul_tree(emul_lib, scene, p, len(to["cx"]))
    assert k == len(to["cx"])
    for q in to:
        assert (tg[q] == to[q]).all(), q


def test_closed_form_keys_on_cell_boundaries(oracle_lib, emul_lib):
    """bh_morton_key_grid == the literal descent of BH.kt:153-155 for bodies ON cell boundaries of
    every level and one ulp to either side (half = 1202 is not a power of two); a
    root box whose grid is not exactly representable falls back to the descent."""
    rng = np.random.default_rng(5)
    half, cx, cy = 1202.0, 1200.0, 400.0
    xs, ys = [], []
    for level in range(1, 22):
        w = 2 * half / 2 ** level
        k = rng.integers(0, 2 ** level, 40)
        bx = (cx - half) + k * w
        by = (cy - half) + rng.integers(0, 2 ** level, 40) * w
        for d in (-1, 0, 1):
            xs.append(np.nextafter(bx, bx + d) if d else bx)
            ys.append(np.nextafter(by, by + d) if d else by)
    x = np.concatenate(xs); y = np.concatenate(ys)
    scene = (x, y, np.zeros_like(x), np.zeros_like(x), np.ones_like(x))
    o = make_engine(oracle_lib, scene, 2400, 800, flags=1, theta=0.5)
    g = emul_acc(emul_lib, scene, o.params, 0.5)
    assert g["stats"][6] == 1 and g["stats"][7] == 0
    o.compute_accelerations()
    depth, path = leaf_paths(oracle_lib, o)
    inb = depth >= 0
    L = key_levels(o.params.root_half)
    assert ((g["key"][inb] >> (2 * (L - depth[inb])).astype(np.uint64)) == path[inb]).all()
    # not exactly representable grid -> literal descent
    o.set_params(root_cx=1200.1, root_half=1000.3)
    g2 = emul_acc(emul_lib, scene, o.params, 0.5)
    assert g2["stats"][6] == 0


@pytest.mark.parametrize("seed", range(48))
def test_fuzzed_tiny_scenes_through_the_algorithm_core(oracle_lib, emul_lib, seed):
    """The adversarial tiny scenes of tests/test_oracle_second_reading.py (tight clusters and exact duplicates in the
    jitter regime, bodies on cell edges and outside the box, zero and heavy masses, four window shapes) through the
    PRODUCT's algorithm core (bh_core.h on the host): mutated positions, dropped bodies, leaf depths, per-body
    decisions and every visitQuads cell equal to the oracle's."""
    from test_oracle_second_reading import _fuzz_scene
    scene, W, H = _fuzz_scene(seed)
    theta = [0.0, 0.3, 0.5, 1.0, 1.6][seed % 5]
    o = make_engine(oracle_lib, scene, W, H, flags=1, theta=theta)
    p = o.params
    ax, ay = o.compute_accelerations()
    ci, co = o.body_counts()
    depth, path = leaf_paths(oracle_lib, o)
    ox, oy, *_ = o.get_bodies()
    x, y, vx, vy, m = (np.ascontiguousarray(a, np.float64) for a in scene)
    n = len(x)
    gx, gy, st = np.empty(n), np.empty(n), np.zeros(2, np.int64)
    emul_lib.bh_emul_positions_after_build(n, _dp(x), _dp(y), _dp(m), C.c_double(p.root_cx), C.c_double(p.root_cy),
                                           C.c_double(p.root_half), _dp(gx), _dp(gy), st.ctypes.data_as(I64))
    assert st[1] == 0
    assert (gx == ox).all() and (gy == oy).all()
    cx, cy, h = p.root_cx, p.root_cy, p.root_half
    outside = ~((x >= cx - h) & (x < cx + h) & (y >= cy - h) & (y < cy + h))           # BH.kt:61-62 on the ORIGINAL positions
    assert st[0] == int((depth < 0).sum()) - int(outside.sum())                         # bodies the jitter replay pushed out of their cell
    g = emul_acc(emul_lib, scene, p, theta)
    assert (g["depth"] == depth).all()
    assert (g["ci"] == ci).all() and (g["co"] == co).all()
    ok = np.isfinite(ax)
    assert (np.isnan(g["ax"]) == ~ok).all()
    to = o.tree()
    k, tg = emul_tree(emul_lib, scene, p, len(to["cx"]))
    assert k == len(to["cx"])
    for q in to:
        assert (tg[q] == to[q]).all(), q
