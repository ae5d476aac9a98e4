"""Committed golden vectors (tests/golden/*.npz, made from the oracle by make_golden.py).
CPU: the oracle still reproduces them bit-for-bit.  GPU: the CUDA engine matches them."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT, assert_acc_parity, key_levels, leaf_paths, make_engine

FILES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
IDS = [os.path.basename(f)[:-4] for f in FILES]


def _scene(g):
    return (g["x"], g["y"], g["vx"], g["vy"], g["m"])


@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_oracle_reproduces_golden(oracle_lib, path):
    g = np.load(path)
    e = make_engine(oracle_lib, _scene(g), int(g["W"]), int(g["H"]), flags=1, theta=float(g["theta"]))
    ax, ay = e.compute_accelerations()
    ci, co = e.body_counts()
    assert (ax == g["ax"]).all() and (ay == g["ay"]).all()
    assert (ci == g["interactions"]).all() and (co == g["opened"]).all()
    depth, pth = leaf_paths(oracle_lib, e)
    assert (depth == g["depth"]).all() and (pth == g["path"]).all()
    t = e.tree()
    for k in ("cx", "cy", "h", "mass", "comx", "comy", "body"):
        assert (t[k] == g["tree_" + k]).all(), k
    e.step(int(g["steps"]))
    x, y, vx, vy, m = e.get_bodies()
    assert (x == g["fx"]).all() and (y == g["fy"]).all() and (vx == g["fvx"]).all() and (vy == g["fvy"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=IDS)
def test_cuda_matches_golden(cuda_lib, path):
    g = np.load(path)
    W, H = int(g["W"]), int(g["H"])
    e = make_engine(cuda_lib, _scene(g), W, H, flags=1, theta=float(g["theta"]))
    ax, ay = e.compute_accelerations()
    ci, co = e.body_counts()
    assert (ci == g["interactions"]).all() and (co == g["opened"]).all()       # identical decisions
    assert_acc_parity(g["ax"], g["ay"], ax, ay, os.path.basename(path))
    key, depth, order = e.morton()
    assert (depth == g["depth"]).all()
    L = key_levels(e.params.root_half)
    inb = g["depth"] >= 0
    assert ((key[inb] >> (2 * (L - g["depth"][inb])).astype(np.uint64)) == g["path"][inb]).all()
    t = e.tree()
    for k in ("cx", "cy", "h", "mass", "comx", "comy", "body"):
        assert (t[k] == g["tree_" + k]).all(), k                               # cells + f64 COM bit-exact
    e.step(int(g["steps"]))
    x, y, vx, vy, m = e.get_bodies()
    scale = max(W, H)
    assert np.abs(x - g["fx"]).max() <= 1e-5 * scale * 1e-2 and np.abs(y - g["fy"]).max() <= 1e-5 * scale * 1e-2
    vs = np.sqrt(np.mean(g["fvx"] ** 2 + g["fvy"] ** 2))
    assert np.hypot(vx - g["fvx"], vy - g["fvy"]).max() <= 1e-5 * max(vs, 1e-12)
