"""Generates tests/golden/*.npz from the ORACLE (oracle/libbh_ref.so).

The reference itself cannot run in the build container (no JVM) and ships no golden
vectors, so these fixtures pin the oracle's current behaviour: they catch regressions of
the restatement and give the GPU tests a committed, reference-free target.  Re-run with
    python tests/golden/make_golden.py
only when oracle/bh_ref.cpp changes on purpose."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bh_b200  # noqa: E402
from bh_b200 import scenes  # noqa: E402
from conftest import leaf_paths, make_engine  # noqa: E402

CASES = {
    # name: (scene, W, H, theta, steps)
    "two_disks_2k": (lambda: scenes.snap_f32(scenes.default_two_disks(n1=1600, n2=400, seed=21)), 2400, 800, 0.5, 5),
    "cloud_3k_theta03": (lambda: scenes.snap_f32(scenes.make_uniform_random(3000, 0.5, seed=22)), 2400, 800, 0.3, 3),
    "mixed_bigbox": (lambda: scenes.mixed_mass_stress(4, 500, 6, 16384, 16384, seed=23), 16384, 16384, 0.8, 2),
}


def main():
    lib = bh_b200.bind(os.path.join(ROOT, "oracle", "libbh_ref.so"))
    for name, (gen, W, H, theta, steps) in CASES.items():
        scene = gen()
        e = make_engine(lib, scene, W, H, flags=1, theta=theta)
        ax, ay = e.compute_accelerations()
        ci, co = e.body_counts()
        depth, path = leaf_paths(lib, e)
        t = e.tree()
        e.step(steps)
        fx, fy, fvx, fvy, fm = e.get_bodies()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=scene[0], y=scene[1], vx=scene[2], vy=scene[3], m=scene[4],
                            W=W, H=H, theta=theta, steps=steps, ax=ax, ay=ay, interactions=ci, opened=co, depth=depth,
                            path=path, tree_cx=t["cx"], tree_cy=t["cy"], tree_h=t["h"], tree_mass=t["mass"],
                            tree_comx=t["comx"], tree_comy=t["comy"], tree_body=t["body"],
                            fx=fx, fy=fy, fvx=fvx, fvy=fvy)
        print(name, len(scene[0]), "bodies,", len(t["cx"]), "cells")


if __name__ == "__main__":
    main()
