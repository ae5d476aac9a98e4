#!/usr/bin/env python
"""A few force evaluations on the bench cloud (for `ncu --metrics gpu__time_duration.sum`: per-kernel times of a build)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bh_b200  # noqa: E402
from bh_b200 import scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
s = (n / 1_000_000) ** 0.5
W, H = int(round(2400 * s)), int(round(800 * s))
scene = scenes.make_uniform_random(n, 0.5, W, H, seed=3)
e = bh_b200.NativeEngine(device=0, capacity_hint=n)
e.set_window(W, H)
e.set_params(theta=0.5, merge_min_dist=0.0)
e.set_bodies(*scene)
for _ in range(4):
    e.compute_accelerations()
c = e.counters()
print("build_ms", c["ms_build"] / 4, "walk_ms", c["ms_walk"] / 4)
