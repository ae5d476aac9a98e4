#!/usr/bin/env python
"""Launch each walk variant twice on the 1M-body bench cloud (for `ncu -k regex:k_walk`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bh_b200  # noqa: E402
from bh_b200 import scenes  # noqa: E402

variants = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(1, 0), (2, 0), (4, 0), (2, 1)]
n = 1_000_000
scene = scenes.make_uniform_random(n, 0.5, 2400, 800, seed=3)
for g, acc in variants:
    os.environ["BH_WALK_G"], os.environ["BH_WALK_ACC"] = str(g), str(acc)
    e = bh_b200.NativeEngine(device=0, capacity_hint=n)
    e.set_params(theta=0.5, merge_min_dist=0.0)
    e.set_bodies(*scene)
    for _ in range(2):
        e.compute_accelerations()
    print("variant", g, acc, e.counters()["ms_walk"] / 2, flush=True)
    e.close()
