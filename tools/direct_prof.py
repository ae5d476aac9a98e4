#!/usr/bin/env python
"""Two launches of the tiled all-pairs direct sum (k_direct) on 100k bodies of the bench cloud (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bh_b200  # noqa: E402
from bh_b200 import scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
scene = scenes.make_uniform_random(n, 0.5, 2400, 800, seed=3)
e = bh_b200.NativeEngine(device=0, capacity_hint=n)
e.set_params(theta=0.5, merge_min_dist=0.0)
e.set_bodies(*scene)
e.direct_sum()
e.direct_sum()
print("k_direct ms", e.counters()["ms_direct"])
