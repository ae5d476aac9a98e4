#!/usr/bin/env python
"""A/B timing of the walk variants (k_walk<G, ACC>) on one B200: 1M-body bench cloud, theta 0.5.
    python tools/walk_ab.py [--bodies N] > gpurun_out/walk_ab.json
Prints one JSON line per variant: ms per launch (CUDA events around the kernel, mean of 8 after 2 warm-up),
interactions, and the error of the accelerations against the G = 1 / fold variant."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bh_b200  # noqa: E402
from bh_b200 import scenes  # noqa: E402

n = int(sys.argv[sys.argv.index("--bodies") + 1]) if "--bodies" in sys.argv else 1_000_000
s = (n / 1_000_000) ** 0.5
W, H = int(round(2400 * s)), int(round(800 * s))
scene = scenes.make_uniform_random(n, 0.5, W, H, seed=3)
base = None
for g, acc, aff in [(1, 0, 1), (1, 0, 0), (1, 1, 1), (2, 1, 1), (2, 1, 0)]:
    os.environ["BH_WALK_G"], os.environ["BH_WALK_ACC"], os.environ["BH_WALK_AFFINE"] = str(g), str(acc), str(aff)
    e = bh_b200.NativeEngine(device=0, capacity_hint=n)
    e.set_window(W, H)
    e.set_params(theta=0.5, merge_min_dist=0.0)
    e.set_bodies(*scene)
    for _ in range(2):
        ax, ay = e.compute_accelerations()
    e.reset_counters()
    for _ in range(8):
        e.compute_accelerations()
    c = e.counters()
    if base is None:
        base = (ax, ay)
    a = np.hypot(*base)
    rel = np.hypot(ax - base[0], ay - base[1]) / np.maximum(a, 1e-3 * np.sqrt(np.mean(a * a)))
    print(json.dumps({"G": g, "acc": "f64" if acc else "fold", "affine": aff, "bodies": n, "walk_ms": c["ms_walk"] / 8, "build_ms": c["ms_build"] / 8,
                      "interactions": c["interactions"], "opened": c["opened"],
                      "max_rel_vs_G1_fold": float(rel.max()), "bit_identical_to_G1_fold": bool((ax == base[0]).all() and (ay == base[1]).all())}), flush=True)
    e.close()
