"""One-GPU run at a given size (default 100M bodies, density-scaled window): per-phase device times."""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bh_b200
from bh_b200 import scenes

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
s = math.sqrt(n / 1e6)
W, H = int(round(2400 * s)), int(round(800 * s))
t0 = time.time()
scene = scenes.make_uniform_random(n, 0.5, W, H, seed=3)
t_gen = time.time() - t0
e = bh_b200.NativeEngine(capacity_hint=n)
e.set_window(W, H)
e.set_params(theta=0.5, merge_min_dist=0.0)
t0 = time.time()
e.set_bodies(*scene)
t_up = time.time() - t0
e.step(1)
e.reset_counters()
e.step(steps)
c = e.counters()
print(json.dumps({"n": n, "window": [W, H], "ms_per_step": c["ms_step_call"] / steps, "build_ms_per_eval": c["ms_build"] / (2 * steps),
                  "walk_ms_per_eval": c["ms_walk"] / (2 * steps), "interactions_per_s": c["total_interactions"] / (c["ms_step_call"] * 1e-3),
                  "interactions_per_body": c["interactions"] / n, "cells": c["n_cells"], "max_depth": c["max_depth"],
                  "jitter_bodies": c["n_jitter_bodies"], "gen_s": t_gen, "upload_s": t_up}))
