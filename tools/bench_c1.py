#!/usr/bin/env python
"""BASELINE.json configs[0]: the reference's own scene (12,500 bodies, merge rule on, theta 0.5) — steps/s of the
CUDA engine with per-step CUDA graphs (default), with sync-free builds but single launches (BH_GRAPH=0) and with
the host round trip per build (BH_SYNCFREE=0)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bh_b200  # noqa: E402
from bh_b200 import scenes  # noqa: E402

scene = scenes.snap_f32(scenes.default_two_disks(seed=1))
for label, env in (("graph per step", {}), ("sync-free builds, single launches", {"BH_GRAPH": "0"}), ("host round trip per build", {"BH_SYNCFREE": "0"})):
    for k in ("BH_GRAPH", "BH_SYNCFREE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for merge in (8.0, 0.0):
        e = bh_b200.NativeEngine(device=0)
        e.set_params(theta=0.5, merge_min_dist=merge)
        e.set_bodies(*scene)
        e.step(20)
        t0 = time.perf_counter()
        e.step(400)
        dt = time.perf_counter() - t0
        c = e.counters()
        print(json.dumps({"mode": label, "merge_rule": merge > 0, "steps_per_s": 400 / dt, "ms_per_step": dt / 400 * 1e3, "bodies_left": e.n,
                          "kernel_launches_per_step": c["kernel_launches"] / c["total_steps"]}), flush=True)
        e.close()
