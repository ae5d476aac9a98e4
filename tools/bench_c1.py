"""configs[0] of BASELINE.json: the reference's own two-disk scene (12,500 bodies, merge rule on,
reference defaults except theta = 0.5) — steps/s of the CUDA engine and of the CPU oracle."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bh_b200
from bh_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(lib, steps, warm):
    e = bh_b200.NativeEngine(lib=lib)
    e.set_params(theta=0.5)                      # merge 4000 / 8 px stays on (reference default)
    e.set_bodies(*scenes.snap_f32(scenes.default_two_disks(seed=1)))
    e.step(warm)
    e.reset_counters()
    t0 = time.perf_counter()
    e.step(steps)
    dt = time.perf_counter() - t0
    c = e.counters()
    return {"steps_per_s": steps / dt, "ms_per_step": dt / steps * 1e3, "interactions_per_s": c["total_interactions"] / dt,
            "bodies_left": e.n, "merged": c["total_merged"], "launches_per_step": c["kernel_launches"] / steps,
            "device_ms_per_step": {k: c[k] / steps for k in ("ms_build", "ms_walk", "ms_integrate", "ms_merge")},
            "device_ms_per_step_total": c["ms_step_call"] / steps}


if __name__ == "__main__":
    out = {"gpu": run(bh_b200.load_cuda_library(), 500, 20)}
    if "--cpu" in sys.argv:
        out["cpu_oracle"] = run(bh_b200.bind(os.path.join(ROOT, "oracle", "libbh_ref.so")), 30, 2)
    print(json.dumps(out))
