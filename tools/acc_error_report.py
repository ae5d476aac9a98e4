#!/usr/bin/env python
"""Distribution of the per-body acceleration error of the CUDA walk against the f64 oracle (the C++ port of
BarnesHutAlg.kt:215-259, identical accept/open decisions), per config:
    python tools/acc_error_report.py [C1 C2 C3 ...] > profiles/r02_acc_error.json
Error of body i: |a_gpu - a_ref| / max(|a_ref|, floor * rms|a_ref|), reported unfloored and for floors 1e-3
and 0.05; plus the theta sweep of C2 against the DEVICE direct sum at the full 1M bodies."""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bh_b200  # noqa: E402
from bh_b200 import scenes  # noqa: E402

oracle = bh_b200.bind(os.path.join(ROOT, "oracle", "libbh_ref.so"))


def dist(ax, ay, gx, gy):
    ok = np.isfinite(ax) & np.isfinite(ay)
    ax, ay, gx, gy = ax[ok], ay[ok], gx[ok], gy[ok]
    a = np.hypot(ax, ay)
    err = np.hypot(gx - ax, gy - ay)
    rms = float(np.sqrt(np.mean(a * a)))
    rel = err / np.maximum(a, 1e-300)
    out = {"bodies": int(len(a)), "rms_acc": rms, "normwise": float(np.sqrt((err ** 2).sum() / (a ** 2).sum())),
           "abs_err_over_rms": {"median": float(np.median(err) / rms), "p99": float(np.quantile(err, 0.99) / rms), "max": float(err.max() / rms)},
           "unfloored": {"median": float(np.median(rel)), "p90": float(np.quantile(rel, 0.9)), "p99": float(np.quantile(rel, 0.99)),
                         "p99.9": float(np.quantile(rel, 0.999)), "max": float(rel.max()),
                         "fraction_above_1e-5": float((rel > 1e-5).mean()), "fraction_above_1e-4": float((rel > 1e-4).mean())}}
    for fl in (1e-3, 0.05):
        r = err / np.maximum(a, fl * rms)
        out[f"floor_{fl:g}_rms"] = {"max": float(r.max()), "p99.9": float(np.quantile(r, 0.999)), "fraction_above_1e-5": float((r > 1e-5).mean())}
    return out


def engines(scene, W, H, theta):
    g = bh_b200.NativeEngine(device=0, flags=bh_b200.BH_FLAG_BODY_COUNTS)
    o = bh_b200.NativeEngine(lib=oracle, flags=bh_b200.BH_FLAG_BODY_COUNTS)
    for e in (g, o):
        e.set_window(W, H)
        e.set_params(theta=theta, merge_min_dist=0.0)
        e.set_bodies(*scene)
    return g, o


def vs_oracle(name, scene, W, H, theta):
    g, o = engines(scene, W, H, theta)
    gx, gy = g.compute_accelerations()
    ox, oy = o.compute_accelerations()
    same = bool((g.body_counts()[0] == o.body_counts()[0]).all() and (g.body_counts()[1] == o.body_counts()[1]).all())
    rec = {"config": name, "theta": theta, "window": [W, H], "decisions_equal_to_oracle": same, "interactions": g.counters()["interactions"]}
    rec.update(dist(ox, oy, gx, gy))
    g.close(); o.close()
    return rec


want = sys.argv[1:] or ["C1", "C2", "C3"]
out = []
if "C1" in want:
    out.append(vs_oracle("C1 reference two-disk scene, 12,500 bodies", scenes.snap_f32(scenes.default_two_disks(seed=1)), 2400, 800, 0.5))
if "C2" in want:
    cloud = scenes.make_uniform_random(1_000_000, 0.5, 2400, 800, seed=3)
    out.append(vs_oracle("C2 1M-body uniform cloud", cloud, 2400, 800, 0.5))
    g = bh_b200.NativeEngine(device=0)
    g.set_window(2400, 800)
    g.set_params(theta=0.5, merge_min_dist=0.0)
    g.set_bodies(*cloud)
    dx, dy = g.direct_sum()
    sweep = []
    for th in (0.2, 0.3, 0.5, 0.8, 1.0, 1.3, 1.6):
        g.set_params(theta=th)
        bx, by = g.compute_accelerations()
        d = dist(dx, dy, bx, by)
        sweep.append({"theta": th, "interactions_per_body": g.counters()["interactions"] / 1e6, "median": d["unfloored"]["median"],
                      "p99": d["unfloored"]["p99"], "max_floor_1e-3_rms": d["floor_0.001_rms"]["max"], "normwise": d["normwise"]})
    out.append({"config": "C2 theta sweep at 1M bodies: Barnes-Hut vs the device direct sum (k_direct)", "sweep": sweep})
    g.close()
if "C3" in want:
    k = math.sqrt(800.0)
    W = H = 32768
    a = scenes.make_galaxy_disk(8_000_000, x=W * 0.5, y=H * 0.5, r=300.0 * k, central_mass=50_000.0, total_satellite_mass=5_000.0, seed=4)
    b = scenes.make_galaxy_disk(2_000_000, x=W * 0.5, y=H * 0.5 - 240.0 * k, vx=-50.0, r=100.0 * k, central_mass=5_000.0, total_satellite_mass=500.0, seed=5)
    scene = scenes.snap_f32(tuple(np.concatenate([p, q]) for p, q in zip(a, b)))
    out.append(vs_oracle("C3 10M-body two-disk merger", scene, W, H, 0.5))
print(json.dumps({"what": "per-body acceleration error of the CUDA walk (FP32 interactions) vs the f64 oracle on identical inputs", "records": out}, indent=1))
