"""Quick GPU-vs-oracle sanity run (development aid; the real gates are tests/ -m gpu)."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bh_b200
from bh_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = bh_b200.bind(os.path.join(ROOT, "oracle", "libbh_ref.so"))


def stats(ax, ay, gx, gy):
    a = np.hypot(ax, ay)
    err = np.hypot(gx - ax, gy - ay)
    rms = np.sqrt((a ** 2).mean())
    rel = err / np.maximum(a, 1e-30)
    relf = err / np.maximum(a, 0.03 * rms)
    return dict(median=float(np.median(rel)), p99=float(np.quantile(rel, .99)), max=float(rel.max()),
                max_floor3pct=float(relf.max()), normwise=float(np.sqrt((err ** 2).sum() / (a ** 2).sum())))


def run(scene, W, H, theta, name, check_tree=True):
    n = len(scene[0])
    o = bh_b200.NativeEngine(lib=ref, flags=1)
    g = bh_b200.NativeEngine(flags=1)
    for e in (o, g):
        e.set_window(W, H)
        e.set_params(theta=theta, merge_min_dist=0)
        e.set_bodies(*scene)
    t = time.time(); ax, ay = o.compute_accelerations(); t_o = time.time() - t
    t = time.time(); gx, gy = g.compute_accelerations(); t_g = time.time() - t
    t = time.time(); gx, gy = g.compute_accelerations(); t_g2 = time.time() - t
    oc, gc = o.counters(), g.counters()
    print(f"[{name}] n={n} theta={theta} oracle {t_o*1e3:.1f} ms, gpu first {t_g*1e3:.1f} ms, second {t_g2*1e3:.2f} ms")
    print("   gpu counters", {k: gc[k] for k in ("n_in_tree", "n_out_of_box", "n_cells", "n_internal", "max_depth", "key_levels",
                                                 "n_jitter_bodies", "interactions", "opened", "exact_retests", "ms_build", "ms_walk")})
    print("   oracle interactions", oc["interactions"], "opened", oc["opened"])
    ok = oc["interactions"] == gc["interactions"] and oc["opened"] == gc["opened"]
    oi, oo = o.body_counts(); gi, go = g.body_counts()
    ok &= bool((oi == gi).all() and (oo == go).all())
    print("   counts equal:", ok, " acc err:", stats(ax, ay, gx, gy))
    if check_tree:
        to, tg = o.tree(), g.tree()
        same = all(len(to[k]) == len(tg[k]) and (to[k] == tg[k]).all() for k in to)
        print("   tree cells", len(to["cx"]), "bit-exact:", same)
        ok &= same
        depth = np.empty(n, np.int32); path = np.empty(n, np.uint64)
        ref.bh_ref_get_leaf_paths(o._h, depth.ctypes.data_as(C.POINTER(C.c_int32)), path.ctypes.data_as(C.POINTER(C.c_uint64)))
        key, gd, order = g.morton()
        L = gc["key_levels"]
        inb = depth >= 0
        pre = key[inb] >> (2 * (L - depth[inb])).astype(np.uint64)
        mo = bool((gd == depth).all() and (pre == path[inb]).all())
        srt = np.argsort(key, kind="stable")
        mo &= bool((srt == order).all())
        print("   morton depth/path/order exact:", mo)
        ok &= mo
    return ok


if __name__ == "__main__":
    ok = True
    ok &= run(scenes.snap_f32(scenes.default_two_disks()), 2400, 800, 0.5, "two-disk")
    ok &= run(scenes.make_uniform_random(20000, 0.5), 2400, 800, 0.3, "uniform20k")
    s = scenes.make_uniform_random(3000, 0.5, seed=5); s[0][:50] += 3000
    ok &= run(s, 2400, 800, 1.0, "out-of-box")
    ok &= run(scenes.make_uniform_random(1, 0.5), 2400, 800, 0.5, "n=1")
    ok &= run(scenes.make_uniform_random(2, 0.5), 2400, 800, 0.5, "n=2")
    ok &= run(scenes.make_uniform_random(200000, 0.5, seed=9), 2400, 800, 0.5, "uniform200k")
    ok &= run(scenes.make_uniform_random(1000000, 0.5, seed=3), 2400, 800, 0.5, "uniform1M", check_tree=False)
    # stepping
    sc = scenes.snap_f32(scenes.default_two_disks())
    o = bh_b200.NativeEngine(lib=ref); g = bh_b200.NativeEngine()
    for e in (o, g):
        e.set_params(theta=0.5, merge_min_dist=0); e.set_bodies(*sc)
    e0 = g.energy(); print("energy gpu", e0, "oracle", o.energy())
    o.step(10); g.step(10)
    so, sg = o.get_bodies(), g.get_bodies()
    for nm, a, b in zip("x y vx vy m".split(), so, sg):
        print("   step10", nm, "max abs diff", float(np.abs(a - b).max()), "rel", float(np.abs(a - b).max() / (np.abs(a).max() + 1e-300)))
    # timing of steps at 1M
    g = bh_b200.NativeEngine()
    g.set_params(theta=0.5, merge_min_dist=0); g.set_bodies(*scenes.make_uniform_random(1000000, 0.5, seed=3))
    g.step(3); g.reset_counters()
    t = time.time(); g.step(10); dt = time.time() - t
    c = g.counters()
    print(f"1M uniform: {dt/10*1e3:.2f} ms/step, {c['total_interactions']/dt:.3e} interactions/s; build {c['ms_build']/20:.3f} walk {c['ms_walk']/20:.3f} integ {c['ms_integrate']/10:.3f} ms")
    # direct sum check
    g = bh_b200.NativeEngine(); o = bh_b200.NativeEngine(lib=ref)
    sc = scenes.snap_f32(scenes.default_two_disks())
    for e in (o, g):
        e.set_bodies(*sc)
    ax, ay = o.direct_sum(); gx, gy = g.direct_sum()
    print("direct-sum err", stats(ax, ay, gx, gy))
    print("ALL OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)
