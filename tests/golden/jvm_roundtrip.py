"""Round trip that pins the oracle against the JVM reference (run by someone who has a JDK).

    python tests/golden/jvm_roundtrip.py make  case.bin [--steps 3] [--merge 1]
    (run integration/kotlin/HeadlessDump.kt inside the reference:  case.bin -> case.out)
    python tests/golden/jvm_roundtrip.py check case.bin case.out

`check` replays case.bin through oracle/libbh_ref.so and compares every f64 of the state and every
visitQuads cell with the JVM dump, bit for bit.  `selftest` writes the dump from the oracle itself
(format check; used by tests/test_oracle_known_answers.py)."""
import argparse
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HEAD = "<qiidddii"


def write_case(path, scene, W=2400, H=800, theta=0.5, G=80.0, dt=0.005, steps=3, merge=1):
    n = len(scene[0])
    with open(path, "wb") as f:
        f.write(struct.pack(HEAD, n, W, H, theta, G, dt, steps, merge))
        for a in scene:
            f.write(np.ascontiguousarray(a, "<f8").tobytes())


def read_case(path):
    raw = open(path, "rb").read()
    hs = struct.calcsize(HEAD)
    n, W, H, theta, G, dt, steps, merge = struct.unpack(HEAD, raw[:hs])
    a = np.frombuffer(raw, "<f8", 5 * n, hs).reshape(5, n)
    return tuple(a[k].copy() for k in range(5)), dict(W=W, H=H, theta=theta, G=G, dt=dt, steps=steps, merge=merge)


def read_dump(path):
    raw = open(path, "rb").read()
    n = struct.unpack("<q", raw[:8])[0]
    st = np.frombuffer(raw, "<f8", 5 * n, 8).reshape(5, n)
    off = 8 + 40 * n
    k = struct.unpack("<q", raw[off:off + 8])[0]
    q = np.frombuffer(raw, "<f8", 3 * k, off + 8).reshape(3, k)
    return st, q


def oracle_dump(scene, p):
    import bh_b200
    lib = bh_b200.bind(os.path.join(ROOT, "oracle", "libbh_ref.so"))
    e = bh_b200.NativeEngine(lib=lib)
    e.set_window(p["W"], p["H"])
    e.set_params(theta=p["theta"], G=p["G"], dt=p["dt"], merge_min_dist=8.0 if p["merge"] else 0.0)
    e.set_bodies(*scene)
    e.step(p["steps"])
    st = np.stack(e.get_bodies())
    t = e.tree()
    return st, np.stack([t["cx"], t["cy"], t["h"]])


def write_dump(path, st, q):
    with open(path, "wb") as f:
        f.write(struct.pack("<q", st.shape[1])); f.write(np.ascontiguousarray(st, "<f8").tobytes())
        f.write(struct.pack("<q", q.shape[1])); f.write(np.ascontiguousarray(q, "<f8").tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["make", "check", "selftest"])
    ap.add_argument("case")
    ap.add_argument("dump", nargs="?")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--merge", type=int, default=1)
    a = ap.parse_args()
    if a.cmd == "make":
        from bh_b200 import scenes
        write_case(a.case, scenes.default_two_disks(seed=1), steps=a.steps, merge=a.merge)
        print("wrote", a.case)
        return 0
    scene, p = read_case(a.case)
    st, q = oracle_dump(scene, p)
    if a.cmd == "selftest":
        write_dump(a.dump, st, q)
        return 0
    jst, jq = read_dump(a.dump)
    ok = jst.shape == st.shape and jq.shape == q.shape and (jst.view(np.uint64) == st.view(np.uint64)).all() \
        and (jq.view(np.uint64) == q.view(np.uint64)).all()
    print("bodies", st.shape[1], "cells", q.shape[1], "bit-identical to the JVM reference:", bool(ok))
    if not ok and jst.shape == st.shape:
        print("max |dstate|", np.abs(jst - st).max())
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
