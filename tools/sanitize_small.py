"""Small end-to-end run for compute-sanitizer (memcheck): build, jitter replay, walk, merge,
re-homing, read-backs on a few thousand bodies."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bh_b200
from bh_b200 import scenes

s = list(scenes.snap_f32(scenes.default_two_disks(n1=3000, n2=800, seed=3)))
s[0][5:40] = s[0][0] + np.linspace(1.0, 7.0, 35)          # victims of the merge rule
s[1][5:40] = s[1][0]
s[0][100], s[1][100] = s[0][101], s[1][101]                # a jitter cluster
e = bh_b200.NativeEngine(flags=bh_b200.BH_FLAG_BODY_COUNTS, rehome_interval=2)
e.set_params(theta=0.5)
e.set_bodies(*s)
e.step(5)
e.compute_accelerations()
e.tree(); e.morton(); e.get_positions_f32(); e.body_counts(); e.energy(); e.direct_sum()
e.set_bodies(*e.get_bodies())
e.step(2)
z = np.zeros(0)
e.set_bodies(z, z, z, z, z)
e.step(1)
print("sanitize run ok, merged", e.counters()["total_merged"])
