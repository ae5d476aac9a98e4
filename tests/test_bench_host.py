"""bench.py's host logic that only runs at N > 1 on the GPU box, driven here at world 1 over the oracle library
(same C ABI): the self-check has to be safe in the reference's jitter regime (BarnesHutAlg.kt:145-156), where
every build moves the coincident bodies it finds."""
import importlib.util
import os

import numpy as np

from conftest import ROOT, make_engine


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _scene(n, pairs, seed=3):
    rng = np.random.default_rng(seed)
    x, y = rng.uniform(0, 2400, n), rng.uniform(0, 800, n)
    x[1:2 * pairs:2], y[1:2 * pairs:2] = x[0:2 * pairs:2], y[0:2 * pairs:2]      # coincident pairs -> jitter in every build
    return x, y, rng.normal(0, 1, n), rng.normal(0, 1, n), rng.uniform(1, 2, n)


def test_parity_check_restores_state_between_evaluations(oracle_lib):
    bench = _bench()
    for pairs in (0, 40):
        e = make_engine(oracle_lib, _scene(4000, pairs), theta=0.5)
        e.step(2)
        r = bench.parity_check(e, 1, lambda v: v, lambda v: v)
        assert r["checked"] and r["interactions_equal"] and r["acc_bit_identical"], r
        assert r["bodies_checked_per_rank"] == 4000
        e.close()


def test_slice_io_zero_steps_round_trips_the_state(oracle_lib):
    e = make_engine(oracle_lib, _scene(1500, 10), theta=0.5)
    e.step(1)
    n = e.n
    snap = [np.empty(n) for _ in range(5)]
    k = e.step_io_slice(0, out=snap)
    assert k == n
    ref = [a.copy() for a in e.get_bodies()]
    perm = e.slice_index()
    for a, b in zip(snap, ref):
        assert np.array_equal(a[:k], b[perm])
    e.compute_accelerations()                     # a build in the jitter regime: positions move
    e.step_io_slice(0, inputs=[a[:k] for a in snap])
    for a, b in zip(e.get_bodies(), ref):
        assert np.array_equal(a, b)
    e.close()


def test_watchdog_prints_the_line_and_leaves_when_a_phase_hangs(tmp_path):
    """A phase after the headline that never returns (ranks waiting in a collective) must not lose the headline."""
    import json
    import subprocess
    import sys
    code = (
        "import importlib.util, time\n"
        f"spec = importlib.util.spec_from_file_location('b', r'{os.path.join(ROOT, 'bench.py')}')\n"
        "b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n"
        "wd = b.Watchdog(0); wd.line = {'metric': 'body_interactions_per_s', 'value': 1.0, 'e2e': None}\n"
        "wd.phase('quick', 30); wd.done()\n"
        "time.sleep(1.2)\n"                        # a finished phase never fires
        "wd.phase('stuck', 1)\n"
        "time.sleep(30)\n"
        "print('not reached')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["value"] == 1.0 and "stuck" in d["incomplete"]
