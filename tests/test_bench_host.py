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


def test_reference_arm_prints_the_contract_line_with_the_same_config_object():
    """`bench.py --impl reference`: the CPU arm on a small instance — same metric / unit / config object as the CUDA arm
    prints for the same command line, a cpu_baseline describing the run, an e2e that repeats the line's value."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--bodies", "20000", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    bench = _bench()
    assert d["impl"] == "reference" and d["metric"] == "body_interactions_per_s" and d["unit"] == "interactions/s" and d["higher_is_better"] is True
    assert set(d["config"]) == set(bench.bench_config(20000, 20000, 1, 1))            # the keys the CUDA arm prints
    assert d["config"]["workload"].startswith("20000-body uniform") and "20000 bodies per GPU" in d["config"]["workload"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]


def test_clock_sampler_reports_rows_from_the_timed_region(tmp_path, monkeypatch):
    """A stand-in `nvidia-smi` that takes 0.3 s to come up (8 ranks starting one each take longer) and prints a row every
    100 ms: a 40 ms timed region still gets its samples, taken from its start to one period after its end."""
    import stat
    import time
    fake = tmp_path / "nvidia-smi"
    fake.write_text('#!/bin/bash\nsleep 0.3\nwhile true; do echo "0, 1830, 1965, 400.0, 0x4, Not Active, Not Active, Not Active, Active"; sleep 0.1; done\n')
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    bench = _bench()
    s = bench.ClockSampler(0)
    s.start()
    s.wait_ready()
    s.begin()
    time.sleep(0.04)
    r = s.stop()
    assert r["samples"] >= 1 and r["sm_mhz"] == 1830.0 and r["sm_max_mhz"] == 1965.0 and r["reasons"] == ["sw_power_cap"], r
    assert "timed region" in r["window"]


def test_bench_main_prints_one_contract_line():
    """bench.py's main() end to end at world 1 over the oracle library (tests/bench_dry_run.py): ONE JSON line with every
    key of the driver's contract, filled phase by phase."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bench_dry_run.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "body_interactions_per_s" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None
    assert set(d["config"]) == {"workload", "step", "l2"}
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] == 5 * 8 * 20000 and "incomplete" not in d and "side_measurement_errors" not in d
