#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200: body-interactions/s (and steps/s)
of PhysicsEngine.step() (2 tree builds + 2 force evaluations + kick-drift-kick) at θ = 0.5.

    python bench.py --gpus N --steps K --warmup W            # CUDA engine (the product)
    python bench.py --impl reference --steps K --warmup W    # the reference CPU path (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU

A "step" is one PhysicsEngine.step() on the synthetic workload.  At N = 1 the workload is
BASELINE.json configs[1]: the 1M-body uniform 'C' cloud (BodyFactory.makeUniformRandom
semantics, m = 0.5, 2400x800 window, θ = 0.5, G = 80, Δt = 0.005, merge off).  For N > 1
GPUs the cloud has 1M bodies PER GPU (window area scaled by N: constant density), the tree
is replicated, every GPU walks and integrates its Morton slice, and the drifted positions
are all-gathered over NCCL (weak scaling).  One JSON line is printed by rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BODIES_PER_GPU = 1_000_000
THETA = 0.5


def workload(n_gpus, bodies_per_gpu):
    from bh_b200 import scenes
    n = bodies_per_gpu * n_gpus
    s = math.sqrt(n_gpus)
    W, H = int(round(2400 * s)), int(round(800 * s))
    return scenes.make_uniform_random(n, 0.5, W, H, seed=3), W, H


class ClockSampler:
    """nvidia-smi clocks / throttle reasons around the timed region (B200_PROFILING.md).  The query loop is started
    before the warm-up (nvidia-smi needs a few hundred ms to come up, longer with 8 ranks starting one each) and
    its rows are time-stamped; the report uses the rows that fall between the start of the timed region and one
    sampling period after its end (a 20-step region is shorter than the 100 ms period), else all rows, and says which."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    PERIOD = 0.1

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu, self.t_begin = [], None, gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(int(self.PERIOD * 1000))],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_ready(self, timeout=5.0):
        t_end = time.perf_counter() + timeout
        while self.proc and not self.rows and time.perf_counter() < t_end and self.proc.poll() is None:
            time.sleep(0.01)

    def begin(self):
        self.t_begin = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_stop = time.perf_counter()
        time.sleep(self.PERIOD + 0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = list(self.rows)
        t0 = self.t_begin if self.t_begin is not None else -1.0
        near = [r for t, r in rows if t0 <= t <= t_stop + self.PERIOD + 0.05]
        window = "from the start of the timed region to one 100 ms sampling period after its end"
        if not near:
            near, window = [r for _, r in rows], "whole run (no sample fell near the timed region)"
        sm = [float(r[1]) for r in near if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in near if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in near if len(r) >= 9 for k in range(4) if r[5 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


class Watchdog:
    """The headline is measured first; everything after it (end-to-end loop, self-check, the larger configs, the side
    measurements) runs under a per-phase deadline.  A phase that does not finish (a multi-rank run waiting in a
    collective, say) must not take the measured headline with it: rank 0 prints the line as far as it got, naming the
    phase under "incomplete", and every rank leaves with os._exit(0) — each rank runs its own timer from the same
    barrier-aligned phase starts."""

    def __init__(self, rank):
        self.rank, self.line, self.deadline, self.name, self.limit = rank, None, None, None, None
        self.lock = threading.Lock()
        threading.Thread(target=self._run, daemon=True).start()

    def phase(self, name, seconds):
        with self.lock:
            self.name, self.limit, self.deadline = name, seconds, time.monotonic() + seconds

    def done(self):
        with self.lock:
            self.deadline = None

    def _run(self):
        while True:
            time.sleep(0.5)
            with self.lock:
                fired = self.deadline is not None and time.monotonic() > self.deadline
                if fired and self.rank == 0 and self.line is not None:
                    self.line["incomplete"] = (f"phase '{self.name}' did not finish within {self.limit} s: the run was cut there, "
                                               f"the fields it and the later phases fill are null / missing")
                    emit(self.line)
                if fired:
                    sys.stderr.write(f"[bench rank {self.rank}] watchdog: phase '{self.name}' exceeded {self.limit} s\n")
                    sys.stderr.flush()
                    os._exit(0)


REFERENCE_BUDGET_S = 150.0     # CPU time the whole reference arm may take (warm-up + timed steps)


def bench_config(n_total, per_gpu, W, H):
    """The `config` object of the JSON line: the workload only, so that both arms (this implementation and
    --impl reference) print the SAME object for the same command line."""
    return {"workload": f"{n_total}-body uniform 'C' cloud ({per_gpu} bodies per GPU), theta={THETA}, {W}x{H} window, G=80, dt=0.005, merge off",
            "step": "PhysicsEngine.step(): 2 builds + 2 evaluations + KDK",
            "l2": "no flush between steps: the per-step working set (~210 B/body of state, sort buffers and tree) exceeds the last-level "
                  "cache (126 MB L2 on B200) and is rewritten by every build"}


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (JVM unavailable: the literal C++
    port in oracle/) on all host threads, same config/metric as the CUDA arm.  Every step is a
    full PhysicsEngine.step(); when K + W steps of the whole workload would not fit the time
    budget, the steps run on a BOUNDED SAMPLE of it: a uniform random subset of the bodies (same
    window), sized from one calibration step.  The metric is a rate, so it stays comparable."""
    if rank != 0:
        return
    import bh_b200
    lib = bh_b200.bind(os.path.join(ROOT, "oracle", "libbh_ref.so"))
    scene, W, H = workload(max(1, args.gpus), args.bodies)
    n_full = len(scene[0])

    def engine(sc):
        e = bh_b200.NativeEngine(lib=lib)
        e.set_window(W, H)
        e.set_params(theta=THETA, merge_min_dist=0.0)
        e.set_bodies(*sc)
        return e

    # calibration: one step on (at most) 250k bodies; a step costs ~ n log n
    n_cal = min(n_full, 250_000)
    rng = np.random.default_rng(11)
    pick = np.sort(rng.choice(n_full, n_cal, replace=False)) if n_cal < n_full else None
    ec = engine(tuple(a[pick] for a in scene) if pick is not None else scene)
    t0 = time.perf_counter()
    ec.step(1)
    t_cal = time.perf_counter() - t0
    ec.close()
    total_steps = args.steps + args.warmup
    per_body = t_cal / (n_cal * math.log2(max(n_cal, 2)))
    n_s = n_full
    while n_s > 20_000 and per_body * n_s * math.log2(n_s) * total_steps > REFERENCE_BUDGET_S - t_cal:
        n_s = int(n_s * 0.8)
    if n_s < n_full:
        pick = np.sort(rng.choice(n_full, n_s, replace=False))
        scene = tuple(np.ascontiguousarray(a[pick]) for a in scene)
        sample = (f"{args.steps} full steps on a uniform random subset of {n_s} of the workload's {n_full} bodies "
                  f"(same window; sized from a {n_cal}-body calibration step to fit {REFERENCE_BUDGET_S:.0f} s)")
    else:
        sample = f"{args.steps} full steps of the whole workload"
    e = engine(scene)
    cores = lib.bh_ref_threads(e._h)
    e.step(args.warmup)
    e.reset_counters()
    t0 = time.perf_counter()
    e.step(args.steps)
    dt = time.perf_counter() - t0
    c = e.counters()
    val = c["total_interactions"] / dt
    out = {
        "impl": "reference", "metric": "body_interactions_per_s", "value": val, "unit": "interactions/s",
        "steps_per_s": args.steps / dt, "n_gpus": max(1, args.gpus), "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": bench_config(n_full, args.bodies, W, H),
        "cpu_baseline": {"value": val, "unit": "interactions/s", "cores": int(cores), "kind": "port",
                         "sample": sample + " (C++ port of BarnesHutAlg.kt; no JVM in the image)", "bodies_in_the_sample": len(scene[0])},
        "e2e": {"value": val, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "phases_ms_per_step": {"build": c["ms_build"] / args.steps, "walk": c["ms_walk"] / args.steps,
                               "integrate": c["ms_integrate"] / args.steps},
    }
    emit(out)



# ---- BASELINE.json configs[2..4] as device-generated scenes -------------------
def build_config_scene(eng, name):
    """Appends the bodies of config `name` with the device generators (bh_append_disk: BodyFactory.makeGalaxyDisk
    laws, counter-based RNG keyed by the seed, so every rank generates the same list) and returns (W, H, label)."""
    def disk(n, x, y, r, vx=0.0, vy=0.0, central=50_000.0, sat=5_000.0, seed=1):
        p = eng.disk_params(2400, 800, x=x, y=y, r=r, vx=vx, vy=vy, central_mass=central, total_satellite_mass=sat)
        eng.append_disk(n, p, seed=seed)
    if name == "C3":      # 10M two-disk merger with central black holes: the reference scene (NBodyPanel.kt:83-99) x 800 bodies
        W = H = 32768
        eng.set_window(W, H)
        k = math.sqrt(800.0)
        disk(8_000_000, W * 0.5, H * 0.5, 300.0 * k, seed=4)
        disk(2_000_000, W * 0.5, H * 0.5 - 240.0 * k, 100.0 * k, vx=-50.0, central=5_000.0, sat=500.0, seed=5)
        return W, H, "10M-body two-disk galaxy merger with central black holes (8M + 2M, radii x sqrt(800))"
    if name == "C4":      # 100M disk collision
        W = H = 131072
        eng.set_window(W, H)
        r = 300.0 * math.sqrt(5000.0)
        disk(50_000_000, W * 0.5 - 30_000.0, H * 0.5, r, vx=+50.0, seed=6)
        disk(50_000_000, W * 0.5 + 30_000.0, H * 0.5, r, vx=-50.0, seed=7)
        return W, H, "100M-body disk collision (2 x 50M, radii 21,213, closing at 100 px/t)"
    if name == "C5":      # mixed-mass stress: 8 disks + 16 black holes, 50M bodies
        W = H = 65536
        eng.set_window(W, H)
        rng = np.random.Generator(np.random.PCG64(8))
        r = 300.0 * math.sqrt(6_250_000 / 10_000.0)
        for d in range(8):
            cx, cy = (0.15 + 0.7 * rng.random()) * W, (0.15 + 0.7 * rng.random()) * H
            ang, v = rng.random() * 2 * math.pi, 50.0 * rng.random()
            disk(6_250_000, cx, cy, r, vx=v * math.cos(ang), vy=v * math.sin(ang), seed=800 + d)
        for b in range(16):   # the right-mouse-button "black hole": a lone 50,000-mass body (NBodyPanel.kt:171)
            disk(1, (0.1 + 0.8 * rng.random()) * W, (0.1 + 0.8 * rng.random()) * H, 8.0, seed=900 + b)
        return W, H, "mixed-mass stress: 8 disks x 6.25M + 16 black holes (50M bodies), theta stepped 0.5 -> 0.8 like Z/X key presses"
    raise ValueError(name)


def run_config(name, steps, warm, local_rank, rank, world, dist, allmax, allsum, barrier):
    """One strong-scaling record of config `name` at this run's GPU count (total work fixed)."""
    import bh_b200
    n_total = {"C3": 10_000_000, "C4": 100_000_000, "C5": 50_000_016}[name]
    eng = bh_b200.NativeEngine(device=local_rank, capacity_hint=n_total, flags=bh_b200.BH_FLAG_LET if world > 1 else 0)
    try:
        eng.set_params(theta=THETA, merge_min_dist=0.0)
        if world > 1:
            from bh_b200.distributed import init_nccl_engine
            init_nccl_engine(eng, dist, rank, world)
        W, H, label = build_config_scene(eng, name)
        n = eng.n
        rec = {"config": name, "workload": label, "bodies": n, "window": [W, H], "n_gpus": world, "scaling": "strong", "steps": steps, "warmup": warm}
        e0 = None
        if name == "C5":
            e0 = eng.energy_tree(0.35)          # tree potential (O(N log N)); theta 0.35 for the diagnostic itself
        eng.step(warm)
        eng.reset_counters()
        barrier()
        dev_ms = 0.0
        if name == "C5":                        # "adaptive theta": the Z/X keys step theta by 0.05 (NBodyPanel.kt:247-248)
            theta, drift, done = THETA, [], 0
            while done < steps:
                k = min(10, steps - done)
                eng.set_params(theta=theta)
                eng.step(k)
                dev_ms += eng.counters()["ms_step_call"]
                done += k
                theta = min(1.6, theta + 0.05)
                if done % 50 == 0 or done == steps:
                    e1 = eng.energy_tree(0.35)
                    drift.append({"step": warm + done, "dE_over_E0": (e1["total"] - e0["total"]) / abs(e0["total"])})
            rec["theta_schedule"] = "0.5 +0.05 every 10 steps"
            rec["energy_drift_tree_potential"] = drift
        else:
            eng.step(steps)
            dev_ms = eng.counters()["ms_step_call"]
        barrier()
        c = eng.counters()
        dev_ms = allmax(dev_ms)
        inter = allsum(float(c["total_interactions"]))
        ne = max(1, c["total_evaluations"])
        rec.update({"ms_per_step": dev_ms / steps, "steps_per_s": steps / (dev_ms * 1e-3), "interactions_per_s": inter / (dev_ms * 1e-3),
                    "interactions_per_body_per_evaluation": inter / (2.0 * steps) / n,
                    "phases_ms_per_evaluation": {"build": allmax(c["ms_build"]) / ne, "walk": allmax(c["ms_walk"]) / ne},
                    "ms_per_step_exchange": allmax(c["ms_comm"]) / steps, "jitter_bodies_last_build": c["n_jitter_bodies"],
                    "cells": c["n_cells"], "max_depth": c["max_depth"]})
        if world > 1:
            ls = eng.let_stats()
            rec["mode"] = "domain (LET)" if ls["let_evaluations"] > 0 else "replicated tree"
            rec["domain_mode_rank0"] = {k: ls[k] for k in ("let_evaluations", "fallbacks", "fallbacks_guest_in_jitter_cluster", "fallbacks_stray_overflow",
                                                           "fallbacks_cell_overflow", "stray_capacity", "let_cells", "cells_imported", "own_strays", "jitter_positions_returned",
                                                           "stray_leaf_descents", "stray_leaf_scans")}
        return rec
    finally:
        eng.close()


def parity_check(eng, world, allsum, allmin):
    """N > 1 self-check on the state the timed region left behind: this rank's slice evaluated in DOMAIN mode
    (locally essential tree) and over the REPLICATED tree must give bit-identical accelerations, and the
    interaction counts summed over the ranks must equal what ONE rank counts walking every body over the
    replicated tree.  A build may MUTATE positions (the reference's jitter regime, BarnesHutAlg.kt:145-156: a few
    bodies per build at 8M bodies in this window), so every evaluation starts from the same saved slice state
    (bh_step_io_slice with 0 steps restores it without touching the partition)."""
    eng.evaluate_slice()                               # absorbs a pending re-homing: the slices are stable from here on
    n = eng.n
    snap = [np.empty(n) for _ in range(5)]
    k = eng.step_io_slice(0, out=snap)                 # this rank's slice as it is now
    snap = [a[:k].copy() for a in snap]
    epoch = eng.slice_epoch()                          # (the same number on every rank: re-homings are collective)

    def skipped(why):
        eng.set_domain_mode(True)
        return {"mode": "not checked: " + why, "checked": False, "interactions_equal": None, "acc_bit_identical": None}

    def restore():
        eng.step_io_slice(0, inputs=snap)

    ls0 = eng.let_stats()
    eng.reset_counters()
    ax_d, ay_d, ui_d = eng.evaluate_slice()
    cd = eng.counters()
    ls1 = eng.let_stats()
    domain_ran = ls1["let_evaluations"] > ls0["let_evaluations"] and ls1["fallbacks"] == ls0["fallbacks"]
    if eng.slice_epoch() != epoch:                     # a fall-back re-homing re-cut the slices: the saved state no longer fits
        return skipped("the domain-mode evaluation fell back to a re-homing")
    restore()
    eng.set_domain_mode(False)
    eng.reset_counters()
    ax_r, ay_r, ui_r = eng.evaluate_slice()
    cr = eng.counters()
    if eng.slice_epoch() != epoch:
        return skipped("a re-homing happened during the replicated evaluation")
    restore()
    eng.reset_counters()
    fx, fy = eng.compute_accelerations()              # every body, on every rank, over the replicated tree
    cf = eng.counters()
    if eng.slice_epoch() != epoch:
        return skipped("a re-homing happened during the evaluation of all bodies")
    restore()
    eng.set_domain_mode(True)
    same_slice = len(ui_d) == len(ui_r) and bool((ui_d == ui_r).all())
    bit_dr = same_slice and bool(np.array_equal(ax_d, ax_r, equal_nan=True) and np.array_equal(ay_d, ay_r, equal_nan=True))
    bit_full = bool(np.array_equal(ax_d, fx[ui_d], equal_nan=True) and np.array_equal(ay_d, fy[ui_d], equal_nan=True))
    inter_d, inter_r = allsum(float(cd["interactions"])), allsum(float(cr["interactions"]))
    open_d, open_r = allsum(float(cd["opened"])), allsum(float(cr["opened"]))
    all_dr = allmin(1.0 if bit_dr else 0.0) == 1.0
    all_full = allmin(1.0 if bit_full else 0.0) == 1.0
    return {"mode": "domain (LET) vs replicated tree, same ranks, same state" if domain_ran else "replicated tree only (domain mode did not run)",
            "checked": True, "bodies_checked_per_rank": int(len(ui_d)),
            "jitter_bodies_in_the_builds": int(cf["n_jitter_bodies"]),
            "interactions_equal": bool(inter_d == inter_r == float(cf["interactions"]) and open_d == open_r == float(cf["opened"])),
            "interactions": {"domain_sum_over_ranks": inter_d, "replicated_sum_over_ranks": inter_r, "one_rank_all_bodies": float(cf["interactions"])},
            "opened": {"domain_sum_over_ranks": open_d, "replicated_sum_over_ranks": open_r, "one_rank_all_bodies": float(cf["opened"])},
            "acc_bit_identical": bool(all_dr and all_full),
            "acc_bit_identical_domain_vs_replicated_slice": bool(all_dr),
            "acc_bit_identical_vs_one_rank_walking_all_bodies": bool(all_full)}


def main():
    # rank 0 must print ONE JSON line on stdout: everything libraries print there (NCCL's version banner, ...)
    # is diverted to stderr; the line itself goes to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bodies", type=int, default=BODIES_PER_GPU, help="bodies per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="C3,C4,C5", help="BASELINE.json configs measured after the headline (strong scaling); '' = none")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args, rank, world)
        return
    args.steps = args.steps if args.steps is not None else 20
    args.warmup = max(3, args.warmup if args.warmup is not None else 3)

    import torch
    import torch.distributed as dist
    import bh_b200
    from bh_b200 import scenes as scenes_mod

    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allmin(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())

    def allsum(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    scene, W, H = workload(n_gpus, args.bodies)
    n = len(scene[0])
    # N > 1: domain mode (BH_FLAG_LET) — every rank builds its own Morton range and walks a locally essential tree
    eng = bh_b200.NativeEngine(device=local_rank, capacity_hint=n, flags=bh_b200.BH_FLAG_LET if world > 1 else 0)
    eng.set_window(W, H)
    eng.set_params(theta=THETA, merge_min_dist=0.0)
    if world > 1:
        from bh_b200.distributed import init_nccl_engine
        init_nccl_engine(eng, dist, rank, world)   # NCCL communicator inside the engine
    eng.set_bodies(*scene)

    # ---- device-resident throughput: K steps, inputs already in HBM ------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_ready()                            # (before the warm-up: no idle gap between warm-up and timed region)
    eng.step(args.warmup)
    eng.reset_counters()
    barrier()
    sampler.begin()
    t0 = time.perf_counter()
    eng.step(args.steps)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    c = eng.counters()
    let_stats = eng.let_stats() if world > 1 else None
    dev_ms = allmax(c["ms_step_call"])              # CUDA events on the engine's stream, max over ranks
    wall = allmax(wall)
    inter = allsum(float(c["total_interactions"]))
    opened = allsum(float(c["total_opened"]))
    launches = int(c["kernel_launches"])
    value = inter / (dev_ms * 1e-3)
    n_walks = 2 * args.steps
    walk_ms = allmax(c["ms_walk"]) / n_walks
    build_ms = allmax(c["ms_build"]) / n_walks

    # ---- roofline of the dominant kernel (k_walk): FP32-issue bound -------------------------
    peaks, peak_src = measured_peaks()
    fp32 = np.zeros(1)
    import ctypes as C
    eng.lib.bh_measure_fp32_tflops(local_rank, fp32.ctypes.data_as(C.POINTER(C.c_double)))
    my_inter, my_open = float(c["total_interactions"]) / n_walks, float(c["total_opened"]) / n_walks
    walk_flops = 14.0 * my_inter + 8.0 * my_open           # 14 flop/interaction + 8 flop/rejected test
    ach = walk_flops / (c["ms_walk"] / n_walks * 1e-3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of one k_walk launch: taken from the ncu capture of THIS code on
    # THIS workload when one is committed (profiles/walk_traffic.json names the capture it summarises), else null
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "walk_traffic.json")
    if world == 1 and args.bodies == BODIES_PER_GPU and os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic, traffic_src = float(tj["dram_bytes_per_launch"]), tj.get("source")
        except Exception:
            pass
    roofline = {"kernel": "k_walk", "bound": "fp32", "achieved": ach, "peak": float(fp32[0]), "unit": "TFLOP/s",
                "frac": ach / float(fp32[0]) if fp32[0] > 0 else None, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": "measured live: FFMA microbenchmark (bh_measure_fp32_tflops) — MEASURED_PEAKS.json has no FP32 figure; "
                               "the walk is neither HBM- nor tensor-bound (ncu summaries under profiles/, DESIGN.md §4)",
                "hbm_bytes_algorithmic": 32.0 * n / world + 32.0 * c["n_cells"],
                "ms_per_launch": c["ms_walk"] / n_walks,
                "algorithmic": "14 flop x interactions + 8 flop x rejected opening tests per launch"}
    # the HBM-bound phase: keygen + onesweep sort + scan + emit + climb (algorithmic bytes/body)
    passes = (2 * c["key_levels"] + 1 + 7) // 8
    cells_per_body = c["n_cells"] / max(1, c["n_in_tree"])
    bytes_per_body = 24 + 8 + passes * 24 + 12 + cells_per_body * (13 + 73) + 24
    bw = bytes_per_body * n / (build_ms * 1e-3) / 1e9
    roofline_build = {"kernel": "build phase (k_keygen, k_histogram, k_onesweep_pass x%d, k_count_scan, k_emit, k_climb)" % passes,
                      "bound": "hbm", "achieved": bw, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": bw / peaks["hbm_gbs"],
                      "traffic": None, "peak_source": peak_src, "ms_per_build": build_ms,
                      "algorithmic_bytes_per_body": bytes_per_body}

    line = {
        "metric": "body_interactions_per_s", "value": value, "unit": "interactions/s",
        "steps_per_s": args.steps / (dev_ms * 1e-3), "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "wall_ms_per_step": wall / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 interactions, f64 state/COM/integrator", "data": "synthetic",
        "config": bench_config(n, args.bodies, W, H),
        "parallelism": "single GPU" if n_gpus == 1 else (
            f"domain mode: {n_gpus} Morton ranges, local tree per rank + locally essential tree (all-reduced level summaries, "
            f"boundary subtrees {'read over NVLink peer memory (CUDA IPC)' if let_stats.get('enabled') == 2 else 'over NCCL send/recv'}), "
            f"re-homing every 8 steps" if let_stats and let_stats["let_evaluations"] > 0 else
            f"replicated tree, {n_gpus} Morton slices of targets, NCCL all-gather of positions"),
        "interactions_per_step": inter / args.steps, "opened_per_step": opened / args.steps,
        "phases_ms_per_evaluation": {"build": build_ms, "walk": walk_ms},
        "ms_per_step_other": {"exchange": c["ms_comm"] / args.steps, "integrate_and_rest": c["ms_integrate"] / args.steps},
        "roofline": roofline, "roofline_build": roofline_build, "cpu_baseline": None, "e2e": None, "configs0": None, "reuse_acc_mode": None,
        "parity_check": None, "configs": [], "roofline_direct": None,
        "gpu_launches": launches, "clocks": clocks,
    }
    if let_stats:
        line["domain_mode_rank0"] = let_stats
    wd = Watchdog(rank)
    wd.line = line

    # ---- end to end through the C ABI with HOST buffers --------------------------------------
    wd.phase("end to end (host buffers through the C ABI)", 120)
    e2e = None
    try:
        e2e_steps = max(3, min(args.steps, 10))
        if world == 1:
            host = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in scene]
            hnp = [t.numpy() for t in host]
            out = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(5)]
            onp = tuple(t.numpy() for t in out)
            eng.step_io(1, inputs=hnp, out=onp)
            eng.reset_counters()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                # resetBodies(host arrays) + step() + getBodies(host arrays) in ONE C-ABI call: H2D of this
                # step's inputs and D2H of its result are inside the call (and inside the timed region),
                # overlapped with the compute where the data dependences allow (bh_step_io)
                eng.step_io(1, inputs=hnp, out=onp)
            barrier()
            e2e_wall = allmax(time.perf_counter() - t0)
            e2e_inter = allsum(float(eng.counters()["total_interactions"]))
            # the same three calls made separately (no overlap), for comparison
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                eng.set_bodies(*hnp)
                eng.step(1)
                eng.get_bodies(out=onp)
            barrier()
            seq_wall = allmax(time.perf_counter() - t0)
            e2e = {"value": e2e_inter / e2e_wall, "unit": "interactions/s", "steps_per_s": e2e_steps / e2e_wall,
                   "h2d_bytes_per_step": 5 * 8 * n, "d2h_bytes_per_step": 5 * 8 * n, "steps": e2e_steps,
                   "api": "bh_step_io(1, host in, host out) per step: resetBodies + step + getBodies, pinned host arrays, copies overlapped with compute",
                   "steps_per_s_separate_calls": e2e_steps / seq_wall}
        else:
            # N > 1: SHARDED I/O — every rank moves only the bodies of its own slice (bh_step_io_slice): the slice's state goes
            # up from pinned host memory, one step runs, the slice's new state comes down
            # (like PhysicsEngine's own loop — getBodies, step, getBodies — the slice that comes down is the slice that goes
            # up next: the host owns the state between steps, and bodies that migrated to another rank at a re-homing
            # simply arrive in that rank's output)
            hslice = [torch.empty(n, dtype=torch.float64).pin_memory().numpy() for _ in range(5)]
            oslice = [torch.empty(n, dtype=torch.float64).pin_memory().numpy() for _ in range(5)]
            k = eng.step_io_slice(0, out=hslice)                  # this rank's slice after the device-resident steps
            k2 = eng.step_io_slice(1, inputs=[a[:k] for a in hslice], out=oslice)
            hslice, oslice, k = oslice, hslice, k2
            eng.reset_counters()
            h2d = d2h = 0
            epoch0 = eng.slice_epoch()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                k2 = eng.step_io_slice(1, inputs=[a[:k] for a in hslice], out=oslice)
                h2d += 5 * 8 * k
                d2h += 5 * 8 * k2
                hslice, oslice, k = oslice, hslice, k2
            barrier()
            e2e_wall = allmax(time.perf_counter() - t0)
            e2e_inter = allsum(float(eng.counters()["total_interactions"]))
            e2e = {"value": e2e_inter / e2e_wall, "unit": "interactions/s", "steps_per_s": e2e_steps / e2e_wall,
                   "h2d_bytes_per_step": allsum(float(h2d)) / e2e_steps, "d2h_bytes_per_step": allsum(float(d2h)) / e2e_steps, "steps": e2e_steps,
                   "api": "bh_step_io_slice(1, slice in, slice out) per step on every rank: only the rank's own bodies cross PCIe (pinned host arrays, "
                          "transfers overlapped with the compute); the slice that comes down is the slice that goes up at the next step",
                   "re_homings_in_timed_region": eng.slice_epoch() - epoch0}
    except Exception as ex:                                       # (reported in the line; the headline stands)
        e2e = {"value": None, "error": f"{type(ex).__name__}: {ex}"[:300]}
    line["e2e"] = e2e

    # ---- N > 1: the run checks itself (domain mode vs replicated tree vs one rank walking every body) -----------
    pcheck = None
    if world > 1:
        wd.phase("parity_check (domain mode vs replicated tree vs one rank walking every body)", 180)
        barrier()
        try:
            pcheck = parity_check(eng, world, allsum, allmin)
        except Exception as ex:
            pcheck = {"mode": "not checked: the check itself raised", "checked": False, "error": f"{type(ex).__name__}: {ex}"[:300],
                      "interactions_equal": None, "acc_bit_identical": None}
        line["parity_check"] = pcheck
        if pcheck.get("checked") and not (pcheck["interactions_equal"] and pcheck["acc_bit_identical"]):
            line["parity_failed"] = True                         # a multi-GPU run that disagrees with itself: flagged at the top level
    try:
        eng.close()
    except Exception:
        pass

    # ---- BASELINE.json configs[2..4] at this GPU count (strong scaling: the total is fixed) --------------------
    cfg_records = line["configs"]
    for name in [c for c in args.configs.split(",") if c]:
        steps_c, warm_c = {"C3": (20, 3), "C4": (8, 2), "C5": (60, 2)}[name]
        wd.phase(f"config {name}", {"C3": 150, "C4": 300, "C5": 300}[name])
        barrier()
        try:
            cfg_records.append(run_config(name, steps_c, warm_c, local_rank, rank, world, dist, allmax, allsum, barrier))
        except Exception as ex:                                   # a config must never take the headline down with it
            cfg_records.append({"config": name, "n_gpus": world, "error": f"{type(ex).__name__}: {ex}"[:300]})
            torch.cuda.synchronize()
    wd.phase("side measurements (reuse mode, configs[0], direct sum, CPU baseline)", 300)

    try:
        # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            lib = bh_b200.bind(os.path.join(ROOT, "oracle", "libbh_ref.so"))
            o = bh_b200.NativeEngine(lib=lib)
            o.set_window(W, H)
            o.set_params(theta=THETA, merge_min_dist=0.0)
            o.set_bodies(*scene)
            o.reset_counters()
            t0 = time.perf_counter()
            o.step(2)
            dt = time.perf_counter() - t0
            oc = o.counters()
            cpu = {"value": oc["total_interactions"] / dt, "unit": "interactions/s", "cores": int(lib.bh_ref_threads(o._h)),
                   "kind": "port", "steps_per_s": 2 / dt,
                   "sample": "2 full steps of the same 1M-body workload (C++ port of BarnesHutAlg.kt; the JVM reference cannot run in this image)"}
            o.close()
        line["cpu_baseline"] = cpu
    except Exception as ex:                                       # (reported in the line; the headline stands)
        line.setdefault("side_measurement_errors", {})["cpu_baseline"] = f"{type(ex).__name__}: {ex}"[:300]

    try:
        # ---- opt-in BH_FLAG_REUSE_ACC (result-identical, one evaluation per step): reported, not the headline
        reuse = None
        if world == 1:
            er = bh_b200.NativeEngine(device=local_rank, capacity_hint=n, flags=bh_b200.BH_FLAG_REUSE_ACC)
            er.set_window(W, H)
            er.set_params(theta=THETA, merge_min_dist=0.0)
            er.set_bodies(*scene)
            er.step(args.warmup)
            er.reset_counters()
            er.step(args.steps)
            cr = er.counters()
            reuse = {"steps_per_s": args.steps / (cr["ms_step_call"] * 1e-3), "ms_per_step": cr["ms_step_call"] / args.steps,
                     "evaluations_per_step": cr["total_evaluations"] / args.steps,
                     "note": "step n+1 reuses a(t+dt) of step n (bit-identical state); not the reference's cost model, so not the headline"}
            er.close()
        line["reuse_acc_mode"] = reuse
    except Exception as ex:                                       # (reported in the line; the headline stands)
        line.setdefault("side_measurement_errors", {})["reuse_acc_mode"] = f"{type(ex).__name__}: {ex}"[:300]

    try:
        # ---- BASELINE.json configs[0]: the reference's own scene (12,500 bodies, merge on) ---------
        c1 = None
        if rank == 0 and world == 1:
            e1 = bh_b200.NativeEngine(device=local_rank)
            e1.set_params(theta=THETA)                      # merge 4000 / 8 px stays on (reference default)
            e1.set_bodies(*scenes_mod.snap_f32(scenes_mod.default_two_disks(seed=1)))
            e1.step(20)
            t0 = time.perf_counter()
            e1.step(300)
            c1 = {"workload": "reference two-disk scene, 12,500 bodies, theta=0.5, merge rule on", "steps_per_s": 300 / (time.perf_counter() - t0),
                  "bodies_left": e1.n}
            e1.close()
        line["configs0"] = c1
    except Exception as ex:                                       # (reported in the line; the headline stands)
        line.setdefault("side_measurement_errors", {})["configs0"] = f"{type(ex).__name__}: {ex}"[:300]

    try:
        # ---- the device accuracy oracle (tiled all-pairs direct sum, BH.kt:250-259 over every pair): FP32-pipe roofline
        direct = None
        if rank == 0 and world == 1:
            direct = []
            for nd in (100_000, 1_000_000):
                if nd > n:
                    continue
                ed = bh_b200.NativeEngine(device=local_rank, capacity_hint=nd)
                ed.set_window(W, H)
                ed.set_params(theta=THETA, merge_min_dist=0.0)
                ed.set_bodies(*[np.ascontiguousarray(a[:nd]) for a in scene])
                ed.direct_sum()
                ed.direct_sum()
                ms = ed.counters()["ms_direct"]
                tf = 14.0 * nd * (nd - 1) / (ms * 1e-3) / 1e12
                direct.append({"kernel": "k_direct", "bodies": nd, "ms": ms, "pair_interactions_per_s": nd * (nd - 1) / (ms * 1e-3), "bound": "fp32",
                               "achieved": tf, "peak": float(fp32[0]), "unit": "TFLOP/s", "frac": tf / float(fp32[0]) if fp32[0] > 0 else None,
                               "algorithmic": "14 flop x N(N-1) pair interactions (SURVEY.md 8(d)); same measured FFMA peak as the walk"})
                ed.close()
        line["roofline_direct"] = direct
    except Exception as ex:                                       # (reported in the line; the headline stands)
        line.setdefault("side_measurement_errors", {})["roofline_direct"] = f"{type(ex).__name__}: {ex}"[:300]

    wd.done()
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    if line.get("parity_failed"):                                 # the line is out (flagged "parity_failed"); the exit status says so too
        sys.stderr.write("[bench] parity_check FAILED: see \"parity_check\" in the JSON line\n")
        sys.exit(3)


if __name__ == "__main__":
    main()
