"""N > 1: replicated tree, sliced targets, one exchange of the drifted positions per step
(SURVEY.md §8(e)).  The acceptance test is the strongest possible one: every rank's final
state is BIT-IDENTICAL to a single-process run on the same inputs.

CPU (`-m "not gpu"`): world_size 2 over gloo, the host-staged transport driving the oracle.
GPU: the same protocol on the CUDA engine (gloo-staged), the engine's own NCCL transport, and the
DOMAIN mode (BH_FLAG_LET: local trees + locally essential trees, no replicated tree) when the box has
>= 2 GPUs — always against the same bit-identity bar."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import bh_b200
from bh_b200 import scenes
from conftest import ROOT, make_engine

WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_world(world, which, transport, scene_path, out_prefix, steps, merge, timeout=600, extra_env=None):
    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        env.update(extra_env or {})
        procs.append(subprocess.Popen([sys.executable, WORKER, which, transport, scene_path, out_prefix, str(steps), str(merge)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            out, _ = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append(out)
    for p, out in zip(procs, outs):
        assert p.returncode == 0, out[-3000:]
    return [np.load(f"{out_prefix}.{r}.npz") for r in range(world)]


def _scene_file(tmp_path, merge):
    if merge:
        # two heavy bodies with satellites inside the 8 px merge radius + a background disk
        s = list(scenes.snap_f32(scenes.default_two_disks(n1=900, n2=300, seed=21)))
        rng = np.random.default_rng(5)
        k = 40
        ang, rad = rng.uniform(0, 2 * np.pi, k), rng.uniform(1.0, 7.5, k)
        s[0][2:2 + k] = s[0][0] + rad * np.cos(ang)
        s[1][2:2 + k] = s[1][0] + rad * np.sin(ang)
        scene = tuple(s)
    else:
        scene = scenes.snap_f32(scenes.default_two_disks(n1=1500, n2=500, seed=22))
    path = str(tmp_path / "scene.npz")
    np.savez(path, x=scene[0], y=scene[1], vx=scene[2], vy=scene[3], m=scene[4], W=2400, H=800, theta=0.5)
    return scene, path


def _single(lib, scene, steps, merge):
    e = make_engine(lib, scene, theta=0.5, merge_min_dist=8.0 if merge else 0.0)
    e.step(steps)
    return e.get_bodies(), e.get_origin(), e.counters()


def _assert_identical(ranks, single, origin, ctr):
    for z in ranks:
        for k, name in enumerate(("x", "y", "vx", "vy", "m")):
            assert z[name].shape == single[k].shape, name
            assert (z[name] == single[k]).all(), name          # bit-identical
        assert (z["origin"] == origin).all()
        assert int(z["merged"]) == ctr["total_merged"]
    assert sum(int(z["interactions"]) for z in ranks) == ctr["total_interactions"]


@pytest.mark.parametrize("merge", [0, 1], ids=["merge-off", "merge-on"])
def test_gloo_world2_host_staged_oracle_is_bit_identical(oracle_lib, tmp_path, merge):
    scene, path = _scene_file(tmp_path, merge)
    ranks = _run_world(2, "oracle", "staged", path, str(tmp_path / "out"), 4, merge)
    single, origin, ctr = _single(oracle_lib, scene, 4, merge)
    if merge:
        assert ctr["total_merged"] > 0
    _assert_identical(ranks, single, origin, ctr)


def test_slice_bounds_partition_the_bodies(oracle_lib):
    e = bh_b200.NativeEngine(lib=oracle_lib)
    for n in (0, 1, 7, 8, 1000, 12_500):
        for world in (1, 2, 3, 8):
            spans = [e.slice_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == (n + world - 1) // world


@pytest.mark.gpu
@pytest.mark.parametrize("merge", [0, 1], ids=["merge-off", "merge-on"])
def test_cuda_host_staged_world2_is_bit_identical_to_one_gpu(cuda_lib, tmp_path, merge):
    """Two processes (sharing cuda:0 if the box has one GPU), slices exchanged through gloo."""
    scene, path = _scene_file(tmp_path, merge)
    ranks = _run_world(2, "cuda", "staged", path, str(tmp_path / "out"), 7, merge)
    single, origin, ctr = _single(cuda_lib, scene, 7, merge)
    _assert_identical(ranks, single, origin, ctr)


@pytest.mark.gpu
@pytest.mark.parametrize("merge", [0, 1], ids=["merge-off", "merge-on"])
def test_cuda_nccl_is_bit_identical_to_one_gpu(cuda_lib, tmp_path, merge):
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    scene, path = _scene_file(tmp_path, merge)
    ranks = _run_world(world, "cuda", "nccl", path, str(tmp_path / "out"), 7, merge)
    single, origin, ctr = _single(cuda_lib, scene, 7, merge)
    _assert_identical(ranks, single, origin, ctr)


def _jitter_stray_scene():
    """Two disks + 2000 PAIRS of bodies 2e-4 px apart (they share a depth-21 cell: the jitter regime of BH.kt:145-156
    mutates them in every build) that fly at 400 px/t across the code boundaries of the partition, so that pairs become
    strays of their home rank and guests — inside a jitter cluster — of the rank that hosts their cell."""
    s = [a.copy() for a in scenes.snap_f32(scenes.default_two_disks(n1=32000, n2=8000, seed=35))]
    rng = np.random.default_rng(36)
    k = 2000
    half = 1202.0
    w = 2.0 * half / 16.0                                  # side of a level-4 cell (the coarsest cut)
    col = rng.integers(2, 14, k)
    bx = -2.0 + col * w - rng.uniform(0.5, 6.0, k)         # just left of a cell boundary, moving right
    by = rng.uniform(-300.0, 1100.0, k)
    for j in range(k):
        for t in range(2):
            i = 1000 + 2 * j + t                              # (satellites of the first disk are replaced)
            s[0][i], s[1][i] = np.float32(bx[j]) + t * 2.0e-4, np.float32(by[j])
            s[2][i], s[3][i] = 400.0, 0.0
    return tuple(np.ascontiguousarray(a) for a in s)


LET_CASES = [
    ("jitter clusters that cross rank boundaries (positions sent back to the home rank)", _jitter_stray_scene, 0.5, 9, 4),
    ("two-disk 40k", lambda: scenes.snap_f32(scenes.default_two_disks(n1=32000, n2=8000, seed=31)), 0.5, 9, 3),
    ("two-disk 40k, stray overflow -> fallback", lambda: scenes.snap_f32(scenes.default_two_disks(n1=32000, n2=8000, seed=31)), 0.5, 9, 4),
    ("cloud 100k θ0.8", lambda: scenes.snap_f32(scenes.make_uniform_random(100_000, 0.5, seed=32)), 0.8, 6, 2),
    ("two-disk 2k (coarsest cut)", lambda: scenes.snap_f32(scenes.default_two_disks(n1=1500, n2=500, seed=22)), 0.3, 7, 3),
    ("cloud 100k θ0.5, blocks over ncclSend/ncclRecv", lambda: scenes.snap_f32(scenes.make_uniform_random(100_000, 0.5, seed=33)), 0.5, 5, 2),
]


@pytest.mark.gpu
@pytest.mark.parametrize("name,gen,theta,steps,rehome", LET_CASES, ids=[c[0] for c in LET_CASES])
def test_cuda_domain_mode_is_bit_identical_to_one_gpu(cuda_lib, tmp_path, name, gen, theta, steps, rehome):
    """BH_FLAG_LET: every rank builds only its own Morton range and walks a locally essential tree
    (no replicated tree, no per-step all-gather of positions).  State after several steps — re-homing
    builds, LET evaluations with strays in between — bit-identical to one GPU; interaction and
    opened-cell totals equal."""
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    scene = gen()
    path = str(tmp_path / "scene.npz")
    np.savez(path, x=scene[0], y=scene[1], vx=scene[2], vy=scene[3], m=scene[4], W=2400, H=800, theta=theta)
    env = {"BH_TEST_FLAGS": str(bh_b200.BH_FLAG_LET), "BH_TEST_REHOME": str(rehome), "BH_LET_MIN_WORLD": "2"}
    overflow = "overflow" in name
    if "ncclSend" in name:
        env["BH_LET_IPC"] = "0"                # no peer-memory mapping: the packed-block exchange
    if overflow:
        env["BH_LET_STRAY_CAP"] = "3"          # more strays than a segment holds: the evaluation is redone after a re-homing
    ranks = _run_world(world, "cuda", "nccl", path, str(tmp_path / "out"), steps, 0, extra_env=env)
    e = make_engine(cuda_lib, scene, theta=theta, merge_min_dist=0.0)
    e.step(steps)
    single, ctr = e.get_bodies(), e.counters()
    for z in ranks:
        for k, nm in enumerate(("x", "y", "vx", "vy", "m")):
            assert (z[nm] == single[k]).all(), nm
        st = dict(zip([str(s) for s in z["let_keys"]], [int(v) for v in z["let"]]))
        # enabled: 2 = blocks imported over NVLink peer memory (CUDA IPC), 1 = over ncclSend/ncclRecv
        assert st["enabled"] == (1 if "ncclSend" in name else 2) and st["let_evaluations"] > 0 and (st["fallbacks"] > 0) == overflow, st
        if "jitter" in name:
            assert st["fallbacks_guest_in_jitter_cluster"] == 0, st
    if "jitter" in name:
        returned = sum(dict(zip([str(s) for s in z["let_keys"]], [int(v) for v in z["let"]]))["jitter_positions_returned"] for z in ranks)
        assert returned > 0, returned                      # guests inside jitter clusters did occur, and were sent back
    assert sum(int(z["interactions"]) for z in ranks) == ctr["total_interactions"]
    assert sum(int(z["opened"]) for z in ranks) == ctr["total_opened"]
