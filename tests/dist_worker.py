"""Worker of the multi-process tests: one rank of a `world`-rank run driven through the
host-staged transport (bh_step_begin / export / import / bh_step_end / bh_step_finish) or
through the CUDA engine's own NCCL transport.  Launched by tests/test_distributed.py with
RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the environment.

argv: <lib: oracle|cuda> <transport: staged|nccl> <scene .npz> <out prefix> <steps> <merge 0|1>
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    which, transport, scene_path, out_prefix, steps, merge = sys.argv[1:7]
    steps, merge = int(steps), int(merge)
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    import torch
    import torch.distributed as dist
    import bh_b200
    from bh_b200.distributed import HostStagedStepper, init_nccl_engine

    cuda = which == "cuda"
    if cuda:
        torch.cuda.set_device(rank % torch.cuda.device_count())
    backend = "nccl" if (cuda and transport == "nccl") else "gloo"
    dist.init_process_group(backend, rank=rank, world_size=world)
    lib = bh_b200.load_cuda_library() if cuda else bh_b200.bind(os.path.join(ROOT, "oracle", "libbh_ref.so"))
    z = np.load(scene_path)
    flags = int(os.environ.get("BH_TEST_FLAGS", "0"))
    e = bh_b200.NativeEngine(lib=lib, device=(rank % torch.cuda.device_count()) if cuda else 0, threads=2,
                             rehome_interval=int(os.environ.get("BH_TEST_REHOME", "3")), flags=flags)
    e.set_window(int(z["W"]), int(z["H"]))
    e.set_params(theta=float(z["theta"]), merge_min_dist=8.0 if merge else 0.0)
    if transport == "nccl":
        init_nccl_engine(e, dist, rank, world)
    e.set_bodies(z["x"], z["y"], z["vx"], z["vy"], z["m"])
    if transport == "nccl":
        e.step(steps)
    else:
        HostStagedStepper(e, dist, rank, world).step(steps)
    x, y, vx, vy, m = e.get_bodies()
    ls = e.let_stats()
    np.savez(f"{out_prefix}.{rank}.npz", x=x, y=y, vx=vx, vy=vy, m=m, origin=e.get_origin(),
             merged=e.counters()["total_merged"], interactions=e.counters()["total_interactions"],
             opened=e.counters()["total_opened"], let=np.array([ls[k] for k in sorted(ls)], np.int64),
             let_keys=np.array(sorted(ls)))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
