#!/usr/bin/env python
"""One BASELINE.json config (C3 / C4 / C5, bench.py::run_config) at the launch's GPU count, without the headline:
    torchrun --nproc-per-node N tools/config_run.py C3 [steps warmup]      (or plain python for N = 1)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else {"C3": 20, "C4": 8, "C5": 60}[name]
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))


def _all(op):
    def f(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())
    return f


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


rec = bench.run_config(name, steps, warm, local_rank, rank, world, dist, _all(dist.ReduceOp.MAX) if world > 1 else (lambda v: v),
                       _all(dist.ReduceOp.SUM) if world > 1 else (lambda v: v), barrier)
if rank == 0:
    print(json.dumps(rec), flush=True)
if world > 1:
    dist.destroy_process_group()
