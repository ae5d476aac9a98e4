"""One process per GPU (or per host) over ``torch.distributed``.

The reference is single-process; this is the partition BASELINE.json's north star asks for:
the tree is replicated, every rank walks and integrates a contiguous slice of the (Morton
ordered) bodies, and the drifted positions are exchanged once per step.  Results are
bit-identical to a single process because every body's walk sees the same replicated tree.

With ``flags=BH_FLAG_LET`` and the NCCL transport the engine switches, from 4 ranks up, to the DOMAIN
mode: every rank builds only the tree of its own Morton range and walks a locally essential tree
 — same results, no replicated build, no per-step all-gather.

Two transports (include/bh_engine.h):

* :func:`init_nccl_engine` — the CUDA engine joins an NCCL communicator itself and
  ``bh_step`` runs whole steps on the device (all-gather over NVLink / NVSwitch);
  ``torch.distributed`` only ships the 128-byte NCCL id.
* :class:`HostStagedStepper` — ``bh_step_begin / bh_export_slice / bh_import_slices /
  bh_step_end / bh_step_finish`` with the slices moved by ``torch.distributed.all_gather``
  on whatever backend the process group uses (gloo on CPU boxes).  Works with ANY library
  exporting the ABI; the CPU tests run it over the oracle with gloo.
"""
from __future__ import annotations

import numpy as np

from . import _abi
from .engine import NativeEngine


def init_nccl_engine(engine: NativeEngine, dist, rank: int, world: int) -> None:
    """Broadcast an NCCL unique id from rank 0 and join the in-engine communicator."""
    uid = [engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    engine.comm_init(rank, world, uid[0])


class HostStagedStepper:
    """Drives PhysicsEngine.step() (BarnesHutAlg.kt:405-439) across ranks with host-staged
    exchanges.  Every rank must have called ``set_bodies`` with the same full body list."""

    def __init__(self, engine: NativeEngine, dist, rank: int, world: int, device: str = "cpu"):
        import torch
        self.torch = torch
        self.e, self.dist, self.rank, self.world, self.device = engine, dist, rank, world, device
        engine.comm_init_external(rank, world)

    def _all_gather(self, field: int):
        """All ranks' slices of `field`, concatenated in rank order = home order."""
        torch = self.torch
        n = self.e.n
        a, b, lo, hi = self.e.export_slice(field)
        per = (n + self.world - 1) // self.world        # bh_slice_bounds: equal padded slices
        mine = torch.zeros(2, max(per, 1), dtype=torch.float64, device=self.device)
        if hi > lo:
            mine[0, :hi - lo] = torch.from_numpy(a).to(self.device)
            mine[1, :hi - lo] = torch.from_numpy(b).to(self.device)
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(parts, mine)
        full = torch.cat(parts, dim=1)[:, :n].cpu().numpy()
        self.e.import_slices(field, np.ascontiguousarray(full[0]), np.ascontiguousarray(full[1]))

    def step(self, nsteps: int = 1):
        for _ in range(nsteps):
            self.e.step_begin()
            if self.world > 1:
                self._all_gather(_abi.BH_FIELD_POS)
            self.e.step_end()
            if self.world > 1:
                self._all_gather(_abi.BH_FIELD_VEL)
            self.e.step_finish()
