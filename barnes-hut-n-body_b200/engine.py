"""Host side of the drop-in boundary.

Two layers:

* :class:`NativeEngine` — thin numpy/ctypes wrapper over one ``bh_engine*`` handle of
  ``include/bh_engine.h`` (SoA f64 arrays in and out).
* :class:`Config`, :class:`Body`, :class:`Quad`, :class:`BHTree`, :class:`PhysicsEngine` —
  the reference's Kotlin class API (``BarnesHutAlg.kt`` / ``Config.kt``) with the same
  names, argument meaning and behaviour, driving the CUDA engine.  ``NBodyPanel.kt``'s
  calls (``step()``, ``getBodies()``, ``resetBodies()``, ``getTreeForDebug().visitQuads``)
  map 1:1; the Kotlin/JNA twin of this file is shown in INTEGRATION.md.

Paths cited as ``BH.kt:a-b`` are /root/reference/src/main/kotlin/BarnesHutAlg.kt.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _abi
from ._abi import BhConfig, BhCounters, BhDiskParams, BhError, BhParams


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class NativeEngine:
    """One ``bh_engine*``.  `lib` defaults to the CUDA product library (no CPU fallback)."""

    def __init__(self, lib: Optional[C.CDLL] = None, device: int = 0, threads: int = 0,
                 flags: int = 0, capacity_hint: int = 0, rehome_interval: int = 0):
        self.lib = lib if lib is not None else _abi.load_cuda_library()
        cfg = BhConfig(C.sizeof(BhConfig), device, threads, flags, capacity_hint, rehome_interval, 0)
        h = C.c_void_p()
        rc = self.lib.bh_create(C.byref(cfg), C.byref(h))
        if rc != _abi.BH_OK:
            raise BhError(rc, "bh_create", (self.lib.bh_last_error(None) or b"").decode())
        self._h = h
        self.backend = self.lib.bh_backend_name().decode()

    # -- plumbing ---------------------------------------------------------------------
    def _check(self, rc: int, where: str):
        if rc != _abi.BH_OK:
            raise BhError(rc, where, (self.lib.bh_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.bh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- parameters -------------------------------------------------------------------
    def default_params(self, width_px: int = 2400, height_px: int = 800) -> BhParams:
        p = BhParams()
        self._check(self.lib.bh_default_params(width_px, height_px, C.byref(p)), "bh_default_params")
        return p

    @property
    def params(self) -> BhParams:
        p = BhParams()
        self._check(self.lib.bh_get_params(self._h, C.byref(p)), "bh_get_params")
        return p

    def set_params(self, p: Optional[BhParams] = None, **kw):
        q = self.params if p is None else p
        for k, v in kw.items():
            if not hasattr(q, k):
                raise AttributeError(k)
            setattr(q, k, float(v))
        self._check(self.lib.bh_set_params(self._h, C.byref(q)), "bh_set_params")

    def set_window(self, width_px: int, height_px: int):
        """Root box of buildTree() for a W x H window (BH.kt:360-361)."""
        d = self.default_params(width_px, height_px)
        self.set_params(root_cx=d.root_cx, root_cy=d.root_cy, root_half=d.root_half)

    # -- state ------------------------------------------------------------------------
    def set_bodies(self, x, y, vx, vy, m):
        x, y, vx, vy, m = map(_f64, (x, y, vx, vy, m))
        n = x.shape[0]
        if not all(a.shape == (n,) for a in (y, vx, vy, m)):
            raise ValueError("x, y, vx, vy, m must be 1-D arrays of equal length")
        self._check(self.lib.bh_set_bodies(self._h, n, _dp(x), _dp(y), _dp(vx), _dp(vy), _dp(m)), "bh_set_bodies")

    @property
    def n(self) -> int:
        return int(self.lib.bh_num_bodies(self._h))

    def get_bodies(self, out=None):
        n = self.n
        arrs = out if out is not None else tuple(np.empty(n, np.float64) for _ in range(5))
        n_out = C.c_int64()
        self._check(self.lib.bh_get_bodies(self._h, n, *[_dp(a) for a in arrs], C.byref(n_out)), "bh_get_bodies")
        return arrs

    def slice_index(self) -> np.ndarray:
        """Positions in the `bodies` list of the bodies of this rank's slice, in home order (bh_get_slice_index)."""
        n = self.n
        ui = np.empty(n, np.int32)
        k = C.c_int64()
        self._check(self.lib.bh_get_slice_index(self._h, n, _ip(ui), C.byref(k)), "bh_get_slice_index")
        return ui[:k.value].copy()

    def slice_epoch(self) -> int:
        return int(self.lib.bh_slice_epoch(self._h))

    def step_io_slice(self, nsteps: int = 1, inputs=None, out=None) -> int:
        """[this rank's slice in] ; nsteps x step() ; [slice out] — every rank moves only its own bodies
        (bh_step_io_slice).  `inputs` / `out`: 5 float64 arrays in slice order; returns the slice length after the steps."""
        ins, n_in = [None] * 5, 0
        if inputs is not None:
            ins = [_dp(a) for a in inputs]
            n_in = len(inputs[0])
        outs, cap = [None] * 5, 0
        if out is not None:
            outs = [_dp(a) for a in out]
            cap = len(out[0])
        k = C.c_int64()
        self._check(self.lib.bh_step_io_slice(self._h, nsteps, n_in, *ins, cap, *outs, C.byref(k)), "bh_step_io_slice")
        return int(k.value)

    def evaluate_slice(self):
        """One evaluation of this rank's slice in the engine's current multi-GPU mode: (ax, ay, user_index)."""
        n = self.n
        ax, ay, ui = np.empty(n, np.float64), np.empty(n, np.float64), np.empty(n, np.int32)
        k = C.c_int64()
        self._check(self.lib.bh_evaluate_slice(self._h, n, _dp(ax), _dp(ay), _ip(ui), C.byref(k)), "bh_evaluate_slice")
        return ax[:k.value], ay[:k.value], ui[:k.value]

    def set_domain_mode(self, enabled: bool) -> None:
        self._check(self.lib.bh_set_domain_mode(self._h, 1 if enabled else 0), "bh_set_domain_mode")

    def rebase_origin(self) -> None:
        """Make the current list the reference list of ``get_origin`` (after the caller has dropped
        the merged-away ``Body`` objects from its own list, BH.kt:519)."""
        self._check(self.lib.bh_rebase_origin(self._h), "bh_rebase_origin")

    def get_origin(self) -> np.ndarray:
        n = self.n
        o = np.empty(n, np.int32)
        n_out = C.c_int64()
        self._check(self.lib.bh_get_origin(self._h, n, _ip(o), C.byref(n_out)), "bh_get_origin")
        return o

    def get_positions_f32(self):
        n = self.n
        xy = np.empty((n, 2), np.float32)
        m = np.empty(n, np.float32)
        n_out = C.c_int64()
        self._check(self.lib.bh_get_positions_f32(self._h, n, xy.ctypes.data_as(C.POINTER(C.c_float)),
                                                  m.ctypes.data_as(C.POINTER(C.c_float)), C.byref(n_out)),
                    "bh_get_positions_f32")
        return xy, m

    # -- scene generators on the device (BodyFactory.kt) -------------------------------------
    def disk_params(self, width_px: int = 2400, height_px: int = 800, **kw) -> BhDiskParams:
        p = BhDiskParams()
        self._check(self.lib.bh_default_disk_params(width_px, height_px, C.byref(p)), "bh_default_disk_params")
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, type(getattr(p, k))(v))
        return p

    def append_disk(self, n_total: int, params: BhDiskParams, seed: int = 1):
        """bodies += BodyFactory.makeGalaxyDisk / makeKeplerDisk(n_total, ...) generated on the device."""
        self._check(self.lib.bh_append_disk(self._h, n_total, C.byref(params), seed), "bh_append_disk")

    def append_uniform_random(self, n: int, m: float, width_px: int = 2400, height_px: int = 800, seed: int = 1):
        """bodies += BodyFactory.makeUniformRandom(n, m) generated on the device."""
        self._check(self.lib.bh_append_uniform_random(self._h, n, float(m), width_px, height_px, seed), "bh_append_uniform_random")

    def request_positions_f32(self):
        """Start an asynchronous float snapshot of (x, y, m); overlaps the next step()."""
        self._check(self.lib.bh_request_positions_f32(self._h), "bh_request_positions_f32")

    def wait_positions_f32(self):
        """(xy[n,2], m[n]) of the last requested snapshot (copies out of the engine's pinned buffers)."""
        pxy, pm, n = C.POINTER(C.c_float)(), C.POINTER(C.c_float)(), C.c_int64()
        self._check(self.lib.bh_wait_positions_f32(self._h, C.byref(pxy), C.byref(pm), C.byref(n)), "bh_wait_positions_f32")
        k = n.value
        if k == 0:
            return np.empty((0, 2), np.float32), np.empty(0, np.float32)
        xy = np.ctypeslib.as_array(pxy, shape=(k, 2)).copy()
        m = np.ctypeslib.as_array(pm, shape=(k,)).copy()
        return xy, m

    # -- compute ----------------------------------------------------------------------
    def step(self, nsteps: int = 1):
        self._check(self.lib.bh_step(self._h, nsteps), "bh_step")

    def step_io(self, nsteps: int = 1, inputs=None, out=None):
        """resetBodies(inputs); nsteps x step(); getBodies() in one call with the transfers
        overlapped with the compute (bh_step_io).  `inputs` / `out`: 5 float64 arrays (pinned host
        memory makes the copies asynchronous); returns the output arrays (or None)."""
        n_in = 0
        ins = [None] * 5
        if inputs is not None:
            ins = [_f64(a) for a in inputs]
            n_in = ins[0].shape[0]
        n_after = n_in if inputs is not None else self.n
        outs = [None] * 5
        if out is not None:
            outs = list(out)
        elif out is None and inputs is not None:
            outs = [np.empty(n_after, np.float64) for _ in range(5)]
        cap = outs[0].shape[0] if outs[0] is not None else 0
        n_out = C.c_int64()
        self._check(self.lib.bh_step_io(self._h, nsteps, n_in, *[_dp(a) for a in ins], cap, *[_dp(a) for a in outs], C.byref(n_out)),
                    "bh_step_io")
        if outs[0] is None:
            return None
        k = n_out.value
        return tuple(a[:k] for a in outs)

    def compute_accelerations(self):
        n = self.n
        ax, ay = np.empty(n, np.float64), np.empty(n, np.float64)
        self._check(self.lib.bh_compute_accelerations(self._h, _dp(ax), _dp(ay)), "bh_compute_accelerations")
        return ax, ay

    def direct_sum(self):
        n = self.n
        ax, ay = np.empty(n, np.float64), np.empty(n, np.float64)
        self._check(self.lib.bh_direct_sum(self._h, _dp(ax), _dp(ay)), "bh_direct_sum")
        return ax, ay

    def energy(self):
        v = [C.c_double() for _ in range(4)]
        self._check(self.lib.bh_energy(self._h, *[C.byref(t) for t in v]), "bh_energy")
        ke, pe, px, py = (t.value for t in v)
        return {"kinetic": ke, "potential": pe, "total": ke + pe, "px": px, "py": py}

    def energy_tree(self, theta: float = 0.0):
        """Energy / momentum with the potential from the tree (O(N log N)); theta <= 0: the engine's."""
        v = [C.c_double() for _ in range(4)]
        self._check(self.lib.bh_energy_tree(self._h, float(theta), *[C.byref(t) for t in v]), "bh_energy_tree")
        ke, pe, px, py = (t.value for t in v)
        return {"kinetic": ke, "potential": pe, "total": ke + pe, "px": px, "py": py}

    # -- introspection ----------------------------------------------------------------
    def build_tree(self):
        self._check(self.lib.bh_build_tree(self._h), "bh_build_tree")

    def morton(self):
        n = self.n
        key = np.empty(n, np.uint64)
        depth = np.empty(n, np.int32)
        order = np.empty(n, np.int32)
        self._check(self.lib.bh_get_morton(self._h, key.ctypes.data_as(C.POINTER(C.c_uint64)), _ip(depth), _ip(order)),
                    "bh_get_morton")
        return key, depth, order

    def tree(self):
        """visitQuads preorder over all cells (BH.kt:265-274) as a dict of arrays."""
        ncells = C.c_int64()
        self._check(self.lib.bh_get_tree(self._h, 0, C.byref(ncells), *([None] * 7)), "bh_get_tree")
        k = ncells.value
        a = {name: np.empty(k, np.float64) for name in ("cx", "cy", "h", "mass", "comx", "comy")}
        body = np.empty(k, np.int32)
        self._check(self.lib.bh_get_tree(self._h, k, C.byref(ncells), _dp(a["cx"]), _dp(a["cy"]), _dp(a["h"]),
                                         _dp(a["mass"]), _dp(a["comx"]), _dp(a["comy"]), _ip(body)), "bh_get_tree")
        a["body"] = body
        return a

    def tree_root(self) -> dict:
        """mass / centre of mass of the root cell (BHTree.mass/comX/comY, BH.kt:103-109) without
        exporting the whole tree; ``n_cells`` = internal + body-leaf cells."""
        v = [C.c_double(), C.c_double(), C.c_double()]
        k = C.c_int64()
        self._check(self.lib.bh_get_tree_root(self._h, C.byref(v[0]), C.byref(v[1]), C.byref(v[2]), C.byref(k)), "bh_get_tree_root")
        return {"mass": v[0].value, "comx": v[1].value, "comy": v[2].value, "n_cells": int(k.value)}

    def counters(self) -> dict:
        c = BhCounters()
        self._check(self.lib.bh_get_counters(self._h, C.byref(c)), "bh_get_counters")
        return c.as_dict()

    def reset_counters(self):
        self._check(self.lib.bh_reset_counters(self._h), "bh_reset_counters")

    def body_counts(self):
        n = self.n
        i, o = np.empty(n, np.int32), np.empty(n, np.int32)
        self._check(self.lib.bh_get_body_counts(self._h, _ip(i), _ip(o)), "bh_get_body_counts")
        return i, o

    def let_stats(self) -> dict:
        """Domain-mode (BH_FLAG_LET) statistics of the last evaluation on this rank."""
        v = np.zeros(26, np.int64)
        self._check(self.lib.bh_get_let_stats(self._h, v.ctypes.data_as(C.POINTER(C.c_int64)), 26), "bh_get_let_stats")
        # [0] 0 = off, 1 = on (blocks over ncclSend/ncclRecv), 2 = on (blocks over NVLink peer memory)
        names = ("enabled", "partition_valid", "cut_level", "let_evaluations", "fallbacks", "let_cells", "cells_imported",
                 "cells_sent", "own_strays", "top_items", "fallbacks_guest_in_jitter_cluster", "fallbacks_stray_overflow",
                 "fallbacks_cell_overflow", "stray_capacity", "jitter_positions_returned", "stray_leaf_descents", "stray_leaf_scans")
        return {k: int(x) for k, x in zip(names, v)}

    # -- multi-GPU --------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(_abi.BH_COMM_ID_BYTES)
        rc = self.lib.bh_comm_unique_id(buf, _abi.BH_COMM_ID_BYTES)
        if rc != _abi.BH_OK:
            raise BhError(rc, "bh_comm_unique_id")
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, _abi.BH_COMM_ID_BYTES)
        self._check(self.lib.bh_comm_init(self._h, rank, world, buf, _abi.BH_COMM_ID_BYTES), "bh_comm_init")

    def comm_init_external(self, rank: int, world: int):
        """Host-staged transport: the caller exchanges slices (see :mod:`.distributed`)."""
        self._check(self.lib.bh_comm_init_external(self._h, rank, world), "bh_comm_init_external")

    def step_begin(self):
        self._check(self.lib.bh_step_begin(self._h), "bh_step_begin")

    def step_end(self):
        self._check(self.lib.bh_step_end(self._h), "bh_step_end")

    def step_finish(self):
        self._check(self.lib.bh_step_finish(self._h), "bh_step_finish")

    def export_slice(self, field: int):
        """(a, b, lo, hi): this rank's slice of (x, y) or (vx, vy) in the engine's home order."""
        n = self.n
        lo, hi = C.c_int64(), C.c_int64()
        a, b = np.empty(n, np.float64), np.empty(n, np.float64)
        self._check(self.lib.bh_export_slice(self._h, field, n, _dp(a), _dp(b), C.byref(lo), C.byref(hi)), "bh_export_slice")
        k = hi.value - lo.value
        return a[:k], b[:k], lo.value, hi.value

    def import_slices(self, field: int, a, b):
        a, b = _f64(a), _f64(b)
        self._check(self.lib.bh_import_slices(self._h, field, a.shape[0], _dp(a), _dp(b)), "bh_import_slices")

    def slice_bounds(self, n: int, world: int, rank: int):
        lo, hi = C.c_int64(), C.c_int64()
        rc = self.lib.bh_slice_bounds(n, world, rank, C.byref(lo), C.byref(hi))
        if rc != _abi.BH_OK:
            raise BhError(rc, "bh_slice_bounds")
        return lo.value, hi.value


# =====================================================================================
# The reference's Kotlin API, same names (Config.kt, BarnesHutAlg.kt)
# =====================================================================================
class _Config:
    """`object Config` — Config.kt:2-40.  Process-global mutable singleton; the engine
    re-reads G/DT/theta/WIDTH_PX/HEIGHT_PX at every step like BH.kt:256,360-361,378,412."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.FULL_SCREEN_MODE = True
        self.WIDTH_PX = 2400          # Config.kt:5
        self.HEIGHT_PX = 800          # Config.kt:8
        self.G = 80.0                 # Config.kt:11
        self.DT = 0.005               # Config.kt:14
        self.SOFTENING = 1.0          # Config.kt:17
        self._SOFT2 = self.SOFTENING * self.SOFTENING  # Config.kt:20 — a `val`, fixed at init
        self.theta = 0.30             # Config.kt:23
        self.R = 100.0                # Config.kt:26
        self.N = 5_000                # Config.kt:29

    @property
    def SOFT2(self) -> float:
        return self._SOFT2

    CENTRAL_MASS = 50_000.0           # Config.kt:32
    MIN_R = 8.0                       # Config.kt:35
    TOTAL_SATELLITE_MASS = 5_000.0    # Config.kt:38


Config = _Config()


@dataclass
class Body:
    """BH.kt:21-25 — mutable particle state."""
    x: float
    y: float
    vx: float
    vy: float
    m: float


@dataclass(frozen=True)
class Quad:
    """BH.kt:53-82 — cell [cx-h, cx+h) x [cy-h, cy+h)."""
    cx: float
    cy: float
    h: float

    def contains(self, b: Body) -> bool:  # BH.kt:61-62
        return (b.x >= self.cx - self.h and b.x < self.cx + self.h
                and b.y >= self.cy - self.h and b.y < self.cy + self.h)

    def child(self, which: int) -> "Quad":  # BH.kt:73-81
        hh = self.h / 2.0
        if which == 0:
            return Quad(self.cx - hh, self.cy - hh, hh)
        if which == 1:
            return Quad(self.cx + hh, self.cy - hh, hh)
        if which == 2:
            return Quad(self.cx - hh, self.cy + hh, hh)
        return Quad(self.cx + hh, self.cy + hh, hh)


class Acc:
    """BH.kt:33-41 — per-worker force accumulator."""
    __slots__ = ("fx", "fy")

    def __init__(self):
        self.fx = 0.0
        self.fy = 0.0

    def reset(self):
        self.fx = 0.0
        self.fy = 0.0


class BHTree:
    """Read-only host view of the device quadtree (BH.kt:95-275).

    The tree lives on the GPU as a flattened preorder array; this view holds the export
    of ``bh_get_tree`` (all cells incl. empty leaves, ``visitQuads`` order; ``body[k]`` = index of
    the body in a leaf, -1 empty leaf, -2 internal cell, which always has 4 children)."""

    def __init__(self, cells: dict, bodies: Optional[List["Body"]] = None):
        self._c = cells
        self._bodies = bodies

    def insert(self, b: "Body"):   # BH.kt:125-137
        raise NotImplementedError("BHTree is a read-only view of the device tree (built from the engine's body list, "
                                  "BH.kt:359-366); use PhysicsEngine.resetBodies(old + new)")

    def computeMass(self):          # BH.kt:173-202: masses / centres of mass arrive computed (bit-identical f64)
        return None

    def accumulateForce(self, b: "Body", theta2: float, acc: Acc) -> None:
        """BH.kt:215-239 in f64 over the exported cells, same expression order (diagnostics / tests).
        `b` is skipped by identity (BH.kt:219) when the view knows the engine's body list."""
        c = self._c
        body, mass, comx, comy, h = c["body"], c["mass"], c["comx"], c["comy"], c["h"]
        if len(body) == 0:
            return
        soft2, G = Config.SOFT2, Config.G
        bodies = self._bodies

        def skip(p):
            if body[p] != -2:
                return p + 1
            q = p + 1
            for _ in range(4):
                q = skip(q)
            return q

        def point(px, py, m):       # BH.kt:250-259
            dx = px - b.x
            dy = py - b.y
            r2 = dx * dx + dy * dy + soft2
            invR = 1.0 / math.sqrt(r2)
            invR2 = 1.0 / r2
            f = G * b.m * m * invR2
            acc.fx += f * dx * invR
            acc.fy += f * dy * invR

        def walk(p):
            if mass[p] == 0.0:      # BH.kt:216
                return skip(p)
            if body[p] != -2:       # leaf, BH.kt:218-221
                k = int(body[p])
                if k >= 0 and not (bodies is not None and bodies[k] is b):
                    point(float(comx[p]), float(comy[p]), float(mass[p]))
                return p + 1
            dx = float(comx[p]) - b.x
            dy = float(comy[p]) - b.y
            dist2 = dx * dx + dy * dy + soft2
            side = float(h[p]) * 2.0
            if side * side < theta2 * dist2:
                point(float(comx[p]), float(comy[p]), float(mass[p]))
                return skip(p)
            q = p + 1
            for _ in range(4):
                q = walk(q)
            return q

        walk(0)

    @property
    def mass(self) -> float:   # BH.kt:103
        return float(self._c["mass"][0]) if len(self._c["mass"]) else 0.0

    @property
    def comX(self) -> float:   # BH.kt:106
        return float(self._c["comx"][0]) if len(self._c["comx"]) else 0.0

    @property
    def comY(self) -> float:   # BH.kt:109
        return float(self._c["comy"][0]) if len(self._c["comy"]) else 0.0

    def cells(self) -> dict:
        return self._c

    def visitQuads(self, visit: Callable[[Quad], None]) -> None:  # BH.kt:265-274
        c = self._c
        for cx, cy, h in zip(c["cx"].tolist(), c["cy"].tolist(), c["h"].tolist()):
            visit(Quad(cx, cy, h))


class PhysicsEngine:
    """BH.kt:287-533 with the CUDA engine behind it.

    Keeps the reference's contract: bodies are held by reference, ``step()`` blocks and
    on return the SAME ``Body`` objects carry the state at t+dt (merged-away bodies are
    removed from the list), ``Config`` is re-read at every call."""

    def __init__(self, initialBodies: List[Body], *, lib: Optional[C.CDLL] = None, device: int = 0):
        self._native = NativeEngine(lib=lib, device=device)
        self.mergeMaxMass: float = 4_000.0          # BH.kt:315
        self.mergeMinDist: float = Config.MIN_R     # BH.kt:321
        self._bodies: List[Body] = initialBodies
        self._lastTree: Optional[BHTree] = None     # BH.kt:304
        self._upload()

    # -- helpers ----------------------------------------------------------------------
    def _push_config(self):
        d = self._native.default_params(Config.WIDTH_PX, Config.HEIGHT_PX)   # BH.kt:360-361
        d.G, d.dt, d.theta, d.soft2 = Config.G, Config.DT, Config.theta, Config.SOFT2
        d.merge_max_mass, d.merge_min_dist = self.mergeMaxMass, self.mergeMinDist
        self._native.set_params(d)

    def _upload(self):
        bs = self._bodies
        n = len(bs)
        a = np.empty((5, n), np.float64)
        for i, b in enumerate(bs):
            a[0, i], a[1, i], a[2, i], a[3, i], a[4, i] = b.x, b.y, b.vx, b.vy, b.m
        self._native.set_bodies(a[0], a[1], a[2], a[3], a[4])

    def _download(self):
        x, y, vx, vy, m = self._native.get_bodies()
        origin = self._native.get_origin()
        bs = self._bodies
        n_before = len(bs)
        if len(origin) != n_before:  # the merge rule removed bodies (BH.kt:514-520)
            keep = [bs[k] for k in origin.tolist()]
            bs[:] = keep             # same list object, like bodies.removeAt
        for b, xi, yi, vxi, vyi, mi in zip(bs, x.tolist(), y.tolist(), vx.tolist(), vy.tolist(), m.tolist()):
            b.x, b.y, b.vx, b.vy, b.m = xi, yi, vxi, vyi, mi
        if len(origin) != n_before:
            self._native.rebase_origin()   # origin[] now indexes the shrunk list (nothing is re-uploaded)

    # -- public API -------------------------------------------------------------------
    def getTreeForDebug(self) -> BHTree:            # BH.kt:329-332
        if self._lastTree is None:
            self._push_config()
            # bh_get_tree exports the tree cached by the last step, or builds a fresh
            # one if a reset/merge dropped it — exactly `lastTree ?: buildTree()`.
            self._lastTree = BHTree(self._native.tree(), self._bodies)
        return self._lastTree

    def getBodies(self) -> List[Body]:              # BH.kt:335
        return self._bodies

    def resetBodies(self, newBodies: List[Body]):   # BH.kt:342-349
        self._bodies = newBodies
        self._upload()
        self._lastTree = None

    def step(self):                                 # BH.kt:405-439
        self._push_config()
        self._native.step(1)
        self._download()
        self._lastTree = None   # exported lazily on the next getTreeForDebug()

    @property
    def native(self) -> NativeEngine:
        return self._native
