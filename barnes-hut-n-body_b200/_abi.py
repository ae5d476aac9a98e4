"""ctypes declarations of include/bh_engine.h.

The product library is ``csrc/libbh_b200.so`` (hand-written sm_100a CUDA).  There is
no CPU fallback: :func:`load_cuda_library` raises if the library is missing or is not
the CUDA backend.  The same declarations can be bound to any other library exporting
the ABI (the tests bind them to the oracle) through :func:`bind`.
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 1

BH_OK = 0
BH_E_ARG = -1
BH_E_CUDA = -2
BH_E_NCCL = -3
BH_E_OOM = -4
BH_E_STATE = -5
BH_E_UNSUPPORTED = -6
_ERR_NAMES = {
    BH_E_ARG: "BH_E_ARG", BH_E_CUDA: "BH_E_CUDA", BH_E_NCCL: "BH_E_NCCL", BH_E_OOM: "BH_E_OOM",
    BH_E_STATE: "BH_E_STATE", BH_E_UNSUPPORTED: "BH_E_UNSUPPORTED",
}

BH_FLAG_BODY_COUNTS = 1
BH_FLAG_REUSE_ACC = 2
BH_FLAG_LET = 4
BH_COMM_ID_BYTES = 128
BH_FIELD_POS = 0
BH_FIELD_VEL = 1


class BhConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("threads", C.c_int32),
                ("flags", C.c_uint32), ("capacity_hint", C.c_int64),
                ("rehome_interval", C.c_int32), ("reserved", C.c_int32)]


class BhParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("G", "dt", "theta", "soft2", "root_cx", "root_cy", "root_half",
                                          "merge_max_mass", "merge_min_dist")]


class BhDiskParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("x", "y", "vx", "vy", "r", "min_r", "central_mass", "total_satellite_mass",
                                          "eps_m2", "phi0", "bar_taper_r", "radial_scale", "speed_jitter", "radial_jitter")] + \
               [("clockwise", C.c_int32), ("kepler", C.c_int32)]


class BhCounters(C.Structure):
    _fields_ = [("n_bodies", C.c_int64), ("n_in_tree", C.c_int64), ("n_out_of_box", C.c_int64),
                ("n_jitter_bodies", C.c_int64), ("n_cells", C.c_int64), ("n_internal", C.c_int64),
                ("key_levels", C.c_int32), ("max_depth", C.c_int32),
                ("interactions", C.c_int64), ("opened", C.c_int64), ("exact_retests", C.c_int64),
                ("total_interactions", C.c_int64), ("total_opened", C.c_int64),
                ("total_evaluations", C.c_int64), ("total_steps", C.c_int64), ("total_merged", C.c_int64),
                ("ms_build", C.c_double), ("ms_walk", C.c_double), ("ms_integrate", C.c_double),
                ("ms_merge", C.c_double), ("ms_comm", C.c_double),
                ("ms_step_call", C.c_double), ("kernel_launches", C.c_int64),
                ("bbox_min_x", C.c_double), ("bbox_max_x", C.c_double), ("bbox_min_y", C.c_double), ("bbox_max_y", C.c_double),
                ("ms_direct", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_H = C.c_void_p
_D = C.POINTER(C.c_double)
_F = C.POINTER(C.c_float)
_I32 = C.POINTER(C.c_int32)
_I64 = C.POINTER(C.c_int64)
_U64 = C.POINTER(C.c_uint64)

# name -> (restype, argtypes).  Every symbol include/bh_engine.h declares.
SYMBOLS = {
    "bh_create": (C.c_int, [C.POINTER(BhConfig), C.POINTER(_H)]),
    "bh_destroy": (None, [_H]),
    "bh_last_error": (C.c_char_p, [_H]),
    "bh_backend_name": (C.c_char_p, []),
    "bh_abi_version": (C.c_int, []),
    "bh_default_params": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(BhParams)]),
    "bh_set_params": (C.c_int, [_H, C.POINTER(BhParams)]),
    "bh_get_params": (C.c_int, [_H, C.POINTER(BhParams)]),
    "bh_set_bodies": (C.c_int, [_H, C.c_int64, _D, _D, _D, _D, _D]),
    "bh_get_bodies": (C.c_int, [_H, C.c_int64, _D, _D, _D, _D, _D, _I64]),
    "bh_num_bodies": (C.c_int64, [_H]),
    "bh_get_origin": (C.c_int, [_H, C.c_int64, _I32, _I64]),
    "bh_rebase_origin": (C.c_int, [_H]),
    "bh_get_positions_f32": (C.c_int, [_H, C.c_int64, _F, _F, _I64]),
    "bh_default_disk_params": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(BhDiskParams)]),
    "bh_append_disk": (C.c_int, [_H, C.c_int64, C.POINTER(BhDiskParams), C.c_uint64]),
    "bh_append_uniform_random": (C.c_int, [_H, C.c_int64, C.c_double, C.c_int32, C.c_int32, C.c_uint64]),
    "bh_request_positions_f32": (C.c_int, [_H]),
    "bh_wait_positions_f32": (C.c_int, [_H, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.POINTER(C.c_float)), _I64]),
    "bh_step": (C.c_int, [_H, C.c_int32]),
    "bh_step_io": (C.c_int, [_H, C.c_int32, C.c_int64, _D, _D, _D, _D, _D, C.c_int64, _D, _D, _D, _D, _D, _I64]),
    "bh_compute_accelerations": (C.c_int, [_H, _D, _D]),
    "bh_direct_sum": (C.c_int, [_H, _D, _D]),
    "bh_energy": (C.c_int, [_H, _D, _D, _D, _D]),
    "bh_energy_tree": (C.c_int, [_H, C.c_double, _D, _D, _D, _D]),
    "bh_get_morton": (C.c_int, [_H, _U64, _I32, _I32]),
    "bh_get_tree": (C.c_int, [_H, C.c_int64, _I64, _D, _D, _D, _D, _D, _D, _I32]),
    "bh_get_tree_root": (C.c_int, [_H, _D, _D, _D, _I64]),
    "bh_build_tree": (C.c_int, [_H]),
    "bh_get_counters": (C.c_int, [_H, C.POINTER(BhCounters)]),
    "bh_reset_counters": (C.c_int, [_H]),
    "bh_get_let_stats": (C.c_int, [_H, _I64, C.c_int32]),
    "bh_get_body_counts": (C.c_int, [_H, _I32, _I32]),
    "bh_comm_unique_id": (C.c_int, [C.c_void_p, C.c_int32]),
    "bh_comm_init": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "bh_comm_init_external": (C.c_int, [_H, C.c_int32, C.c_int32]),
    "bh_slice_bounds": (C.c_int, [C.c_int64, C.c_int32, C.c_int32, _I64, _I64]),
    "bh_step_begin": (C.c_int, [_H]),
    "bh_step_end": (C.c_int, [_H]),
    "bh_step_finish": (C.c_int, [_H]),
    "bh_export_slice": (C.c_int, [_H, C.c_int32, C.c_int64, _D, _D, _I64, _I64]),
    "bh_import_slices": (C.c_int, [_H, C.c_int32, C.c_int64, _D, _D]),
    "bh_get_slice_index": (C.c_int, [_H, C.c_int64, _I32, _I64]),
    "bh_slice_epoch": (C.c_int64, [_H]),
    "bh_step_io_slice": (C.c_int, [_H, C.c_int32, C.c_int64, _D, _D, _D, _D, _D, C.c_int64, _D, _D, _D, _D, _D, _I64]),
    "bh_evaluate_slice": (C.c_int, [_H, C.c_int64, _D, _D, _I32, _I64]),
    "bh_set_domain_mode": (C.c_int, [_H, C.c_int32]),
    "bh_measure_fp32_tflops": (C.c_int, [C.c_int32, _D]),
}


class BhError(RuntimeError):
    def __init__(self, code, where, text=""):
        self.code = code
        super().__init__(f"{where}: {_ERR_NAMES.get(code, code)}{(' — ' + text) if text else ''}")


class CudaLibraryMissing(ImportError):
    """The sm_100a CUDA library is not built; there is deliberately no CPU fallback."""


def bind(path: str) -> C.CDLL:
    """dlopen `path` and attach the signatures of every ABI symbol (raises if one is missing)."""
    # RTLD_LOCAL: the product library and the oracle export the SAME symbol names (one ABI); global
    # binding would let one library's internal calls land in the other (both are also linked with
    # -Bsymbolic-functions for that reason)
    lib = C.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0) | getattr(os, "RTLD_NOW", 2))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.bh_abi_version() != ABI_VERSION:
        raise BhError(BH_E_ARG, "bind", f"ABI version {lib.bh_abi_version()} != {ABI_VERSION}")
    lib._bh_path = path
    return lib


# BH_B200_LIB: development override (A/B-testing a differently tuned build of the same library)
CUDA_LIB_PATH = os.environ.get("BH_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libbh_b200.so")
_cuda_lib = None


def load_cuda_library() -> C.CDLL:
    """The product library.  Fails loudly when the CUDA extension has not been built."""
    global _cuda_lib
    if _cuda_lib is None:
        if not os.path.exists(CUDA_LIB_PATH):
            raise CudaLibraryMissing(
                f"{CUDA_LIB_PATH} not found: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        lib = bind(CUDA_LIB_PATH)
        name = lib.bh_backend_name().decode()
        if name != "b200-cuda":
            raise CudaLibraryMissing(f"{CUDA_LIB_PATH} reports backend '{name}', expected 'b200-cuda'")
        _cuda_lib = lib
    return _cuda_lib
