"""B200-native Barnes–Hut physics step (drop-in for qwertukg/Barnes-Hut-N-Body's
PhysicsEngine path).  The compute lives in csrc/ (hand-written sm_100a CUDA behind
the C ABI of include/bh_engine.h); this package is the host-side mirror of the
reference's Kotlin API plus scene generators.  No CPU fallback exists."""
from ._abi import (BH_FLAG_BODY_COUNTS, BH_FLAG_LET, BH_FLAG_REUSE_ACC, BhConfig, BhCounters, BhDiskParams, BhError, BhParams, CudaLibraryMissing,
                   CUDA_LIB_PATH, bind, load_cuda_library)
from .engine import Acc, BHTree, Body, Config, NativeEngine, PhysicsEngine, Quad
from . import scenes
from . import distributed
from ._abi import BH_FIELD_POS, BH_FIELD_VEL

__all__ = ["BH_FLAG_BODY_COUNTS", "BH_FLAG_LET", "BH_FLAG_REUSE_ACC", "BhConfig", "BhCounters", "BhDiskParams", "BhError", "BhParams", "CudaLibraryMissing",
           "CUDA_LIB_PATH", "bind", "load_cuda_library", "Acc", "BHTree", "Body", "Config", "NativeEngine",
           "PhysicsEngine", "Quad", "scenes", "distributed", "BH_FIELD_POS", "BH_FIELD_VEL"]
