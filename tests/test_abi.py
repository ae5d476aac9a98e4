"""The drop-in boundary: both libraries export every symbol include/bh_engine.h declares,
the host mirror keeps the reference's Kotlin names, and the product refuses to run
without CUDA (no CPU fallback).  CPU only — no compute calls on the CUDA library."""
import ctypes as C
import os
import re

import pytest

import bh_b200
from bh_b200 import _abi
from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "bh_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bh_[a-z0-9_]+)\s*\(", src)))


def test_binding_declares_every_header_symbol():
    assert header_symbols() == sorted(_abi.SYMBOLS)


def test_oracle_exports_abi(oracle_lib):
    for name in header_symbols():
        assert hasattr(oracle_lib, name), name
    assert oracle_lib.bh_backend_name() == b"reference-port"
    assert oracle_lib.bh_abi_version() == _abi.ABI_VERSION


def test_cuda_library_loads_and_exports_abi():
    """libbh_b200.so must exist in-tree (built by __graft_entry__.build) and export the ABI."""
    assert os.path.exists(bh_b200.CUDA_LIB_PATH), "run `python __graft_entry__.py build`"
    lib = bh_b200.load_cuda_library()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.bh_backend_name() == b"b200-cuda"
    p = _abi.BhParams()
    assert lib.bh_default_params(2400, 800, C.byref(p)) == 0 and p.root_half == 1202.0
    lo, hi = C.c_int64(), C.c_int64()
    assert lib.bh_slice_bounds(10, 4, 3, C.byref(lo), C.byref(hi)) == 0 and (lo.value, hi.value) == (9, 10)


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device bh_create must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bh_b200.BhError) as ei:
        bh_b200.NativeEngine()
    assert ei.value.code == _abi.BH_E_CUDA and "no CPU fallback" in str(ei.value)


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "barnes-hut-n-body_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "libbh_ref" not in txt and "oracle/" not in txt.replace("accuracy oracle", ""), f


def test_kotlin_api_mirror_names():
    for cls, names in ((bh_b200.PhysicsEngine, ["step", "getBodies", "resetBodies", "getTreeForDebug"]),
                       (bh_b200.BHTree, ["visitQuads", "mass", "comX", "comY"]),
                       (bh_b200.Quad, ["contains", "child"])):
        for nme in names:
            assert hasattr(cls, nme), (cls, nme)
    c = bh_b200.Config
    assert (c.WIDTH_PX, c.HEIGHT_PX, c.G, c.DT, c.SOFT2, c.theta, c.MIN_R) == (2400, 800, 80.0, 0.005, 1.0, 0.30, 8.0)
    q = bh_b200.Quad(10.0, 20.0, 4.0)
    assert q.child(0) == bh_b200.Quad(8.0, 18.0, 2.0) and q.child(3) == bh_b200.Quad(12.0, 22.0, 2.0)
    assert q.contains(bh_b200.Body(6.0, 16.0, 0, 0, 1)) and not q.contains(bh_b200.Body(14.0, 16.0, 0, 0, 1))
