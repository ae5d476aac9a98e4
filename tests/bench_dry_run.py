"""Helper of tests/test_bench_host.py: runs bench.py's main() at world 1 WITHOUT a GPU by binding the engine class to the
oracle library (same C ABI) and stubbing the three torch.cuda calls main() makes — host logic and JSON contract only; the
numbers it prints mean nothing.  (Test infrastructure: the product never routes through the oracle.)"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bh_b200  # noqa: E402
from conftest import _build_oracle  # noqa: E402

olib = bh_b200.bind(_build_oracle())
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self, *a, **k: self


class _Lib:                      # the oracle exports everything but the FP32 probe
    def __getattr__(self, k):
        if k == "bh_measure_fp32_tflops":
            def probe(dev, p):
                p[0] = 72.5
                return 0
            return probe
        return getattr(olib, k)


_lib = _Lib()
_init = bh_b200.NativeEngine.__init__
bh_b200.NativeEngine.__init__ = lambda self, lib=None, **kw: _init(self, lib=lib if lib is not None else _lib, **kw)

spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)
sys.argv = ["bench.py", "--bodies", "20000", "--steps", "2", "--warmup", "1", "--configs", ""]
bench.main()
