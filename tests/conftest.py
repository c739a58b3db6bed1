import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def lib():
    """The product shared library (built in-tree by ucnerf_b200.build)."""
    from ucnerf_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build(verbose=False)
    return _lib.load()


@pytest.fixture(scope="session")
def harness():
    """CPU instantiation of the device algorithm templates (tests/cpu_harness.cpp), test-only."""
    import ctypes
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "cpu_harness.so")
    src = os.path.join(ROOT, "tests", "cpu_harness.cpp")
    hdrs = [os.path.join(ROOT, "ucnerf_b200", "csrc", f) for f in ("ray_algos.cuh", "common.cuh", "pooled_algos.cuh", "train_algos.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I/usr/local/cuda/include",
                               src, "-o", so])
    return ctypes.CDLL(so)


def load_golden(name):
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=False))
