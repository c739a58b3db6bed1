"""TEST / BASELINE INFRASTRUCTURE - builds oracle/grid_cpu.c (the C + OpenMP restatement of the reference's CUDA-only
hash-grid forward) into oracle/grid_cpu.so and loads it with ctypes.  The .so is git-ignored and travels to the GPU box
with gpurun like the product library; the product never loads it."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "grid_cpu.c")
SO = os.path.join(HERE, "grid_cpu.so")
_LIB = None


def build(verbose=False):
    if os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(SRC):
        return SO
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", SRC, "-o", SO, "-lm"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return SO


def load():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        f = lib.ucnerf_oracle_grid_forward_f32
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_uint32] * 4 + [ctypes.c_float, ctypes.c_uint32] + [ctypes.c_int] * 3
        _LIB = lib
    return _LIB


def grid_encode_forward(inputs, embeddings, offsets, B, D, C, L, S, H, gridtype=0, align_corners=False, interp=0):
    """numpy in / numpy out ([L,B,C] fp32), same contract as ucnerf_oracle.grid_encode_forward without dy_dx."""
    import numpy as np
    x = np.ascontiguousarray(inputs, dtype=np.float32)
    e = np.ascontiguousarray(embeddings, dtype=np.float32)
    o = np.ascontiguousarray(offsets, dtype=np.int32)
    out = np.empty((L, B, C), dtype=np.float32)
    if C > 64:
        raise ValueError("grid_cpu.c: C <= 64")
    rc = load().ucnerf_oracle_grid_forward_f32(x.ctypes.data, e.ctypes.data, o.ctypes.data, out.ctypes.data, B, D, C, L,
                                               float(S), int(H), int(gridtype), int(bool(align_corners)), int(interp))
    if rc != 0:
        raise ValueError("grid_cpu.c: unsupported input dimension")
    return out


if __name__ == "__main__":
    print(build(verbose=True))
