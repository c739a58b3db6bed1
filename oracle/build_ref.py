"""TEST INFRASTRUCTURE — builds the *reference's own* CUDA extension as a GPU-side checker.

Compiles /root/reference/nerf/gridencoder/src/{gridencoder.cu,bindings.cpp} where they lie
(no sources are copied) for sm_100a into oracle/_ref/_gridencoder_ref.so.  The only change to
the reference recipe (gridencoder/setup.py:L7-10) is -std=c++14 -> -std=c++17, which torch 2.11
headers require.  oracle/_ref/ is git-ignored but travels to the GPU box with gpurun, where
tests/test_gpu_grid_vs_ref.py uses it to pin our kernel (and the oracle restatement) against the
reference kernel itself.  It is never imported by the product path.

Usage:  python oracle/build_ref.py      (≈5 min on 8 cores; no-op if /root/reference is absent)
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/nerf/gridencoder/src"
OUT = os.path.join(HERE, "_ref")


def build(verbose=True):
    if not os.path.isdir(REF_SRC):
        print("[oracle/build_ref] /root/reference not present - using prebuilt oracle/_ref if any")
        return None
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "_gridencoder_ref.so")
    srcs = [os.path.join(REF_SRC, f) for f in ("gridencoder.cu", "bindings.cpp")]
    if os.path.exists(so) and all(os.path.getmtime(so) > os.path.getmtime(s) for s in srcs):
        return so
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load
    nvcc_flags = ["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__",
                  "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__"]
    load(name="_gridencoder_ref", sources=srcs, extra_cflags=["-O3", "-std=c++17"],
         extra_cuda_cflags=nvcc_flags, build_directory=OUT, verbose=verbose, is_python_module=False)
    return so


if __name__ == "__main__":
    print(build())
