"""Build libucnerf_b200.so (sm_100a only) in-tree with nvcc.  No torch involved: the library is a
plain C-ABI CUDA shared object (include/ucnerf_b200.h).  `python -m ucnerf_b200.build` or
__graft_entry__.build() call this; the .so is git-ignored but travels to the GPU box via gpurun."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libucnerf_b200.so")
OBJ = os.path.join(HERE, "csrc", "build")
SOURCES = ["grid_encode.cu", "grid_adam.cu", "pooled_encode.cu", "resample_op.cu", "composite_train.cu", "cast_rays_op.cu", "ray_gen.cu", "peer.cu", "gemm3_tc.cu", "ray_march.cu", "sample_encode.cu", "color_mlp_tc.cu", "sky_mlp_tc.cu", "sky_mlp_tc2.cu", "sky_model.cu", "model.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-diag-suppress", "128"]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(verbose=True, force=False, extra_flags=(), out=None, obj_dir=None):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    OUT_ = out or OUT
    OBJ_ = obj_dir or OBJ
    os.makedirs(OBJ_, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "ucnerf_b200.h"))
    jobs = []
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ_, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or not _newer(obj, [src] + headers):
            jobs.append([nvcc, *NVCC_FLAGS, *extra_flags, "-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        if verbose and (r.stdout or r.stderr):
            print(r.stdout + r.stderr)

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    if jobs or force or not os.path.exists(OUT_):
        run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT_, *objs])
    return OUT_


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
