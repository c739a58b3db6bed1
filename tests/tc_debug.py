"""Debug driver for the tensor-core colour MLP (run on the GPU box): prints the watchdog record and errors."""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import cases
from test_gpu_render import build_renderer, run
from ucnerf_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg, params, batch = cases.make_case("waymo", n)
r = build_renderer(cfg, params)
lib = _lib.load()
r.set_option("color_mlp", 0)
simt = run(r, batch)
print("simt done", flush=True)
r.set_option("color_mlp", 1)
tc = run(r, batch)
st = (ctypes.c_uint32 * 32)()
lib.ucnerf_debug_tc_status(st)
print("watchdog:", list(st)[:8], flush=True)
d = np.abs(tc["sample_rgb"] - simt["sample_rgb"])
print("tc vs simt sample_rgb: max", d.max(), "mean", d.mean(), "nan", np.isnan(tc["sample_rgb"]).sum())
print("rows with err>1e-4:", (d.max(-1) > 1e-4).sum(), "of", d.shape[0] * d.shape[1])
print(tc["sample_rgb"][0, :4], simt["sample_rgb"][0, :4])
