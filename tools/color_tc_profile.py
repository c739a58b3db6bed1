#!/usr/bin/env python
"""Barrier-wait breakdown of color_mlp_tc_kernel (debug_flags bit 2, CTA 0): kilo-cycles per role over one 131,072-ray chunk.
    python tools/color_tc_profile.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ucnerf_b200 import _lib, synthetic  # noqa: E402

dev = torch.device("cuda:0")
wl = synthetic.WORKLOADS["eval_800x600_waymo_gin"]
sd = {k: v.to(dev) for k, v in synthetic.synthetic_state_dict(wl, seed=0).items()}
r = synthetic.make_renderer(wl, sd, dev)
rays = {k: v.to(dev)[:131072].contiguous() for k, v in synthetic.pinhole_rays(wl.height, wl.width, seed=0).items()}
r.render_rays(rays, 1.0, rays["rand_vec"], ("packed",))
r.set_option("tc_debug", 4)
r.render_rays(rays, 1.0, rays["rand_vec"], ("packed",))
torch.cuda.synchronize()
lib = _lib.load()
out = (C.c_uint32 * 32)()
lib.ucnerf_debug_tc_status.argtypes = [C.POINTER(C.c_uint32)]
lib.ucnerf_debug_tc_status(out)
w = list(out)
print("mma thread: total %d kcyc, waits acc3_empty %d acc4_empty %d b_full %d a_full %d, tiles %d" % tuple(w[8:14]))
for g in (0, 1):
    d = w[16 + 8 * g: 24 + 8 * g]
    print("group %d: total %d kcyc, waits a_empty %d acc3_full %d, epilogue %d, h1 chunk %d, tmem chunks %d" % (g, d[0], d[1], d[2], d[3], d[4], d[6]))
