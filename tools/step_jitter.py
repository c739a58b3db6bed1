#!/usr/bin/env python
"""Per-step times of the bench frame (device-resident rays), to see whether single steps are outliers.
    python tools/step_jitter.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ucnerf_b200 import synthetic  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda:0")
wl = synthetic.WORKLOADS["eval_800x600_waymo_gin"]
sd = {k: v.to(dev) for k, v in synthetic.synthetic_state_dict(wl, seed=0).items()}
r = synthetic.make_renderer(wl, sd, dev)
rays = {k: v.to(dev) for k, v in synthetic.pinhole_rays(wl.height, wl.width, seed=0).items()}
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
torch.cuda.synchronize()
ev[0].record()
for i in range(steps):
    r.render_rays(rays, 1.0, rays["rand_vec"], ("packed",))
    ev[i + 1].record()
torch.cuda.synchronize()
print(" ".join(f"{ev[i].elapsed_time(ev[i + 1]):.1f}" for i in range(steps)))
