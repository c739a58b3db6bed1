#!/usr/bin/env python
"""A/B experiment (B200): how the order of the rays inside a batch (row-major vs. WxH pixel tiles) and the warp shape of
sample_encode_kernel (2^k rays x 2^(5-k) samples, option warp_rays_*) change the gather kernels' time.  One JSON line
per combination with the per-family CUDA-event times; every variant is checked bit-for-bit against the default."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from ucnerf_b200 import synthetic  # noqa: E402


def tile_perm(H, W, tw, th):
    return torch.arange(H * W).view(H // th, th, W // tw, tw).permute(0, 2, 1, 3).reshape(-1)


def main():
    wl = synthetic.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "eval_800x600_waymo_gin"]
    steps = 3
    dev = torch.device("cuda:0")
    sd = synthetic.synthetic_state_dict(wl, seed=0)
    r = synthetic.make_renderer(wl, sd, dev)
    rays_h = synthetic.pinhole_rays(wl.height, wl.width, seed=0)
    rays0 = {k: v.to(dev) for k, v in rays_h.items()}
    base = None
    combos = []
    for tile in ((0, 0), (8, 4), (4, 8), (16, 2), (4, 4), (4, 2), (2, 2)):
        for rw in (32, 16, 8, 4):
            combos.append((tile, rw))
    for (tw, th), rw in combos:
        if tw and (wl.width % tw or wl.height % th):
            continue
        perm = tile_perm(wl.height, wl.width, tw, th).to(dev) if tw else None
        rays = {k: (v[perm].contiguous() if perm is not None else v) for k, v in rays0.items()}
        r.set_option("warp_rays_prop", rw)
        r.set_option("warp_rays_nerf", rw)
        for _ in range(2):
            out = r.render_rays(rays, 1.0, rays["rand_vec"], ("packed",))
        torch.cuda.synchronize()
        packed = out["packed"].clone()
        if perm is not None:
            un = torch.empty_like(packed)
            un[perm] = packed
            packed = un
        if base is None:
            base = packed
        same = bool(torch.equal(packed, base))
        r.set_option("timing", 1)
        r.timing(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r.render_rays(rays, 1.0, rays["rand_vec"], ("packed",))
        e1.record()
        torch.cuda.synchronize()
        fam = r.timing(reset=True)
        r.set_option("timing", 0)
        print(json.dumps({"tile": f"{tw}x{th}" if tw else "row", "warp_rays": rw, "bit_identical": same,
                          "ms_per_frame": round(e0.elapsed_time(e1) / steps, 3),
                          **{k: round(v[0] / steps, 3) for k, v in fam.items()}}), flush=True)


if __name__ == "__main__":
    main()
