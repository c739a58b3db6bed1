"""Small end-to-end run for `compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py` on the GPU box: the render
path on two configurations, the sky head and the ray generator (profiles/r1_sanitizer_memcheck.txt: 0 errors)."""
import sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import cases, ucnerf_oracle as O
from test_gpu_render import build_renderer, run
from ucnerf_b200.render import SkyHead, generate_rays
from ucnerf_b200 import synthetic
for name, n in (("config1", 96), ("waymo", 40)):
    cfg, params, batch = cases.make_case(name, n)
    r = build_renderer(cfg, params)
    out = run(r, batch)
    print(name, float(out["rgb"].mean()))
heads = cases.make_heads(seed=3)
b = O.synthetic_rays(40, seed=5)
sky = SkyHead(heads).render(b["origins"].cuda(), b["directions"].cuda(), b["far"].cuda(), b["cam_dirs"].cuda())
print("sky", float(sky.mean()))
g = generate_rays(*synthetic.pinhole_camera(8, 12, seed=1))
print("rays", float(g["directions"].mean()))
torch.cuda.synchronize()
