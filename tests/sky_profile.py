"""Profiling driver (GPU box): where do the MMA-issuing thread and the producers of the sky tensor-core kernel wait?
Run with UCNERF_SKY_DEBUG=4."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import cases
from ucnerf_b200 import synthetic
from ucnerf_b200.render import SkyHead
head = SkyHead(cases.make_heads(seed=3))
rays = synthetic.pinhole_rays(600, 800, seed=0)
o, d, far, cam = (rays[k][:240000].cuda() for k in ("origins", "directions", "far", "cam_dirs"))
for _ in range(2):
    head.render(o, d, far, cam)
torch.cuda.synchronize()
st = (ctypes.c_uint32 * 32)()
head.lib.ucnerf_debug_sky_status(st)
it = max(st[12], 1)
print(f"MMA thread: tiles/CTA={st[12]} total={st[8]}k cycles ({st[8] / it:.1f}k per tile); waits: epi_done={st[9]}k b_full={st[10]}k "
      f"a_full={st[11]}k ({st[11] / it:.1f}k per tile) -> issue+other={(st[8] - st[9] - st[10] - st[11]) / it:.1f}k per tile")
it = max(st[20], 1)
print(f"producer (group 0, row 0): total={st[16]}k; per tile: wait acc_full={st[17] / it:.1f}k wait a_empty={st[18] / it:.1f}k "
      f"epilogue={st[19] / it:.1f}k work={(st[16] - st[17] - st[18] - st[19]) / it:.1f}k")
