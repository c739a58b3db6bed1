"""Profiling driver (GPU box): where do the MMA-issuing thread and the producer groups of the tensor-core colour MLP
spend their cycles?  flags: 4 = in-kernel wait profiler, 8 = disable the L2 prefetch of the next tile's h1 rows."""
import ctypes, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ucnerf_b200 import _lib, synthetic
wl = synthetic.WORKLOADS["eval_800x600_waymo_gin"]
sd = synthetic.synthetic_state_dict(wl, seed=0)
r = synthetic.make_renderer(wl, sd)
rays = {k: v.cuda() for k, v in synthetic.pinhole_rays(wl.height, wl.width, seed=0).items()}
lib = _lib.load()
for flags in (4,):
    r.set_option("tc_debug", flags)
    for _ in range(2):
        r.render_rays(rays, 1.0, rays["rand_vec"], ("packed",))
    st = (ctypes.c_uint32 * 32)()
    lib.ucnerf_debug_tc_status(st)
    tot, a3, a4, b, a, it = st[8], st[9], st[10], st[11], st[12], st[13]
    print(f"flags={flags}: tiles/CTA={it} total={tot}k cycles; MMA thread waits: acc3_empty={a3}k acc4_empty={a4}k b_full={b}k a_full={a}k "
          f"-> issue+other={tot - a3 - a4 - b - a}k; per tile total={tot / max(it, 1):.1f}k")
    for g in (0, 1):
        d = list(st)[16 + 8 * g: 16 + 8 * g + 8]
        n = max(it, 1)
        print(f"   producer group {g}: total={d[0]}k wait a_empty={d[1]}k acc3_full={d[2]}k epilogue(incl. waits)={d[3]}k "
              f"work h1={d[4]}k tmem={d[6]}k  (per tile: a_empty {d[1]/n:.2f}k h1 {d[4]/n:.2f}k tmem {d[6]/n:.2f}k epi {d[3]/n:.2f}k)")
