"""Profiling driver (GPU box): where does the MMA-issuing thread of the tensor-core colour MLP wait?"""
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
    print(f"flags={flags}: tiles/CTA={it} total={tot}k cycles; waits: acc3_empty={a3}k acc4_empty={a4}k b_full={b}k a_full={a}k "
          f"-> issue+other={tot - a3 - a4 - b - a}k; per tile total={tot / max(it, 1):.1f}k")
    for g in (0, 1):
        d = list(st)[16 + 8 * g: 16 + 8 * g + 8]
        print(f"      store_a_row={d[7] & 0xffff}k fence.proxy.async={d[7] >> 16}k (all 6 chunks per tile: store {(d[7] & 0xffff)/max(it,1):.2f}k fence {(d[7] >> 16)/max(it,1):.2f}k)")
        print(f"   producer group {g}: total={d[0]}k wait a_empty={d[1]}k acc3_full={d[2]}k epilogue(incl. waits)={d[3]}k "
              f"work h1={d[4]}k tmem_ld+bias_wait={d[5]}k tmem_total={d[6]}k  (per tile: h1 {d[4]/max(it,1):.2f}k tmem_ld {d[5]/max(it,1):.2f}k tmem_total {d[6]/max(it,1):.2f}k epi {d[3]/max(it,1):.2f}k)")
