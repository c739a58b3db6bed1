#!/usr/bin/env python
"""Secondary benchmark (not the driver contract): the pooled hash-grid encode for training
(ucnerf_pooled_encode_forward/backward, SURVEY.md section 8a rows R4 + R10) against the reference-shaped chain on the
same GPU - contract_mean_std as torch ops -> GridEncoder (this repo's drop-in kernels behind the reference's autograd
Function) -> erf weights -> multiply -> mean, and its autograd - at one GPU's share of a 65,536-ray train batch
(config 5: 8,192 rays x 128 proposal intervals, 8,192 x 32 NeRF intervals, 6 points each).  One JSON line.

    python bench_pooled.py [--reps 10]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch

from ucnerf_b200 import synthetic
from ucnerf_b200.gridencoder import GridEncoder
from ucnerf_b200.gridencoder.pooled import pooled_encode


def contract_mean_std(x, std):     # coord.py:L60-72 as the reference evaluates it (separate ATen kernels)
    eps = torch.finfo(x.dtype).eps
    m2 = torch.sum(x ** 2, dim=-1, keepdim=True).clamp_min(eps)
    ms = torch.sqrt(m2)
    mask = m2 <= 1
    z = torch.where(mask, x, ((2 * torch.sqrt(m2) - 1) / m2) * x)
    det = (torch.pow(2 * ms - 1, 1 / 3) / ms) ** 2
    return z, torch.where(mask[..., 0], std, det[..., 0] * std)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--rays", type=int, default=8192)
    a = ap.parse_args()
    wl = synthetic.WORKLOADS["eval_800x600_waymo_gin"]
    sd = synthetic.synthetic_state_dict(wl, seed=0)
    peaks_p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_p)).get("hbm_gbs", 6650.0)) if os.path.exists(peaks_p) else 6650.0

    def timeit(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {"bench": "pooled_encode (training front end of MLP.predict_density)", "rays": a.rays, "levels": {}}
    for tag, prefix, desired, S in (("prop", "prop_mlp_0.encoder", wl.prop_desired[0], wl.num_prop_samples),
                                    ("nerf", "nerf_mlp.encoder", wl.nerf_desired, wl.num_nerf_samples)):
        L = wl.grid_levels(desired)
        enc = GridEncoder(3, L, 4, base_resolution=16, desired_resolution=desired, log2_hashmap_size=wl.log2_hashmap_size).cuda()
        with torch.no_grad():
            enc.embeddings.copy_(sd[prefix + ".embeddings"])
        B = a.rays * S
        g = torch.Generator(device="cuda").manual_seed(1)
        # multisample points of neighbouring intervals along rays through the contracted scene
        t = torch.rand((a.rays, S, 1, 1), generator=g, device="cuda") ** 2 * 7.5 + 0.02
        d = torch.nn.functional.normalize(torch.randn((a.rays, 1, 1, 3), generator=g, device="cuda"), dim=-1)
        means = (d * t + 1e-3 * t * torch.randn((a.rays, S, 6, 3), generator=g, device="cuda")).reshape(B, 6, 3).contiguous()
        stds = (5e-4 * t.expand(-1, -1, 6, 1)).reshape(B, 6).contiguous()
        gf = torch.randn((B, L * 4), generator=g, device="cuda")

        def fused():
            enc.embeddings.grad = None
            f, _ = pooled_encode(enc, means, stds)
            f.backward(gf)

        def fused_merge():                      # backward variant: runs of points in one cell reduced once
            enc.embeddings.grad = None
            f, _ = pooled_encode(enc, means, stds, merge_runs=True)
            f.backward(gf)

        def fused_merge_ray():                  # ... and across 4 consecutive intervals of a ray
            enc.embeddings.grad = None
            f, _ = pooled_encode(enc, means, stds, merge_runs='ray')
            f.backward(gf)

        def fused_fwd():
            with torch.no_grad():
                pooled_encode(enc, means, stds)

        def chain():
            enc.embeddings.grad = None
            with torch.no_grad():
                m, s = contract_mean_std(means.reshape(-1, 3), stds.reshape(-1))
                m, s = m.reshape(B, 6, 3) / 2, s.reshape(B, 6) / 2
            f = enc(m, bound=1).unflatten(-1, (L, -1))
            w = torch.erf(1 / torch.sqrt(8 * s[..., None] ** 2 * enc.grid_sizes ** 2))
            f = (f * w[..., None]).mean(dim=-3).flatten(-2, -1)
            f.backward(gf)

        ms_f = timeit(fused, a.reps)
        ms_ff = timeit(fused_fwd, a.reps)
        ms_fm = timeit(fused_merge, a.reps)
        ms_fr = timeit(fused_merge_ray, a.reps)
        ms_c = timeit(chain, max(3, a.reps // 2))
        alg = B * 6 * L * 8 * 16              # gathered (forward) or reduced (backward) table bytes per pass
        out["levels"][tag] = {"intervals": B, "grid_levels": L, "fused_fwd_bwd_ms": ms_f, "fused_fwd_ms": ms_ff,
                              "fused_merge_runs_fwd_bwd_ms": ms_fm, "fused_merge_ray_runs_fwd_bwd_ms": ms_fr,
                              "chain_fwd_bwd_ms": ms_c, "speedup": ms_c / ms_f,
                              "algorithmic_table_bytes_per_pass": alg,
                              "fwd_gather_gbs": alg / ms_ff / 1e6, "fwd_gather_frac_of_hbm_peak": alg / ms_ff / 1e6 / peak,
                              "bwd_reduce_gbs": alg / max(ms_f - ms_ff, 1e-6) / 1e6}
    out["hbm_peak_gbs"] = peak
    out["note"] = ("chain = contract (torch) -> GridEncoder drop-in kernels ([L,B*6,C] out + permute) -> erf weights -> "
                   "mean, and its autograd, on the same B200; algorithmic bytes = 6 points x L x 8 corners x 16 B per interval")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
