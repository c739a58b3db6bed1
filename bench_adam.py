#!/usr/bin/env python
"""Secondary benchmark (not the driver contract): the fused hash-decay + Adam step (ucnerf_grid_adam_step, SURVEY.md
section 8f N2) on the two waymo.gin tables against the reference's recipe on the same GPU (hash-decay loss term +
autograd + grad.nan_to_num_() + torch.optim.Adam + zero_grad).  One JSON line with the HBM roofline of the kernel.

    python bench_adam.py [--reps 20]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch

from ucnerf_b200 import synthetic
from ucnerf_b200.gridencoder.optim import GridAdam


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    wl = synthetic.WORKLOADS["eval_800x600_waymo_gin"]
    sd = synthetic.synthetic_state_dict(wl, seed=0)

    class Enc(torch.nn.Module):
        def __init__(self, prefix):
            super().__init__()
            self.embeddings = torch.nn.Parameter(sd[prefix + ".embeddings"].clone().cuda())
            self.register_buffer("offsets", sd[prefix + ".offsets"].clone())
            off = self.offsets.long()
            self.register_buffer("idx", torch.repeat_interleave(torch.arange(off.numel() - 1), off[1:] - off[:-1]).cuda())

    encs = [Enc("prop_mlp_0.encoder"), Enc("nerf_mlp.encoder")]
    floats = sum(e.embeddings.numel() for e in encs)
    grads = [torch.randn_like(e.embeddings) * 1e-3 for e in encs]

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

    opt = GridAdam(encs, lr=0.01, betas=(0.9, 0.99), eps=1e-15, hash_decay_mult=0.1, zero_grad=True)

    def ours():
        for e, g in zip(encs, grads):
            e.embeddings.grad = g          # (zeroed by the step; re-pointing is free)
        opt.step()

    ms = timeit(ours, a.reps)
    ref_params = [torch.nn.Parameter(e.embeddings.detach().clone()) for e in encs]
    topt = torch.optim.Adam(ref_params, lr=0.01, betas=(0.9, 0.99), eps=1e-15)

    def reference_recipe():
        topt.zero_grad()
        loss = 0.
        for p, e in zip(ref_params, encs):        # models.py:L297-306 x hash_decay_mults
            n = int(e.offsets.numel() - 1)
            sums = torch.zeros(n, 4, device="cuda").index_add_(0, e.idx, p ** 2)
            cnt = (e.offsets[1:] - e.offsets[:-1]).float().cuda()
            loss = loss + 0.1 * (sums / cnt[:, None]).mean()
        loss.backward()
        for p, g in zip(ref_params, grads):
            p.grad += g
            p.grad.nan_to_num_()
        topt.step()

    ms_ref = timeit(reference_recipe, max(3, a.reps // 4))
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    gbs = floats * 32 / ms / 1e6
    print(json.dumps({"bench": "grid_adam_step", "tables": "waymo.gin proposal (6.6 M entries) + NeRF (15.0 M entries), C = 4",
                      "floats": floats, "ms": ms, "algorithmic_bytes": floats * 32,
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                   "bytes_per_float": 32},
                      "reference_recipe_ms": ms_ref, "speedup": ms_ref / ms,
                      "note": "reference recipe = hash-decay loss term + autograd + nan_to_num_ + torch.optim.Adam (foreach) "
                              "+ zero_grad on the same B200"}))


if __name__ == "__main__":
    main()
