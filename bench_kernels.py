#!/usr/bin/env python
"""Secondary benchmark (not the driver contract): the `_gridencoder` drop-in kernels against the reference's own
CUDA kernels (oracle/_ref/_gridencoder_ref.so, compiled for sm_100a from /root/reference where it lies) on the
per-chunk loads of waymo.gin.  Prints one JSON line per case; used for profiles/r1_grid_kernels.json.

    python bench_kernels.py [--reps 10]"""
import argparse
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from ucnerf_b200.gridencoder import backend as mine
from ucnerf_b200.synthetic import _layout


def load_ref():
    so = os.path.join(ROOT, "oracle", "_ref", "_gridencoder_ref.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("_gridencoder_ref", so)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    ref = load_ref()
    peak = 6539.9
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    # (name, points, levels, desired): one 15,000-ray chunk of waymo.gin
    cases = [("proposal_chunk", 15000 * 128 * 6, 6, 512), ("nerf_chunk", 15000 * 32 * 6, 10, 8192)]
    for name, B, L, desired in cases:
        offsets, _ = _layout(L, desired, 21)
        off = torch.from_numpy(offsets).cuda()
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.rand((B, 3), device="cuda", generator=g)
        emb = (torch.rand((int(offsets[-1]), 4), device="cuda", generator=g) - 0.5)
        out = torch.empty(L, B, 4, device="cuda")
        grad = torch.randn(L, B, 4, device="cuda", generator=g)
        gemb = torch.zeros_like(emb)
        res = {"case": name, "points": B, "levels": L, "gather_bytes": B * L * 8 * 16}
        for tag, be in (("ours", mine), ("reference", ref)):
            if be is None:
                continue
            f = lambda: be.grid_encode_forward(x, emb, off, out, B, 3, 4, L, 1.0, 16, None, 0, False, 0)
            ms = timeit(f, a.reps)
            res[f"fwd_ms_{tag}"] = ms
            res[f"fwd_gather_GBs_{tag}"] = res["gather_bytes"] / ms / 1e6
            b = lambda: be.grid_encode_backward(grad, x, emb, off, gemb, B, 3, 4, L, 1.0, 16, None, None, 0, False, 0)
            ms = timeit(b, a.reps)
            res[f"bwd_ms_{tag}"] = ms
        if ref is not None:
            res["fwd_speedup"] = res["fwd_ms_reference"] / res["fwd_ms_ours"]
            res["bwd_speedup"] = res["bwd_ms_reference"] / res["bwd_ms_ours"]
        res["fwd_frac_of_hbm_peak_ours"] = res["fwd_gather_GBs_ours"] / peak
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
