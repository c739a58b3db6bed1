"""TEST INFRASTRUCTURE - golden vectors for the stand-alone resampling op (training and eval) from the REFERENCE's own
stepfun.py running on CPU (oracle/ref_shim.py): for the waymo.gin NeRF level (128 proposal bins -> 32 intervals) and a
three-level case, the reference's `max_dilate_weights` + slice + logits + `sample_intervals` (models.py:L156-205) with
rand=False, rand=True / single_jitter=True and rand=True / single_jitter=False; `torch.rand` is patched to return the
stored draw.

    python oracle/make_resample_golden.py        # writes tests/golden/resample_op.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases, ref_shim, ucnerf_oracle as O  # noqa: E402


def reference_level(R, sdist, weights, S, dilation, anneal, padding, rand01, single_jitter):
    """models.py:L165-199 verbatim in structure, calling the reference's stepfun."""
    sdist, weights = R.stepfun.max_dilate_weights(sdist, weights, dilation, domain=(0., 1.), renormalize=True)
    sdist, weights = sdist[..., 1:-1], weights[..., 1:-1]
    logits = torch.where(sdist[..., 1:] > sdist[..., :-1], anneal * torch.log(weights + padding),
                         torch.full_like(sdist[..., :-1], -torch.inf))
    orig = torch.rand
    if rand01 is not None:
        torch.rand = lambda *a, **k: rand01.clone()
    try:
        return R.stepfun.sample_intervals(rand01 is not None, sdist, logits, S, single_jitter=single_jitter, domain=(0., 1.))
    finally:
        torch.rand = orig


def main():
    R = ref_shim.load_reference()
    out = {}
    g = torch.Generator().manual_seed(21)
    for name, n_rays in (("waymo", 24), ("three_level", 12)):
        cfg, params, batch = cases.make_case(name, n_rays)
        _, hist = O.model_forward(params, cfg, batch)
        lvl = cfg.num_levels - 1
        t_prev, w_prev = hist[lvl - 1]["sdist"], hist[lvl - 1]["weights"]
        S = hist[lvl]["weights"].shape[1]
        prod = 1
        for l in range(lvl):
            prod *= hist[l]["weights"].shape[1]
        dilation = float(np.float32(cfg.dilation_bias + cfg.dilation_multiplier / prod))
        out[f"{name}_t_prev"], out[f"{name}_w_prev"] = t_prev.numpy(), w_prev.numpy()
        out[f"{name}_S"], out[f"{name}_dilation"] = np.int64(S), np.float32(dilation)
        r1 = torch.rand((n_rays, 1), generator=g)
        rS = torch.rand((n_rays, S), generator=g)
        out[f"{name}_rand_single"], out[f"{name}_rand_each"] = r1.numpy(), rS.numpy()
        for tag, r01, single, anneal, padding in (("det", None, True, 1.0, 0.0), ("single", r1, True, 0.7, 0.0),
                                                  ("each", rS, False, 1.0, 0.01)):
            sd = reference_level(R, t_prev, w_prev, S, dilation, anneal, padding, r01, single)
            out[f"{name}_sdist_{tag}"] = sd.numpy()
            out[f"{name}_anneal_{tag}"], out[f"{name}_padding_{tag}"] = np.float32(anneal), np.float32(padding)
            mine = O.resample_level(t_prev, w_prev, S, dilation, True, anneal, padding, r01)
            print(name, tag, "oracle == reference:", bool(torch.equal(mine, sd)), float((mine - sd).abs().max()))
    path = os.path.join(ROOT, "tests", "golden", "resample_op.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
