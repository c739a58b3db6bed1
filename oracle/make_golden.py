"""TEST INFRASTRUCTURE - generate tests/golden/<case>.npz by running the UNMODIFIED reference Python
(through oracle/ref_shim.py) on seeded synthetic rays / weights, after asserting that the oracle restatement
reproduces it bit-for-bit.  Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden waymo      # one case

What is stored: the ray batch, the reference outputs of every level (sdist, weights) and of the final level
(rgb, depth, acc, distance_*), `depth_raw` (the oracle's pre-threshold depth; equals reference depth wherever
acc >= 0.6) and fp64 checksums of the regenerated weights so a test can verify it rebuilt the same model."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import cases, ref_shim, ucnerf_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def run(name):
    cfg, params, batch = cases.make_case(name)
    rr, rh, _, _ = ref_shim.reference_forward(cfg, params, batch)
    orr, oh = O.model_forward(params, cfg, batch)
    out = {"batch_" + k: v.numpy() for k, v in batch.items()}
    for lvl in range(cfg.num_levels):
        for k in rr[lvl]:
            assert torch.equal(rr[lvl][k], orr[lvl][k]), (name, lvl, k)
        for k in ("sdist", "weights", "density", "rgb", "coord"):
            assert torch.equal(rh[lvl][k], oh[lvl][k]), (name, lvl, k)
        out[f"sdist_{lvl}"] = rh[lvl]["sdist"].numpy()
        out[f"weights_{lvl}"] = rh[lvl]["weights"].numpy()
    last = rr[-1]
    for k in ("rgb", "depth", "acc", "distance_mean", "distance_median", "distance_percentile_5",
              "distance_percentile_95"):
        out[k] = last[k].numpy()
    out["depth_raw"] = orr[-1]["depth_raw"].numpy()
    assert np.array_equal(out["depth_raw"][out["acc"] >= 0.6], out["depth"][out["acc"] >= 0.6])
    out["sample_rgb"] = rh[-1]["rgb"].numpy()
    out["sample_density"] = rh[-1]["density"].numpy()
    out["sample_coord"] = rh[-1]["coord"].numpy()
    cs = cases.param_checksums(params)
    out["checksum_keys"] = np.array(sorted(cs))
    out["checksum_vals"] = np.array([cs[k] for k in sorted(cs)], dtype=np.float64)
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: reference == oracle bit-for-bit on {batch['origins'].shape[0]} rays; wrote {path} "
          f"({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    if not ref_shim.available():
        sys.exit("needs /root/reference")
    for n in (sys.argv[1:] or list(cases.CASES)):
        run(n)
