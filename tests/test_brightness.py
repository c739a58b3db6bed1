"""Brightness-correction affine folded into the compositing epilogue against vectors
produced by the reference's own Model.forward with config.brightness_correction=True
(tests/golden/brightness.npz, oracle/make_brightness_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import cases, ucnerf_oracle as O


def test_oracle_apply_brightness_matches_reference_vectors():
    g = load_golden("brightness")
    out = O.apply_brightness(torch.from_numpy(g["rgb_plain"]), torch.from_numpy(g["affine"]))
    assert np.array_equal(out.numpy(), g["rgb_corrected"])


@pytest.mark.gpu
def test_gpu_affine_epilogue_matches_reference_vectors():
    from test_gpu_render import build_renderer, run
    g = load_golden("brightness")
    cfg, params, batch = cases.make_case("waymo", int(g["n_rays"]))
    r = build_renderer(cfg, params)
    plain = run(r, batch)
    assert np.abs(plain["rgb"] - g["rgb_plain"]).max() < 1e-4
    r.set_rgb_affine(torch.from_numpy(g["affine"]))
    corr = run(r, batch)
    r.set_rgb_affine(None)
    again = run(r, batch)
    assert np.abs(corr["rgb"] - g["rgb_corrected"]).max() < 1e-4
    # the epilogue itself: exactly the affine of this kernel's own plain rgb, to fp32 round-off
    want = O.apply_brightness(torch.from_numpy(plain["rgb"]), torch.from_numpy(g["affine"])).numpy()
    assert np.abs(corr["rgb"] - want).max() < 5e-7
    assert np.array_equal(corr["packed"][:, :3], corr["rgb"])
    for k in ("acc", "depth", "weights_1", "sample_rgb"):
        assert np.array_equal(corr[k], plain[k]), k
    assert np.array_equal(again["rgb"], plain["rgb"])
