"""Stand-alone render.cast_rays op (ucnerf_cast_rays / ucnerf_b200.render_train.cast_rays; render.py:L94-152), rand=False and
rand=True.  Vectors: the REFERENCE's own render.cast_rays on CPU with torch.rand_like / randn_like patched
(oracle/make_cast_rays_golden.py).  CPU: train_algos.cuh::cast_interval - the code the kernel runs per (ray, interval) -
through tests/cpu_harness.cpp.  GPU: the CUDA op through the C ABI."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_golden


def _fp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def gold():
    return load_golden("cast_rays")


def _check(means, stds, ts, g, tag):
    rm, rs, rt = g[f"means_{tag}"], g[f"stds_{tag}"], g[f"ts_{tag}"]
    # the reference evaluates t_m ** 4 with a <= 1-ulp pow (ray_algos.cuh::make_cone_interval): ~0.2 % of the values move by one ulp
    assert (means == rm).mean() > 0.99 and np.abs(means - rm).max() <= 2.4e-7 * max(1.0, np.abs(rm).max())
    assert (ts == rt).mean() > 0.99 and np.abs(ts - rt).max() <= 4.8e-7
    assert (np.abs(stds - rs) / rs).max() < 5e-7


@pytest.mark.parametrize("tag", ["det", "rand"])
def test_per_interval_function_matches_reference_cast_rays_on_cpu(harness, gold, tag):
    g = gold
    N, S = g["tdist"].shape[0], g["tdist"].shape[1] - 1
    arr = lambda k: np.ascontiguousarray(g[k], np.float32)
    means = np.zeros((N, S, 6, 3), np.float32)
    stds = np.zeros((N, S, 6), np.float32)
    ts = np.zeros((N, S, 6), np.float32)
    rot = arr("rot01") if tag == "rand" else None
    flip = arr("flip01") if tag == "rand" else None
    harness.h_cast_rays(N, S, _fp(arr("tdist")), _fp(arr("origins")), _fp(arr("directions")), _fp(arr("cam_dirs")),
                        _fp(arr("radii").reshape(-1).copy()), _fp(arr("rand_vec")), _fp(rot), _fp(flip), ctypes.c_float(0.5),
                        _fp(means), _fp(stds), _fp(ts))
    _check(means, stds, ts, g, tag)
    if tag == "rand":   # the random pattern really differs from the deterministic one
        assert np.abs(means - g["means_det"]).max() > 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["det", "rand"])
def test_cuda_op_matches_reference_cast_rays(gold, tag):
    from ucnerf_b200.render_train import cast_rays
    g = gold
    c = lambda k: torch.from_numpy(g[k]).cuda()
    draws = (c("flip01") if tag == "rand" else None, c("rot01") if tag == "rand" else None, c("rand_vec"))
    means, stds, ts = cast_rays(c("tdist"), c("origins"), c("directions"), c("cam_dirs"), c("radii"), tag == "rand",
                                std_scale=0.5, draws=draws)
    _check(means.cpu().numpy(), stds.cpu().numpy(), ts.cpu().numpy(), g, tag)


@pytest.mark.gpu
def test_cuda_op_train_batch_seeded_stream_and_errors(gold):
    """15,000 rays x 32 intervals: shapes, finiteness, ordered distances with their mean inside the interval, the six points
    at radius * t / sqrt(2) from the axis; a seeded generator reproduces the draw; CPU tensors raise."""
    from ucnerf_b200.render_train import cast_rays
    N, S = 15000, 32
    gg = torch.Generator().manual_seed(12)
    d = torch.nn.functional.normalize(torch.randn((N, 3), generator=gg), dim=-1)
    cam = torch.nn.functional.normalize(d + 0.1 * torch.randn((N, 3), generator=gg), dim=-1)
    o = torch.rand((N, 3), generator=gg) * 0.2 - 0.1
    r = torch.full((N, 1), 5e-4)
    t = torch.sort(torch.rand((N, S + 1), generator=gg) * 8, dim=-1).values
    args = [x.cuda() for x in (t, o, d, cam, r)]
    outs = []
    for seed in (3, 3, 4):
        outs.append(cast_rays(*args, True, generator=torch.Generator(device="cuda").manual_seed(seed)))
    (m0, s0, t0), (m1, _, _), (m2, _, _) = outs
    assert m0.shape == (N, S, 6, 3) and s0.shape == (N, S, 6) and t0.shape == (N, S, 6)
    assert bool(torch.isfinite(m0).all()) and bool(torch.isfinite(s0).all())
    assert torch.equal(m0, m1) and not torch.equal(m0, m2)
    tc = args[0]
    # the six distances increase with j and their mean (the frustum's mean distance) lies inside the interval; single
    # points of a wide interval may overshoot t1, in the reference too
    assert bool((t0[..., 1:] >= t0[..., :-1]).all())
    mean_t = t0.double().mean(-1)
    assert bool((mean_t >= tc[:, :-1].double() - 1e-5).all()) and bool((mean_t <= tc[:, 1:].double() + 1e-5).all())
    axis = args[1][:, None, None, :] + t0[..., None] * args[2][:, None, None, :]
    assert float(((m0 - axis).norm(dim=-1) / (5e-4 * t0.clamp_min(1e-6))).max()) < 0.72      # radius * t / sqrt(2)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        cast_rays(t, o, d, cam, r, False)
