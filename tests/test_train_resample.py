"""Stand-alone resampling op (ucnerf_resample_intervals / ucnerf_b200.stepfun.resample_level): the level-loop stretch
models.py:L156-205 with rand=False and rand=True (single / per-sample jitter).  Vectors: the REFERENCE's own stepfun.py
on CPU with torch.rand patched (oracle/make_resample_golden.py).  CPU: oracle bit-identical, device algorithm template
(ray_algos.cuh::resample_ray via tests/cpu_harness.cpp).  GPU: the CUDA op through the C ABI."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import ucnerf_oracle as O

CASES = ["waymo", "three_level"]
TAGS = ["det", "single", "each"]
cf = ctypes.c_float


def _fp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _rand01(g, name, tag):
    return None if tag == "det" else torch.from_numpy(g[f"{name}_rand_{tag}"])


@pytest.fixture(scope="module")
def gold():
    return load_golden("resample_op")


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tag", TAGS)
def test_oracle_is_bit_identical_to_reference_stepfun(gold, name, tag):
    g = gold
    out = O.resample_level(torch.from_numpy(g[f"{name}_t_prev"]), torch.from_numpy(g[f"{name}_w_prev"]),
                           int(g[f"{name}_S"]), float(g[f"{name}_dilation"]), True, float(g[f"{name}_anneal_{tag}"]),
                           float(g[f"{name}_padding_{tag}"]), _rand01(g, name, tag))
    assert np.array_equal(out.numpy(), g[f"{name}_sdist_{tag}"])


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tag", TAGS)
def test_device_algorithm_matches_reference_vectors_on_cpu(harness, gold, name, tag):
    g = gold
    t_prev = np.ascontiguousarray(g[f"{name}_t_prev"], np.float32)
    w_prev = np.ascontiguousarray(g[f"{name}_w_prev"], np.float32)
    N, n = w_prev.shape
    S = int(g[f"{name}_S"])
    base, jit = O.jitter_grid(S, _rand01(g, name, tag))
    base = base.numpy().copy()
    jit = None if jit is None else np.ascontiguousarray(jit.numpy(), np.float32)
    out = np.zeros((N, S + 1), np.float32)
    harness.h_resample_jitter(N, n, _fp(t_prev), _fp(w_prev), 1, cf(float(g[f"{name}_dilation"])),
                              cf(float(g[f"{name}_anneal_{tag}"])), cf(float(g[f"{name}_padding_{tag}"])), S, _fp(base),
                              _fp(jit), 0 if jit is None else jit.shape[1], _fp(out))
    ref = g[f"{name}_sdist_{tag}"]
    # the fp32 reference itself is ~1e-6 from the exact answer (softmax / log round-off), as in test_device_algos_cpu.py
    assert np.abs(out - ref).max() < 4e-6, np.abs(out - ref).max()
    assert np.all(np.diff(out, axis=1) >= 0) and out.min() >= 0 and out.max() <= 1


def test_first_level_with_jitter_on_cpu(harness):
    """models.py:L143-147: sdist = [0, 1], weights = [1]; rand=True shifts every center by the ray's jitter."""
    S, N = 16, 5
    r01 = torch.rand((N, 1), generator=torch.Generator().manual_seed(2))
    base, jit = O.jitter_grid(S, r01)
    ref = O.resample_level(torch.tensor([[0., 1.]]).repeat(N, 1), torch.ones((N, 1)), S, 0.0, False, 1.0, 0.0, r01).numpy()
    out = np.zeros((N, S + 1), np.float32)
    harness.h_resample_jitter(N, 1, None, None, 0, cf(0), cf(1), cf(0), S, _fp(base.numpy().copy()),
                              _fp(np.ascontiguousarray(jit.numpy(), np.float32)), 1, _fp(out))
    assert np.abs(out - ref).max() < 1e-6
    assert len({tuple(r) for r in out.round(6)}) == N          # every ray got its own jitter


# ---------------------------------------------------------------------------------------------------------------------


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tag", TAGS)
def test_cuda_op_matches_reference_vectors(gold, name, tag):
    from ucnerf_b200.stepfun import resample_level
    g = gold
    r01 = _rand01(g, name, tag)
    out = resample_level(torch.from_numpy(g[f"{name}_t_prev"]).cuda(), torch.from_numpy(g[f"{name}_w_prev"]).cuda(),
                         int(g[f"{name}_S"]), float(g[f"{name}_dilation"]), True, float(g[f"{name}_anneal_{tag}"]),
                         float(g[f"{name}_padding_{tag}"]), rand=tag != "det", single_jitter=tag != "each",
                         rand01=None if r01 is None else r01.cuda())
    ref = torch.from_numpy(g[f"{name}_sdist_{tag}"])
    err = float((out.cpu() - ref).abs().max())
    assert err < 4e-6, err


@pytest.mark.gpu
def test_cuda_op_training_batch_properties_and_errors():
    """One waymo.gin train batch (15,000 rays, 128 -> 32): sorted fenceposts inside [0, 1], reproducible with a seeded
    generator, different draws differ, rand=False equals the seeded golden path; CPU tensors raise."""
    from ucnerf_b200.stepfun import resample_level
    N, n, S = 15000, 128, 32
    g = torch.Generator().manual_seed(4)
    t = torch.sort(torch.rand((N, n + 1), generator=g), dim=-1).values
    t[:, 0], t[:, -1] = 0.0, 1.0
    # a well-conditioned histogram: with near-empty bins the inverse CDF amplifies fp32 round-off and the fp32 reference
    # itself is 1e-4 from an fp64 evaluation of the same formulas (checked with the oracle in float64)
    w = torch.rand((N, n), generator=g) + 0.1
    w = w / w.sum(-1, keepdim=True)
    t, w = t.cuda(), w.cuda()
    outs = []
    for seed in (7, 7, 8):
        gg = torch.Generator(device="cuda").manual_seed(seed)
        outs.append(resample_level(t, w, S, 0.0064, True, 1.0, 0.0, rand=True, single_jitter=True, generator=gg))
    a, b, c = outs
    assert a.shape == (N, S + 1) and bool(torch.isfinite(a).all())
    assert bool((a[:, 1:] >= a[:, :-1]).all()) and float(a.min()) >= 0 and float(a.max()) <= 1
    assert torch.equal(a, b) and not torch.equal(a, c)
    det = resample_level(t, w, S, 0.0064, True)
    ref = O.resample_level(t[:64].cpu(), w[:64].cpu(), S, 0.0064, True)
    assert float((det[:64].cpu() - ref).abs().max()) < 1e-5      # serial CPU instantiation on the same rays: 2.3e-6
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        resample_level(t.cpu(), w.cpu(), S)
    with pytest.raises(RuntimeError):
        resample_level(t, w[:, :-1], S)
