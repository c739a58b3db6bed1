"""CPU: host-side logic of the product package (no GPU): constants the C++ host code derives must equal what
torch computes in the reference, the GridEncoder mirror must lay tables out like the reference, errors must
surface like the reference's TORCH_CHECKs, and there must be no silent CPU fallback."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import ucnerf_oracle as O


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("S", [2, 3, 32, 64, 128, 512, 1000])
def test_u_grid_equals_torch_linspace(lib, S):
    """stepfun.py:L203-204: u = linspace(1/2S, 1 - 1/2S - eps, S) - bit-identical to torch."""
    u = np.zeros(S, np.float32)
    assert lib.ucnerf_debug_u_grid(S, _fp(u)) == 0
    pad = 1 / (2 * S)
    ref = torch.linspace(pad, 1. - pad - O.EPS, S).numpy()
    assert np.array_equal(u, ref)


def test_cone_table_equals_torch(lib):
    """render.py:L116-131 constants as torch evaluates them."""
    ct = np.zeros(30, np.float32)
    assert lib.ucnerf_debug_cone_table(_fp(ct)) == 0
    deg = torch.pi / 3 * torch.tensor([0, 2, 4, 3, 5, 1], dtype=torch.float)
    odd = torch.pi * 5 / 3 - (deg + torch.pi / 6)
    j = torch.arange(6)
    assert np.array_equal(ct[0:6], torch.cos(deg).numpy()) and np.array_equal(ct[6:12], torch.cos(odd).numpy())
    assert np.array_equal(ct[12:18], torch.sin(deg).numpy()) and np.array_equal(ct[18:24], torch.sin(odd).numpy())
    assert np.array_equal(ct[24:30], (3 / 7 ** 0.5 * (2 * j / 5 - 1)).numpy())


@pytest.mark.parametrize("kw", [dict(num_levels=10, desired_resolution=8192, log2_hashmap_size=21),
                                dict(num_levels=6, desired_resolution=512, log2_hashmap_size=21),
                                dict(num_levels=4, desired_resolution=128, log2_hashmap_size=15),
                                dict(num_levels=5, desired_resolution=None, log2_hashmap_size=12)])
def test_grid_encoder_mirror_layout(kw):
    from ucnerf_b200.gridencoder import GridEncoder
    enc = GridEncoder(input_dim=3, level_dim=4, base_resolution=16, **kw)
    lay = O.grid_layout(kw["num_levels"], 4, 16, kw["desired_resolution"], kw["log2_hashmap_size"])
    assert np.array_equal(enc.offsets.numpy(), lay["offsets"])
    assert np.array_equal(enc.grid_sizes.numpy(), lay["grid_sizes"])
    assert enc.embeddings.shape == (int(lay["offsets"][-1]), 4)
    assert enc.idx.shape[0] == int(lay["offsets"][-1]) and enc.idx.dtype == torch.long
    for l in range(kw["num_levels"]):
        assert torch.all(enc.idx[lay["offsets"][l]:lay["offsets"][l + 1]] == l)
    assert set(dict(enc.state_dict())) == {"embeddings", "offsets", "idx", "grid_sizes"}
    assert float(enc.embeddings.detach().abs().max()) <= 1e-4 and enc.output_dim == kw["num_levels"] * 4


def test_backend_rejects_cpu_tensors_like_the_reference():
    """gridencoder.cu:L15,L449: 'inputs must be a CUDA tensor' -> RuntimeError; no CPU path."""
    from ucnerf_b200.gridencoder import GridEncoder
    enc = GridEncoder(input_dim=3, num_levels=2, level_dim=2, desired_resolution=32, log2_hashmap_size=8)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        enc(torch.zeros(4, 3))


def test_renderer_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ucnerf_b200 import _lib
    from ucnerf_b200.render import HotPathModel
    with pytest.raises(_lib.UcnerfError, match="no CPU fallback"):
        HotPathModel({}, num_prop_samples=8, num_nerf_samples=8, num_prop_levels=1)


def test_product_package_never_imports_the_oracle():
    import os
    from conftest import ROOT
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ucnerf_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f


def test_shard_bounds_cover_image_exactly():
    from ucnerf_b200.render import shard_bounds
    for n in (1, 7, 480000, 2457600, 15001):
        for world in (1, 2, 3, 8):
            seen = 0
            for rank in range(world):
                per, a, b = shard_bounds(n, world, rank)
                assert a == min(rank * per, n) and b - a <= per
                seen += b - a
            assert seen == n and per * world >= n
