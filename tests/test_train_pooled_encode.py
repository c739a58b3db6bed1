"""Pooled hash-grid encode (training front end of MLP.predict_density, models.py:L485-496; SURVEY.md section 8a rows
R4 + R10).  CPU: the oracle restatement and the device algorithm templates (pooled_algos.cuh through
tests/cpu_harness.cpp) against vectors produced by the REFERENCE's own modules under autograd
(oracle/make_pooled_golden.py).  GPU: the CUDA kernels through the C ABI / autograd Function against the same vectors,
the oracle on fresh inputs, and - at full training size - the adjoint identity <F(E), dF> = <E, dE(dF)> (the
features are linear in the embeddings).

Tolerances.  On the reference's vectors and wherever the grid positions are given (contract=False) features agree to
2e-6.  With the contraction inside the kernel a position can differ by one ulp from torch's (torch reduces |x|^2 in its
own order on CPU and on CUDA); one ulp of a unit-cube coordinate is 5e-4 of a cell on the finest NeRF level, i.e. up to
~1e-4 of feature value there (measured on B200: 1.1e-5 .. 2.3e-5 on the NeRF grid, 8e-6 on the proposal grid, the same
effect the hot path's 1e-4 bar absorbs; the serial CPU instantiation of the same template differs from torch by exactly
the same 1.98e-5 / 2.28e-5 / 1.12e-5 on those inputs, gradients by 2.3e-5 .. 3.6e-5 of the largest entry).  Those cases
are therefore held to 1e-4 (features) / 2e-4 of the largest gradient entry, and the adjoint identity - which does not
depend on how positions round - pins the backward to 1e-5."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import cases, ucnerf_oracle as O

TAGS = {"prop": ("prop_mlp_0", lambda cfg: cfg.prop_grids[0]), "nerf": ("nerf_mlp", lambda cfg: cfg.nerf_grid)}


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def case():
    g = load_golden("pooled_encode")
    cfg, params, _ = cases.make_case("waymo", int(g["n_rays"]))
    return g, cfg, params


def _dense_grad(g, tag, n_rows):
    ge = np.zeros((n_rows, 4), np.float32)
    ge[g[f"{tag}_grad_rows"]] = g[f"{tag}_grad_vals"]
    return ge


@pytest.mark.parametrize("tag", ["prop", "nerf"])
def test_oracle_matches_reference_autograd_vectors(case, tag):
    g, cfg, params = case
    prefix, gsf = TAGS[tag]
    gs = gsf(cfg)
    means, stds = torch.from_numpy(g["means"]), torch.from_numpy(g["stds"])
    feats, coord, _, _ = O.pooled_encode_forward(params, prefix, gs, means, stds)
    assert np.array_equal(feats.numpy(), g[f"{tag}_features"])          # same ops, same order: bit-identical
    assert np.array_equal(coord.numpy(), g[f"{tag}_coord"])
    ge = O.pooled_encode_backward(params, prefix, gs, means, stds, torch.from_numpy(g[f"{tag}_grad_features"])).numpy()
    ref = _dense_grad(g, tag, ge.shape[0])
    assert np.abs(ge - ref).max() <= 2e-6 * np.abs(ref).max()
    assert np.array_equal(np.nonzero(np.abs(ge).sum(-1))[0], g[f"{tag}_grad_rows"])


def _harness_run(harness, params, prefix, gs, means, stds, grad_feats, contract=1, merge_runs=False):
    emb = params[prefix + ".encoder.embeddings"].numpy()
    offs = np.ascontiguousarray(params[prefix + ".encoder.offsets"].numpy(), np.int32)
    gsz = np.ascontiguousarray(params[prefix + ".encoder.grid_sizes"].numpy(), np.int32)
    L = gs.num_levels
    S = float(np.log2(gs.layout()["per_level_scale"]))
    M = means.shape[-2]
    m = np.ascontiguousarray(means.reshape(-1, M, 3), np.float32)
    s = np.ascontiguousarray(stds.reshape(-1, M), np.float32)
    B = m.shape[0]
    feats = np.zeros((B, L * 4), np.float32)
    coord = np.zeros((B, 3), np.float32)
    harness.h_pooled_forward(B, M, contract, L, _fp(offs), _fp(gsz), ctypes.c_float(S), gs.base_resolution, _fp(emb),
                             _fp(m), _fp(s), _fp(feats), _fp(coord))
    ge = np.zeros((emb.shape[0], 4), np.float64)
    gf = np.ascontiguousarray(grad_feats.reshape(B, L * 4), np.float32)
    harness.h_pooled_backward(B, M, contract | (4 if merge_runs == 'ray' else (2 if merge_runs else 0)), L, _fp(offs), _fp(gsz), ctypes.c_float(S), gs.base_resolution, _fp(gf),
                              _fp(m), _fp(s), _fp(ge))
    return feats, coord, ge


@pytest.mark.parametrize("tag", ["prop", "nerf"])
def test_device_algorithm_matches_reference_vectors_on_cpu(harness, case, tag):
    g, cfg, params = case
    prefix, gsf = TAGS[tag]
    gs = gsf(cfg)
    feats, coord, ge = _harness_run(harness, params, prefix, gs, g["means"], g["stds"], g[f"{tag}_grad_features"])
    ref_f = g[f"{tag}_features"].reshape(feats.shape)
    assert np.abs(feats - ref_f).max() < 2e-6, np.abs(feats - ref_f).max()
    assert np.abs(coord - g[f"{tag}_coord"].reshape(-1, 3)).max() < 3e-7
    ref = _dense_grad(g, tag, ge.shape[0])
    assert np.abs(ge - ref).max() <= 3e-6 * np.abs(ref).max()
    assert np.array_equal(np.nonzero(np.abs(ge).sum(-1))[0], g[f"{tag}_grad_rows"])


@pytest.mark.parametrize("tag", ["prop", "nerf"])
def test_run_merging_backward_gives_the_same_gradient_on_cpu(harness, case, tag):
    """pooled_level_backward_runs: consecutive points of one cell reduced once; same rows, same values up to fp32 order -
    and fewer reductions (counted through the number of distinct (row, value) adds is not observable here, so the
    gradient itself and the reference vectors are the check)."""
    g, cfg, params = case
    prefix, gsf = TAGS[tag]
    gs = gsf(cfg)
    _, _, ge0 = _harness_run(harness, params, prefix, gs, g["means"], g["stds"], g[f"{tag}_grad_features"])
    ref = _dense_grad(g, tag, ge0.shape[0])
    for variant in (True, 'ray'):        # within an interval / across 4 consecutive intervals of a ray
        _, _, ge1 = _harness_run(harness, params, prefix, gs, g["means"], g["stds"], g[f"{tag}_grad_features"], merge_runs=variant)
        assert np.abs(ge1 - ge0).max() <= 2e-6 * np.abs(ref).max()
        assert np.abs(ge1 - ref).max() <= 3e-6 * np.abs(ref).max()
        assert np.array_equal(np.nonzero(np.abs(ge1).sum(-1))[0], g[f"{tag}_grad_rows"])
    # out-of-range points between in-range ones, no contraction, M = 3
    rng = np.random.default_rng(8)
    means = rng.uniform(-1.2, 1.2, (64, 3, 3)).astype(np.float32)
    stds = rng.uniform(1e-4, 5e-2, (64, 3)).astype(np.float32)
    gf = rng.standard_normal((64, gs.num_levels * 4)).astype(np.float32)
    _, _, a = _harness_run(harness, params, prefix, gs, means, stds, gf, contract=0)
    for variant in (True, 'ray'):
        _, _, b = _harness_run(harness, params, prefix, gs, means, stds, gf, contract=0, merge_runs=variant)
        assert np.abs(a - b).max() <= 2e-6 * np.abs(a).max()


def test_device_algorithm_edge_cases_on_cpu(harness, case):
    """No contraction (warp_fn=None), M != 6, points outside the unit cube (zero features / no gradient,
    gridencoder.cu:L110-135,L276-281) and the adjoint identity."""
    _, cfg, params = case
    gs = cfg.prop_grids[0]
    rng = np.random.default_rng(5)
    means = rng.uniform(-1.2, 1.2, (40, 3, 3)).astype(np.float32)     # some points leave [-1,1]^3 -> out of range
    stds = rng.uniform(1e-4, 5e-2, (40, 3)).astype(np.float32)
    gf = rng.standard_normal((40, gs.num_levels * 4)).astype(np.float32)
    feats, _, ge = _harness_run(harness, params, "prop_mlp_0", gs, means, stds, gf, contract=0)
    m, s = torch.from_numpy(means), torch.from_numpy(stds)
    L = gs.num_levels
    ref = O.encoder_forward(params, "prop_mlp_0.encoder", gs, m).unflatten(-1, (L, -1))
    w = torch.erf(1 / torch.sqrt(8 * s[..., None] ** 2 * params["prop_mlp_0.encoder.grid_sizes"] ** 2))
    ref = (ref * w[..., None]).mean(dim=-3).flatten(-2, -1).numpy()
    assert np.abs(feats - ref).max() < 2e-6
    assert (np.abs(means) > 1).any(-1).any()                      # the case does contain out-of-range points
    emb = params["prop_mlp_0.encoder.embeddings"].numpy().astype(np.float64)
    lhs = float((feats.astype(np.float64) * gf).sum())
    rhs = float((emb * ge).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs)), (lhs, rhs)


# ---------------------------------------------------------------------------------------------------------------------
# GPU: the CUDA instantiation through the C ABI / autograd Function


def _gpu_encoder(params, prefix, gs):
    from ucnerf_b200.gridencoder import GridEncoder
    enc = GridEncoder(3, gs.num_levels, gs.level_dim, base_resolution=gs.base_resolution,
                      desired_resolution=gs.desired_resolution, log2_hashmap_size=gs.log2_hashmap_size).cuda()
    assert torch.equal(enc.offsets.cpu(), params[prefix + ".encoder.offsets"])
    with torch.no_grad():
        enc.embeddings.copy_(params[prefix + ".encoder.embeddings"])
    return enc


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["prop", "nerf"])
def test_cuda_matches_reference_autograd_vectors(case, tag):
    from ucnerf_b200.gridencoder.pooled import pooled_encode
    g, cfg, params = case
    prefix, gsf = TAGS[tag]
    enc = _gpu_encoder(params, prefix, gsf(cfg))
    means, stds = torch.from_numpy(g["means"]).cuda(), torch.from_numpy(g["stds"]).cuda()
    feats, coord = pooled_encode(enc, means, stds)
    assert feats.shape == g[f"{tag}_features"].shape and coord.shape == g[f"{tag}_coord"].shape
    err = float((feats.detach().cpu() - torch.from_numpy(g[f"{tag}_features"])).abs().max())
    assert err < 2e-6, err
    assert float((coord.cpu() - torch.from_numpy(g[f"{tag}_coord"])).abs().max()) < 3e-7
    feats.backward(torch.from_numpy(g[f"{tag}_grad_features"]).cuda())
    ge = enc.embeddings.grad.cpu().numpy()
    ref = _dense_grad(g, tag, ge.shape[0])
    assert np.abs(ge - ref).max() <= 1e-5 * np.abs(ref).max()      # fp32 atomics in arbitrary order
    assert np.array_equal(np.nonzero(np.abs(ge).sum(-1))[0], g[f"{tag}_grad_rows"])


@pytest.mark.gpu
@pytest.mark.parametrize("M,contract", [(6, True), (3, False), (1, True), (8, True)])
def test_cuda_matches_oracle_on_fresh_inputs(case, M, contract):
    from ucnerf_b200.gridencoder.pooled import pooled_encode
    _, cfg, params = case
    gs = cfg.nerf_grid
    enc = _gpu_encoder(params, "nerf_mlp", gs)
    g = torch.Generator().manual_seed(100 + M)
    B = 777                                                     # does not tile the 256-thread blocks
    scale = 6.0 if contract else 1.2
    means = (torch.rand((B, M, 3), generator=g) * 2 - 1) * scale
    means[:50] *= 0.1                                           # inside the unit ball: identity branch of contract
    stds = torch.rand((B, M), generator=g) * 0.05 + 1e-5
    gf = torch.randn((B, gs.num_levels * 4), generator=g)
    feats, coord = pooled_encode(enc, means.cuda(), stds.cuda(), contract=contract)
    feats.backward(gf.cuda())
    if contract:
        ref_f, ref_c, _, _ = O.pooled_encode_forward(params, "nerf_mlp", gs, means, stds)
        ref_g = O.pooled_encode_backward(params, "nerf_mlp", gs, means, stds, gf).numpy()
        assert float((coord.cpu() - ref_c).abs().max()) < 5e-7
    else:
        L = gs.num_levels
        f = O.encoder_forward(params, "nerf_mlp.encoder", gs, means).unflatten(-1, (L, -1))
        w = torch.erf(1 / torch.sqrt(8 * stds[..., None] ** 2 * params["nerf_mlp.encoder.grid_sizes"] ** 2))
        ref_f = (f * w[..., None]).mean(dim=-3).flatten(-2, -1)
        ref_g = None
    tol_f, tol_g = (1e-4, 2e-4) if contract else (2e-6, 1e-5)   # see the module docstring
    assert float((feats.detach().cpu() - ref_f).abs().max()) < tol_f
    ge = enc.embeddings.grad.cpu().numpy()
    if ref_g is not None:
        assert np.abs(ge - ref_g).max() <= tol_g * np.abs(ref_g).max()
    # adjoint identity (features are linear in the embeddings)
    prod = feats.detach().double() * gf.cuda().double()
    lhs = float(prod.sum())
    rhs = float((enc.embeddings.detach().double() * enc.embeddings.grad.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * float(prod.abs().sum()), (lhs, rhs)


@pytest.mark.gpu
def test_cuda_equals_the_unfused_gridencoder_chain_and_errors(case):
    """The fused op against the reference-shaped chain (contract in torch -> GridEncoder drop-in kernels -> erf weights
    -> mean) on the same GPU, forward and embeddings.grad; plus the error behaviour."""
    from ucnerf_b200.gridencoder.pooled import pooled_encode
    _, cfg, params = case
    gs = cfg.prop_grids[0]
    enc = _gpu_encoder(params, "prop_mlp_0", gs)
    g = torch.Generator().manual_seed(9)
    means = ((torch.rand((4096, 6, 3), generator=g) * 2 - 1) * 3).cuda()
    stds = (torch.rand((4096, 6), generator=g) * 0.02 + 1e-5).cuda()
    gf = torch.randn((4096, gs.num_levels * 4), generator=g).cuda()
    feats, _ = pooled_encode(enc, means, stds)
    feats.backward(gf)
    g_fused = enc.embeddings.grad.clone()
    enc.embeddings.grad = None
    m, s = O.contract_mean_std(means.reshape(-1, 3), stds.reshape(-1))          # torch ops, run on the GPU tensors
    m, s = m.reshape(4096, 6, 3) / 2, s.reshape(4096, 6) / 2
    f = enc(m, bound=1).unflatten(-1, (gs.num_levels, -1))
    w = torch.erf(1 / torch.sqrt(8 * s[..., None] ** 2 * enc.grid_sizes ** 2))
    chain = (f * w[..., None]).mean(dim=-3).flatten(-2, -1)
    chain.backward(gf)
    assert float((feats.detach() - chain.detach()).abs().max()) < 1e-4          # see the module docstring
    ref = enc.embeddings.grad
    assert float((g_fused - ref).abs().max()) <= 2e-4 * float(ref.abs().max())
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        pooled_encode(enc, means.cpu(), stds.cpu())
    with pytest.raises(RuntimeError):
        pooled_encode(enc, means[..., :2], stds)


@pytest.mark.gpu
@pytest.mark.parametrize("M", [6, 4])
def test_cuda_run_merging_backward_equals_the_plain_backward(case, M):
    from ucnerf_b200.gridencoder.pooled import pooled_encode
    _, cfg, params = case
    gs = cfg.prop_grids[0]
    enc = _gpu_encoder(params, "prop_mlp_0", gs)
    g = torch.Generator().manual_seed(40 + M)
    B = 3000
    t = torch.rand((B, 1, 1), generator=g) ** 2 * 7.5 + 0.02               # points of one interval close together
    d = torch.nn.functional.normalize(torch.randn((B, 1, 3), generator=g), dim=-1)
    means = (d * t + 2e-2 * t * torch.randn((B, M, 3), generator=g)).cuda()
    stds = (5e-4 * t.expand(-1, M, 1)).reshape(B, M).contiguous().cuda()
    gf = torch.randn((B, gs.num_levels * 4), generator=g).cuda()
    grads = []
    for merge in (False, True, 'ray'):
        enc.embeddings.grad = None
        feats, _ = pooled_encode(enc, means, stds, merge_runs=merge)
        feats.backward(gf)
        grads.append(enc.embeddings.grad.clone())
    for gm in grads[1:]:
        assert float((grads[0] - gm).abs().max()) <= 1e-5 * float(grads[0].abs().max())
        assert torch.equal(grads[0].abs().sum(-1) > 0, gm.abs().sum(-1) > 0)


@pytest.mark.gpu
def test_cuda_full_training_size_properties(case):
    """One GPU's share of a 65,536-ray train batch on the proposal level (8,192 rays x 128 intervals = 1,048,576
    intervals x 6 points): finite outputs, zero gradient rows stay zero where no point lands, adjoint identity."""
    from ucnerf_b200.gridencoder.pooled import pooled_encode
    _, cfg, params = case
    gs = cfg.prop_grids[0]
    enc = _gpu_encoder(params, "prop_mlp_0", gs)
    B = 8192 * 128
    g = torch.Generator(device="cuda").manual_seed(3)
    means = (torch.rand((B, 6, 3), generator=g, device="cuda") * 2 - 1) * 4
    stds = torch.rand((B, 6), generator=g, device="cuda") * 0.03 + 1e-5
    gf = torch.randn((B, gs.num_levels * 4), generator=g, device="cuda")
    feats, coord = pooled_encode(enc, means, stds)
    feats.backward(gf)
    assert bool(torch.isfinite(feats).all()) and bool(torch.isfinite(coord).all())
    assert float(coord.abs().max()) <= 1.0 + 1e-6                  # contracted means / 2 live in the unit ball
    prod = feats.detach().double() * gf.double()
    lhs = float(prod.sum())
    rhs = float((enc.embeddings.detach().double() * enc.embeddings.grad.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * float(prod.abs().sum()), (lhs, rhs)


def test_level_whose_stride_product_wraps_is_indexed_like_the_reference_on_cpu(harness):
    """desired_resolution 65536 -> level 12 has resolution + 1 = 65537: the reference's 32-bit stride product wraps
    (gridencoder.cu:L72-77), the level then counts as dense although it does not fit its table and the index is reduced
    modulo the table size.  The product's level descriptor (make_grid_level) + pooled_level_forward against the oracle."""
    gs = O.GridSpec(65536)
    assert gs.num_levels == 13
    cfg = O.HotPathConfig(prop_grids=[gs], nerf_grid=O.GridSpec(64))
    params = O.init_params(cfg, seed=5)
    gen = torch.Generator().manual_seed(6)
    means = (torch.rand((40, 6, 3), generator=gen) * 2 - 1) * 0.9
    stds = torch.rand((40, 6), generator=gen) * 1e-3 + 1e-4
    feats, coord, _, _ = O.pooled_encode_forward(params, "prop_mlp_0", gs, means, stds)
    got, gcoord, _ = _harness_run(harness, params, "prop_mlp_0", gs, means.numpy(), stds.numpy(),
                                  np.zeros((40, gs.num_levels * 4), np.float32))
    ref = feats.numpy().reshape(40, gs.num_levels, 4)
    err = np.abs(got.reshape(ref.shape) - ref).max(axis=(0, 2))
    assert err[:11].max() < 1e-5, err
    # levels 11 and 12: pos = x * 65535 + 0.5 leaves 7 - 8 fraction bits in fp32, so the rounding of that one operation
    # (fused or not) moves the interpolation weights by up to 2^-7; a wrongly indexed entry would be off by the table
    # amplitude (0.5) times the erf weight (about 0.06 here), a hundred times the bar
    assert err[11:].max() < 1e-3, err
    assert np.abs(ref[:, 12]).mean() > 3e-3
