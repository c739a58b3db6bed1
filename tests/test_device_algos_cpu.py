"""CPU: the device algorithm templates (ucnerf_b200/csrc/ray_algos.cuh), instantiated serially by
tests/cpu_harness.cpp, against the oracle.  This checks tie semantics, op order and index arithmetic of the
CUDA code in a container without a GPU; the CUDA instantiation itself is checked by the -m gpu tests."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import cases, ucnerf_oracle as O


def _fp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


cf = ctypes.c_float


@pytest.fixture(scope="module")
def waymo():
    cfg, params, batch = cases.make_case("waymo", 64)
    rend, hist = O.model_forward(params, cfg, batch)
    return cfg, params, batch, rend, hist


def _u(S):
    pad = 1 / (2 * S)
    return torch.linspace(pad, 1. - pad - O.EPS, S).numpy().copy()


def test_first_level_resample_is_bit_exact(harness, waymo):
    cfg, _, _, _, hist = waymo
    S = cfg.num_prop_samples
    out = np.zeros((64, S + 1), np.float32)
    harness.h_resample(64, 1, None, None, 0, cf(0), cf(1), cf(0), S, _fp(_u(S)), _fp(out))
    assert np.array_equal(out, hist[0]["sdist"].numpy())


@pytest.mark.parametrize("case,n", [("waymo", 64), ("three_level", 32), ("target1024", 8), ("config1", 64)])
def test_resample_with_dilation_matches_oracle(harness, case, n):
    cfg, params, batch = cases.make_case(case, n)
    _, hist = O.model_forward(params, cfg, batch)
    prod = 1
    for lvl in range(1, cfg.num_levels):
        n_prev = hist[lvl - 1]["weights"].shape[1]
        prod *= n_prev
        S = hist[lvl]["weights"].shape[1]
        dil = np.float32(cfg.dilation_bias + cfg.dilation_multiplier / prod)
        t_prev = hist[lvl - 1]["sdist"].numpy().copy()
        w_prev = hist[lvl - 1]["weights"].numpy().copy()
        out = np.zeros((n, S + 1), np.float32)
        harness.h_resample(n, n_prev, _fp(t_prev), _fp(w_prev), 1, cf(dil), cf(1), cf(0), S, _fp(_u(S)), _fp(out))
        ref = hist[lvl]["sdist"].numpy()
        # the fp32 reference itself is ~1e-6 from the exact answer (softmax/log round-off)
        assert np.abs(out - ref).max() < 4e-6, (case, lvl, np.abs(out - ref).max())
        assert np.all(np.diff(out, axis=1) >= 0) and out.min() >= 0 and out.max() <= 1


def test_resample_tie_semantics_zero_width_and_zero_weight_bins(harness):
    """Zero-width bins get -inf logits (models.py:L188-191); empty-weight runs must be skipped by the inverse CDF
    with the reference's 'last xp<=x / first xp>x' rule (math.py:L88-107)."""
    t = torch.tensor([[0., 0.1, 0.1, 0.1, 0.4, 0.4, 0.7, 1.0, 1.0]])
    w = torch.tensor([[0.2, 0.0, 0.0, 0.0, 0.0, 0.5, 0.3, 0.0]])
    for S in (4, 16, 33):
        logits = torch.where(t[..., 1:] > t[..., :-1], torch.log(w), torch.full_like(w, -torch.inf))
        ref = O.sample_intervals_det(t, logits, S, (0., 1.)).numpy()
        out = np.zeros((1, S + 1), np.float32)
        harness.h_resample(1, 8, _fp(t.numpy().copy()), _fp(w.numpy().copy()), 0, cf(0), cf(1), cf(0), S, _fp(_u(S)),
                           _fp(out))
        np.testing.assert_allclose(out, ref, atol=1e-6)
    # with dilation: duplicates in t make ties in the 3-way merge
    td, wd = O.max_dilate_weights(t, w / w.sum(), 0.05, (0., 1.))
    td, wd = td[..., 1:-1], wd[..., 1:-1]
    logits = torch.where(td[..., 1:] > td[..., :-1], torch.log(wd), torch.full_like(wd, -torch.inf))
    ref = O.sample_intervals_det(td, logits, 16, (0., 1.)).numpy()
    out = np.zeros((1, 17), np.float32)
    wn = (w / w.sum()).numpy().copy()
    harness.h_resample(1, 8, _fp(t.numpy().copy()), _fp(wn), 1, cf(0.05), cf(1), cf(0), 16, _fp(_u(16)), _fp(out))
    np.testing.assert_allclose(out, ref, atol=1e-6)


def test_cone_points_and_contraction(harness, waymo):
    """render.cast_rays + coord.contract + /2 + (x+1)/2: >99.5% of coordinates bit-identical, rest within 1 ulp."""
    cfg, _, batch, _, hist = waymo
    b = {k: v.numpy().copy() for k, v in batch.items()}
    for lvl in range(cfg.num_levels):
        sd = hist[lvl]["sdist"]
        S = sd.shape[1] - 1
        g = np.zeros((64, S, 6, 3), np.float32)
        sg = np.zeros((64, S, 6), np.float32)
        harness.h_cone_points(64, S, _fp(b["origins"]), _fp(b["directions"]), _fp(b["cam_dirs"]), _fp(b["rand_vec"]),
                              _fp(b["radii"]), _fp(b["near"]), _fp(b["far"]), _fp(sd.numpy().copy()), cf(0.5),
                              _fp(g), _fp(sg))
        tdist = sd * batch["far"] + (1 - sd) * batch["near"]
        means, stds, _ = O.cast_rays_det(tdist, batch["origins"], batch["directions"], batch["cam_dirs"],
                                         batch["radii"], batch["rand_vec"])
        m, s = O.contract_mean_std(means.reshape(-1, 3), stds.reshape(-1))
        gr = ((m.reshape(64, S, 6, 3) / 2 + 1) / 2).numpy()
        sr = (s.reshape(64, S, 6) / 2).numpy()
        assert (g == gr).mean() > 0.995
        assert np.abs(g - gr).max() <= 1.2e-7
        assert (np.abs(sg - sr) / sr).max() < 1e-6


def test_composite_matches_oracle(harness, waymo):
    cfg, _, batch, rend, hist = waymo
    b = {k: v.numpy().copy() for k, v in batch.items()}
    lvl = cfg.num_levels - 1
    S = cfg.num_nerf_samples
    ow = np.zeros((64, S), np.float32)
    orr = np.zeros((64, 10), np.float32)
    harness.h_composite(64, S, _fp(hist[lvl]["sdist"].numpy().copy()), _fp(hist[lvl]["density"].numpy().copy()),
                        _fp(hist[lvl]["rgb"].numpy().copy()), _fp(b["directions"]), _fp(b["near"]), _fp(b["far"]),
                        cf(1.0), 1, _fp(ow), _fp(orr))
    r = rend[lvl]
    np.testing.assert_allclose(ow, r["weights"].numpy(), atol=2e-7)
    np.testing.assert_allclose(orr[:, 0:3], r["rgb"].numpy(), atol=1e-6)
    for i, k in enumerate(["depth", "depth_raw", "acc", "distance_mean", "distance_median", "distance_percentile_5",
                           "distance_percentile_95"], start=3):
        np.testing.assert_allclose(orr[:, i], r[k].numpy(), atol=3e-6, err_msg=k)
    # proposal level: no colours -> background only
    harness.h_composite(64, cfg.num_prop_samples, _fp(hist[0]["sdist"].numpy().copy()),
                        _fp(hist[0]["density"].numpy().copy()), None, _fp(b["directions"]), _fp(b["near"]),
                        _fp(b["far"]), cf(1.0), 0, _fp(np.zeros((64, cfg.num_prop_samples), np.float32)), _fp(orr))
    np.testing.assert_allclose(orr[:, 0], rend[0]["rgb"].numpy()[:, 0], atol=1e-6)


def test_fused_path_hash_lookup_matches_kernel_restatement(harness):
    """level_index / cell_of with host-precomputed level constants == the literal restatement of kernel_grid."""
    gs = O.GridSpec(8192)
    lay = gs.layout()
    offs = lay["offsets"]
    L = gs.num_levels
    rng = np.random.default_rng(0)
    emb = rng.standard_normal((int(offs[-1]), 4)).astype(np.float32)
    x = rng.random((512, 3), dtype=np.float32)
    x[:4] = [[0, 0, 0], [1, 1, 1], [0, 1, 0.5], [0.999999, 0.5, 0.25]]
    desc = np.zeros((L, 6), np.uint32)
    for l in range(L):
        scale, res = O.level_geometry(l, 1.0, 16)
        hs = int(offs[l + 1] - offs[l])
        stride, d = 1, 0
        while d < 3 and stride <= hs:
            stride = (stride * (int(res) + 1)) & 0xFFFFFFFF
            d += 1
        desc[l] = [offs[l], hs, int(res) + 1, int(stride > hs), hs - 1 if hs & (hs - 1) == 0 else 0,
                   np.float32(scale).view(np.uint32)]
    out = np.zeros((512, L, 4), np.float32)
    harness.h_grid_features(512, L, _fp(desc), _fp(emb), _fp(x), _fp(out))
    ref, _ = O.grid_encode_forward(x, emb, offs, 512, 3, 4, L, 1.0, 16)
    assert np.array_equal(out, ref.transpose(1, 0, 2))


def test_sample_coord_matches_reference_golden(harness):
    """`coord` of the NeRF level (models.py:L512,L677) from the device algorithm against the reference's own values."""
    from conftest import load_golden
    g = load_golden("waymo")
    n, S = g["sample_coord"].shape[:2]
    arr = lambda k: np.ascontiguousarray(g["batch_" + k], np.float32)
    sd = np.ascontiguousarray(g[f"sdist_{1}"], np.float32)
    out = np.zeros((n, S, 3), np.float32)
    harness.h_sample_coord(n, S, _fp(arr("origins")), _fp(arr("directions")), _fp(arr("cam_dirs")), _fp(arr("rand_vec")),
                           _fp(arr("radii").reshape(-1)), _fp(arr("near").reshape(-1)), _fp(arr("far").reshape(-1)), _fp(sd),
                           cf(0.5), _fp(out))
    err = np.abs(out - g["sample_coord"]).max()
    assert err < 3e-7, err
