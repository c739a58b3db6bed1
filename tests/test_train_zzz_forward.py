"""`ucnerf_b200.train_forward.level_loop` - Model.forward's level loop for the training step through the native ops -
against vectors from the REFERENCE's own Model.forward(rand=True) in train() mode with autograd on CPU
(oracle/make_train_forward_golden.py: draws patched in, loss over rgb / weights / acc / hash decay of both levels,
gradients of every parameter).  The model handed to level_loop is a mirror with the reference's attribute and
parameter names (the reference itself does not travel to the GPU box).  Sorted last: it runs every training op."""
import types

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import load_golden
from oracle import cases


class MirrorMLP(nn.Module):
    """Attribute / parameter names of the reference's MLP (models.py:L367-483) under configs/waymo.gin."""
    deg_view, net_depth_viewdirs, skip_layer_dir = 4, 2, 0
    density_bias, density_noise, bottleneck_noise = -1., 0., 0.
    rgb_premultiplier, rgb_bias, rgb_padding = 1., 0., 0.001
    disable_density_normals, warp_fn, num_glo_features, scale_featurization = True, 'contract', 0, False

    def __init__(self, gs, disable_rgb, bottleneck_width=256, width=256):
        super().__init__()
        from ucnerf_b200.gridencoder import GridEncoder
        self.disable_rgb, self.bottleneck_width = disable_rgb, bottleneck_width
        self.encoder = GridEncoder(3, gs.num_levels, gs.level_dim, base_resolution=gs.base_resolution,
                                   desired_resolution=gs.desired_resolution, log2_hashmap_size=gs.log2_hashmap_size)
        self.density_layer = nn.Sequential(nn.Linear(self.encoder.output_dim, 64), nn.ReLU(),
                                           nn.Linear(64, 1 if disable_rgb else bottleneck_width))
        if not disable_rgb:
            d_in = bottleneck_width + 27
            self.lin_second_stage_0 = nn.Linear(d_in, width)
            self.lin_second_stage_1 = nn.Linear(width + d_in, width)
            self.rgb_layer = nn.Linear(width, 3)


class MirrorModel(nn.Module):
    """Attribute names of the reference's Model (models.py:L28-55)."""
    bg_intensity_range, anneal_slope, stop_level_grad, use_viewdirs, raydist_fn = (1., 1.), 10, True, True, None
    single_jitter, num_glo_features, near_anneal_rate, single_mlp, distinct_prop = True, 0, None, False, True
    resample_padding, opaque_background, std_scale, learned_exposure_scaling = 0.0, False, 0.5, False

    def __init__(self, cfg):
        super().__init__()
        self.num_levels, self.num_prop_samples, self.num_nerf_samples = cfg.num_levels, cfg.num_prop_samples, cfg.num_nerf_samples
        self.dilation_multiplier, self.dilation_bias = cfg.dilation_multiplier, cfg.dilation_bias
        self.config = types.SimpleNamespace(brightness_correction=False, model_sky=False)
        for i, gs in enumerate(cfg.prop_grids):
            self.register_module(f"prop_mlp_{i}", MirrorMLP(gs, True))
        self.nerf_mlp = MirrorMLP(cfg.nerf_grid, False, cfg.bottleneck_width, cfg.net_width_viewdirs)


def loss_fn(renderings, ray_history, target, Gs):
    loss = 0.
    for l, (r, h) in enumerate(zip(renderings, ray_history)):
        loss = loss + (0.5 + l) * ((r['rgb'] - target) ** 2).sum() + (h['weights'] * Gs[l]).sum() \
            + 0.05 * r['acc'].sum() + 0.1 * h['loss_hash_decay']
    return loss


def projections(g, seed):
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(3):
        R = torch.randn(g.shape, generator=gen, dtype=torch.float64)
        out.append(float((g.double().cpu() * R).sum()))
    return np.array(out + [float(g.double().abs().sum())])


def test_gradient_scaler_equals_train_utils_formula():
    """CPU: train_utils.GradientScaler (train_utils.py:L101-111): identity forward, grads x clamp(ray_dist^2, 0, 1)."""
    from ucnerf_b200.train_forward import _GradientScaler
    g = torch.Generator().manual_seed(0)
    c = torch.rand((5, 7, 3), generator=g, requires_grad=True)
    s = torch.rand((5, 7), generator=g, requires_grad=True)
    dist = torch.rand((5, 7), generator=g) * 2
    c2, s2 = _GradientScaler.apply(c, s, dist)
    assert torch.equal(c2, c) and torch.equal(s2, s)
    (c2.sum() * 2 + (s2 * 3).sum()).backward()
    k = torch.square(dist).clamp(0, 1)
    assert torch.equal(c.grad, (2 * k)[..., None].expand_as(c)) and torch.equal(s.grad, 3 * k)


def test_unsupported_configurations_raise_before_any_kernel():
    from ucnerf_b200.train_forward import _check_mlp, _check_model
    m = types.SimpleNamespace(raydist_fn=None, near_anneal_rate=0.1)
    with pytest.raises(NotImplementedError):
        _check_model(m, {}, False)
    with pytest.raises(NotImplementedError):
        _check_model(types.SimpleNamespace(), {}, True)                       # compute_extras
    with pytest.raises(NotImplementedError):
        _check_mlp(types.SimpleNamespace(disable_density_normals=False))      # the reference's class default
    _check_mlp(types.SimpleNamespace(disable_density_normals=True))


@pytest.mark.gpu
def test_level_loop_matches_the_reference_training_forward_and_backward():
    from ucnerf_b200.train_forward import level_loop
    g = load_golden("train_forward")
    cfg, params, batch = cases.make_case("waymo", g["target"].shape[0])
    model = MirrorModel(cfg)
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not unexpected and all(k.endswith(".idx") for k in missing), (missing, unexpected)
    model = model.cuda().train()
    b = {k: v.cuda() for k, v in batch.items() if k != "rand_vec"}
    draws = [{k: torch.from_numpy(g[f"draw{l}_{k}"]).cuda() for k in ("jitter01", "flip01", "rot01", "rand_vec")}
             for l in range(cfg.num_levels)]
    renderings, ray_history = level_loop(model, True, b, float(g["train_frac"]), compute_extras=False, draws=draws)
    target = torch.from_numpy(g["target"]).cuda()
    Gs = [torch.from_numpy(g[f"G{l}"]).cuda() for l in range(cfg.num_levels)]
    loss = loss_fn(renderings, ray_history, target, Gs)
    loss.backward()
    assert len(renderings) == cfg.num_levels == len(ray_history)
    for l, (r, h) in enumerate(zip(renderings, ray_history)):
        assert float((h["sdist"].cpu() - torch.from_numpy(g[f"h{l}_sdist"])).abs().max()) < 4e-6
        # downstream of sdist the comparison inherits its round-off (5e-7 of s moves a point by a hundredth of a
        # finest-level cell).  Rehearsed on CPU with the serial instantiation of the same templates
        # (tests/dryrun_train_gpu_tests_on_cpu.py): density 1.7e-6, weights / rgb / acc 2.4e-7, loss 8e-8 relative,
        # gradients <= 6e-4 of the largest entry; the bars below leave >= 10x for the GPU's own round-off
        assert float((h["density"].detach().cpu() - torch.from_numpy(g[f"h{l}_density"])).abs().max()) < 1e-4
        assert float((r["weights"].detach().cpu() - torch.from_numpy(g[f"r{l}_weights"])).abs().max()) < 2e-5
        assert float((r["rgb"].detach().cpu() - torch.from_numpy(g[f"r{l}_rgb"])).abs().max()) < 2e-5
        assert float((r["acc"].detach().cpu() - torch.from_numpy(g[f"r{l}_acc"])).abs().max()) < 2e-5
        assert abs(float(h["loss_hash_decay"].detach()) - float(g[f"h{l}_loss_hash_decay"])) <= 1e-5 * float(g[f"h{l}_loss_hash_decay"])
        assert set(h) >= {"coord", "density", "rgb", "sdist", "weights", "loss_hash_decay", "normals", "roughness"}
    assert abs(float(loss.detach()) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    checked = 0
    for name, p in model.named_parameters():
        if "grad_" + name in g:
            ref = torch.from_numpy(g["grad_" + name])
            err = float((p.grad.cpu() - ref).abs().max())
            assert err <= 1e-2 * float(ref.abs().max()) + 1e-6, (name, err, float(ref.abs().max()))
            checked += 1
        elif "gproj_" + name in g:
            ref = g["gproj_" + name]
            got = projections(p.grad, seed=len(name))
            assert abs(got[3] - ref[3]) <= 5e-3 * ref[3], (name, got, ref)           # sum |grad|
            scale = ref[3] / np.sqrt(p.numel())                                      # typical size of a random projection
            assert np.abs(got[:3] - ref[:3]).max() <= 0.05 * max(np.abs(ref[:3]).max(), scale), (name, got, ref)
            checked += 1
    assert checked == len(list(model.named_parameters()))
