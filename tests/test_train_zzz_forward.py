"""`ucnerf_b200.train_forward.level_loop` - Model.forward's level loop for the training step through the native ops -
against vectors from the REFERENCE's own Model.forward(rand=True) in train() mode with autograd on CPU
(oracle/make_train_forward_golden.py: draws patched in, loss over rgb / weights / acc / hash decay of both levels,
gradients of every parameter).  The model handed to level_loop is the product's mirror of the reference classes
(ucnerf_b200/models.py: same attribute and parameter names; the reference itself does not travel to the GPU box).  Sorted last: it runs every training op."""
import types

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import cases


def make_model(cfg):
    """The product's mirror of the reference Model under configs/waymo.gin (ucnerf_b200/models.py)."""
    from ucnerf_b200.models import Model
    return Model(config=types.SimpleNamespace(brightness_correction=False, model_sky=False, vis_num_rays=16),
                 num_levels=cfg.num_levels, num_prop_samples=cfg.num_prop_samples, num_nerf_samples=cfg.num_nerf_samples,
                 prop_desired_grid_size=[gs.desired_resolution for gs in cfg.prop_grids],
                 dilation_multiplier=cfg.dilation_multiplier, dilation_bias=cfg.dilation_bias,
                 nerf_mlp_kwargs=dict(grid_disired_resolution=cfg.nerf_grid.desired_resolution,
                                      bottleneck_width=cfg.bottleneck_width, net_width_viewdirs=cfg.net_width_viewdirs))


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


def test_mirror_model_has_the_reference_state_dict_and_rejects_what_it_cannot_run():
    """CPU: parameter / buffer names and shapes of ucnerf_b200.models.Model == the reference's state_dict (the golden
    parameter set is keyed by the reference's names, oracle/ref_shim.py loads the same dict into the real Model)."""
    from ucnerf_b200.models import MLP, Model
    cfg, params, _ = cases.make_case("waymo", 2)
    model = make_model(cfg)
    sd = model.state_dict()
    assert set(sd) - set(params) == {k for k in sd if k.endswith(".idx")}        # level-id buffer, not in the parameter set
    assert set(params) <= set(sd)
    for k, v in params.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not unexpected and all(k.endswith(".idx") for k in missing)
    with pytest.raises(TypeError):
        Model(num_levles=2)
    with pytest.raises(NotImplementedError):
        MLP(disable_density_normals=False)
    with pytest.raises(NotImplementedError):
        Model(single_mlp=True)


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
@pytest.mark.parametrize("native_mlp", [True, False], ids=["tensor-core layers", "nn.Linear layers"])
def test_level_loop_matches_the_reference_training_forward_and_backward(native_mlp):
    from ucnerf_b200.train_forward import level_loop
    g = load_golden("train_forward")
    cfg, params, batch = cases.make_case("waymo", g["target"].shape[0])
    model = make_model(cfg)
    missing, unexpected = model.load_state_dict(params, strict=False)
    assert not unexpected and all(k.endswith(".idx") for k in missing), (missing, unexpected)
    model = model.cuda().train()
    b = {k: v.cuda() for k, v in batch.items() if k != "rand_vec"}
    draws = [{k: torch.from_numpy(g[f"draw{l}_{k}"]).cuda() for k in ("jitter01", "flip01", "rot01", "rand_vec")}
             for l in range(cfg.num_levels)]
    renderings, ray_history = level_loop(model, True, b, float(g["train_frac"]), compute_extras=False, draws=draws,
                                         native_mlp=native_mlp)
    assert model.training
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
    # embedding gradients ENTRY-WISE on the stored sample of table entries (8,192 touched by the 256 rays + untouched
    # ones, where only the hash-decay term acts): atomic-order / run-merging round-off only
    for name, p in model.named_parameters():
        if "gsel_idx_" + name not in g:
            continue
        sel = torch.from_numpy(g["gsel_idx_" + name].astype(np.int64))
        ref = torch.from_numpy(g["gsel_val_" + name])
        got = p.grad.cpu()[sel]
        scale = float(ref.abs().max())
        err = float((got - ref).abs().max())
        rel_l2 = float((got - ref).double().norm() / ref.double().norm())
        print(name, 'entry-wise: max err / max', err / scale, 'relative L2', rel_l2)
        # The first level's fenceposts are exact, so the proposal table sees atomic-order / run-merging round-off only.
        # The NeRF level's fenceposts carry the resampler's round-off (< 4e-6 of s = 6 % of a finest-level cell at
        # far = 8), which moves trilinear weights of single entries by percents while the field stays continuous:
        # a loose per-entry bar, a tight one on the whole sampled vector.
        if name.startswith("prop_mlp_0"):
            assert err <= 1e-4 * scale and rel_l2 < 1e-5, (name, err, scale, rel_l2)
        else:
            # (per grid level, printed with -s: a fencepost that differs by 4e-6 moves a point by 6 % of a finest-level cell
            # and, rarely, across a cell face on any level - single entries then differ by percents)
            offs = model.get_submodule(name.rsplit(".", 1)[0]).offsets.cpu().long()
            lvl = torch.bucketize(sel, offs[1:], right=True)
            for l in range(offs.numel() - 1):
                m = lvl == l
                if int(m.sum()) < 16:
                    continue
                l2 = float((got[m] - ref[m]).double().norm() / ref[m].double().norm())
                print("   level", l, "entries", int(m.sum()), "relative L2", l2)
            assert err <= 5e-2 * scale and rel_l2 < 2e-2, (name, err, scale, rel_l2)
        assert int(g["gsel_touched_" + name]) > 10000


@pytest.mark.gpu
def test_mirror_model_forward_dispatch_eval_is_the_fused_path_and_follows_weight_updates():
    """Model.forward: rand / training -> level_loop, eval -> the fused render path (== a HotPathModel built from the same
    state_dict), refreshed when the parameters change in place (an optimiser step)."""
    from ucnerf_b200.render import HotPathModel
    cfg, params, batch = cases.make_case("waymo", 96)
    model = make_model(cfg)
    model.load_state_dict(params, strict=False)
    model = model.cuda().eval()
    b = {k: v.cuda() for k, v in batch.items()}
    r1, h1 = model(False, b, 1.0, True)
    ref = HotPathModel.from_reference_model(model, model.config)
    r2, _ = ref.forward(False, b, 1.0, True)
    assert torch.equal(r1[-1]["rgb"], r2[-1]["rgb"]) and torch.equal(r1[-1]["depth"], r2[-1]["depth"])
    with torch.no_grad():
        model.nerf_mlp.rgb_layer.bias.add_(0.5)
    r3, _ = model(False, b, 1.0, True)
    assert float((r3[-1]["rgb"] - r1[-1]["rgb"]).abs().max()) > 1e-3
    model.train()
    rays = {k: v for k, v in b.items() if k != "rand_vec"}
    r4, h4 = model(True, rays, 0.5, False)
    assert r4[-1]["rgb"].requires_grad and h4[0]["weights"].shape == (96, cfg.num_prop_samples)
    # the two native routes agree: the training ops with rand=False against the fused eval kernels on the same weights
    # and cone basis (both sit within ~3e-7 of the reference-pinned oracle; 1e-4 is the eval path's parity bar)
    from ucnerf_b200.train_forward import level_loop
    with torch.no_grad():
        r5, h5 = level_loop(model, False, rays, 1.0, draws=[{"rand_vec": b["rand_vec"]} for _ in range(cfg.num_levels)])
    _, h3 = model.eval()(False, b, 1.0, True)
    for l in range(cfg.num_levels):
        assert float((h5[l]["sdist"] - h3[l]["sdist"]).abs().max()) < 4e-6
        assert float((h5[l]["weights"] - h3[l]["weights"]).abs().max()) < 1e-4
    assert float((r5[-1]["rgb"] - r3[-1]["rgb"]).abs().max()) < 1e-4
    assert float((r5[-1]["acc"] - r3[-1]["acc"]).abs().max()) < 1e-4
