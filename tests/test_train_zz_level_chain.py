"""One NeRF level of the reference's TRAINING forward + backward through the native seams chained together
(INTEGRATION.md seams 7 -> 9 -> 6 -> nn.Linear layers in torch -> 8) against the reference-shaped chain evaluated with the
oracle on CPU (models.py:L156-283 + MLP.forward L514-685 under waymo.gin; gradients by torch autograd and the oracle's
kernel_grid_backward restatement).  GPU only; sorted last on purpose - it exercises every training op at once."""
import numpy as np
import pytest
import torch

from oracle import cases, ucnerf_oracle as O

MLP_KEYS = ["density_layer.0.weight", "density_layer.0.bias", "density_layer.2.weight", "density_layer.2.bias",
            "lin_second_stage_0.weight", "lin_second_stage_0.bias", "lin_second_stage_1.weight", "lin_second_stage_1.bias",
            "rgb_layer.weight", "rgb_layer.bias"]


def mlp_tail(feats, P, viewdirs, cfg):
    """MLP.forward after the pooled features (models.py:L507-508, L581, L587-674 under waymo.gin), device-agnostic."""
    F = torch.nn.functional
    x = F.linear(torch.relu(F.linear(feats, P["density_layer.0.weight"], P["density_layer.0.bias"])),
                 P["density_layer.2.weight"], P["density_layer.2.bias"])
    density = F.softplus(x[..., 0] + cfg.density_bias)
    scales = 2.0 ** torch.arange(0, cfg.deg_view, device=viewdirs.device, dtype=viewdirs.dtype)      # coord.py:L214-225
    xb = (viewdirs[..., None, :] * scales[:, None]).reshape(viewdirs.shape[:-1] + (-1,))
    enc = torch.cat([viewdirs, torch.sin(torch.cat([xb, xb + 0.5 * torch.pi], dim=-1))], dim=-1)
    enc = torch.broadcast_to(enc[..., None, :], x.shape[:-1] + (enc.shape[-1],))
    inp = torch.cat([x, enc], dim=-1)
    h = torch.relu(F.linear(inp, P["lin_second_stage_0.weight"], P["lin_second_stage_0.bias"]))
    h = torch.relu(F.linear(torch.cat([h, inp], dim=-1), P["lin_second_stage_1.weight"], P["lin_second_stage_1.bias"]))
    rgb = torch.sigmoid(F.linear(h, P["rgb_layer.weight"], P["rgb_layer.bias"]))
    return density, rgb * (1 + 2 * cfg.rgb_padding) - cfg.rgb_padding


def test_mlp_tail_equals_the_oracle_mlp_forward():
    """CPU: the helper above is the oracle's MLP tail (so the GPU chain below is compared like with like)."""
    cfg, params, batch = cases.make_case("waymo", 4)
    g = torch.Generator().manual_seed(1)
    t = torch.sort(torch.rand((4, 9), generator=g) * 6 + 0.1, dim=-1).values
    means, stds, _ = O.cast_rays_det(t, batch["origins"], batch["directions"], batch["cam_dirs"], batch["radii"], batch["rand_vec"])
    ref = O.mlp_forward(params, "nerf_mlp", cfg.nerf_grid, cfg, means, stds, batch["viewdirs"], False)
    P = {k: params["nerf_mlp." + k] for k in MLP_KEYS}
    density, rgb = mlp_tail(ref["features"], P, batch["viewdirs"], cfg)
    assert torch.equal(density, ref["density"]) and torch.equal(rgb, ref["rgb"])


@pytest.mark.gpu
def test_training_level_through_the_native_seams_matches_the_reference_shaped_chain():
    from ucnerf_b200.gridencoder import GridEncoder
    from ucnerf_b200.gridencoder.pooled import pooled_encode
    from ucnerf_b200.render_train import cast_rays, composite
    from ucnerf_b200.stepfun import resample_level
    N = 48
    cfg, params, batch = cases.make_case("waymo", N)
    _, hist = O.model_forward(params, cfg, batch)
    t_prev, w_prev = hist[0]["sdist"], hist[0]["weights"]
    S, gs, prefix = cfg.num_nerf_samples, cfg.nerf_grid, "nerf_mlp"
    dilation = float(np.float32(cfg.dilation_bias + cfg.dilation_multiplier / cfg.num_prop_samples))
    g = torch.Generator().manual_seed(5)
    rand01 = torch.rand((N, 1), generator=g)
    target = torch.rand((N, 3), generator=g)
    Gw = torch.randn((N, S), generator=g) * 0.1
    near, far = batch["near"], batch["far"]

    # ---- native seams on the GPU -------------------------------------------------------------------------------------
    cu = lambda x: x.cuda()
    sdist_g = resample_level(cu(t_prev), cu(w_prev), S, dilation, True, 1.0, 0.0, rand=True, single_jitter=True, rand01=cu(rand01))
    tdist_g = sdist_g * cu(far) + (1 - sdist_g) * cu(near)                                   # coord.py:L176, fn = None
    means_g, stds_g, _ = cast_rays(tdist_g, cu(batch["origins"]), cu(batch["directions"]), cu(batch["cam_dirs"]),
                                   cu(batch["radii"]), False, std_scale=0.5, draws=(None, None, cu(batch["rand_vec"])))
    enc = GridEncoder(3, gs.num_levels, gs.level_dim, base_resolution=gs.base_resolution,
                      desired_resolution=gs.desired_resolution, log2_hashmap_size=gs.log2_hashmap_size).cuda()
    with torch.no_grad():
        enc.embeddings.copy_(params[prefix + ".encoder.embeddings"])
    Pg = {k: params[prefix + "." + k].cuda().requires_grad_(True) for k in MLP_KEYS}
    feats_g, _ = pooled_encode(enc, means_g, stds_g)
    density_g, rgbs_g = mlp_tail(feats_g, Pg, cu(batch["viewdirs"]), cfg)
    w_g, rgb_g, acc_g = composite(density_g, rgbs_g, tdist_g, cu(batch["directions"]), 1.0)
    loss_g = ((rgb_g - cu(target)) ** 2).sum() + (w_g * cu(Gw)).sum()
    loss_g.backward()

    # ---- reference-shaped chain with the oracle on CPU -------------------------------------------------------------
    sdist = O.resample_level(t_prev, w_prev, S, dilation, True, 1.0, 0.0, rand01)
    assert float((sdist_g.cpu() - sdist).abs().max()) < 4e-6
    sdist = sdist_g.cpu()              # continue from the same fenceposts: 4e-6 of s is a tenth of a finest-level cell
    tdist = sdist * far + (1 - sdist) * near
    means, stds, _ = O.cast_rays_det(tdist, batch["origins"], batch["directions"], batch["cam_dirs"], batch["radii"], batch["rand_vec"])
    assert float((means_g.cpu() - means).abs().max()) < 1e-6
    f0, _, _, _ = O.pooled_encode_forward(params, prefix, gs, means, stds)
    feats = f0.clone().requires_grad_(True)
    P = {k: params[prefix + "." + k].clone().requires_grad_(True) for k in MLP_KEYS}
    density, rgbs = mlp_tail(feats, P, batch["viewdirs"], cfg)
    w = O.compute_alpha_weights(density, tdist, batch["directions"])
    acc = w.sum(dim=-1)
    rgb = (w[..., None] * rgbs).sum(dim=-2) + (1 - acc[..., None]).clamp_min(0.) * 1.0
    loss = ((rgb - target) ** 2).sum() + (w * Gw).sum()
    loss.backward()
    emb_grad = O.pooled_encode_backward(params, prefix, gs, means, stds, feats.grad)

    # ---- compare ---------------------------------------------------------------------------------------------------
    assert float((feats_g.detach().cpu() - f0).abs().max()) < 1e-4
    assert float((w_g.detach().cpu() - w.detach()).abs().max()) < 2e-5
    assert float((rgb_g.detach().cpu() - rgb.detach()).abs().max()) < 5e-5
    assert abs(float(loss_g.detach()) - float(loss.detach())) <= 2e-4 * abs(float(loss.detach()))
    for k in MLP_KEYS:
        ref = P[k].grad
        err = float((Pg[k].grad.cpu() - ref).abs().max())
        assert err <= 2e-3 * float(ref.abs().max()) + 1e-7, (k, err, float(ref.abs().max()))
    err = float((enc.embeddings.grad.cpu() - emb_grad).abs().max())
    assert err <= 2e-3 * float(emb_grad.abs().max()), (err, float(emb_grad.abs().max()))
