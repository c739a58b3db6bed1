"""The level loop of `Model.forward` for the TRAINING step (internal/models.py:L126-311, `rand=True` or False) through the
native ops of this package, returning the reference's `(renderings, ray_history)` for train.py:L166-216.

    renderings, ray_history = level_loop(model, rand, batch, train_frac, compute_extras=False)

`model` is the reference's own `Model` (or any module tree with the same attribute / parameter names): its
hyper-parameters are read as attributes, the weights of its `nn.Linear` layers feed fp32-accurate tensor-core GEMMs
(gemm.tc_linear; `native_mlp=False` calls the modules themselves, i.e. cuBLAS), and everything between them runs in
libucnerf_b200.so as well:

    resampling incl. dilation + jitter       stepfun.resample_level        (models.py:L156-205)
    render.cast_rays, rand pattern            render_train.cast_rays        (L208-217, render.py:L94-152)
    contraction + hash-grid encode + pooling  gridencoder.pooled            (L485-496, forward and backward)
    alpha weights + acc / rgb                 render_train.composite        (L229-283, render.py:L155-205, fwd and bwd)

The sky / brightness heads that follow the loop in Model.forward (L326-363) are not part of it: a maintainer replaces
the loop and keeps those lines (INTEGRATION.md seam 10).  Supported configuration = configs/waymo.gin / nuscenes.gin:
`raydist_fn=None`, `near_anneal_rate=None`, constant background, no GLO, `disable_density_normals=True`, no
reflections / diffuse / tint / roughness / n_dot_v, `warp_fn='contract'`, `scale_featurization=False`,
`opaque_background=False`, `compute_extras=False`; anything else raises NotImplementedError instead of diverging
silently.  There is no CPU path."""
import torch
import torch.nn.functional as F

from .gridencoder.pooled import pooled_encode
from .render_train import cast_rays, composite
from .stepfun import resample_level


class _GradientScaler(torch.autograd.Function):
    """train_utils.GradientScaler (train_utils.py:L101-111): identity forward, gradients x clamp(ray_dist^2, 0, 1)."""

    @staticmethod
    def forward(ctx, colors, sigmas, ray_dist):
        ctx.save_for_backward(ray_dist)
        return colors, sigmas

    @staticmethod
    def backward(ctx, g_colors, g_sigmas):
        (ray_dist,) = ctx.saved_tensors
        scaling = torch.square(ray_dist).clamp(0, 1)
        return g_colors * scaling[..., None], g_sigmas * scaling, None


def _unsupported(cond, what):
    if cond:
        raise NotImplementedError(f"level_loop: {what} is not supported on the native training path")


def _check_model(model, batch, compute_extras):
    _unsupported(getattr(model, "raydist_fn", None) is not None, "raydist_fn")
    _unsupported(getattr(model, "near_anneal_rate", None) is not None, "near_anneal_rate")
    _unsupported(getattr(model, "num_glo_features", 0) > 0, "GLO features")
    _unsupported(getattr(model, "opaque_background", False), "opaque_background")
    _unsupported(getattr(model, "learned_exposure_scaling", False) or batch.get('exposure_idx') is not None, "exposure scaling")
    _unsupported(not getattr(model, "use_viewdirs", True), "use_viewdirs=False")
    _unsupported(not getattr(model, "stop_level_grad", True), "stop_level_grad=False")
    bg = getattr(model, "bg_intensity_range", (1., 1.))
    _unsupported(bg[0] != bg[1], "a random background colour")
    _unsupported(compute_extras, "compute_extras=True")


def _check_mlp(mlp):
    for flag in ("use_reflections", "use_directional_enc", "enable_pred_roughness", "use_diffuse_color",
                 "use_specular_tint", "use_n_dot_v", "enable_pred_normals", "scale_featurization"):
        _unsupported(getattr(mlp, flag, False), flag)
    _unsupported(not getattr(mlp, "disable_density_normals", False), "density normals (set disable_density_normals=True)")
    _unsupported(getattr(mlp, "warp_fn", "contract") != "contract", "warp_fn != 'contract'")
    _unsupported(getattr(mlp, "num_glo_features", 0) > 0, "GLO features")
    _unsupported(getattr(mlp, "bottleneck_width", 256) <= 0 and not getattr(mlp, "disable_rgb", False), "bottleneck_width=0")


def _pos_enc(x, max_deg):
    """coord.pos_enc(x, 0, max_deg, append_identity=True) (coord.py:L214-225)."""
    scales = 2.0 ** torch.arange(0, max_deg, device=x.device, dtype=x.dtype)
    xb = (x[..., None, :] * scales[:, None]).reshape(x.shape[:-1] + (-1,))
    return torch.cat([x, torch.sin(torch.cat([xb, xb + 0.5 * torch.pi], dim=-1))], dim=-1)


def _native_mlp_ok(mlp):
    """The tensor-core layers (ucnerf_b200.gemm.tc_linear) cover widths up to 256 and the default layer layout."""
    if mlp.density_layer[0].out_features > 256 or mlp.density_layer[2].out_features > 256:
        return False
    if getattr(mlp, "disable_rgb", False):
        return True
    return (mlp.net_depth_viewdirs == 2 and mlp.skip_layer_dir == 0 and mlp.net_width_viewdirs <= 256
            and mlp.bottleneck_width <= 256 and 3 + 6 * mlp.deg_view <= 256)


def _mlp_forward_native(mlp, rand, means, stds, viewdirs, merge_runs=True):
    """MLP.forward (models.py:L514-685) with every dense layer on the tensor cores (gemm.tc_linear: 3xTF32, fp32 accuracy,
    forward and backward) instead of nn.Linear's cuBLAS SGEMMs; the reference's `torch.cat` copies become K segments."""
    from .gemm import tc_linear
    features, coord = pooled_encode(mlp.encoder, means, stds, merge_runs=merge_runs)   # L487-496, L512
    d0, d2 = mlp.density_layer[0], mlp.density_layer[2]
    x = tc_linear([tc_linear([features], d0.weight, d0.bias, relu=True)], d2.weight, d2.bias)   # L507
    raw_density = x[..., 0]                                                       # L508
    if rand and getattr(mlp, "density_noise", 0.) > 0:                            # L510-511
        raw_density = raw_density + mlp.density_noise * torch.randn_like(raw_density)
    density = F.softplus(raw_density + mlp.density_bias)                          # L581
    if getattr(mlp, "disable_rgb", False):
        return dict(coord=coord, density=density, rgb=None)
    bottleneck = x                                                                # L601
    if rand and getattr(mlp, "bottleneck_noise", 0.) > 0:                         # L604-605
        bottleneck = bottleneck + mlp.bottleneck_noise * torch.randn_like(bottleneck)
    dir_enc = _pos_enc(viewdirs, mlp.deg_view)                                    # L620-627
    dir_enc = torch.broadcast_to(dir_enc[..., None, :], bottleneck.shape[:-1] + (dir_enc.shape[-1],)).contiguous()
    l0, l1 = mlp.lin_second_stage_0, mlp.lin_second_stage_1
    y0 = tc_linear([bottleneck, dir_enc], l0.weight, l0.bias, relu=True)          # L643-647, i = 0 (then cat with inputs)
    y1 = tc_linear([y0, bottleneck, dir_enc], l1.weight, l1.bias, relu=True)      # i = 1
    rgb = torch.sigmoid(mlp.rgb_premultiplier * tc_linear([y1], mlp.rgb_layer.weight, mlp.rgb_layer.bias) + mlp.rgb_bias)
    rgb = rgb * (1 + 2 * mlp.rgb_padding) - mlp.rgb_padding                       # L665
    return dict(coord=coord, density=density, rgb=rgb)


def _mlp_forward(mlp, rand, means, stds, viewdirs, merge_runs=True, native_mlp=True):
    """MLP.forward (models.py:L514-685) for the supported configuration; predict_density's front end is the fused op."""
    if native_mlp and _native_mlp_ok(mlp):
        return _mlp_forward_native(mlp, rand, means, stds, viewdirs, merge_runs)
    features, coord = pooled_encode(mlp.encoder, means, stds, merge_runs=merge_runs)   # L487-496, L512
    x = mlp.density_layer(features)                                               # L507
    raw_density = x[..., 0]                                                       # L508
    if rand and getattr(mlp, "density_noise", 0.) > 0:                            # L510-511
        raw_density = raw_density + mlp.density_noise * torch.randn_like(raw_density)
    density = F.softplus(raw_density + mlp.density_bias)                          # L581
    if getattr(mlp, "disable_rgb", False):
        return dict(coord=coord, density=density, rgb=None)                       # L584-585: zeros, never differentiated
    bottleneck = x                                                                # L601
    if rand and getattr(mlp, "bottleneck_noise", 0.) > 0:                         # L604-605
        bottleneck = bottleneck + mlp.bottleneck_noise * torch.randn_like(bottleneck)
    dir_enc = _pos_enc(viewdirs, mlp.deg_view)                                    # L620-627
    dir_enc = torch.broadcast_to(dir_enc[..., None, :], bottleneck.shape[:-1] + (dir_enc.shape[-1],))
    x = torch.cat([bottleneck, dir_enc], dim=-1)
    inputs = x
    for i in range(mlp.net_depth_viewdirs):                                       # L643-647
        x = F.relu(mlp.get_submodule(f"lin_second_stage_{i}")(x))
        if i == mlp.skip_layer_dir:
            x = torch.cat([x, inputs], dim=-1)
    rgb = torch.sigmoid(mlp.rgb_premultiplier * mlp.rgb_layer(x) + mlp.rgb_bias)  # L650-652
    rgb = rgb * (1 + 2 * mlp.rgb_padding) - mlp.rgb_padding                       # L665
    return dict(coord=coord, density=density, rgb=rgb)


def _hash_decay(encoder):
    """models.py:L297-306 without torch_scatter: mean over levels and channels of the per-level mean of param^2."""
    param, idx = encoder.embeddings, encoder.idx
    n = int(encoder.offsets.shape[0] - 1)
    sums = torch.zeros((n, param.shape[-1]), device=param.device, dtype=param.dtype).index_add_(0, idx, param ** 2)
    counts = (encoder.offsets[1:] - encoder.offsets[:-1]).to(param.dtype)
    return (sums / counts[:, None]).mean()


def level_loop(model, rand, batch, train_frac, compute_extras=False, hash_decay=True, generator=None, draws=None,
               merge_runs='auto', native_mlp=True):
    """-> (renderings, ray_history), the lists Model.forward builds in its level loop.  `draws` (optional): one dict per
    level with the uniform / normal draws `jitter01`, `flip01`, `rot01`, `rand_vec` (testing / reproducibility); else
    they are drawn on the device with `generator`, in the reference's order.  `merge_runs`: backward variant of the
    pooled encode (False / True / 'ray', see gridencoder.pooled.pooled_encode); 'auto' = the measured best per level:
    'ray' on the proposal levels, True on the NeRF level.  `hash_decay`: True = the reference's differentiable term
    (torch), 'fused' = its value from one native read pass without autograd (pair it with GridAdam(hash_decay_mult=...),
    which applies the term's gradient inside the optimiser kernel), False = omit.  `native_mlp`: dense layers on the
    tensor cores (gemm.tc_linear, forward + backward) instead of the module's nn.Linear / cuBLAS calls."""
    _check_model(model, batch, compute_extras)
    lead = batch['origins'].shape[:-1]
    flat = lambda k, c: batch[k].reshape(-1, c)
    origins, directions, viewdirs, cam_dirs = flat('origins', 3), flat('directions', 3), flat('viewdirs', 3), flat('cam_dirs', 3)
    radii, near, far = flat('radii', 1), flat('near', 1), flat('far', 1)
    N = origins.shape[0]
    dev = origins.device
    sdist = torch.cat([torch.zeros_like(near), torch.ones_like(far)], dim=-1)     # L143-146 with init_s_near = 0
    weights = torch.ones_like(near)                                               # L147
    prod_num_samples = 1
    use_dilation = model.dilation_bias > 0 or model.dilation_multiplier > 0        # L166
    bg = float(model.bg_intensity_range[0])
    scale_grads = bool(getattr(getattr(model, "config", None), "brightness_correction", False))
    renderings, ray_history = [], []
    for i_level in range(model.num_levels):
        is_prop = i_level < model.num_levels - 1
        num_samples = model.num_prop_samples if is_prop else model.num_nerf_samples
        dilation = model.dilation_bias + model.dilation_multiplier * (1.0 - 0.0) / prod_num_samples      # L158-159
        prod_num_samples *= num_samples
        if model.anneal_slope > 0:                                                # L179-184
            anneal = (model.anneal_slope * train_frac) / ((model.anneal_slope - 1) * train_frac + 1)
        else:
            anneal = 1.
        d = draws[i_level] if draws is not None else {}
        sdist = resample_level(sdist, weights, num_samples, dilation, i_level > 0 and use_dilation, anneal,
                               model.resample_padding, rand=bool(rand), single_jitter=model.single_jitter,
                               generator=generator, rand01=d.get("jitter01"))     # L165-204 (detached)
        tdist = sdist * far + (1 - sdist) * near                                  # L207, coord.py:L176 with fn = None
        cast_draws = (d.get("flip01"), d.get("rot01"), d["rand_vec"]) if "rand_vec" in d else None
        means, stds, ts = cast_rays(tdist, origins, directions, cam_dirs, radii, bool(rand), std_scale=model.std_scale,
                                    generator=generator, draws=cast_draws)        # L210-217
        if is_prop:                                                               # L220-221
            mlp = model.get_submodule(f'prop_mlp_{i_level}') if getattr(model, "distinct_prop", True) else model.prop_mlp
            if getattr(model, "single_mlp", False):
                mlp = model.nerf_mlp
        else:
            mlp = model.nerf_mlp
        _check_mlp(mlp)
        mr = ('ray' if is_prop else True) if merge_runs == 'auto' else merge_runs
        ray_results = _mlp_forward(mlp, rand, means, stds, viewdirs, mr, native_mlp)   # L222-229
        density, rgbs = ray_results['density'], ray_results['rgb']
        if scale_grads:                                                           # L232-234
            if rgbs is None:
                _, density = _GradientScaler.apply(density.new_zeros(density.shape + (3,)), density, ts.mean(dim=-1))
            else:
                rgbs, density = _GradientScaler.apply(rgbs, density, ts.mean(dim=-1))
        weights, rgb, acc = composite(density, rgbs, tdist, directions, bg)       # L237-283 (alpha weights, acc, rgb)
        with torch.no_grad():                                                     # render.py:L206-214 (no loss uses depth)
            t_mids = 0.5 * (tdist[..., :-1] + tdist[..., 1:])
            eps = torch.finfo(torch.float32).eps
            depth = torch.clip(torch.nan_to_num((weights * t_mids).sum(dim=-1) / acc.clamp_min(eps), torch.inf),
                               tdist[..., 0], tdist[..., -1])
            depth[acc < 0.6] = 300
        rendering = dict(rgb=rgb.reshape(lead + (3,)), depth=depth.reshape(lead), acc=acc.reshape(lead),
                         weights=weights.reshape(lead + (num_samples,)))          # L283-284
        S = num_samples
        ray_results['rgb'] = (torch.zeros((N, S, 3), device=dev) if rgbs is None else rgbs).reshape(lead + (S, 3))
        ray_results['density'] = density.reshape(lead + (S,))
        ray_results['coord'] = ray_results['coord'].reshape(lead + (S, 3))
        for k in ('raw_grad_density', 'grad_pred', 'normals', 'normals_pred', 'roughness'):      # L676-685
            ray_results[k] = None
        if hash_decay and getattr(model, "training", False):                      # L297-306
            if hash_decay == 'fused':   # value only (one native read pass); its gradient lives in GridAdam(hash_decay_mult)
                from .gridencoder.optim import hash_decay_loss
                ray_results['loss_hash_decay'] = hash_decay_loss(mlp.encoder)
            else:
                ray_results['loss_hash_decay'] = _hash_decay(mlp.encoder)
        renderings.append(rendering)
        ray_results['sdist'] = sdist.reshape(lead + (S + 1,)).clone()             # L309-311
        ray_results['weights'] = rendering['weights'].clone()
        ray_history.append(ray_results)
    return renderings, ray_history
