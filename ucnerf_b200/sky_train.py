"""The sky head under TRAINING (`Config.model_sky = True`, internal/models.py:L326-337): the reference renders a second,
classic NeRF behind every ray - `render_rays(ray_batch, network_fn=self.skynerf)` (L849-904) with 120 samples through
the 8 x 256 MLP `NeRF.forward` (L797-820) and `raw2outputs` (L822-847) - as fp32 `nn.Linear` layers under autograd:
562,688 MAC per sample forward, about 2.5 x the rest of the optimisation step in FLOPs.  Here every dense layer of that
MLP runs on the tensor cores, forward and backward, through `gemm.tc_linear` (3xTF32 on tcgen05, fp32 accuracy); the
skip connection `cat([pts, h])` and the view branch `cat([feature, emb(view)])` are K segments instead of copies.

    ret = sky_render_rays(ray_batch, model.skynerf)          # drop-in for models.render_rays in Model.forward (seam 11)

Bug-compatible with the reference (decreasing sample depths `near (1 - t) + t / far`, the 1e10 last interval), like the
eval kernel (csrc/sky_mlp_tc.cu).  The eval path keeps using that fused forward-only kernel."""
import torch

from .gemm import tc_linear


def sky_nerf_forward(net, pts, views):
    """NeRF.forward (models.py:L797-820) of the module `Model` builds (L84-92: D = 8, W = 256, raw xyz input, skips = [4],
    4-frequency view embedding); `net` is the reference module itself (its parameters and embedders are used as they are)."""
    if not getattr(net, "use_viewdirs", True):
        raise NotImplementedError("sky_nerf_forward: use_viewdirs=False is not supported")
    lead = pts.shape[:-1]
    x = pts.reshape(-1, pts.shape[-1])
    if net.embed_fn is not None:
        x = net.embed_fn(x)
    v = views.reshape(-1, views.shape[-1])
    if net.embed_fn_view is not None:
        v = net.embed_fn_view(v)
    x, v = x.contiguous(), v.contiguous()
    h = [x]                                             # the current activation as K segments
    for i, lin in enumerate(net.pts_linears):           # L802-806
        y = tc_linear(h, lin.weight, lin.bias, relu=True)
        h = [x, y] if i in net.skips else [y]           # torch.cat([input_pts, h], -1)
    alpha = tc_linear(h, net.alpha_linear.weight, net.alpha_linear.bias)            # L809
    feature = tc_linear(h, net.feature_linear.weight, net.feature_linear.bias)      # L810
    hv = [feature, v]                                   # torch.cat([feature, input_views], -1)
    for lin in net.views_linears:                       # L812-814
        hv = [tc_linear(hv, lin.weight, lin.bias, relu=True)]
    rgb = tc_linear(hv, net.rgb_linear.weight, net.rgb_linear.bias)                 # L816
    return alpha.reshape(*lead, 1), rgb.reshape(*lead, 3)


def sky_render_rays(ray_batch, network_fn, N_samples=120, white_bkgd=False):
    """models.render_rays (L849-904) + raw2outputs (L822-847) for the arguments Model.forward uses (no perturbation, no
    noise, no importance sampling): ray_batch [N, 11] = origins | directions | near | far | view -> dict(rgb_map,
    depth_map, acc_map)."""
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:]
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    t_vals = torch.linspace(0., 1., steps=N_samples, device=ray_batch.device)
    z_vals = near * (1. - t_vals) + 1. / far * t_vals                                       # L872
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    alpha, rgb = sky_nerf_forward(network_fn, pts, viewdirs.unsqueeze(1).expand(-1, N_samples, -1))
    dists = z_vals[..., 1:] - z_vals[..., :-1]                                              # raw2outputs
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e10)], -1) * torch.norm(rays_d[..., None, :], dim=-1)
    a = 1. - torch.exp(-torch.relu(alpha[..., 0]) * dists)
    weights = a * torch.cumprod(torch.cat([torch.ones_like(a[:, :1]), 1. - a + 1e-10], -1), -1)[:, :-1]
    rgb_map = torch.sum(weights[..., None] * torch.sigmoid(rgb), -2)
    acc_map = torch.sum(weights, -1)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])
    return {"rgb_map": rgb_map, "depth_map": torch.sum(weights * z_vals, -1), "acc_map": acc_map}
