#!/usr/bin/env python
"""Secondary benchmark (not the driver contract): one optimisation step of the reference's training loop through the
native ops - BASELINE.json configs[4] ("train.py one optimisation step (forward+backward through gridencoder/MLP)
65536-ray batch, 8xB200"), i.e. 8,192 rays per GPU with the waymo.gin shapes (128 proposal + 32 NeRF samples):

    forward   ucnerf_b200.models.Model(rand=True) = train_forward.level_loop: native resampling, cast_rays, pooled
              hash-grid encode and compositing around the model's nn.Linear layers (cuBLAS)
    loss      Charbonnier data loss on both levels (train_utils.compute_data_loss, data_loss_type='charb') + a
              weights-dependent term standing in for the interlevel / distortion losses (they are the reference's own
              Python and not part of this package)
    backward  autograd through the native backward kernels
    exchange  N > 1: all-reduce of the gradients (what DDP does for the reference)
    step      torch.optim.Adam for the nn.Linear layers + the fused hash-decay + Adam + zero_grad kernel for the tables

    python bench_train.py [--steps 10 --warmup 3 --rays 8192]            # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench_train.py --gpus 8

One JSON line from rank 0: rays/s over all ranks, ms per step (CUDA events, max over ranks) and the per-phase times."""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch


def build_model(dev):
    from ucnerf_b200 import synthetic
    from ucnerf_b200.models import Model
    wl = synthetic.WORKLOADS["eval_800x600_waymo_gin"]
    model = Model(config=types.SimpleNamespace(brightness_correction=False, model_sky=False, vis_num_rays=16),
                  num_levels=2, num_prop_samples=wl.num_prop_samples, num_nerf_samples=wl.num_nerf_samples,
                  prop_desired_grid_size=list(wl.prop_desired),
                  nerf_mlp_kwargs=dict(grid_disired_resolution=wl.nerf_desired, bottleneck_width=wl.bottleneck_width,
                                       net_width_viewdirs=wl.net_width_viewdirs,
                                       grid_log2_hashmap_size=wl.log2_hashmap_size),
                  prop_mlp_kwargs=dict(grid_log2_hashmap_size=wl.log2_hashmap_size))
    sd = synthetic.synthetic_state_dict(wl, seed=0)            # same weights on every rank (replicas)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith(".idx") for k in missing), (missing, unexpected)
    return model.to(dev).train(), wl


def make_batch(n_rays, seed, dev):
    from ucnerf_b200 import synthetic
    height = 1 << (max(n_rays, 1).bit_length() - 1) // 2          # 8,192 rays -> a 64 x 128 pixel patch
    rays = synthetic.pinhole_rays(height, max(n_rays // height, 1), seed=seed)
    n = rays["origins"].shape[0]
    g = torch.Generator().manual_seed(seed)
    batch = {k: v.to(dev) for k, v in rays.items() if k != "rand_vec"}
    batch["rgb"] = torch.rand((n, 3), generator=g).to(dev)
    batch["lossmult"] = torch.ones((n, 1), device=dev)
    return batch, n


def compute_loss(batch, renderings, ray_history, charb_padding=0.001, coarse_mult=0.0, interlevel_like=0.01):
    """train_utils.compute_data_loss (L171-230, data_loss_type='charb', data_coarse_loss_mult as given) + a term that puts
    a gradient on the proposal weights like the interlevel loss does."""
    lossmult = torch.broadcast_to(batch['lossmult'], batch['rgb'].shape)
    denom = lossmult.sum()
    data = []
    for r in renderings:
        resid_sq = (r['rgb'] - batch['rgb']) ** 2
        data.append((lossmult * torch.sqrt(resid_sq + charb_padding ** 2)).sum() / denom)
    loss = coarse_mult * sum(data[:-1]) + data[-1]
    for h in ray_history[:-1]:
        s = h['sdist']
        loss = loss + interlevel_like * ((h['weights'] ** 2) / (s[..., 1:] - s[..., :-1]).clamp_min(1e-5)).mean()
    return loss


class Phases:
    """CUDA-event timing of named phases, accumulated over the timed steps (no sync on the step path)."""

    def __init__(self):
        self.events = []

    def mark(self, name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.events.append((name, e))

    def totals(self):
        out = {}
        for (n0, e0), (n1, e1) in zip(self.events[:-1], self.events[1:]):
            if n1 != "start":
                out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        return out


class SkyNet(torch.nn.Module):
    """Module with the reference NeRF's attributes / parameter names (models.py:L743-795; D = 8, W = 256, skips = [4],
    4-frequency view embedding) holding synthetic weights - the sky head `Model` builds with `Config.model_sky`."""

    def __init__(self, heads):
        super().__init__()
        L = torch.nn.Linear
        self.skips, self.use_viewdirs, self.embed_fn = [4], True, None
        self.pts_linears = torch.nn.ModuleList([L(3, 256)] + [L(259 if i == 4 else 256, 256) for i in range(7)])
        self.views_linears = torch.nn.ModuleList([L(283, 128)])
        self.feature_linear, self.alpha_linear, self.rgb_linear = L(256, 256), L(256, 1), L(128, 3)
        self.load_state_dict({k[len("skynerf."):]: v for k, v in heads.items() if k.startswith("skynerf.")})

    @staticmethod
    def embed_fn_view(v):            # get_embedder(4) (models.py:L689-727): [x, sin(x f), cos(x f)] for f = 1, 2, 4, 8
        out = [v]
        for f in (1.0, 2.0, 4.0, 8.0):
            out += [torch.sin(v * f), torch.cos(v * f)]
        return torch.cat(out, -1)

    def forward(self, pts, views):   # NeRF.forward (models.py:L797-820) with nn.Linear: the cuBLAS arm of the A/B
        v = self.embed_fn_view(views)
        h = pts
        for i, l in enumerate(self.pts_linears):
            h = torch.relu(l(h))
            if i in self.skips:
                h = torch.cat([pts, h], -1)
        alpha, feature = self.alpha_linear(h), self.feature_linear(h)
        h = torch.relu(self.views_linears[0](torch.cat([feature, v], -1)))
        return alpha, self.rgb_linear(h)


def sky_torch(ray_batch, net, n_samples=120):
    """models.render_rays + raw2outputs (models.py:L822-904) in plain torch on `net.forward` (reference semantics)."""
    o, d, near, far, views = ray_batch[:, 0:3], ray_batch[:, 3:6], ray_batch[:, 6:7], ray_batch[:, 7:8], ray_batch[:, -3:]
    t = torch.linspace(0., 1., steps=n_samples, device=o.device)
    z = near * (1. - t) + 1. / far * t
    pts = o[..., None, :] + d[..., None, :] * z[..., :, None]
    alpha, rgb = net(pts, views.unsqueeze(1).expand(-1, n_samples, -1))
    dists = torch.cat([z[..., 1:] - z[..., :-1], torch.full_like(z[..., :1], 1e10)], -1) * torch.norm(d[..., None, :], dim=-1)
    a = 1. - torch.exp(-torch.relu(alpha[..., 0]) * dists)
    w = a * torch.cumprod(torch.cat([torch.ones_like(a[:, :1]), 1. - a + 1e-10], -1), -1)[:, :-1]
    return {"rgb_map": torch.sum(w[..., None] * torch.sigmoid(rgb), -2)}


def run(dev, world, rank, steps=10, warmup=3, rays=8192, merge_runs="auto", native_mlp="auto", sky=False):
    """One training step of config 5 timed on `dev` (process group already initialised when world > 1) -> result dict."""
    import torch.distributed as dist
    from ucnerf_b200 import _lib
    from ucnerf_b200.gridencoder.optim import GridAdam
    from ucnerf_b200.parallel_train import OverlappedGradientExchange
    from ucnerf_b200.train_forward import level_loop
    model, wl = build_model(dev)
    batch, n = make_batch(rays, seed=rank, dev=dev)
    encoders = [m.encoder for m in (model.prop_mlp_0, model.nerf_mlp)]
    table_ids = {id(e.embeddings) for e in encoders}
    dense = [p for p in model.parameters() if id(p) not in table_ids]
    skynet = None
    if sky:     # scripts/train_waymo.sh: model_sky = True (models.py:L84-92, L326-337): + 120 samples x 8x256 MLP per ray
        from ucnerf_b200 import synthetic
        skynet = SkyNet(synthetic.synthetic_heads(seed=0)).to(dev)
        dense += list(skynet.parameters())
        far = batch["far"].reshape(-1, 1)
        sky_batch = torch.cat([batch["origins"], batch["directions"], far, torch.full_like(far, float(far[0]) * 1.5),
                               batch["cam_dirs"]], -1)
    opt = torch.optim.Adam(dense, lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    grid_opt = GridAdam(encoders, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, hash_decay_mult=0.1, zero_grad=True)
    for e in encoders:                                   # the fused step zeroes these in place every step
        e.embeddings.grad = torch.zeros_like(e.embeddings)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    exchange = OverlappedGradientExchange(dense, [e.embeddings for e in encoders])     # table all-reduces start inside backward
    ph = Phases()
    mr = {"auto": "auto", "off": False, "interval": True, "ray": "ray"}[merge_runs]
    kw = {} if native_mlp == "auto" else {"native_mlp": native_mlp == "on"}

    def step(timed):
        if timed:
            ph.mark("start")
        opt.zero_grad(set_to_none=True)
        renderings, ray_history = level_loop(model, True, batch, 0.5, compute_extras=False, hash_decay='fused', generator=gen,
                                              merge_runs=mr, **kw)
        if skynet is not None:      # models.py:L336-337, L351-354 (without the brightness affines)
            if native_mlp == "off":
                sky_rgb = sky_torch(sky_batch, skynet)["rgb_map"]
            else:
                from ucnerf_b200.sky_train import sky_render_rays
                sky_rgb = sky_render_rays(sky_batch, skynet)["rgb_map"]
            last = renderings[-1]
            last["rgb"] = last["rgb"] + (1 - last["weights"].sum(-1, keepdim=True)) * sky_rgb
        loss = compute_loss(batch, renderings, ray_history)
        if timed:
            ph.mark("forward")
        loss.backward()
        if timed:
            ph.mark("backward")
        if world > 1:
            exchange.finish()
            if timed:
                ph.mark("exchange")
        opt.step()
        grid_opt.step()
        if timed:
            ph.mark("optimizer")
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step(False)
    sync_all()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step(True)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    phases = {k: v / steps for k, v in ph.totals().items()}
    grad_bytes = sum(p.numel() for p in dense) * 4 + sum(e.embeddings.numel() for e in encoders) * 4
    res = {"bench": "train_step (BASELINE.json configs[4])", "metric": "train_rays_per_sec",
           "value": world * n * steps / (ms * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": steps,
           "warmup": warmup, "ms_per_step": ms / steps, "rays_per_gpu_per_step": n, "samples_per_ray": wl.samples_per_ray,
           "scaling": "weak", "dtype": "f32", "pooled_backward": merge_runs, "data": "synthetic", "phase_ms_per_step": phases,
           "native_launches_per_step": (_lib.launch_count() - launches0) / steps, "final_loss": float(loss.detach()),
           "gradient_allreduce_bytes_per_step": grad_bytes if world > 1 else 0, "sky_head": bool(sky),
           "dense_layers": "cuBLAS fp32 (nn.Linear)" if native_mlp == "off" else "tcgen05 3xTF32 (gemm.tc_linear)",
           "exchange": "table all-reduces (NCCL AVG) start from post-accumulate-grad hooks during backward; `exchange` = what is left",
           "note": "forward / backward through ucnerf_b200.train_forward.level_loop (native resample, cast_rays, pooled "
                   "encode, composite and MLP kernels), torch Adam for the dense layers, fused hash-decay + Adam + zero_grad "
                   "for the tables; N > 1 adds the gradient all-reduce DDP would do"}
    exchange.close()
    del model, opt, grid_opt
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rays", type=int, default=8192, help="rays per GPU per step (config 5: 65,536 / 8)")
    ap.add_argument("--merge-runs", default="auto", choices=["auto", "off", "interval", "ray"], help="pooled-encode backward variant")
    ap.add_argument("--native-mlp", default="auto", choices=["auto", "on", "off"], help="native MLP kernels vs nn.Linear (cuBLAS)")
    ap.add_argument("--sky", action="store_true", help="add the sky head (model_sky=True: 120 samples x 8x256 MLP per ray)")
    a = ap.parse_args()
    import torch.distributed as dist
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        sys.exit("bench_train.py needs a CUDA device: the training ops have no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    res = run(dev, world, rank, a.steps, a.warmup, a.rays, a.merge_runs, a.native_mlp, a.sky)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
