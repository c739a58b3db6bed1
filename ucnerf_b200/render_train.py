"""Differentiable alpha compositing for the training step: `render.compute_alpha_weights` and the acc / rgb lines of
`render.volumetric_rendering` (internal/render.py:L155-174, L202-205) as one autograd Function over
`ucnerf_composite_train_forward/backward` (csrc/composite_train.cu).

    weights, rgb, acc = composite(density, rgbs, tdist, dirs, bg=1.0)     # rgbs may be None on proposal levels

`weights` feeds the interlevel / distortion losses and the next level's resampling, `rgb` the data loss, `acc` the sky
/ opacity losses - all three are differentiable with respect to `density` and `rgbs`.  `tdist` and `dirs` carry no
gradient (sdist is detached by `stop_level_grad`, models.py:L203-204; rays are data).  Constant background only
(`bg_intensity_range` with equal ends, models.py:L244-246), `opaque_background=False` as in configs/*.gin.
There is no CPU path: CPU tensors raise."""
import torch
from torch.autograd import Function

from . import _lib


def _ptr(t):
    return None if t is None else t.data_ptr()


class _composite(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, density, rgbs, tdist, dirs, bg):
        if density.device.type != "cuda":
            raise RuntimeError("composite: tensors must be CUDA tensors (no CPU path)")
        N, S = density.shape
        if tuple(tdist.shape) != (N, S + 1) or tuple(dirs.shape) != (N, 3) or (rgbs is not None and tuple(rgbs.shape) != (N, S, 3)):
            raise RuntimeError("composite: expected density [N,S], rgbs [N,S,3] or None, tdist [N,S+1], dirs [N,3]")
        d = density.detach().contiguous().float()
        c = None if rgbs is None else rgbs.detach().contiguous().float()
        t = tdist.detach().contiguous().float()
        dr = dirs.detach().contiguous().float()
        weights = torch.empty((N, S), device=d.device, dtype=torch.float32)
        rgb = torch.empty((N, 3), device=d.device, dtype=torch.float32)
        acc = torch.empty((N,), device=d.device, dtype=torch.float32)
        lib = _lib.load()
        with torch.cuda.device(d.device):
            rc = lib.ucnerf_composite_train_forward(t.data_ptr(), d.data_ptr(), _ptr(c), dr.data_ptr(), N, S, float(bg),
                                                    weights.data_ptr(), rgb.data_ptr(), acc.data_ptr(),
                                                    torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "composite_train_forward")
        ctx.set_materialize_grads(False)          # unused outputs arrive as None, not as zero tensors
        ctx.save_for_backward(t, d, c, dr, weights, acc)
        ctx.bg = float(bg)
        ctx.has_rgb = c is not None
        return weights, rgb, acc

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g_weights, g_rgb, g_acc):
        t, d, c, dr, weights, acc = ctx.saved_tensors
        N, S = d.shape
        gw = None if g_weights is None else g_weights.contiguous().float()
        gr = None if g_rgb is None else g_rgb.contiguous().float()
        ga = None if g_acc is None else g_acc.contiguous().float()
        d_density = torch.empty_like(d)
        d_rgbs = torch.empty_like(c) if ctx.has_rgb and ctx.needs_input_grad[1] else None
        lib = _lib.load()
        with torch.cuda.device(d.device):
            rc = lib.ucnerf_composite_train_backward(t.data_ptr(), d.data_ptr(), _ptr(c), dr.data_ptr(), weights.data_ptr(),
                                                     acc.data_ptr(), _ptr(gw), _ptr(gr), _ptr(ga), N, S, ctx.bg,
                                                     d_density.data_ptr(), _ptr(d_rgbs),
                                                     torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "composite_train_backward")
        return d_density, d_rgbs, None, None, None


def composite(density, rgbs, tdist, dirs, bg=1.0):
    """density [N,S], rgbs [N,S,3] or None, tdist [N,S+1], dirs [N,3] -> (weights [N,S], rgb [N,3], acc [N])."""
    return _composite.apply(density, rgbs, tdist, dirs, bg)


@torch.no_grad()
def cast_rays(tdist, origins, directions, cam_dirs, radii, rand=True, n=7, m=3, std_scale=0.5, generator=None, draws=None,
              **kwargs):
    """Drop-in for the reference's `render.cast_rays` (internal/render.py:L94-152, same positional arguments and return
    values): tdist [N,S+1] + rays -> (means [N,S,6,3], stds [N,S,6], t [N,S,6]) through `ucnerf_cast_rays`, one kernel.
    The random numbers are drawn with torch on the device in the reference's order - flip mask (L121), rotation (L122),
    `rand_vec` (L140) - so a seeded generator reproduces the reference's stream; `draws = (flip01, rot01, rand_vec)`
    supplies them explicitly.  No gradient flows (tdist is detached by the caller, models.py:L203-204)."""
    if tdist.device.type != "cuda":
        raise RuntimeError("cast_rays: tensors must be CUDA tensors (no CPU path)")
    if tdist.dim() != 2:
        raise RuntimeError("cast_rays: expected tdist [N, S+1] (flatten the leading dimensions first)")
    N, S = tdist.shape[0], tdist.shape[1] - 1
    dev = tdist.device
    f = lambda x, cols: x.detach().reshape(N, cols).contiguous().float()
    t = tdist.detach().contiguous().float()
    o, d, c, r = f(origins, 3), f(directions, 3), f(cam_dirs, 3), f(radii, 1)
    if draws is not None:
        flip01, rot01, rand_vec = draws
    else:
        flip01 = torch.rand((N, S), device=dev, generator=generator) if rand else None
        rot01 = torch.rand((N, S), device=dev, generator=generator) if rand else None
        rand_vec = torch.randn((N, 3), device=dev, generator=generator)
    if rand and (flip01 is None or rot01 is None):
        raise RuntimeError("cast_rays: rand=True needs the flip and rotation draws")
    flip01 = f(flip01, S) if rand else None
    rot01 = f(rot01, S) if rand else None
    rand_vec = f(rand_vec, 3)
    means = torch.empty((N, S, 6, 3), device=dev, dtype=torch.float32)
    stds = torch.empty((N, S, 6), device=dev, dtype=torch.float32)
    ts = torch.empty((N, S, 6), device=dev, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(dev):
        rc = lib.ucnerf_cast_rays(t.data_ptr(), o.data_ptr(), d.data_ptr(), c.data_ptr(), r.data_ptr(), rand_vec.data_ptr(),
                                  _ptr(rot01), _ptr(flip01), N, S, float(std_scale), means.data_ptr(), stds.data_ptr(),
                                  ts.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "cast_rays")
    return means, stds, ts
