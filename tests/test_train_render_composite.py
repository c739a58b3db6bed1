"""Differentiable compositing for the training step (ucnerf_composite_train_forward/backward,
ucnerf_b200.render_train.composite): render.py:L155-174 + L202-205 under autograd.  The checker is torch autograd over the
oracle's restatement of those lines (oracle/ucnerf_oracle.py, pinned bit-for-bit to the reference's forward by
tests/test_oracle_golden.py).  CPU: the per-ray host+device functions of train_algos.cuh - the code the CUDA kernels run
one thread per ray - through tests/cpu_harness.cpp.  GPU: the CUDA op through the C ABI / autograd Function."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import ucnerf_oracle as O

cf = ctypes.c_float


def _fp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def make_inputs(N, S, seed, with_rgb=True, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    t = torch.sort(torch.rand((N, S + 1), generator=g) * 8.0, dim=-1).values
    density = torch.rand((N, S), generator=g) ** 3 * 6.0 * scale
    density[: N // 8] = 0.0                                   # empty rays: acc = 0, background only
    density[N // 8: N // 4] *= 200.0                          # opaque rays: transmittance underflows along the ray
    rgbs = torch.rand((N, S, 3), generator=g) if with_rgb else None
    dirs = torch.randn((N, 3), generator=g) * 1.7             # directions are not unit norm
    gw = torch.randn((N, S), generator=g)
    grgb = torch.randn((N, 3), generator=g)
    gacc = torch.randn((N,), generator=g)
    return t, density, rgbs, dirs, gw, grgb, gacc


def oracle_autograd(t, density, rgbs, dirs, gw, grgb, gacc, bg=1.0):
    d = density.clone().requires_grad_(True)
    c = None if rgbs is None else rgbs.clone().requires_grad_(True)
    w = O.compute_alpha_weights(d, t, dirs)                                   # render.py:L155-174
    acc = w.sum(dim=-1)                                                       # render.py:L203
    bg_w = (1 - acc[..., None]).clamp_min(0.)                                 # L204
    cc = c if c is not None else torch.zeros(w.shape + (3,))                  # models.py:L584-585 (disable_rgb)
    rgb = (w[..., None] * cc).sum(dim=-2) + bg_w * bg                         # L205
    loss = 0.
    if gw is not None:
        loss = loss + (w * gw).sum()
    if grgb is not None:
        loss = loss + (rgb * grgb).sum()
    if gacc is not None:
        loss = loss + (acc * gacc).sum()
    loss.backward()
    return w.detach(), rgb.detach(), acc.detach(), d.grad, None if c is None else c.grad


def harness_run(h, t, density, rgbs, dirs, gw, grgb, gacc, bg=1.0):
    N, S = density.shape
    arr = lambda x: None if x is None else np.ascontiguousarray(x.numpy(), np.float32)
    tn, dn, cn, rn = arr(t), arr(density), arr(rgbs), arr(dirs)
    w = np.zeros((N, S), np.float32); rgb = np.zeros((N, 3), np.float32); acc = np.zeros((N,), np.float32)
    h.h_composite_train_forward(N, S, _fp(tn), _fp(dn), _fp(cn), _fp(rn), cf(bg), _fp(w), _fp(rgb), _fp(acc))
    dd = np.zeros((N, S), np.float32)
    dc = None if rgbs is None else np.zeros((N, S, 3), np.float32)
    h.h_composite_train_backward(N, S, _fp(tn), _fp(dn), _fp(cn), _fp(rn), cf(bg), _fp(w), _fp(acc), _fp(arr(gw)),
                                 _fp(arr(grgb)), _fp(arr(gacc)), _fp(dd), _fp(dc))
    return w, rgb, acc, dd, dc


def check(got, ref, with_rgb):
    w, rgb, acc, dd, dc = got
    rw, rrgb, racc, rdd, rdc = ref
    # serial CPU instantiation vs autograd: w 6e-8, acc / rgb 2.4e-7, d_density 1e-7 of the largest entry; the bars leave
    # room for the GPU's 2-ulp expf
    np.testing.assert_allclose(w, rw.numpy(), atol=1e-6)
    np.testing.assert_allclose(acc, racc.numpy(), atol=4e-6)
    np.testing.assert_allclose(rgb, rrgb.numpy(), atol=4e-6)
    scale = float(rdd.abs().max())
    assert np.abs(dd - rdd.numpy()).max() <= 5e-5 * scale, (np.abs(dd - rdd.numpy()).max(), scale)
    if with_rgb:      # autograd leaves .grad unset when no colour-dependent output has a gradient; the op returns zeros
        np.testing.assert_allclose(dc, rdc.numpy() if rdc is not None else np.zeros_like(dc), atol=5e-6)


@pytest.mark.parametrize("S,with_rgb", [(32, True), (128, False), (1, True), (7, True)])
def test_per_ray_functions_match_autograd_of_the_oracle_on_cpu(harness, S, with_rgb):
    inp = make_inputs(96, S, seed=10 + S, with_rgb=with_rgb)
    check(harness_run(harness, *inp), oracle_autograd(*inp), with_rgb)


def test_absent_incoming_gradients_and_background_on_cpu(harness):
    t, density, rgbs, dirs, gw, grgb, gacc = make_inputs(64, 16, seed=3)
    for sel in ((gw, None, None), (None, grgb, None), (None, None, gacc)):
        check(harness_run(harness, t, density, rgbs, dirs, *sel, bg=0.3), oracle_autograd(t, density, rgbs, dirs, *sel, bg=0.3), True)


# ---------------------------------------------------------------------------------------------------------------------


def cuda_run(t, density, rgbs, dirs, gw, grgb, gacc, bg=1.0):
    from ucnerf_b200.render_train import composite
    d = density.cuda().requires_grad_(True)
    c = None if rgbs is None else rgbs.cuda().requires_grad_(True)
    w, rgb, acc = composite(d, c, t.cuda(), dirs.cuda(), bg)
    loss = 0.
    if gw is not None:
        loss = loss + (w * gw.cuda()).sum()
    if grgb is not None:
        loss = loss + (rgb * grgb.cuda()).sum()
    if gacc is not None:
        loss = loss + (acc * gacc.cuda()).sum()
    loss.backward()
    f = lambda x: x.detach().cpu().numpy()
    return f(w), f(rgb), f(acc), f(d.grad), None if c is None else f(c.grad)


@pytest.mark.gpu
@pytest.mark.parametrize("S,with_rgb", [(32, True), (128, False), (1, True)])
def test_cuda_op_matches_autograd_of_the_oracle(S, with_rgb):
    inp = make_inputs(1000, S, seed=10 + S, with_rgb=with_rgb)
    check(cuda_run(*inp), oracle_autograd(*inp), with_rgb)


@pytest.mark.gpu
def test_cuda_op_partial_gradients_train_batch_and_errors():
    from ucnerf_b200.render_train import composite
    t, density, rgbs, dirs, gw, grgb, gacc = make_inputs(512, 32, seed=5)
    check(cuda_run(t, density, rgbs, dirs, None, grgb, None, bg=0.3), oracle_autograd(t, density, rgbs, dirs, None, grgb, None, bg=0.3), True)
    # one waymo.gin train batch on the proposal level: 15,000 rays x 128 samples
    inp = make_inputs(15000, 128, seed=6, with_rgb=False)
    w, rgb, acc, dd, _ = cuda_run(*inp)
    assert np.isfinite(w).all() and np.isfinite(dd).all() and w.min() >= 0 and acc.max() <= 1 + 1e-5
    for a in (1800, 5000):                                       # empty + opaque rays / ordinary rays (see make_inputs)
        ref = oracle_autograd(*[None if x is None else x[a:a + 256] for x in inp])
        check((w[a:a + 256], rgb[a:a + 256], acc[a:a + 256], dd[a:a + 256], None), ref, False)
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        composite(density, rgbs, t, dirs)                        # CPU tensors
    with pytest.raises(RuntimeError):
        composite(density.cuda(), rgbs.cuda(), t[:, :-1].cuda(), dirs.cuda())
