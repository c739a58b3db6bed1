"""Resampling for callers that keep the rest of a level in PyTorch (the reference's training step): one call replaces
`stepfun.max_dilate_weights` + the slice + the annealed logits + `stepfun.sample_intervals` of Model.forward's level loop
(internal/models.py:L156-205), `rand=True` included, through `ucnerf_resample_intervals` (csrc/resample_op.cu - the eval
path's warp-per-ray algorithm).  The reference detaches the result (`stop_level_grad`, models.py:L203-204); so does this.

    sdist = resample_level(sdist, weights, num_samples, dilation, use_dilation, anneal, resample_padding, rand, single_jitter)

There is no CPU path: CPU tensors raise."""
import numpy as np
import torch

from . import _lib

_EPS = float(np.finfo(np.float32).eps)


def u_grid(num_samples, rand, device):
    """Base grid of stepfun.sample (stepfun.py:L198-211) and the jitter scale (0 for rand=False)."""
    if not rand:
        pad = 1 / (2 * num_samples)
        return torch.linspace(pad, 1. - pad - _EPS, num_samples, device=device), 0.0
    u_max = _EPS + (1 - _EPS) / num_samples
    max_jitter = (1 - u_max) / (num_samples - 1) - _EPS
    return torch.linspace(0, 1 - u_max, num_samples, device=device), max_jitter


@torch.no_grad()
def resample_level(sdist, weights, num_samples, dilation=0.0, use_dilation=False, anneal=1.0, resample_padding=0.0,
                   rand=False, single_jitter=True, generator=None, rand01=None):
    """sdist [N, n+1], weights [N, n] (previous level; domain [0, 1]) -> new sdist [N, num_samples + 1].
    `rand01` (optional) supplies the uniform draw of stepfun.py:L212, shape [N, 1] or [N, num_samples]."""
    if sdist.device.type != "cuda" or weights.device.type != "cuda":
        raise RuntimeError("resample_level: sdist / weights must be CUDA tensors (no CPU path)")
    if sdist.dim() != 2 or weights.dim() != 2 or sdist.shape[1] != weights.shape[1] + 1 or sdist.shape[0] != weights.shape[0]:
        raise RuntimeError("resample_level: expected sdist [N, n+1] and weights [N, n]")
    if num_samples <= 1:
        raise ValueError(f'num_samples must be > 1, is {num_samples}.')          # stepfun.py:L271-272
    t = sdist.detach().contiguous().float()
    w = weights.detach().contiguous().float()
    N, n = w.shape
    base, max_jitter = u_grid(num_samples, rand, t.device)
    jitter, cols = None, 0
    if rand:
        cols = 1 if single_jitter else num_samples
        if rand01 is None:
            rand01 = torch.rand((N, cols), device=t.device, generator=generator)
        if tuple(rand01.shape) != (N, cols):
            raise RuntimeError(f"resample_level: rand01 must have shape {(N, cols)}")
        jitter = (rand01.float() * max_jitter).contiguous()
    out = torch.empty((N, num_samples + 1), device=t.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(t.device):
        rc = lib.ucnerf_resample_intervals(t.data_ptr(), w.data_ptr(), N, n, int(bool(use_dilation)), float(dilation),
                                           float(anneal), float(resample_padding), num_samples, base.data_ptr(),
                                           None if jitter is None else jitter.data_ptr(), cols, out.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "resample_intervals")
    return out
