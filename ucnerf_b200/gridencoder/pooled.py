"""Pooled hash-grid features for training: the front end of the reference's `MLP.predict_density`
(internal/models.py:L485-496) as one autograd Function backed by two kernels of libucnerf_b200.so
(`ucnerf_pooled_encode_forward/backward`, csrc/pooled_encode.cu).

    features, coord = pooled_encode(encoder, means, stds)      # means [...,M,3], stds [...,M]  ->  [..., L*C], [...,3]

replaces, in the reference,

    means, stds = coord.track_linearize(self.warp_fn, means, stds)      # no_grad (coord.py:L75)
    means = means / bound; stds = stds / bound                          # bound = 2
    features = self.encoder(means, bound=1).unflatten(-1, (self.encoder.num_levels, -1))
    weights = torch.erf(1 / torch.sqrt(8 * stds[..., None] ** 2 * self.encoder.grid_sizes ** 2))
    features = (features * weights[..., None]).mean(dim=-3).flatten(-2, -1)
    ... means.mean(dim=-2)                                              # the `coord` output (models.py:L512)

The only tensor with a gradient on that stretch is `encoder.embeddings` (the Gaussians come out of a no_grad block and
normals are disabled under configs/*.gin), so backward returns a gradient for the embeddings alone.  Supported:
fp32 embeddings with level_dim = 4, input_dim = 3, gridtype 'hash', align_corners = False, linear interpolation (what
MLP.__init__ builds, models.py:L425-436); anything else raises.  There is no CPU path."""
import numpy as np
import torch
from torch.autograd import Function

from .. import _lib


def _host_layout(encoder):
    """Host copies of the (constant) offsets / grid_sizes buffers, cached on the module."""
    lay = getattr(encoder, "_ucnerf_host_layout", None)
    if lay is None:
        lay = (np.ascontiguousarray(encoder.offsets.detach().cpu().numpy(), dtype=np.int32),
               np.ascontiguousarray(encoder.grid_sizes.detach().cpu().numpy(), dtype=np.int32))
        encoder._ucnerf_host_layout = lay
    return lay


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


class _pooled_encode(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, means, stds, embeddings, offsets_h, grid_sizes_h, S, H, contract, merge_runs):
        _check(means.device.type == "cuda" and embeddings.device.type == "cuda", "pooled_encode: tensors must be CUDA tensors")
        _check(embeddings.dtype == torch.float32 and embeddings.is_contiguous(), "pooled_encode: embeddings must be contiguous fp32")
        _check(embeddings.shape[1] == 4, "pooled_encode: level_dim must be 4")
        M = means.shape[-2]
        lead = means.shape[:-2]
        _check(means.shape[-1] == 3 and tuple(stds.shape) == tuple(means.shape[:-1]), "pooled_encode: means [...,M,3], stds [...,M]")
        m2 = means.detach().reshape(-1, M, 3).contiguous().float()
        s2 = stds.detach().reshape(-1, M).contiguous().float()
        B = m2.shape[0]
        L = offsets_h.shape[0] - 1
        feats = torch.empty((B, L * 4), device=means.device, dtype=torch.float32)
        coord = torch.empty((B, 3), device=means.device, dtype=torch.float32)
        lib = _lib.load()
        with torch.cuda.device(means.device):
            rc = lib.ucnerf_pooled_encode_forward(m2.data_ptr(), s2.data_ptr(), B, M, int(contract), embeddings.data_ptr(),
                                                  offsets_h.ctypes.data, grid_sizes_h.ctypes.data, L, 4, float(S), int(H),
                                                  feats.data_ptr(), coord.data_ptr(),
                                                  torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "pooled_encode_forward")
        ctx.save_for_backward(m2, s2, embeddings)
        ctx.meta = (offsets_h, grid_sizes_h, S, H, contract, merge_runs, B, M, L)
        feats, coord = feats.view(*lead, L * 4), coord.view(*lead, 3)
        ctx.mark_non_differentiable(coord)
        return feats, coord

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_feats, _grad_coord):
        m2, s2, embeddings = ctx.saved_tensors
        offsets_h, grid_sizes_h, S, H, contract, merge_runs, B, M, L = ctx.meta
        g = grad_feats.reshape(B, L * 4).contiguous().float()
        grad_emb = torch.zeros_like(embeddings)
        lib = _lib.load()
        with torch.cuda.device(g.device):
            # UCNERF_POOLED_CONTRACT | UCNERF_POOLED_MERGE_RUNS (merge_runs=True) | UCNERF_POOLED_MERGE_RAY_RUNS ('ray')
            flags = int(contract) | (4 if merge_runs == 'ray' else (2 if merge_runs else 0))
            rc = lib.ucnerf_pooled_encode_backward(g.data_ptr(), m2.data_ptr(), s2.data_ptr(), B, M, flags,
                                                   offsets_h.ctypes.data, grid_sizes_h.ctypes.data, L, 4, float(S), int(H),
                                                   grad_emb.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "pooled_encode_backward")
        return None, None, grad_emb, None, None, None, None, None, None


def pooled_encode(encoder, means, stds, contract=True, merge_runs=True):
    """`encoder`: a GridEncoder (this package's mirror or the reference's own class - only `embeddings`, `offsets`,
    `grid_sizes`, `per_level_scale`, `base_resolution` and the configuration attributes are read).
    `merge_runs`: backward variants that combine the contributions of consecutive points sharing a cell before the atomic
    reductions - True: within an interval, 'ray': across 4 consecutive intervals (rows must be ordered ray by ray, sample
    by sample, as render.cast_rays produces them) - same gradient up to fp32 summation order.  Measured on B200
    (profiles/r2_pooled_encode.json, forward + backward): proposal level 9.67 ms plain / 2.39 ms True / 2.20 ms 'ray';
    NeRF level 2.67 / 1.08 / 1.13 ms - hence on by default; False keeps the one-reduction-per-point form.
    Returns (features [..., L*C], coord [..., 3])."""
    _check(encoder.input_dim == 3 and encoder.level_dim == 4, "pooled_encode: input_dim 3 / level_dim 4 only")
    _check(getattr(encoder, "gridtype", "hash") == "hash" and not encoder.align_corners
           and getattr(encoder, "interpolation", "linear") == "linear",
           "pooled_encode: hash grid, align_corners=False, linear interpolation only")
    offsets_h, grid_sizes_h = _host_layout(encoder)
    return _pooled_encode.apply(means, stds, encoder.embeddings, offsets_h, grid_sizes_h,
                                np.log2(encoder.per_level_scale), encoder.base_resolution, bool(contract), merge_runs)
