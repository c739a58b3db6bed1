"""Host-side mirror of the reference's gridencoder Python module (gridencoder/grid.py): the same
`GridEncoder` constructor, buffers (`embeddings`, `offsets`, `idx`, `grid_sizes`), attributes and
`forward(inputs, bound)` contract, and the `_grid_encode` autograd Function calling a `_backend` with the
reference's three entry points - here backed by libucnerf_b200.so (sm_100a) instead of `_gridencoder`.
State-dict names and shapes are identical, so reference checkpoints load unchanged."""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import backend as _backend

GRIDTYPE_IDS = {"hash": 0, "tiled": 1}
INTERP_IDS = {"linear": 0, "smoothstep": 1}


def level_table_sizes(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners):
    """Entries per level and python-side resolutions (reference grid.py:L118-135)."""
    cap = 2 ** log2_hashmap_size
    sizes, resolutions = [], []
    for level in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** level))
        if not align_corners:
            res += 1
        n = min(cap, res ** input_dim)
        sizes.append(int(np.ceil(n / 8) * 8))
        resolutions.append(res)
    return sizes, resolutions


class _grid_encode(Function):
    """reference grid.py:L24-89.  forward: inputs [B,D] in [0,1] -> [B, L*C]."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                gridtype=0, align_corners=False, interpolation=0):
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = np.log2(per_level_scale)
        H = base_resolution
        if torch.is_autocast_enabled() and C % 2 == 0:
            embeddings = embeddings.to(torch.half)
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
        dy_dx = (torch.empty(B, L * D * C, device=inputs.device, dtype=embeddings.dtype)
                 if calc_grad_inputs else None)
        _backend.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype,
                                     align_corners, interpolation)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = (B, D, C, L, S, H, gridtype, interpolation)
        ctx.align_corners = align_corners
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        grad_embeddings = torch.zeros_like(embeddings)
        grad_inputs = torch.zeros_like(inputs, dtype=embeddings.dtype) if dy_dx is not None else None
        _backend.grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx,
                                      grad_inputs, gridtype, ctx.align_corners, interpolation)
        if grad_inputs is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class GridEncoder(nn.Module):
    """reference grid.py:L95-198."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash', align_corners=False,
                 interpolation='linear', init_std=1e-4):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = GRIDTYPE_IDS[gridtype]
        self.interpolation = interpolation
        self.interp_id = INTERP_IDS[interpolation]
        self.align_corners = align_corners
        self.init_std = init_std
        self.max_params = 2 ** log2_hashmap_size

        sizes, resolutions = level_table_sizes(input_dim, num_levels, per_level_scale, base_resolution,
                                               log2_hashmap_size, align_corners)
        offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        total = int(offsets[-1])
        self.register_buffer('offsets', torch.from_numpy(offsets))
        self.register_buffer('idx', torch.repeat_interleave(torch.arange(num_levels, dtype=torch.long),
                                                            torch.tensor(sizes, dtype=torch.long)))
        self.register_buffer('grid_sizes', torch.from_numpy(np.array(resolutions, dtype=np.int32)))
        self.n_params = self.offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(total, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-self.init_std, self.init_std)

    def __repr__(self):
        finest = int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {finest} per_level_scale={self.per_level_scale:.4f} "
                f"params={tuple(self.embeddings.shape)} gridtype={self.gridtype} "
                f"align_corners={self.align_corners} interpolation={self.interpolation}")

    def forward(self, inputs, bound=1):
        """inputs [..., input_dim] in [-bound, bound] -> [..., num_levels * level_dim]."""
        inputs = (inputs + bound) / (2 * bound)
        lead = list(inputs.shape[:-1])
        flat = inputs.view(-1, self.input_dim)
        out = grid_encode(flat, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                          flat.requires_grad, self.gridtype_id, self.align_corners, self.interp_id)
        return out.view(lead + [self.output_dim])

    @torch.amp.autocast("cuda", enabled=False)
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        """reference grid.py:L176-198: accumulates the TV gradient into embeddings.grad."""
        D, C = self.input_dim, self.embeddings.shape[1]
        L = self.offsets.shape[0] - 1
        S = np.log2(self.per_level_scale)
        H = self.base_resolution
        if inputs is None:
            inputs = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            inputs = ((inputs + bound) / (2 * bound)).view(-1, self.input_dim)
            B = inputs.shape[0]
        if self.embeddings.grad is None:
            raise ValueError('grad is None, should be called after loss.backward() and before optimizer.step()!')
        _backend.grad_total_variation(inputs.contiguous(), self.embeddings, self.embeddings.grad, self.offsets, weight,
                                      B, D, C, L, S, H, self.gridtype_id, self.align_corners)
