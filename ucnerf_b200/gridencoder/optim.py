"""Fused optimiser for GridEncoder tables.

`GridAdam` is a drop-in for the `torch.optim.Adam(model.parameters(), lr, betas, eps)` the reference creates
(internal/train_utils.py:L347-366) restricted to hash-grid embeddings: one kernel pass per table applies the hash-decay
gradient the reference obtains from its `loss_hash_decay` term (internal/models.py:L297-306 x `Config.hash_decay_mults`,
train_utils.py:L301-305), `grad.nan_to_num_()` (train_utils.py:L344-345), the Adam update and - optionally - zeroes the
gradient for the next backward (train.py:L164).  When it is used, drop the hash-decay term from the loss
(`Config.hash_decay_mults = 0`) and pass the multiplier here instead; every other parameter stays with torch's Adam.
The learning-rate schedule works as in train.py:L154-157 (`param_group['lr'] = lr_fn(step)`).

Gradient clipping (train_utils.clip_gradients, L335-345: `Config.grad_max_norm` / `grad_max_val`, both 0 in the shipped
configs) is part of the pass: `table_stats()` gives each table's squared gradient norm (hash-decay gradient included)
in one read pass, `clip_coefficient()` combines them with the dense parameters' norm exactly as
`torch.nn.utils.clip_grad_norm_` does over `model.parameters()`, and `step(grad_scale=..., grad_max_val=...)` applies the
coefficient and the value clamp before `nan_to_num_()` inside the kernel."""
import torch

from .. import _lib


@torch.no_grad()
def hash_decay_loss(encoder) -> torch.Tensor:
    """Value of the reference's per-level hash-decay term (models.py:L297-306:
    `segment_coo(param ** 2, idx, reduce='mean').mean()`) from ONE streaming read of the table (ucnerf_grid_table_stats),
    instead of the reference's index_add over the 21 M-entry `idx` buffer.  0-dim fp32 CUDA tensor WITHOUT autograd: on
    the fused training path the term's gradient is applied inside GridAdam.step (hash_decay_mult)."""
    p = encoder.embeddings
    lib = _lib.load()
    off = getattr(encoder, "_ucnerf_offsets_host", None)
    if off is None:
        off = encoder.offsets.detach().cpu().to(torch.int32).contiguous()
        encoder._ucnerf_offsets_host = off
    L = off.numel() - 1
    sums = torch.empty(2 * L, dtype=torch.float64, device=p.device)
    with torch.cuda.device(p.device):
        rc = lib.ucnerf_grid_table_stats(p.data_ptr(), None, off.data_ptr(), L, int(p.shape[1]), 0.0, sums.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "grid_table_stats")
    T = (off[1:] - off[:-1]).to(torch.float64).to(p.device)
    return (sums[0::2] / (T * L * p.shape[1])).sum().float()


class GridAdam(torch.optim.Optimizer):
    def __init__(self, encoders, lr=0.01, betas=(0.9, 0.99), eps=1e-15, hash_decay_mult=0.0, zero_grad=False):
        """encoders: iterable of GridEncoder modules (anything with `.embeddings` [sum T, 4] fp32 CUDA and `.offsets`)."""
        self.lib = _lib.load()
        encoders = list(encoders)
        params = []
        self._offsets = {}
        for enc in encoders:
            p = enc.embeddings
            if p.dtype != torch.float32 or p.shape[1] != 4 or not p.is_cuda or not p.is_contiguous():
                raise ValueError("GridAdam needs contiguous fp32 CUDA embeddings with level_dim == 4")
            params.append(p)
            self._offsets[id(p)] = enc.offsets.detach().cpu().to(torch.int32).contiguous()
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, hash_decay_mult=hash_decay_mult, zero_grad=zero_grad))

    @torch.no_grad()
    def table_stats(self):
        """One read pass per table -> list of dicts (one per table, in construction order):
        `loss_hash_decay` (models.py:L297-306, a 0-dim fp64 CUDA tensor, no autograd: its gradient is applied by step())
        and `grad_sq_norm` (sum of (grad + hash-decay gradient)^2, 0-dim fp64 CUDA tensor)."""
        out = []
        for group in self.param_groups:
            for p in group["params"]:
                off = self._offsets[id(p)]
                L = off.numel() - 1
                sums = torch.empty(2 * L, dtype=torch.float64, device=p.device)
                g = p.grad
                if g is not None and not g.is_contiguous():
                    g = g.contiguous()
                with torch.cuda.device(p.device):
                    rc = self.lib.ucnerf_grid_table_stats(p.data_ptr(), None if g is None else g.data_ptr(), off.data_ptr(),
                                                          L, 4, float(group["hash_decay_mult"]), sums.data_ptr(),
                                                          torch.cuda.current_stream().cuda_stream)
                _lib.check(rc, "grid_table_stats")
                T = (off[1:] - off[:-1]).to(torch.float64).to(p.device)
                out.append({"loss_hash_decay": (sums[0::2] / (T * L * 4)).sum(), "grad_sq_norm": sums[1::2].sum()})
        return out

    @torch.no_grad()
    def clip_coefficient(self, other_params=(), max_norm=0.0):
        """torch.nn.utils.clip_grad_norm_ over (tables + `other_params`): clips the gradients of `other_params` in place and
        returns the coefficient (0-dim CUDA tensor) to pass to step(grad_scale=...).  max_norm <= 0: no clipping (1.0)."""
        if max_norm <= 0:
            return 1.0
        sq = sum(s["grad_sq_norm"] for s in self.table_stats())
        others = [p for p in other_params if p.grad is not None]
        if others:
            sq = sq + torch.stack([p.grad.double().square().sum() for p in others]).sum().to(sq.device)
        coef = torch.clamp(max_norm / (sq.sqrt() + 1e-6), max=1.0)
        for p in others:
            p.grad.mul_(coef.to(p.grad.dtype))
        return coef

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0, grad_max_val=0.0):
        loss = closure() if closure is not None else None
        grad_scale = float(grad_scale)      # (a 0-dim tensor from clip_coefficient is read back here: one host sync)
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                off = self._offsets[id(p)]
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                with torch.cuda.device(p.device):
                    rc = self.lib.ucnerf_grid_adam_step_clipped(
                        p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), off.data_ptr(),
                        off.numel() - 1, 4, float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]),
                        float(group["eps"]), int(st["step"]), float(group["hash_decay_mult"]), int(group["zero_grad"]),
                        grad_scale, float(grad_max_val), torch.cuda.current_stream().cuda_stream)
                _lib.check(rc, "grid_adam_step")
                if g is not p.grad and group["zero_grad"]:
                    p.grad.zero_()
        return loss
