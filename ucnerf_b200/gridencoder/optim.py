"""Fused optimiser for GridEncoder tables.

`GridAdam` is a drop-in for the `torch.optim.Adam(model.parameters(), lr, betas, eps)` the reference creates
(internal/train_utils.py:L347-366) restricted to hash-grid embeddings: one kernel pass per table applies the hash-decay
gradient the reference obtains from its `loss_hash_decay` term (internal/models.py:L297-306 x `Config.hash_decay_mults`,
train_utils.py:L301-305), `grad.nan_to_num_()` (train_utils.py:L344-345), the Adam update and - optionally - zeroes the
gradient for the next backward (train.py:L164).  When it is used, drop the hash-decay term from the loss
(`Config.hash_decay_mults = 0`) and pass the multiplier here instead; every other parameter stays with torch's Adam.
The learning-rate schedule works as in train.py:L154-157 (`param_group['lr'] = lr_fn(step)`)."""
import torch

from .. import _lib


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
    def step(self, closure=None):
        loss = closure() if closure is not None else None
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
                    rc = self.lib.ucnerf_grid_adam_step(
                        p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), off.data_ptr(),
                        off.numel() - 1, 4, float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]),
                        float(group["eps"]), int(st["step"]), float(group["hash_decay_mult"]), int(group["zero_grad"]),
                        torch.cuda.current_stream().cuda_stream)
                _lib.check(rc, "grid_adam_step")
                if g is not p.grad and group["zero_grad"]:
                    p.grad.zero_()
        return loss
