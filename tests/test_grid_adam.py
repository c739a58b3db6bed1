"""Fused hash-decay + Adam step for GridEncoder tables against the reference's recipe evaluated
with torch itself: autograd of the hash-decay loss expression (models.py:L297-306), grad.nan_to_num_(), torch.optim.Adam
with the reference's hyper-parameters (train_utils.py:L344-366)."""
import numpy as np
import pytest
import torch

from oracle import ucnerf_oracle as O


def _table(seed=0):
    g = torch.Generator().manual_seed(seed)
    sizes = [48, 200, 1024, 4096]                      # entries per level (multiples of 8 like GridEncoder's)
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32)
    idx = torch.cat([torch.full((s,), l, dtype=torch.long) for l, s in enumerate(sizes)])
    emb = (torch.rand((int(offsets[-1]), 4), generator=g) * 2 - 1) * 0.5
    return emb, offsets, idx, g


def test_hash_decay_gradient_closed_form_matches_autograd():
    """d/dp of mult * loss_hash_decay = mult * 2 p / (T_level * L * C): what the kernel folds into the gradient."""
    emb, offsets, idx, _ = _table()
    p = emb.clone().requires_grad_(True)
    (0.1 * O.hash_decay_loss(p, idx)).backward()
    T = (offsets[1:] - offsets[:-1]).double()
    coef = 0.1 * 2.0 / (T * 4 * 4)
    want = emb.double() * coef[idx][:, None]
    assert torch.allclose(p.grad.double(), want, rtol=2e-6, atol=1e-12)


@pytest.mark.gpu
def test_gpu_fused_step_matches_reference_recipe():
    from ucnerf_b200.gridencoder.optim import GridAdam
    emb, offsets, idx, g = _table(1)

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.embeddings = torch.nn.Parameter(emb.clone().cuda())
            self.register_buffer("offsets", offsets.clone())

    enc = Enc()
    opt = GridAdam([enc], lr=0.01, betas=(0.9, 0.99), eps=1e-15, hash_decay_mult=0.1, zero_grad=True)
    ref = emb.clone()
    state = {}
    for step in range(1, 5):
        lr = 0.01 * (0.7 ** step)                       # a schedule, as train.py:L154-157 sets it per step
        rg = torch.randn(emb.shape, generator=g) * 1e-3
        rg[::7] = 0.0                                   # untouched entries: only the decay term moves them
        if step == 2:
            rg[5, 1] = float("nan")
            rg[9, 0] = float("inf")
        ref = O.reference_train_step_tables(ref, rg.clone(), idx, state, lr, 0.1)
        for grp in opt.param_groups:
            grp["lr"] = lr
        enc.embeddings.grad = rg.clone().cuda()
        opt.step()
        torch.cuda.synchronize()
        got = enc.embeddings.detach().cpu()
        err = (got - ref.detach()).abs().max().item()
        assert err < 2e-6, (step, err)
        assert torch.count_nonzero(enc.embeddings.grad).item() == 0          # zero_grad=True
        st = opt.state[enc.embeddings]
        rs = state["opt"].state[ref]
        m_got, m_ref = st["exp_avg"].cpu(), rs["exp_avg"]
        assert ((m_got - m_ref).abs() <= 1e-6 * m_ref.abs() + 1e-9).all()
        v_got, v_ref = st["exp_avg_sq"].cpu(), rs["exp_avg_sq"]
        fin = torch.isfinite(v_ref)
        assert torch.equal(torch.isfinite(v_got), fin)                       # the +inf gradient saturates both the same way
        rel = ((v_got[fin] - v_ref[fin]).abs() / (v_ref[fin].abs() + 1e-30)).max().item()
        assert rel < 1e-5, rel


@pytest.mark.gpu
def test_gpu_table_stats_and_clipped_step_match_torch_clipping():
    """train_utils.clip_gradients (L335-345): clip_grad_norm_ over ALL parameters, clip_grad_value_, nan_to_num_, then Adam.
    The fused path gets the tables' share of the norm from ucnerf_grid_table_stats (hash-decay gradient included) and
    applies coefficient + value clamp inside the optimiser kernel."""
    from ucnerf_b200.gridencoder.optim import GridAdam
    emb, offsets, idx, g = _table(2)

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.embeddings = torch.nn.Parameter(emb.clone().cuda())
            self.register_buffer("offsets", offsets.clone())

    enc = Enc()
    dense = torch.nn.Parameter(torch.randn(37, 5, generator=g).cuda())
    opt = GridAdam([enc], lr=0.01, betas=(0.9, 0.99), eps=1e-15, hash_decay_mult=0.1, zero_grad=True)
    rg = torch.randn(emb.shape, generator=g) * 2e-2
    dg = torch.randn(37, 5, generator=g) * 1e-2
    max_norm, max_val = 0.5, 0.01
    # reference recipe in torch
    p_ref = emb.clone().requires_grad_(True)
    d_ref = dense.detach().cpu().clone().requires_grad_(True)
    loss_ref = O.hash_decay_loss(p_ref, idx)
    (0.1 * loss_ref).backward()
    p_ref.grad += rg
    d_ref.grad = dg.clone()
    sq_ref = float(p_ref.grad.double().square().sum())
    torch.nn.utils.clip_grad_norm_([p_ref, d_ref], max_norm)
    torch.nn.utils.clip_grad_value_([p_ref, d_ref], max_val)
    ref_opt = torch.optim.Adam([p_ref], lr=0.01, betas=(0.9, 0.99), eps=1e-15)
    ref_opt.step()
    # fused path
    enc.embeddings.grad = rg.clone().cuda()
    dense.grad = dg.clone().cuda()
    st = opt.table_stats()[0]
    assert abs(float(st["loss_hash_decay"]) - float(loss_ref)) < 1e-6 * max(1.0, abs(float(loss_ref)))
    assert abs(float(st["grad_sq_norm"]) - sq_ref) < 1e-6 * sq_ref
    coef = opt.clip_coefficient([dense], max_norm)
    assert float(coef) < 1.0                                   # the clip is active in this case
    assert torch.allclose(dense.grad.cpu(), d_ref.grad if max_val <= 0 else (dg * float(coef)), rtol=1e-5, atol=1e-9)
    opt.step(grad_scale=coef, grad_max_val=max_val)
    torch.cuda.synchronize()
    err = (enc.embeddings.detach().cpu() - p_ref.detach()).abs().max().item()
    assert err < 2e-6, err
    # empty gradient: the loss value alone
    enc.embeddings.grad = None
    st2 = opt.table_stats()[0]
    assert float(st2["grad_sq_norm"]) >= 0 and abs(float(st2["loss_hash_decay"]) - float(O.hash_decay_loss(enc.embeddings.detach().cpu(), idx))) < 1e-6
