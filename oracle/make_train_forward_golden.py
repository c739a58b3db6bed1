"""TEST INFRASTRUCTURE - golden vectors for the training level loop from the REFERENCE's own Model.forward(rand=True)
in train() mode on CPU, with autograd (oracle/ref_shim.py: unmodified internal/models.py, stepfun.py, render.py,
coord.py, gridencoder/grid.py; the native `_gridencoder` is the pinned CPU stand-in).  torch.rand / rand_like /
randn_like are patched to return the stored draws (per level: jitter, flip mask, rotation, rand_vec - the reference's
call order).  Stored: the draws, per-level sdist / weights / rgb / acc / loss_hash_decay, a scalar loss built from all
of them, the gradients of the small layers in full, seeded projections of the large ones and of the embeddings, and the
embedding gradients ENTRY-WISE on a seeded sample of the table entries the rays touch (plus some they do not).

    python oracle/make_train_forward_golden.py        # writes tests/golden/train_forward.npz"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cases, ref_shim  # noqa: E402

N_RAYS, TRAIN_FRAC = 256, 0.3
N_ENTRIES = 8192   # table entries whose gradient is stored entry-wise (seeded sample of the entries the rays touch)
FULL = ("density_layer.0.bias", "density_layer.2.bias", "rgb_layer.weight", "rgb_layer.bias", "density_layer.0.weight")


def loss_fn(renderings, ray_history, target, Gs):
    """A scalar touching everything the real losses touch: rgb of every level (data loss), weights (interlevel /
    distortion), acc (sky / opacity), loss_hash_decay."""
    loss = 0.
    for l, (r, h) in enumerate(zip(renderings, ray_history)):
        loss = loss + (0.5 + l) * ((r['rgb'] - target) ** 2).sum() + (h['weights'] * Gs[l]).sum() \
            + 0.05 * r['acc'].sum() + 0.1 * h['loss_hash_decay']
    return loss


def projections(g, seed):
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(3):
        R = torch.randn(g.shape, generator=gen, dtype=torch.float64)
        out.append(float((g.double() * R).sum()))
    return np.array(out + [float(g.double().abs().sum())])


def main():
    cfg, params, batch = cases.make_case("waymo", N_RAYS)
    model, conf = ref_shim.build_reference_model(cfg, params)
    model.train()
    g = torch.Generator().manual_seed(77)
    out = {}
    draws = []
    for l in range(cfg.num_levels):
        S = cfg.num_prop_samples if l < cfg.num_levels - 1 else cfg.num_nerf_samples
        d = dict(jitter01=torch.rand((N_RAYS, 1), generator=g), flip01=torch.rand((N_RAYS, S), generator=g),
                 rot01=torch.rand((N_RAYS, S), generator=g), rand_vec=torch.randn((N_RAYS, 3), generator=g))
        draws.append(d)
        for k, v in d.items():
            out[f"draw{l}_{k}"] = v.numpy()
    target = torch.rand((N_RAYS, 3), generator=g)
    Gs = [torch.randn((N_RAYS, cfg.num_prop_samples if l < cfg.num_levels - 1 else cfg.num_nerf_samples), generator=g) * 0.1
          for l in range(cfg.num_levels)]
    out["target"] = target.numpy()
    for l, G in enumerate(Gs):
        out[f"G{l}"] = G.numpy()
    q_rand = [d["jitter01"] for d in draws]
    q_like = [x for d in draws for x in (d["flip01"], d["rot01"])]
    q_randn = [d["rand_vec"] for d in draws]
    o = (torch.rand, torch.rand_like, torch.randn_like)
    torch.rand = lambda *a, **k: q_rand.pop(0).clone()
    torch.rand_like = lambda t, *a, **k: q_like.pop(0).clone()
    torch.randn_like = lambda t, *a, **k: q_randn.pop(0).clone()
    try:
        b = {k: v for k, v in batch.items() if k != 'rand_vec'}
        renderings, ray_history = model(True, b, train_frac=TRAIN_FRAC, compute_extras=False, zero_glo=True)
    finally:
        torch.rand, torch.rand_like, torch.randn_like = o
    assert not q_rand and not q_like and not q_randn, "draw order differs from the reference's"
    loss = loss_fn(renderings, ray_history, target, Gs)
    loss.backward()
    out["loss"] = np.float64(loss.item())
    out["train_frac"] = np.float64(TRAIN_FRAC)
    for l, (r, h) in enumerate(zip(renderings, ray_history)):
        for k in ("rgb", "acc", "weights", "depth"):
            out[f"r{l}_{k}"] = r[k].detach().numpy()
        out[f"h{l}_sdist"] = h["sdist"].detach().numpy()
        out[f"h{l}_density"] = h["density"].detach().numpy()
        out[f"h{l}_loss_hash_decay"] = np.float64(h["loss_hash_decay"].item())
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        short = name.split(".", 1)[1]
        if name.endswith("embeddings"):
            # entry-wise: entries whose gradient is more than the hash-decay term 0.1 * 2 p / (T_l L C) (i.e. touched by a
            # ray) - a seeded sample of them - plus a few untouched ones
            enc = model.get_submodule(name.rsplit(".", 1)[0])
            T = (enc.offsets[1:] - enc.offsets[:-1]).double()
            L = T.numel()
            decay = 0.1 * 2.0 * p.detach().double() / (T[enc.idx.long()] * L * p.shape[1])[:, None]
            touched = ((p.grad.double() - decay).abs().max(dim=1).values > 1e-12).nonzero()[:, 0]
            gen = torch.Generator().manual_seed(len(name) + 1)
            pick = touched[torch.randperm(touched.numel(), generator=gen)[:N_ENTRIES]]
            extra = torch.randint(0, p.shape[0], (512,), generator=gen)
            sel = torch.unique(torch.cat([pick, extra]))
            out["gsel_idx_" + name] = sel.to(torch.int32).numpy()
            out["gsel_val_" + name] = p.grad[sel].numpy()
            out["gsel_touched_" + name] = np.int64(touched.numel())
        if name.endswith("embeddings") or short not in FULL:
            out["gproj_" + name] = projections(p.grad, seed=len(name))
        else:
            out["grad_" + name] = p.grad.numpy()
    path = os.path.join(ROOT, "tests", "golden", "train_forward.npz")
    np.savez_compressed(path, **out)
    print("loss", loss.item(), "levels", len(renderings), "wrote", path, os.path.getsize(path), "bytes")
    print(sorted(k for k in out if k.startswith("g")))


if __name__ == "__main__":
    main()
