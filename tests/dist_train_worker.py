"""Worker for the world_size-2 gloo test of the training step's gradient exchange (ucnerf_b200.parallel_train)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ucnerf_b200.parallel_train import allreduce_gradients  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 24), (64,), (1, 64), (3, 256), (0,)]
    dense = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    tables = [torch.nn.Parameter(torch.zeros((1000, 4))), torch.nn.Parameter(torch.zeros((32, 4)))]
    base_d = [torch.randn(s, generator=g) for s in shapes]
    base_t = [torch.randn(p.shape, generator=g) for p in tables]
    for p, b in zip(dense, base_d):
        p.grad = b * (rank + 1)                          # rank r holds (r + 1) * base -> mean = (world + 1) / 2 * base
    for p, b in zip(tables, base_t):
        p.grad = (b * (rank + 1)).t().contiguous().t()   # a non-contiguous gradient must come back in place too
    if rank == 0:
        dense[1].grad = None                             # missing on ONE rank only: it contributes zeros (what DDP does) and
    sent = allreduce_gradients(dense, tables)            # every rank still issues identical collectives
    k = (world + 1) / 2
    for i, (p, b) in enumerate(zip(dense, base_d)):
        if i == 1:
            assert torch.allclose(p.grad, b * (k - 1.0 / world), atol=1e-6)   # rank 0's share (1 * base) is zero
        else:
            assert torch.allclose(p.grad, b * k, atol=1e-6), i
    for p, b in zip(tables, base_t):
        assert torch.allclose(p.grad, b * k, atol=1e-6)
    expect = 4 * (sum(b.numel() for b in base_d) + sum(b.numel() for b in base_t))
    assert sent == expect, (sent, expect)
    # the overlapped variant: table all-reduces start from post-accumulate-grad hooks inside backward()
    from ucnerf_b200.parallel_train import OverlappedGradientExchange
    t2 = [torch.nn.Parameter(torch.ones((50, 4))), torch.nn.Parameter(torch.ones((7, 4)))]
    d2 = [torch.nn.Parameter(torch.ones(5)), torch.nn.Parameter(torch.ones((2, 3)))]
    ex = OverlappedGradientExchange(d2, t2)
    for step in range(2):
        for p in t2 + d2:
            p.grad = None
        loss = (rank + 1.0) * (t2[0].sum() * 2 + (t2[1] ** 2).sum() + d2[0].sum() * 3 + d2[1].sum())
        loss.backward()
        ex.finish()
        assert torch.allclose(t2[0].grad, torch.full((50, 4), 2 * k)) and torch.allclose(t2[1].grad, torch.full((7, 4), 2 * k))
        assert torch.allclose(d2[0].grad, torch.full((5,), 3 * k)) and torch.allclose(d2[1].grad, torch.full((2, 3), k))
    ex.close()
    dist.barrier()
    if rank == 0:
        print(f"TRAIN_EXCHANGE_OK gloo {world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
