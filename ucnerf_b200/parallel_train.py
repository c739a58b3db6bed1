"""Gradient exchange of the training step for N > 1 ("Training (config 5): replicas + gradient
all-reduce"): what DistributedDataParallel does for the reference under `accelerate` (train.py:L119, L222), written
out for the two parameter families of this package - the dense `nn.Linear` parameters travel as ONE flat bucket
(a few hundred kB), each hash table's gradient is reduced in place (it is the fused optimiser's input and is
zeroed by it).  Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU test-suite)."""
import torch
import torch.distributed as dist


def allreduce_gradients(dense_params, table_params=(), group=None):
    """Average the `.grad` of `dense_params` (one flat all-reduce) and of `table_params` (in place, one all-reduce
    each) over the ranks of `group`.  A parameter without a gradient on this rank contributes zeros (what DDP does), so
    every rank issues the same collectives with the same sizes whatever subset of its parameters received a gradient.
    Returns the bytes this rank contributed, for reporting."""
    if not dist.is_available() or not dist.is_initialized():
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    sent = 0
    dense = list(dense_params)
    for p in dense:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    if dense:
        flat = torch.cat([p.grad.reshape(-1) for p in dense])
        dist.all_reduce(flat, group=group)
        flat /= world
        off = 0
        for p in dense:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        sent += flat.numel() * flat.element_size()
    for p in table_params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
        dist.all_reduce(g, group=group)
        g /= world
        if g is not p.grad:
            p.grad.copy_(g)
        sent += g.numel() * g.element_size()
    return sent
