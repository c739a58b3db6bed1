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


class OverlappedGradientExchange:
    """The same exchange with the big part hidden behind the backward pass: a hash table's gradient is final as soon as the
    pooled-encode backward of its level has run (the NeRF level's table - 229 MB of the 346 MB - half a backward before the
    end), so a post-accumulate-grad hook starts its all-reduce right there, asynchronously on the communication stream, and
    `finish()` - called where `allreduce_gradients` was - only waits and exchanges the small dense bucket.

        ex = OverlappedGradientExchange(dense_params, [enc.embeddings for enc in encoders])     # once
        loss.backward(); ex.finish(); optimizer.step(); grid_opt.step()

    Requires the table gradients to be produced on every rank in every step (they are: every ray touches every level)."""

    def __init__(self, dense_params, table_params, group=None):
        self.dense, self.tables, self.group = list(dense_params), list(table_params), group
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self._pending = []
        self._handles = []
        if self.enabled:
            self.world = dist.get_world_size(group)
            self._avg = dist.get_backend(group) == "nccl"       # NCCL averages in the collective; gloo sums
            for p in self.tables:
                self._handles.append(p.register_post_accumulate_grad_hook(self._start))

    def _start(self, p):
        g = p.grad
        if not g.is_contiguous():
            p.grad = g = g.contiguous()
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._pending.append((g, dist.all_reduce(g, op=op, group=self.group, async_op=True)))

    def finish(self):
        """Wait for the table reductions started during backward, exchange the dense bucket; returns bytes sent."""
        if not self.enabled:
            return 0
        sent = allreduce_gradients(self.dense, (), self.group)
        started = {id(g) for g, _ in self._pending}
        for p in self.tables:                                   # a table whose hook did not fire (no gradient this step)
            if p.grad is None or id(p.grad) not in started:
                sent += allreduce_gradients((), [p], self.group)
        for g, work in self._pending:
            work.wait()
            if not self._avg:
                g /= self.world
            sent += g.numel() * g.element_size()
        self._pending = []
        return sent

    def close(self):
        for h in self._handles:
            h.remove()
        self._handles = []
