"""Fused tile exchange of a multi-GPU render over NVLink peer memory.

`render_image` on N ranks renders one contiguous ray tile per rank.  With a `PeerImage` the tiles do not wait for an
all-gather at the end: the compositing kernel of every chunk stores its finished packed pixel rows directly into the
image buffer of EVERY rank (this rank's own and, through cudaIpc peer mappings over NVLink, the others'), so the
exchange overlaps the render chunk by chunk.  What remains per frame is one 4-byte all-reduce on the render stream: it
cannot complete before every rank has entered it, every rank enters it after its own kernels (stream order), and a
kernel's stores - peer stores included - are visible system-wide when it completes.  The reference gathers every leaf of
every 15k-ray chunk with accelerate.gather (models.py:L965-968); the plain NCCL variant of this package is
`render.gather_tiles`.

Frames alternate between two image buffers per rank, so a rank that is still reading frame k (on its render stream) is
never overwritten by a peer that already renders frame k + 1: frame k + 2 reuses frame k's buffer only after the
all-reduce of frame k + 1, which this rank joins after its reads of frame k.

One process per GPU, all ranks on one node (IPC handles travel through torch.distributed's object all-gather)."""
import ctypes as C

import torch

from . import _lib
from .render import PACKED_WIDTH


class _DeviceArray:
    """A raw device allocation seen by torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerImage:
    """Image buffers [rows, 12] fp32 on every rank of `group`, mapped into every other rank."""

    def __init__(self, rows: int, device=None, group=None, buffers: int = 2):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise _lib.UcnerfError("PeerImage needs an initialised torch.distributed process group")
        self.lib = _lib.load()
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise _lib.UcnerfError("PeerImage: at most 16 ranks (UCNERF_MAX_PEERS)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.rows = int(rows)
        self.frame = 0
        nbytes = self.rows * PACKED_WIDTH * 4
        self._own, self._maps, self.images, self._tables = [], [], [], []
        err = None
        with torch.cuda.device(self.device):
            handles = []
            try:
                for _ in range(buffers):
                    ptr, h = C.c_void_p(), C.create_string_buffer(64)
                    _lib.check(self.lib.ucnerf_peer_alloc(nbytes, C.byref(ptr), h), "peer_alloc")
                    self._own.append(ptr)
                    handles.append(h.raw)
            except _lib.UcnerfError as e:
                err, handles = e, None
            everyone = [None] * self.world
            dist.all_gather_object(everyone, handles, group=group)
            if err is None and any(h is None for h in everyone):
                err = _lib.UcnerfError("PeerImage: a peer could not allocate its image buffers")
            if err is None:
                try:
                    for b in range(buffers):
                        table = (C.c_void_p * self.world)()
                        for r in range(self.world):
                            if r == self.rank:
                                table[r] = self._own[b].value
                            else:
                                p = C.c_void_p()
                                _lib.check(self.lib.ucnerf_peer_open(everyone[r][b], C.byref(p)), f"peer_open(rank {r})")
                                self._maps.append(p)
                                table[r] = p.value
                        self._tables.append(table)
                        self.images.append(torch.as_tensor(_DeviceArray(self._own[b].value, (self.rows, PACKED_WIDTH)),
                                                           device=self.device))
                except _lib.UcnerfError as e:
                    err = e
            # every rank learns whether every rank succeeded, so all of them take the same path afterwards
            ok = torch.tensor([0.0 if err is not None else 1.0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if float(ok) < 1.0:
                self._release()
                raise err if err is not None else _lib.UcnerfError("PeerImage: peer mapping failed on another rank")
        self._flag = torch.zeros(1, device=self.device, dtype=torch.float32)

    def render(self, renderer, batch, train_frac, rand_vec, row0: int, want=()):
        """Render `batch` (this rank's tile, rays [row0, row0 + n) of the image) with the exchange fused in; returns
        (image [rows, 12] holding every rank's tile once the call's all-reduce is complete on the current stream, the
        dict of the other requested outputs)."""
        b = self.frame % len(self.images)
        self.frame += 1
        _lib.check(self.lib.ucnerf_set_peer_targets(renderer._handle, self.world, self._tables[b], int(row0)), "set_peer_targets")
        try:
            out = renderer.render_rays(batch, train_frac, rand_vec, tuple(want))
        finally:
            _lib.check(self.lib.ucnerf_set_peer_targets(renderer._handle, 0, None, 0), "set_peer_targets")
        self.dist.all_reduce(self._flag, group=self.group)      # frame complete on every rank (see module docstring)
        return self.images[b], out

    def render_host(self, renderer, host_batch, train_frac, row0: int, want=(), out=None):
        """Same with the ray batch in (pinned) HOST memory: the library's chunk pipeline copies chunk c + 1 in while chunk c
        renders (ucnerf_render_rays_host), the compositing kernel writes the tiles to every rank's image.  Returns the
        image; the call has synchronised the render stream before the all-reduce is enqueued.
        `want` / `out` (pinned host tensors): outputs of THIS rank's tile copied back by the same pipeline, chunk c - 1
        under chunk c's kernels - N ranks then do not all start a whole-tile device-to-host copy at the frame's end."""
        b = self.frame % len(self.images)
        self.frame += 1
        _lib.check(self.lib.ucnerf_set_peer_targets(renderer._handle, self.world, self._tables[b], int(row0)), "set_peer_targets")
        try:
            renderer.render_rays_host(host_batch, train_frac, want=tuple(want), out=out if out is not None else {})
        finally:
            _lib.check(self.lib.ucnerf_set_peer_targets(renderer._handle, 0, None, 0), "set_peer_targets")
        self.dist.all_reduce(self._flag, group=self.group)
        return self.images[b]

    def _release(self):
        self.images = []
        for p in self._maps:
            self.lib.ucnerf_peer_close(p)
        for p in self._own:
            self.lib.ucnerf_peer_free(p)
        self._maps, self._own = [], []

    def close(self):
        """Collective: call on every rank (nobody unmaps while a peer may still write)."""
        if not self._own:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            self.dist.barrier(group=self.group)
            self._release()
