"""Frame sharding over the GPUs of one box: one process per GPU, torch.distributed for the plumbing.

Frames are independent given the fixed template/factor, so rank r reconstructs the
contiguous block ``shard_range(n, r, world)`` on its own GPU with a replicated plan; there is no
collective inside the path.  The only exchange is the optional gather of the vertex buffers
(``all_gather_meshes`` / ``gather_meshes``): NCCL over NVLink on GPUs, gloo in the CPU tests.

The reference has nothing to mirror here: it reconstructs one frame per Python call in a single
process (speech_anime/model/model.py:201-212, viewer/video.py:220-277).
"""
from __future__ import annotations


def shard_range(n_frames: int, rank: int, world: int):
    """Contiguous block of frames of `rank`: sizes differ by at most one, earlier ranks get the extra."""
    base, extra = divmod(int(n_frames), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_frames: int, world: int):
    return [shard_range(n_frames, r, world)[1] - shard_range(n_frames, r, world)[0] for r in range(world)]


def reconstruct_sharded(compute, inputs, n_frames, group=None):
    """Runs `compute` on this rank's frame block of every array in `inputs` (each [n_frames, ...]).

    `compute(*local_inputs) -> [n_local, n_verts, 3]`; in production it is
    ``Reconstructor.get_mesh_batch`` or ``Reconstructor.decode_and_get_mesh`` bound to this rank's GPU.
    Returns (local_vertices, (lo, hi))."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    lo, hi = shard_range(n_frames, rank, world)
    return compute(*[x[lo:hi] for x in inputs]), (lo, hi)


def _padded(local, rows):
    import torch
    if local.shape[0] == rows:
        return local.contiguous()
    pad = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    return pad


def _unpad(stacked, sizes):
    import torch
    return torch.cat([stacked[r, :n] for r, n in enumerate(sizes)], dim=0)


def all_gather_meshes(local, n_frames, group=None):
    """Every rank ends up with the full [n_frames, n_verts, 3] float32 tensor.  Ragged shards (n_frames not a
    multiple of the world size) are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_frames, world)
    if len(set(sizes)) == 1:
        out = torch.empty((n_frames,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    stacked = torch.empty((world, max(sizes)) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(stacked.view((world * max(sizes),) + tuple(local.shape[1:])), _padded(local, max(sizes)), group=group)
    return _unpad(stacked, sizes)


def gather_meshes(local, n_frames, dst=0, group=None):
    """Only `dst` receives the full tensor (returns None elsewhere)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_frames, world)
    send = _padded(local, max(sizes))
    if rank == dst:
        stacked = torch.empty((world, max(sizes)) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.gather(send, list(stacked.unbind(0)), dst=dst, group=group)
        return _unpad(stacked, sizes)
    dist.gather(send, None, dst=dst, group=group)
    return None


class GatherPipeline:
    """Reconstruct this rank's frames chunk by chunk and gather every chunk while the next one is computed.

    SURVEY 8(e): the collective runs on a side stream after the output kernel has written the contiguous send
    buffer and overlaps with the next chunk's kernels.  Only the FREE rows travel (``free_only`` results,
    [frames, n_free, 3]: 15 KB instead of 60 KB per FLAME frame -- the constrained rows are constants every rank
    already holds); ``expand=True`` rebuilds the reference layout [frames, n_verts, 3] on the receiving side with
    ``Reconstructor.expand_free``.

    ``rec`` needs ``decode_and_get_mesh(xs, xr, out=, free_only=True)`` / ``get_mesh_batch(x, out=, free_only=True)``,
    ``expand_free(rows, out=)``, ``n_free`` and ``n_verts`` -- a ``deformation.Reconstructor`` bound to this rank's GPU
    (the CPU tests pass a stand-in and gloo).  Every rank must hold the same number of frames per call (the bench's
    weak-scaling layout); ragged batches go through ``all_gather_meshes``.
    """

    def __init__(self, rec, chunk_frames=9472, group=None, mode="all", dst=0, expand=False):
        import torch.distributed as dist
        assert mode in ("all", "root")
        self.rec, self.chunk, self.group, self.mode, self.dst, self.expand = rec, int(chunk_frames), group, mode, dst, expand
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._side = None
        self._bufs = None

    def _streams(self, device):
        import torch
        if device.type != "cuda":
            return None
        if self._side is None:
            self._side = torch.cuda.Stream(device)
        return self._side

    def _buffers(self, device):
        import torch
        if self._bufs is None or self._bufs[0].device != device:
            rows = (self.chunk, self.rec.n_free, 3)
            self._bufs = [torch.empty(rows, dtype=torch.float32, device=device) for _ in range(2)]          # send, double buffered
            self._recv = [torch.empty((self.world,) + rows, dtype=torch.float32, device=device) for _ in range(2)]
        return self._bufs, self._recv

    def run(self, compute, inputs, out=None):
        """``compute(*chunk_inputs, out=send_rows)`` fills ``send_rows`` [nf, n_free, 3] for a chunk of this rank's
        frames.  Returns, on the receiving ranks, [world * n_local, n_free, 3] (rank-major: rank r's frames at
        ``r * n_local``), or [world * n_local, n_verts, 3] with ``expand=True``; ``None`` on the other ranks in
        ``mode="root"``."""
        import torch
        import torch.distributed as dist
        n_local = inputs[0].shape[0]
        device = inputs[0].device
        rows_out = self.rec.n_verts if self.expand else self.rec.n_free
        receiver = self.mode == "all" or self.rank == self.dst
        if receiver and out is None:
            out = torch.empty((self.world * n_local, rows_out, 3), dtype=torch.float32, device=device)
        if self.world == 1:
            for c0 in range(0, n_local, self.chunk):
                nf = min(self.chunk, n_local - c0)
                if self.expand:
                    (send, _), _ = self._buffers(device), None
                    compute(*[x[c0:c0 + nf] for x in inputs], out=send[:nf])
                    self.rec.expand_free(send[:nf], out=out[c0:c0 + nf])
                else:
                    compute(*[x[c0:c0 + nf] for x in inputs], out=out[c0:c0 + nf])
            return out
        side = self._streams(device)
        send, recv = self._buffers(device)
        main = torch.cuda.current_stream(device) if side is not None else None
        done = [None, None]                   # the side stream's event per buffer: its previous gather has been consumed
        for i, c0 in enumerate(range(0, n_local, self.chunk)):
            nf, b = min(self.chunk, n_local - c0), i & 1
            if side is not None and done[b] is not None:
                main.wait_event(done[b])      # the send buffer is free again
            compute(*[x[c0:c0 + nf] for x in inputs], out=send[b][:nf])
            ctx = torch.cuda.stream(side) if side is not None else _null()
            if side is not None:
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
            with ctx:
                src = send[b][:nf]
                # a ragged last chunk gets its own (side-stream) receive buffer so that the gather output stays contiguous
                fresh = lambda: torch.empty((self.world, nf) + tuple(src.shape[1:]), dtype=src.dtype, device=device)  # noqa: E731
                if self.mode == "all":
                    got = recv[b] if nf == self.chunk else fresh()
                    dist.all_gather_into_tensor(got.view((self.world * nf,) + tuple(src.shape[1:])), src, group=self.group)
                else:
                    if self.rank == self.dst:
                        got = recv[b] if nf == self.chunk else fresh()
                        parts = [got[r] for r in range(self.world)]
                        dist.gather(src, parts, dst=self.dst, group=self.group)
                    else:
                        dist.gather(src, None, dst=self.dst, group=self.group)
                        got = None
                if got is not None:
                    for r in range(self.world):
                        tgt = out[r * n_local + c0: r * n_local + c0 + nf]
                        if self.expand:
                            self.rec.expand_free(got[r], out=tgt, **({"stream": side.cuda_stream} if side is not None else {}))
                        else:
                            tgt.copy_(got[r])
                if side is not None:
                    done[b] = torch.cuda.Event()
                    done[b].record(side)
        if side is not None:
            main.wait_stream(side)
        return out if receiver else None


class PeerGather:
    """The gather of the vertex buffers without a collective library on the data path: every rank's result buffer is
    allocated as symmetric memory (``torch.distributed._symmetric_memory``: CUDA VMM allocations mapped into every peer
    over NVLink), each chunk of free rows is reconstructed straight into this rank's block of its own buffer, and a side
    stream pushes the finished chunk into the same block of every receiver's buffer with peer-to-peer copies (copy
    engines: no SMs, no staging buffer, no unpack) while the next chunk is computed.  A device-side barrier at either end
    of a call.

    ``mode="all"``: every rank ends up with [world * n_local, n_free, 3] (rank-major); ``mode="root"``: only ``dst``.
    The returned tensor is the symmetric buffer itself: consume (or copy) it before the next ``run``.
    Raises at construction if the box cannot map peer memory (then use ``GatherPipeline``: NCCL)."""

    def __init__(self, rec, n_local, chunk_frames=9472, group=None, mode="all", dst=0):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        assert mode in ("all", "root")
        self.rec, self.n_local, self.chunk, self.mode, self.dst = rec, int(n_local), int(chunk_frames), mode, dst
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        dev = torch.device("cuda", rec.device)
        shape = (self.world * self.n_local, rec.n_free, 3)
        self.buf = symm_mem.empty(shape, dtype=torch.float32, device=dev)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.peers = [self.hdl.get_buffer(r, shape, torch.float32) for r in range(self.world)]
        # two pushes in flight (measured at eight GPUs: two streams 13.6 ms per step, four 24.3 ms -- more concurrent pushes
        # put several senders on one receiver's link again)
        self.sides = [torch.cuda.Stream(dev) for _ in range(2)]
        self.bytes_pushed_per_run = 0

    def run(self, compute, inputs):
        import torch
        n, lo = self.n_local, self.rank * self.n_local
        assert inputs[0].shape[0] == n, "every rank holds n_local frames per call"
        main = torch.cuda.current_stream(self.buf.device)
        # staggered order: in phase s every rank pushes to rank + s, so no receiver has two senders at once
        targets = [(self.rank + s) % self.world for s in range(1, self.world)]
        targets = [r for r in targets if self.mode == "all" or r == self.dst]
        self.hdl.barrier()                                 # every receiver is done with the previous result
        pushed = 0
        for c0 in range(0, n, self.chunk):
            nf = min(self.chunk, n - c0)
            mine = self.buf[lo + c0: lo + c0 + nf]
            compute(*[x[c0:c0 + nf] for x in inputs], out=mine)
            ev = torch.cuda.Event()
            ev.record(main)
            for i, r in enumerate(targets):
                side = self.sides[i % len(self.sides)]
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    self.peers[r][lo + c0: lo + c0 + nf].copy_(mine, non_blocking=True)
                pushed += mine.numel() * 4
        for side in self.sides:
            main.wait_stream(side)
        self.hdl.barrier()                                 # everybody's pushes have landed everywhere
        self.bytes_pushed_per_run = pushed
        return self.buf if (self.mode == "all" or self.rank == self.dst) else None


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
