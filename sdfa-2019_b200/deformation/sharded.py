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
