"""The N>1 path on CPU: two gloo ranks shard the frames, each "reconstructs" its block with the oracle
standing in for the GPU call (the host-side sharding/gather logic is what is under test), gather, and
the result must equal the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deformation import sharded, workloads as W
from oracle.dgrad_oracle import TriangleDeformationOracle


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 240, 1000):
        for world in (1, 2, 3, 8):
            r = [sharded.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharded.shard_sizes(n, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    V, F, border = W.grid_mesh()
    o = TriangleDeformationOracle()
    o.set_target(V, F, cnsts=border)
    dg = torch.from_numpy(W.iid_dgrad(n_frames, len(F), sigma=0.05, seed=3))

    def compute(x):
        return torch.from_numpy(np.stack([o.get_mesh(f.numpy().astype(np.float64), vert_cnsts=V[border]) for f in x])
                                if len(x) else np.zeros((0, len(V), 3), np.float32))

    local, (lo, hi) = sharded.reconstruct_sharded(compute, [dg], n_frames)
    assert local.shape[0] == hi - lo
    full = sharded.all_gather_meshes(local, n_frames)
    root = sharded.gather_meshes(local, n_frames, dst=0)
    if rank == 0:
        q.put((full.numpy(), root.numpy()))
    else:
        assert root is None
        q.put((full.numpy(), None))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [6, 7])          # even and ragged split
def test_two_rank_shard_and_gather(n_frames):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    V, F, border = W.grid_mesh()
    o = TriangleDeformationOracle()
    o.set_target(V, F, cnsts=border)
    dg = W.iid_dgrad(n_frames, len(F), sigma=0.05, seed=3)
    ref = np.stack([o.get_mesh(f.astype(np.float64), vert_cnsts=V[border]) for f in dg])
    for full, root in got:
        assert np.array_equal(full, ref)
        if root is not None:
            assert np.array_equal(root, ref)


class _CpuRec:
    """Stand-in for deformation.Reconstructor in the gather pipeline test: the oracle computes, torch CPU tensors."""

    def __init__(self):
        self.V, self.F, self.border = W.grid_mesh()
        self.o = TriangleDeformationOracle()
        self.o.set_target(self.V, self.F, cnsts=self.border)
        self.n_verts = len(self.V)
        self.free = np.setdiff1d(np.arange(self.n_verts), self.border)
        self.n_free = len(self.free)

    def get_mesh_batch(self, x, out=None, free_only=False):
        full = np.stack([self.o.get_mesh(f.numpy().astype(np.float64), vert_cnsts=self.V[self.border]) for f in x])
        res = torch.from_numpy(full[:, self.free] if free_only else full)
        if out is not None:
            out.copy_(res)
            return out
        return res

    def expand_free(self, rows, out=None, stream=None):
        full = torch.from_numpy(np.broadcast_to(self.V, (rows.shape[0],) + self.V.shape).copy())
        full[:, torch.from_numpy(self.free)] = rows
        if out is not None:
            out.copy_(full)
            return out
        return full


def _pipe_worker(rank, world, port, n_local, chunk, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rec = _CpuRec()
    dg = torch.from_numpy(W.iid_dgrad(n_local, len(rec.F), sigma=0.05, seed=3, start=rank * n_local))
    res = {}
    for mode, expand in (("all", False), ("all", True), ("root", True)):
        pipe = sharded.GatherPipeline(rec, chunk_frames=chunk, mode=mode, dst=0, expand=expand)
        got = pipe.run(lambda x, out: rec.get_mesh_batch(x, out=out, free_only=True), [dg])
        res[(mode, expand)] = None if got is None else got.numpy()
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_pipeline_free_rows():
    """GatherPipeline (chunked compute, free rows only on the wire, expansion on the receiver): every receiving rank
    ends up with all frames in rank-major order; a ragged last chunk (5 frames in chunks of 2) included."""
    world, port, n_local, chunk = 2, _free_port(), 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pipe_worker, args=(r, world, port, n_local, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rec = _CpuRec()
    dg = torch.from_numpy(W.iid_dgrad(world * n_local, len(rec.F), sigma=0.05, seed=3))
    ref = rec.get_mesh_batch(dg).numpy()
    for rank in range(world):
        r = got[rank]
        assert np.array_equal(r[("all", False)], ref[:, rec.free])
        assert np.array_equal(r[("all", True)], ref)
        if rank == 0:
            assert np.array_equal(r[("root", True)], ref)
        else:
            assert r[("root", True)] is None
