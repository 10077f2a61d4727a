"""GPU tier, multi-GPU: the north star's only collective -- the NCCL gather of the vertex buffers -- on real devices
(world size 2, one process per GPU; skipped on a one-GPU box; run with ``gpurun --gpus 2``).

Each rank reconstructs its frame shard on its own GPU (replicated plan), the free rows travel over NCCL on a side
stream chunk by chunk (deformation/sharded.py: GatherPipeline) and are expanded on the receiver; the result must
equal the one-GPU reconstruction of all frames bit for bit, and sampled frames must match the reference."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_local, chunk, q):
    import torch
    import torch.distributed as dist
    import deformation as D
    from deformation import sharded, workloads as W
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    V, F, nfv, nft = W.load_flame()
    rec = D.Reconstructor(V, F, cnsts=nfv, device=rank)
    rec.set_pca(*W.random_pca(len(F), seed=1, zero_tris=nft))
    xs, xr = W.random_coeffs(world * n_local, seed=2)
    lo, hi = rank * n_local, (rank + 1) * n_local
    xs_d, xr_d = torch.from_numpy(xs[lo:hi]).cuda(), torch.from_numpy(xr[lo:hi]).cuda()
    res = {}
    for mode, expand in (("all", False), ("all", True), ("root", True)):
        pipe = sharded.GatherPipeline(rec, chunk_frames=chunk, mode=mode, dst=0, expand=expand)
        for _ in range(2):                                    # twice: buffers and streams are reused
            got = pipe.run(lambda a, b, out: rec.decode_and_get_mesh(a, b, out=out, free_only=True), [xs_d, xr_d])
        torch.cuda.synchronize()
        res[(mode, expand)] = None if got is None else got.cpu().numpy()
    # the same gather without NCCL on the data path: symmetric memory + peer-to-peer pushes (deformation/sharded.py: PeerGather)
    for mode in ("all", "root"):
        pg = sharded.PeerGather(rec, n_local, chunk_frames=chunk, mode=mode, dst=0)
        for _ in range(2):
            got = pg.run(lambda a, b, out: rec.decode_and_get_mesh(a, b, out=out, free_only=True), [xs_d, xr_d])
        torch.cuda.synchronize()
        res[("p2p", mode)] = None if got is None else got.cpu().numpy()
    # the plain collectives on device tensors as well (ragged shards: 7 frames over 2 ranks)
    a, b = sharded.shard_range(7, rank, world)
    local = rec.decode_and_get_mesh(torch.from_numpy(xs[a:b]).cuda(), torch.from_numpy(xr[a:b]).cuda())
    res["ragged_all"] = sharded.all_gather_meshes(local, 7).cpu().numpy()
    root = sharded.gather_meshes(local, 7, dst=0)
    res["ragged_root"] = None if root is None else root.cpu().numpy()
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_gather_of_free_rows_two_gpus(flame):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import deformation as D
    from deformation import workloads as W
    from oracle import ref_loader
    world, n_local, chunk = 2, 1000, 384                     # 3 chunks per rank, the last one ragged (232 frames)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_local, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    V, F, nfv, nft = flame["V"], flame["F"], flame["nfv"], flame["nft"]
    rec = D.Reconstructor(V, F, cnsts=nfv, device=0)
    pca = W.random_pca(len(F), seed=1, zero_tris=nft)
    rec.set_pca(*pca)
    xs, xr = W.random_coeffs(world * n_local, seed=2)
    ref = rec.decode_and_get_mesh(xs, xr)                     # one GPU, all frames
    ids = rec.free_vertices
    for rank in range(world):
        r = got[rank]
        assert np.array_equal(r[("all", False)], ref[:, ids])
        assert np.array_equal(r[("all", True)], ref)
        assert np.array_equal(r["ragged_all"], ref[:7])
        assert np.array_equal(r[("p2p", "all")], ref[:, ids])         # peer-to-peer pushes into symmetric memory
        assert (r[("p2p", "root")] is None) == (rank != 0)
        if rank == 0:
            assert np.array_equal(r[("root", True)], ref)
            assert np.array_equal(r["ragged_root"], ref[:7])
            assert np.array_equal(r[("p2p", "root")], ref[:, ids])
        else:
            assert r[("root", True)] is None and r["ragged_root"] is None
    if ref_loader.ref_available():
        from oracle.dgrad_oracle import pca_decode
        o = ref_loader.RefSolver(1)
        o.set_target(V, F, cnsts=nfv)
        dg = pca_decode(xs, *pca[:2], xr, *pca[2:], dtype=np.float32)
        for i in (0, n_local - 1, n_local, world * n_local - 1):
            want = o.get_mesh(dg[i].astype(np.float64), vert_cnsts=V[nfv])
            assert np.abs(got[0][("all", True)][i] - want).max() <= flame["tol"]
