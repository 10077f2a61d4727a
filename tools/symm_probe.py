"""Probe (torchrun, N >= 2): can this box map peer GPU memory through torch symmetric memory, and does the path's output
kernel write its free rows straight into the peers' buffers over NVLink?  Prints timings of the direct-store all-gather
against the NCCL pipeline."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sdfa-2019_b200")]
import numpy as np, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
import deformation as D
from deformation import sharded, workloads as W

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
V, F, nfv, nft = W.load_flame()
rec = D.Reconstructor(V, F, cnsts=nfv, device=local)
rec.set_pca(*W.random_pca(len(F), seed=1, zero_tris=nft))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 75600
xs, xr = W.random_coeffs(n, seed=2 + rank)
xs_d, xr_d = torch.from_numpy(xs).to(dev), torch.from_numpy(xr).to(dev)
t = symm_mem.empty((world * n, rec.n_free, 3), dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(f"rank {rank}: rendezvous ok, multicast {hdl.has_multicast_support}, ptrs {[hex(p) for p in hdl.buffer_ptrs][:4]}", flush=True)
peers = [hdl.get_buffer(r, t.shape, t.dtype) for r in range(world)]
mine = [p[rank * n:(rank + 1) * n] for p in peers]

def step_direct():
    # one decode + reconstruction per destination would redo the work: reconstruct once into the local buffer, then
    # push the finished free rows to every peer with device-side copies over NVLink (torch copy kernels on peer memory)
    rec.decode_and_get_mesh(xs_d, xr_d, out=mine[rank], free_only=True)
    for r in range(world):
        if r != rank:
            mine[r].copy_(mine[rank], non_blocking=True)
    hdl.barrier()

def step_into_root():
    rec.decode_and_get_mesh(xs_d, xr_d, out=mine[0], free_only=True)     # the output kernel itself stores into rank 0's memory
    hdl.barrier()

def timed(fn, k=5):
    for _ in range(2):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    dist.barrier(); torch.cuda.synchronize()
    v = torch.tensor([a.elapsed_time(b) / k], device=dev, dtype=torch.float64)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v)

ms_plain = timed(lambda: rec.decode_and_get_mesh(xs_d, xr_d, out=mine[rank], free_only=True))
ms_direct = timed(step_direct)
# check: rank r's block of everybody's buffer equals r's own reconstruction
torch.cuda.synchronize(); dist.barrier()
ref = rec.decode_and_get_mesh(xs_d, xr_d, free_only=True)
ok = bool(torch.equal(t[rank * n:(rank + 1) * n], ref))
nxt = (rank + 1) % world
xs2, xr2 = W.random_coeffs(n, seed=2 + nxt)
ref2 = rec.decode_and_get_mesh(torch.from_numpy(xs2).to(dev), torch.from_numpy(xr2).to(dev), free_only=True)
ok2 = bool(torch.equal(t[nxt * n:(nxt + 1) * n], ref2))
ms_root = timed(step_into_root)
pipe = sharded.GatherPipeline(rec, chunk_frames=148 * 128, mode="all", dst=0, expand=False)
g_out = torch.empty((world * n, rec.n_free, 3), dtype=torch.float32, device=dev)
ms_nccl = timed(lambda: pipe.run(lambda a, b, out: rec.decode_and_get_mesh(a, b, out=out, free_only=True), [xs_d, xr_d], out=g_out))
if rank == 0:
    f = lambda ms: f"{ms:.3f} ms = {world * n / ms / 1e3:.2f} M frames/s"
    print(f"world {world}, {n} frames per rank: local only {f(ms_plain)}; direct peer copies all-gather {f(ms_direct)} (own block ok {ok}, "
          f"neighbour block ok {ok2}); output kernel storing into rank 0 {f(ms_root)}; NCCL pipeline all-gather {f(ms_nccl)}")
dist.destroy_process_group()
