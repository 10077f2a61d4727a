"""Config 5 at more frames than one chunk holds: the chunked path (8448-frame chunks for this template) returns exactly what a
single pass returns; prints throughput at 9000 frames.  python tools/config5_chunked_check.py (B200)."""
import os, sys
sys.path[:0] = ["/root/repo", "/root/repo/sdfa-2019_b200"]
import numpy as np, torch
import deformation as D
from deformation import workloads as W
V, F, c = W.flame_sub2()
rec = D.Reconstructor(V, F, cnsts=c, device=0)
n = 9000
dg = torch.from_numpy(W.iid_dgrad(16, len(F), sigma=0.01, seed=5)).cuda().repeat((n + 15) // 16, 1)[:n].contiguous()
out = torch.empty((n, len(V), 3), device="cuda")
os.environ["SDFA_PIPE_CHUNK"] = "0"
rec.get_mesh_batch(dg, out=out); torch.cuda.synchronize()
ref = out.clone()
del os.environ["SDFA_PIPE_CHUNK"]
out.zero_()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
rec.get_mesh_batch(dg, out=out); torch.cuda.synchronize()
ev[0].record(); rec.get_mesh_batch(dg, out=out); ev[1].record(); torch.cuda.synchronize()
print("chunked == single pass:", torch.equal(out, ref), " frames/s", n / ev[0].elapsed_time(ev[1]) * 1e3, " peak mem GB", torch.cuda.max_memory_allocated() / 2**30)
