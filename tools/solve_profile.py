"""Cycle breakdown of k_solve's consumer warp 0 (SDFA_SOLVE_PROFILE=1): where a tile's time goes."""
import os, sys
os.environ["SDFA_SOLVE_PROFILE"] = "1"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sdfa-2019_b200")]
import numpy as np, torch
import deformation as D
from deformation import workloads as W
V, F, nfv, nft = W.load_flame()
rec = D.Reconstructor(V, F, cnsts=nfv, device=0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15360
dg = torch.from_numpy(W.iid_dgrad(64, len(F), sigma=0.01, seed=0)).cuda().repeat((n + 63) // 64, 1)[:n].contiguous()
for _ in range(3):
    out = rec.get_mesh_batch(dg)
torch.cuda.synchronize()
p = rec.debug("solve_prof").reshape(-1, 8)
p = p[p[:, 0] > 0]
tiles = (n + 31) // 32
print("CTAs", len(p), "tiles", tiles, "stats", rec.debug("stats")[:6])
names = ["total", "wait_stage", "wait_rows_loaded", "tasks", "level_barrier"]
for i, nm in enumerate(names):
    print(f"{nm:18s} mean {p[:, i].mean():12.0f} cycles   max {p[:, i].max():12.0f}   share {100 * p[:, i].sum() / p[:, 0].sum():5.1f}%")
