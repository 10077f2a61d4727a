"""Config 5 of BASELINE.json: the twice-subdivided FLAME template (79 936 v / 159 616 tris, 20 653 unknowns), synthetic
dgrad resident in HBM; throughput and per-kernel times (SIMT solve plan: the tensor plan declines > 2560 unknowns)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sdfa-2019_b200")]
import numpy as np, torch
import deformation as D
from deformation import workloads as W
V, F, c = W.flame_sub2()
rec = D.Reconstructor(V, F, cnsts=c, device=0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
dg = torch.from_numpy(W.iid_dgrad(16, len(F), sigma=0.01, seed=5)).cuda().repeat((n + 15) // 16, 1)[:n].contiguous()
out = torch.empty((n, len(V), 3), device="cuda")
rec.set_timing(True)
acc = np.zeros(3)
for i in range(5):
    rec.get_mesh_batch(dg, out=out)
    if i >= 2:
        t = rec.last_timing()
        acc += np.array([t["assembly_ms"], t["solve_ms"], t["output_ms"]]) / 3
print(f"config 5: {n} frames, frames per solve tile {int(rec.debug('stats')[14])}, assembly {acc[0]:.2f} ms, solve {acc[1]:.2f} ms, "
      f"output {acc[2]:.2f} ms -> {n / acc.sum() * 1e3:.0f} frames/s; path bytes 2 457 408 B/frame -> "
      f"{2457408 * n / acc.sum() / 1e6:.0f} GB/s")
