"""Config 5 (subdivided FLAME, SIMT sweeps): per-kernel times and the cycle breakdown of k_solve's consumer warp 0
(SDFA_SOLVE_PROFILE=1) under the planner knobs given in the environment (SDFA_FRAMES_PER_TILE, SDFA_PIECE_CAP,
SDFA_SUPERNODE_CAP, SDFA_SUBTREE_CAP)."""
import os, sys
os.environ["SDFA_SOLVE_PROFILE"] = "1"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sdfa-2019_b200")]
import numpy as np, torch
import deformation as D
from deformation import workloads as W
V, F, c = W.flame_sub2()
rec = D.Reconstructor(V, F, cnsts=c, device=0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
dg = torch.from_numpy(W.iid_dgrad(16, len(F), sigma=0.01, seed=5)).cuda().repeat((n + 15) // 16, 1)[:n].contiguous()
out = torch.empty((n, len(V), 3), device="cuda")
rec.set_timing(True)
acc = np.zeros(3)
for i in range(4):
    rec.get_mesh_batch(dg, out=out)
    if i >= 2:
        t = rec.last_timing()
        acc += np.array([t["assembly_ms"], t["solve_ms"], t["output_ms"]]) / 2
st = rec.debug("stats")
knobs = {k: os.environ.get(k) for k in ("SDFA_FRAMES_PER_TILE", "SDFA_PIECE_CAP", "SDFA_SUPERNODE_CAP", "SDFA_SUBTREE_CAP") if os.environ.get(k)}
print(f"knobs {knobs} frames {n} F {int(st[14])} stats {[int(x) for x in st[:15]]}")
print(f"  assembly {acc[0]:.2f} ms, solve {acc[1]:.2f} ms, output {acc[2]:.2f} ms -> {n / acc.sum() * 1e3:.0f} frames/s, solve alone {n / acc[1] * 1e3:.0f} frames/s")
p = rec.debug("solve_prof").reshape(-1, 8)
p = p[p[:, 0] > 0]
for i, nm in enumerate(["total", "wait_stage", "wait_rows_loaded", "tasks", "level_barrier"]):
    print(f"  {nm:18s} mean {p[:, i].mean():12.0f} cycles   share {100 * p[:, i].sum() / p[:, 0].sum():5.1f}%")
