"""Timeline of one tile of k_solve_tc (SDFA_SOLVE_PROFILE=1): per EPI / MMA op start, end-of-wait and end clocks."""
import os, sys
os.environ["SDFA_SOLVE_PROFILE"] = "1"
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sdfa-2019_b200")]
import numpy as np, torch
import deformation as D
from deformation import workloads as W
from tests import tplan_emulator as T
V, F, nfv, nft = W.load_flame()
rec = D.Reconstructor(V, F, cnsts=nfv, device=0, solver="tensor")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 71040
dg = torch.from_numpy(W.iid_dgrad(64, len(F), sigma=0.01, seed=0)).cuda().repeat((n + 63) // 64, 1)[:n].contiguous()
for _ in range(3):
    out = rec.get_mesh_batch(dg)
torch.cuda.synchronize()
pl = T.plan(rec)
ne, nm = len(pl["epi"]), len(pl["mma"])
raw = rec.debug("solve_prof").astype(np.float64)
e6 = raw[:6 * ne].reshape(-1, 6)
m = raw[6 * ne:6 * ne + 3 * nm].reshape(-1, 3)
t0 = min(e6[e6[:, 0] > 0, 0].min(), m[m[:, 0] > 0, 0].min())
e6 -= t0; m -= t0
e = e6[:, [0, 2, 5]]
print(f"tile: EPI stream {e[-1,2]-e[0,0]:.0f} clk, MMA stream {m[-1,2]-m[0,0]:.0f} clk")
print(f"EPI: waiting {np.sum(e[:,1]-e[:,0]):.0f}  body {np.sum(e[:,2]-e[:,1]):.0f}")
print(f"MMA: waiting {np.sum(m[:,1]-m[:,0]):.0f}  issue {np.sum(m[:,2]-m[:,1]):.0f}")
fl = pl["epi"]["flags"]
d = np.diff(e6, axis=1)
print("  per EPI op means [rows from ring, wait mma, tmem ld, global store, tmem st + signal]")
for name, mask in (("load(ring->tmem)", (fl & 2 > 0) & (fl & 1 == 0) & (fl & 4 == 0)), ("sepfin", (fl & 1 > 0) & (fl & 16 > 0) & (fl & 4 == 0)),
                   ("readu(tmem->global)", (fl & 7) == 5), ("xout", (fl & 6) == 6)):
    if mask.any():
        print(f"  {name:22s} n={mask.sum():3d} ", " ".join(f"{x:7.0f}" for x in d[mask].mean(0)))
if len(sys.argv) > 2:
    for i in range(ne):
        print("E", i, int(fl[i]), int(pl["epi"]["n_chunks"][i]), *(int(x) for x in e6[i]))
    for i in range(nm):
        print("M", i, int(pl["mma"]["n"][i]), int(pl["mma"]["k8"][i]), *(int(x) for x in m[i]))
