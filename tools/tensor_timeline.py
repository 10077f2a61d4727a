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
m5 = raw[6 * ne:6 * ne + 5 * nm].reshape(-1, 5)
m = m5[:, [0, 2, 4]]
t0 = min(e6[e6[:, 0] > 0, 0].min(), m[m[:, 0] > 0, 0].min())
e6 -= t0; m -= t0; m5 = m5 - t0
e = e6[:, [0, 2, 5]]
st = pl["epi"]["stream"]
print(f"tile: EPI streams {e[:,2].max()-e[:,0].min():.0f} clk, MMA stream {m[-1,2]-m[0,0]:.0f} clk")
for q in (0, 1):
    k = st == q
    if k.any():
        print(f"EPI stream {q}: {k.sum()} ops, span {e[k,2].max()-e[k,0].min():.0f}  waiting for events {np.sum(e[k,1]-e[k,0]):.0f}  body {np.sum(e[k,2]-e[k,1]):.0f}")
print(f"MMA: waiting {np.sum(m[:,1]-m[:,0]):.0f}  issue {np.sum(m[:,2]-m[:,1]):.0f}")
dm = np.diff(m5, axis=1)
gap = m5[1:, 0] - m5[:-1, 4]
print("MMA per-op sums [event waits, chunk wait, issue loop, commits] + gaps:", " ".join(f"{x:.0f}" for x in dm.sum(0)), f"{gap.sum():.0f}")
mm = pl["mma"]
noev = (mm["wait_epi"] < 0) & (mm["wait_epi2"] < 0)
first = (mm["flags"] & 2) > 0
print(f"  event-wait phase, ops without an event {dm[noev,0].mean():.0f}, with {dm[~noev,0].mean():.0f} (median {np.median(dm[~noev,0]):.0f});"
      f" chunk phase, not first {dm[~first,1].mean():.0f}, first {dm[first,1].mean():.0f} (median {np.median(dm[first,1]):.0f});"
      f" commit phase by commits 0/1/2: " + " ".join(f"{dm[((mm['flags'] & 4) > 0).astype(int) + (mm['commit_mma'] >= 0).astype(int) == k, 3].mean():.0f}" for k in (0, 1, 2))
      + f"; gap mean {gap.mean():.0f}")
fl = pl["epi"]["flags"]
d = np.diff(e6, axis=1)
print("  per EPI op means [-, wait events, tmem ld, rows from ring + global store, tmem st + signal]")
for name, mask in (("load(ring->tmem)", (fl & 2 > 0) & (fl & 1 == 0) & (fl & 4 == 0)), ("sepfin", (fl & 1 > 0) & (fl & 16 > 0) & (fl & 4 == 0)),
                   ("readu(tmem->global)", (fl & 7) == 5), ("xout", (fl & 6) == 6)):
    if mask.any():
        print(f"  {name:22s} n={mask.sum():3d} ", " ".join(f"{x:7.0f}" for x in d[mask].mean(0)))
if len(sys.argv) > 2:
    for i in range(ne):
        print("E", i, int(fl[i]), int(pl["epi"]["n_chunks"][i]), *(int(x) for x in e6[i]))
    for i in range(nm):
        print("M", i, int(pl["mma"]["n"][i]), int(pl["mma"]["k8"][i]), *(int(x) for x in m5[i]))
