"""MMA-stream op periods of one k_solve_tc tile with the least-perturbing profile (build the library with -DTS_PROF_LIGHT=1:
`make -C sdfa-2019_b200 OUT=lib_light EXTRA=-DTS_PROF_LIGHT=1`, then SDFA_LIB=.../lib_light/libsdfa_b200.so python tools/tensor_periods.py)."""
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
n = 75600
dg = torch.from_numpy(W.iid_dgrad(64, len(F), sigma=0.01, seed=0)).cuda().repeat((n + 63) // 64, 1)[:n].contiguous()
for _ in range(3):
    out = rec.get_mesh_batch(dg)
torch.cuda.synchronize()
pl = T.plan(rec)
ne, nm = len(pl["epi"]), len(pl["mma"])
raw = rec.debug("solve_prof").astype(np.float64)
m5 = raw[6 * ne:6 * ne + 5 * nm].reshape(-1, 5)
t = m5[:, 0]
d = np.diff(t)
mm = pl["mma"]
print("M op start-to-start: n", len(d), "sum", d.sum(), "median", np.median(d), "mean", d.mean())
noev = (mm["wait_epi"][:-1] < 0) & (mm["wait_epi2"][:-1] < 0)
print("ops without event wait: median period", np.median(d[noev]), "mean", d[noev].mean(), "n", noev.sum())
k=mm["k8"][:-1]*3
for N in (16,32,48,64):
    sel=noev&(mm["n"][:-1]==N)
    if sel.any(): print("N",N,"n",sel.sum(),"median period",np.median(d[sel]),"median per MMA",np.median(d[sel]/k[sel]))
