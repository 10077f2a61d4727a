"""GPU check of the tensor-core solve (K3T) against the SIMT solve (K3) and the oracle on the same frames."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sdfa-2019_b200"))
import deformation as D                                  # noqa: E402
from deformation import workloads as W                   # noqa: E402
from oracle.dgrad_oracle import TriangleDeformationOracle  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
V, F, nfv, _ = W.load_flame()
dg = W.iid_dgrad(n, len(F), sigma=0.05, seed=0)
rt = D.Reconstructor(V, F, cnsts=nfv, device=0, solver="tensor")
print("ts_stats", rt.debug("ts_stats"), flush=True)
t = time.time()
a = rt.get_mesh_batch(dg)
print("tensor done", time.time() - t, "nan:", int(np.isnan(a).sum()), flush=True)
rs = D.Reconstructor(V, F, cnsts=nfv, device=0, solver="simt")
b = rs.get_mesh_batch(dg)
print("max |tensor - simt| =", float(np.abs(a - b).max()), flush=True)
d = np.abs(a - b).max(axis=(1, 2))
print("worst frames", np.argsort(d)[-5:], d[np.argsort(d)[-5:]])
o = TriangleDeformationOracle()
assert o.set_target(V, F, cnsts=nfv)
tol = 1e-6 * W.bbox_diag(V)
worst = 0.0
for i in sorted(set([0, 1, 31, 32, 127, 128, n - 1]) & set(range(n))):
    ref = o.get_mesh(dg[i].astype(np.float64), vert_cnsts=V[nfv])
    worst = max(worst, float(np.abs(a[i] - ref).max()))
print("max |tensor - oracle| =", worst, "tol", tol, "OK" if worst <= tol else "FAIL")
