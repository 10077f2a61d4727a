"""Target of the compute-sanitizer runs (tools/run_sanitizers.sh): the smoke call plus one batch large enough that
every persistent CTA of K1 / K3T walks more than one tile, through both batched entry points and both solvers."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "sdfa-2019_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import deformation as D  # noqa: E402
from deformation import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 13000
V, F, nfv, nft = W.load_flame()
for solver in ("tensor", "simt"):
    rec = D.Reconstructor(V, F, cnsts=nfv, device=0, solver=solver)
    rec.set_pca(*W.random_pca(len(F), seed=1, zero_tris=nft))
    xs, xr = (torch.from_numpy(a).cuda() for a in W.random_coeffs(n, seed=2))
    out = rec.decode_and_get_mesh(xs, xr)
    free = rec.decode_and_get_mesh(xs, xr, free_only=True)
    back = rec.expand_free(free)
    torch.cuda.synchronize()
    assert torch.equal(back, out)
    m = min(n, 2000)
    dg = torch.from_numpy(W.iid_dgrad(64, len(F), sigma=0.02, seed=3)).cuda().repeat((m + 63) // 64, 1)[:m].contiguous()
    out2 = rec.get_mesh_batch(dg)
    one = rec.get_mesh(np.zeros(len(F) * 9), vert_cnsts=V[nfv])
    torch.cuda.synchronize()
    assert np.abs(one - V).max() < 1e-8 and bool(torch.isfinite(out2).all())
    g = D.get_deform_grad_batch(V, out[:4], F)
    torch.cuda.synchronize()
    rec.close()
    print(f"sanitize target ok: solver={solver}, {n} frames")
