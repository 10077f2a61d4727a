"""Target of the ncu captures (profiles/): the bench's batch (75 600 frames) through both batched entry points --
three decode + reconstruction calls (k_split16 x2, k_decode_tc16, k_assemble, k_solve_tc, k_output), then three calls
on the reference-layout dgrad (k_assemble_gather3, k_solve_tc, k_output).  L2 is flushed between calls like in bench.py."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "sdfa-2019_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import deformation as D  # noqa: E402
from deformation import workloads as W  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 75600
V, F, nfv, nft = W.load_flame()
rec = D.Reconstructor(V, F, cnsts=nfv, device=0)
rec.set_pca(*W.random_pca(len(F), seed=1, zero_tris=nft))
rec.set_option("pipe_chunk", 0)                      # one launch per kernel for the whole batch
xs, xr = (torch.from_numpy(a).cuda() for a in W.random_coeffs(n, seed=2))
out = torch.empty((n, len(V), 3), device="cuda")
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
for _ in range(3):
    flush.zero_()
    rec.decode_and_get_mesh(xs, xr, out=out)
dg = rec.decode_dgrad(xs, xr)
for _ in range(3):
    flush.zero_()
    rec.get_mesh_batch(dg, out=out)
torch.cuda.synchronize()
print("profile target ok:", n, "frames")
