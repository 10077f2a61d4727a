"""Config 3 of BASELINE.json: batched reconstruction sweep over the RHS batch size (frames per call), dgrad resident in
HBM, per-kernel CUDA-event times from the library; reports frames/s and the solve's GB/s against its 30 264 B/frame."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sdfa-2019_b200")]
import numpy as np, torch
import deformation as D
from deformation import workloads as W
V, F, nfv, nft = W.load_flame()
solver = sys.argv[1] if len(sys.argv) > 1 else None
rec = D.Reconstructor(V, F, cnsts=nfv, device=0, solver=solver)
cs, ms, cr, mr = W.random_pca(len(F), seed=1, zero_tris=nft)
rec.set_pca(cs, ms, cr, mr)
print("solver:", "tensor (K3T)" if rec.debug("ts_stats")[0] else "simt (K3)")
print(f"{'batch':>6s} {'frames/s':>12s} {'decode':>8s} {'assembly':>9s} {'solve':>8s} {'output':>8s}  solve GB/s (of 6553)")
rec.set_timing(True)
for n in (64, 128, 256, 512, 1024, 2048, 4096, 10000, 75600):
    xs, xr = W.random_coeffs(n, seed=3)
    xs, xr = torch.from_numpy(xs).cuda(), torch.from_numpy(xr).cuda()
    out = torch.empty((n, 5023, 3), device="cuda")
    acc = np.zeros(4)
    reps = 8
    for i in range(reps + 2):
        rec.decode_and_get_mesh(xs, xr, out=out)
        if i >= 2:
            t = rec.last_timing()
            acc += np.array([t["decode_ms"], t["assembly_ms"], t["solve_ms"], t["output_ms"]]) / reps
    tot = acc.sum()
    print(f"{n:6d} {n / tot * 1e3:12.0f} {acc[0]:8.3f} {acc[1]:9.3f} {acc[2]:8.3f} {acc[3]:8.3f}  {30264 * n / acc[2] / 1e6:8.0f}")
