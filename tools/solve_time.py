"""Per-kernel times (solve, output, gather assembly) of a get_mesh_batch call (library's own CUDA events), for A/B runs: SDFA_LIB / SDFA_TS_* in the environment."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sdfa-2019_b200")]
import numpy as np, torch
import deformation as D
from deformation import workloads as W
V, F, nfv, nft = W.load_flame()
rec = D.Reconstructor(V, F, cnsts=nfv, device=0, solver="tensor")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 75600
dg = torch.from_numpy(W.iid_dgrad(64, len(F), sigma=0.01, seed=0)).cuda().repeat((n + 63) // 64, 1)[:n].contiguous()
out = torch.empty((n, len(V), 3), dtype=torch.float32, device="cuda")
for _ in range(3):
    rec.get_mesh_batch(dg, out=out)
torch.cuda.synchronize()
rec.set_timing(True)
t, to, ta = [], [], []
for _ in range(8):
    rec.get_mesh_batch(dg, out=out)
    lt = rec.last_timing()
    t.append(lt["solve_ms"]); to.append(lt["output_ms"]); ta.append(lt["assembly_ms"])
ref = D.Reconstructor(V, F, cnsts=nfv, device=0, solver="simt")
rec.set_timing(False)
a = rec.get_mesh_batch(dg[:4096]); b = ref.get_mesh_batch(dg[:4096])
print(f"solve_ms min {min(t):.4f} median {sorted(t)[len(t)//2]:.4f}  output_ms min {min(to):.4f}  assembly(gather)_ms min {min(ta):.4f}  |tensor-simt| {float((a-b).abs().max()):.3e}  "
      f"env {dict((k, v) for k, v in os.environ.items() if k.startswith('SDFA_'))}")
