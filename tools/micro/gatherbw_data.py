"""Writes tools/micro/_bin/flame_blocks.bin for gatherbw.cu: the FLAME assembly plan's row blocks as lists of source
triangles (int32: n_blocks, then per block: count, triangles...), in block-walk order."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "sdfa-2019_b200")]
import numpy as np
import deformation as D
from deformation import workloads as W
V, F, nfv, _ = W.load_flame()
rec = D.Reconstructor(V, F, cnsts=nfv, device=-1)
blocks = rec.debug("asm_blocks").reshape(-1, 5)
eq_id, eq_src = rec.debug("asm_eq_id"), rec.debug("eq_src")
out = [np.int32(len(blocks))]
for b in blocks:
    tris = eq_src[eq_id[b[0]:b[1]]]
    tris = tris[tris >= 0]
    out.append(np.int32(len(tris)))
    out.append(tris.astype(np.int32))
np.concatenate([np.atleast_1d(x) for x in out]).astype(np.int32).tofile(os.path.join(ROOT, "tools/micro/_bin/flame_blocks.bin"))
print("blocks", len(blocks), "records", sum(int(b[1] - b[0]) for b in blocks), "spans", [int(eq_src[eq_id[b[0]:b[1]]].max() - eq_src[eq_id[b[0]:b[1]]].min()) for b in blocks])
