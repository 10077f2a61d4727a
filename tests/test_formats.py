"""CPU tier: on-disk formats around the path against the reference's own reader / writer sources
when the reference tree is present, and through round trips always."""
import ast
import os

import numpy as np
import pytest

from deformation import formats as FM
from deformation import workloads as W

REF_IO = "/root/reference/saber/data/mesh/io.py"


def _ref_functions(names):
    """The reference module imports `plyfile` (absent here) at the top: take the wanted functions out of its source."""
    tree = ast.parse(open(REF_IO).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = {"np": np, "os": os}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF_IO, "exec"), ns)
    return [ns[n] for n in names]


def test_obj_round_trip_and_export_layout(tmp_path):
    V, F, _ = W.grid_mesh()
    p = tmp_path / "m.obj"
    FM.write_obj(p, V, F)
    v2, f2 = FM.read_obj(p)
    assert np.array_equal(v2, V) and np.array_equal(f2, F)          # str(float32) round-trips exactly
    vf, ff = FM.read_mesh(str(p), flatten=True)
    assert vf.shape == (V.size,) and ff.shape == (F.size,)
    meshes = np.stack([V, V + np.float32(0.5)])
    dg = np.arange(2 * 9 * len(F), dtype=np.float32).reshape(2, -1)
    FM.export_frames(tmp_path / "out", meshes, F, dgrad=dg, start=7)
    assert sorted(os.listdir(tmp_path / "out")) == ["000007.obj", "000007_dgrad.npy", "000008.obj", "000008_dgrad.npy"]
    assert np.array_equal(FM.read_obj(tmp_path / "out" / "000008.obj")[0], meshes[1])
    assert np.array_equal(np.load(tmp_path / "out" / "000008_dgrad.npy"), dg[1])


@pytest.mark.skipif(not os.path.exists(REF_IO), reason="reference tree not present (GPU box)")
def test_obj_text_and_reader_match_reference_source(tmp_path):
    ref_write, ref_read = _ref_functions(["write_obj", "read_obj"])
    V, F, _ = W.grid_mesh()
    a, b = tmp_path / "a.obj", tmp_path / "b.obj"
    FM.write_obj(a, V, F)
    ref_write(b, V, F)
    assert open(a).read() == open(b).read()
    with open(a, "a") as fp:
        fp.write("\n# a quad with texture indices\nf 1/1 2/2 3/3 4/4\n")
    v1, f1 = FM.read_obj(a)
    v2, f2 = ref_read(str(a), flatten=False)
    assert np.array_equal(v1, v2) and np.array_equal(f1, f2) and len(f1) == len(F) + 2


def test_ply_binary_and_ascii(tmp_path):
    """The speaker templates are binary little-endian `float x,y,z` + `uchar 3, int32 x 3`."""
    V, F, _ = W.grid_mesh()
    hdr = ("ply\nformat {}\ncomment made for the test\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
           "element face %d\nproperty list uchar int vertex_indices\nend_header\n") % (len(V), len(F))
    pb = tmp_path / "b.ply"
    with open(pb, "wb") as fp:
        fp.write(hdr.format("binary_little_endian 1.0").encode())
        fp.write(V.astype("<f4").tobytes())
        for f in F:
            fp.write(bytes([3]) + np.asarray(f, dtype="<i4").tobytes())
    pa = tmp_path / "a.ply"
    with open(pa, "w") as fp:
        fp.write(hdr.format("ascii 1.0"))
        for v in V:
            fp.write(f"{v[0]!r} {v[1]!r} {v[2]!r}\n".replace("np.float32(", "").replace(")", ""))
        for f in F:
            fp.write(f"3 {f[0]} {f[1]} {f[2]}\n")
    for p in (pb, pa):
        v, f = FM.read_mesh(str(p))
        assert v.dtype == np.float32 and f.dtype == np.uint32
        assert np.array_equal(v, V) and np.array_equal(f, F)


def test_pca_directory_layout(tmp_path):
    cs, ms, cr, mr = W.random_pca(12, seed=3, k_scale=5, k_rotat=4)
    FM.save_pca(tmp_path, cs.astype(np.float64), ms, cr, mr)
    assert sorted(os.listdir(tmp_path / "pca")) == ["rotat_compT.npy", "rotat_means.npy", "scale_compT.npy", "scale_means.npy"]
    got = FM.load_pca(tmp_path)
    assert all(g.dtype == np.float32 for g in got)
    assert np.array_equal(got[0], cs) and np.array_equal(got[1], ms) and np.array_equal(got[2], cr) and np.array_equal(got[3], mr)
