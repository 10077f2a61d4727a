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


# ---- checkpoint buffers + hparams.json (second half) ----------------------------------------

REF_API = "/root/reference/speech_anime/api.py"
REF_FRAME = "/root/reference/speech_anime/viewer/frame.py"


def _make_hparams(n_tris, ks, kr):
    """What ConfigDict.dump writes for config/model/dgrad.py:75-92 (only the keys the loader reads, plus the
    __entirety__ flags every level carries)."""
    return {"__entirety__": False, "model": {"__entirety__": True, "output": {
        "__entirety__": True, "layers_scale": [["fc", 520, 512], ["fc", 512, 256], ["fc", 256, ks, "act=linear"]],
        "layers_rotat": [["fc", 520, 512], ["fc", 512, 256], ["fc", 256, kr, "act=linear"]],
        "output_dim_scale": n_tris * 6, "output_dim_rotat": n_tris * 3, "using_pca": True}}}


def test_pca_from_checkpoint_current_and_legacy_keys(tmp_path):
    import json
    import torch
    cs, ms, cr, mr = W.random_pca(12, seed=3, k_scale=5, k_rotat=4)
    t = {k: torch.from_numpy(v) for k, v in dict(cs=cs, ms=ms, cr=cr, mr=mr).items()}
    new = {"_model._output_module._scale_pca.compT": t["cs"], "_model._output_module._scale_pca.means": t["ms"],
           "_model._output_module._rotat_pca.compT": t["cr"], "_model._output_module._rotat_pca.means": t["mr"],
           "_model._output_module._layers.0.weight": torch.zeros(3, 3)}
    old = {"anime_decoder.proj_scale.compT": t["cs"], "anime_decoder.proj_scale.means": t["ms"],
           "anime_decoder.proj_rotat.compT": t["cr"], "anime_decoder.proj_rotat.means": t["mr"],
           "audio_encoder.layers.0.weight": torch.zeros(2), "hamm": torch.zeros(4)}
    with open(tmp_path / "hparams.json", "w") as fp:
        json.dump(_make_hparams(12, 5, 4), fp)
    for name, state in (("new.ckpt", new), ("old.ckpt", old)):
        torch.save({"epoch": 3, "global_step": 77, "state": state}, tmp_path / name)
        got = FM.load_pca_from_checkpoint(tmp_path / name, hparams=str(tmp_path))
        assert all(g.dtype == np.float32 and g.flags.c_contiguous for g in got)
        for g, w in zip(got, (cs, ms, cr, mr)):
            assert np.array_equal(g, w)
    torch.save(new, tmp_path / "bare.ckpt")                      # a bare state dict
    assert np.array_equal(FM.load_pca_from_checkpoint(tmp_path / "bare.ckpt")[2], cr)
    with open(tmp_path / "hparams.json", "w") as fp:
        json.dump(_make_hparams(12, 6, 4), fp)                    # K mismatch -> refuse
    with pytest.raises(ValueError):
        FM.load_pca_from_checkpoint(tmp_path / "new.ckpt", hparams=str(tmp_path / "hparams.json"))
    torch.save({"state": {"x": torch.zeros(1)}}, tmp_path / "nopca.ckpt")
    with pytest.raises(KeyError):
        FM.load_pca_from_checkpoint(tmp_path / "nopca.ckpt")


@pytest.mark.skipif(not os.path.exists(REF_API), reason="reference tree not present (GPU box)")
def test_legacy_key_renaming_matches_reference_source():
    """ckpt_backward_compatible_preprocess (api.py:170-197), run from its own source, must lead to the same PCA keys."""
    tree = ast.parse(open(REF_API).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "ckpt_backward_compatible_preprocess"]
    ns = {"Any": object}
    exec(compile(ast.Module(body=fn, type_ignores=[]), REF_API, "exec"), ns)
    old = {f"anime_decoder.proj_{a}.{b}": (a, b) for a in ("scale", "rotat") for b in ("compT", "means")}
    old["hamm"] = None
    renamed = ns["ckpt_backward_compatible_preprocess"]({"state": dict(old)})["state"]
    assert set(renamed) == set(FM._PCA_KEYS.values())
    for a in ("scale", "rotat"):
        for b in ("compT", "means"):
            assert renamed[FM._PCA_KEYS[f"{b}_{a}"]] == (a, b)


def _write_template_files(tmp_path, n_faces):
    cpath, tpath = tmp_path / "cnst.txt", tmp_path / "corr.txt"
    with open(cpath, "w") as fp:
        fp.write("0 5 7\n  12\n3 9\n")
    recs = [(4, 0), (2, 3), (9, 3), (1, n_faces - 1), (7, 3), (5, 1)]
    with open(tpath, "w") as fp:
        fp.write(f"{len(recs)}\n")
        for s, d in recs:
            fp.write(f"{s},{d},0.5\n")
        fp.write("99,2,0.1\n")                                   # beyond the declared count: ignored (frame.py:67-68)
    return cpath, tpath


def test_constraint_and_tricorres_files(tmp_path):
    V, F, _ = W.grid_mesh()
    cpath, tpath = _write_template_files(tmp_path, len(F))
    c = FM.read_constraints(cpath)
    assert c.dtype == np.uint32 and c.tolist() == [0, 5, 7, 12, 3, 9]
    cor = FM.read_tricorres(tpath, len(F))
    cc, cf = cor["corr_count"], cor["corr_faces"]
    assert cc[0] == 1 and cc[1] == 1 and cc[2] == 0 and cc[3] == 3 and cc[len(F) - 1] == 1
    assert len(cf) == int(np.maximum(cc, 1).sum())                # one placeholder per empty target triangle
    assert cf[:6].tolist() == [4, 5, 0, 2, 9, 7]                   # file order kept inside a target triangle
    p = tmp_path / "t.obj"
    FM.write_obj(p, V, F)
    v, f, ci, co = FM.load_template(str(p), str(cpath), str(tpath))
    assert np.array_equal(v, V) and np.array_equal(f, F) and np.array_equal(ci, c) and np.array_equal(co["corr_faces"], cf)


@pytest.mark.skipif(not os.path.exists(REF_FRAME), reason="reference tree not present (GPU box)")
def test_template_file_parsers_match_reference_source(tmp_path):
    """set_template_mesh (viewer/frame.py:48-96) run from its own source with the mesh reader and set_dgrad_static stubbed."""
    V, F, _ = W.grid_mesh()
    cpath, tpath = _write_template_files(tmp_path, len(F))
    tree = ast.parse(open(REF_FRAME).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "set_template_mesh"]
    seen = {}

    class _Mesh:
        @staticmethod
        def read_mesh(path, dtype=None):
            return V, F

    class _Saber:
        mesh = _Mesh

    ns = {"saber": _Saber, "np": np, "os": os, "renderer": None,
          "set_dgrad_static": lambda verts, faces, c_indices, corres: seen.update(c=c_indices, corres=corres)}
    exec(compile(ast.Module(body=fn, type_ignores=[]), REF_FRAME, "exec"), ns)
    ns["set_template_mesh"]("t.obj", str(cpath), str(tpath))
    assert FM.read_constraints(cpath).tolist() == seen["c"]
    mine = FM.read_tricorres(tpath, len(F))
    assert mine["corr_count"].tolist() == seen["corres"]["corr_count"]
    assert mine["corr_faces"].tolist() == seen["corres"]["corr_faces"]
