"""On-disk formats either side of the dgrad -> mesh path, restated from the reference so that the
new path can read what the reference reads and write what it writes:

* OBJ / PLY templates        saber/data/mesh/io.py:6-84 (read_ply needs `plyfile` there; here a small reader of the
                             binary-little-endian / ascii layouts the shipped templates use)
* per-frame export           speech_anime/model/model.py:207-212: ``%06d.obj`` + ``%06d_dgrad.npy``
* PCA bases                  speech_anime/datasets/vocaset/preload.py:890-893, 939-951:
                             ``<dgrad_root>/pca/{scale,rotat}_{compT,means}.npy``

Host-side Python like the reference's; nothing here is on the GPU path.
"""
import os
import struct

import numpy as np

from .workloads import read_obj as _read_obj

_PLY_TYPES = {"char": "b", "int8": "b", "uchar": "B", "uint8": "B", "short": "h", "int16": "h", "ushort": "H",
              "uint16": "H", "int": "i", "int32": "i", "uint": "I", "uint32": "I", "float": "f", "float32": "f",
              "double": "d", "float64": "d"}


def read_obj(path, dtype=np.float32, flatten=False):
    """io.py:23-68: `v x y z` lines, `f` polygons fan-triangulated, 1-based -> 0-based uint32."""
    verts, faces = _read_obj(path, dtype=dtype)
    return (verts.reshape(-1), faces.reshape(-1)) if flatten else (verts, faces)


def read_ply(path, dtype=np.float32, flatten=False):
    """io.py:6-20: vertex x/y/z -> [n,3] dtype, face vertex_indices -> [m,3] uint32 (triangles only, like np.stack there)."""
    with open(path, "rb") as fp:
        if fp.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements = None, []
        while True:
            t = fp.readline().decode("ascii").split()
            if not t or t[0] == "comment":
                continue
            if t[0] == "format":
                fmt = t[1]
            elif t[0] == "element":
                elements.append({"name": t[1], "count": int(t[2]), "props": []})
            elif t[0] == "property":
                elements[-1]["props"].append(t[1:])
            elif t[0] == "end_header":
                break
        data = {}
        end = {"binary_little_endian": "<", "binary_big_endian": ">"}.get(fmt)
        for el in elements:
            rows = []
            scalar_only = all(p[0] != "list" for p in el["props"])
            if end and scalar_only:                           # one structured read
                dt = np.dtype([(p[1], end + _PLY_TYPES[p[0]]) for p in el["props"]])
                rows = np.frombuffer(fp.read(dt.itemsize * el["count"]), dtype=dt)
                data[el["name"]] = {n: rows[n] for n in dt.names}
                continue
            cols = {p[-1]: [] for p in el["props"]}
            for _ in range(el["count"]):
                tok = None if end else fp.readline().split()
                at = 0
                for p in el["props"]:
                    if p[0] == "list":
                        if end:
                            (k,) = struct.unpack(end + _PLY_TYPES[p[1]], fp.read(struct.calcsize(_PLY_TYPES[p[1]])))
                            vals = struct.unpack(end + str(k) + _PLY_TYPES[p[2]], fp.read(k * struct.calcsize(_PLY_TYPES[p[2]])))
                        else:
                            k = int(tok[at]); vals = [int(x) for x in tok[at + 1:at + 1 + k]]; at += 1 + k
                        cols[p[-1]].append(vals)
                    else:
                        if end:
                            (v,) = struct.unpack(end + _PLY_TYPES[p[0]], fp.read(struct.calcsize(_PLY_TYPES[p[0]])))
                        else:
                            v = float(tok[at]); at += 1
                        cols[p[-1]].append(v)
            data[el["name"]] = cols
    v = data["vertex"]
    verts = np.stack((np.asarray(v["x"]), np.asarray(v["y"]), np.asarray(v["z"])), axis=1).astype(dtype)
    key = "vertex_indices" if "vertex_indices" in data["face"] else "vertex_index"
    faces = np.stack([np.asarray(f) for f in data["face"][key]], axis=0).astype(np.uint32)
    return (verts.reshape(-1), faces.reshape(-1)) if flatten else (verts, faces)


def read_mesh(fname, dtype=np.float32, flatten=False):
    """io.py:78-83."""
    ext = os.path.splitext(fname)[1]
    if ext == ".obj":
        return read_obj(fname, dtype=dtype, flatten=flatten)
    if ext == ".ply":
        return read_ply(fname, dtype=dtype, flatten=flatten)
    raise NotImplementedError(f"Cannot read '{fname}'.")


def write_obj(fname, verts, faces):
    """io.py:71-76, byte for byte: numbers are printed with str() of the array's scalars, faces 1-based."""
    verts, faces = np.asarray(verts).reshape(-1, 3), np.asarray(faces).reshape(-1, 3)
    lines = [f"v {v[0]} {v[1]} {v[2]}\n" for v in verts]
    lines += [f"f {f[0] + 1} {f[1] + 1} {f[2] + 1}\n" for f in faces]
    with open(fname, "w") as fp:
        fp.writelines(lines)


def export_frames(export_dir, verts, faces, dgrad=None, start=0):
    """model.py:207-212 for a whole batch: ``%06d.obj`` per frame of verts [N, n_verts, 3] and, if given,
    ``%06d_dgrad.npy`` per row of dgrad [N, 9*n_tris]."""
    os.makedirs(export_dir, exist_ok=True)
    verts = np.asarray(verts)
    for i in range(len(verts)):
        write_obj(os.path.join(export_dir, f"{start + i:06d}.obj"), verts[i], faces)
        if dgrad is not None:
            np.save(os.path.join(export_dir, f"{start + i:06d}_dgrad.npy"), np.asarray(dgrad[i]))


def save_pca(dgrad_root, compT_scale, means_scale, compT_rotat, means_rotat):
    """preload.py:939-951 file layout."""
    d = os.path.join(dgrad_root, "pca")
    os.makedirs(d, exist_ok=True)
    for name, a in (("scale_compT", compT_scale), ("scale_means", means_scale), ("rotat_compT", compT_rotat), ("rotat_means", means_rotat)):
        np.save(os.path.join(d, name + ".npy"), np.asarray(a))


def load_pca(dgrad_root):
    """preload.py:890-893 -> (compT_scale [6*n_tris, Ks], means_scale, compT_rotat [3*n_tris, Kr], means_rotat) as float32,
    ready for ``set_pca`` (PcaInversion registers them as float buffers, output_module.py:103-113)."""
    d = os.path.join(dgrad_root, "pca")
    return tuple(np.load(os.path.join(d, n + ".npy")).astype(np.float32)
                 for n in ("scale_compT", "scale_means", "rotat_compT", "rotat_means"))


# ---------------------------------------------------------------------------------------------------------------
# A trained model's PCA bases straight from its checkpoint (second half).
# The reference registers them as buffers of the two PcaInversion modules (output_module.py:103-113), so they travel
# inside every ``.ckpt`` under ``state`` (saber/trainer/manager/checkpoints.py:14-25, :52-57); old checkpoints carry
# other key names, which api.py:170-197 renames before ``load_state_dict``.

# api.py:173-188, the pairs that can touch the PCA buffers' keys (the others rename encoder layers)
_CKPT_LEGACY_KEYS = (("anime_decoder.proj_scale", "_model._output_module._scale_pca"),
                     ("anime_decoder.proj_rotat", "_model._output_module._rotat_pca"))
_PCA_KEYS = {"compT_scale": "_model._output_module._scale_pca.compT", "means_scale": "_model._output_module._scale_pca.means",
             "compT_rotat": "_model._output_module._rotat_pca.compT", "means_rotat": "_model._output_module._rotat_pca.means"}


def load_hparams(path):
    """``hparams.json`` as written by ConfigDict.dump (saber/utils/config_dict.py:221-243): plain JSON, every level
    carrying an ``__entirety__`` flag.  ``path`` may be the file or the log directory holding it (experiment.py:32)."""
    import json
    if os.path.isdir(path):
        path = os.path.join(path, "hparams.json")
    with open(path) as fp:
        return json.load(fp)


def pca_dims_from_hparams(hparams):
    """(output_dim_scale, output_dim_rotat, k_scale, k_rotat) the model was built with
    (config/model/dgrad.py:75-92: the last fc of layers_scale / layers_rotat gives K, output_dim_* the rows)."""
    out = hparams["model"]["output"]
    return (int(out["output_dim_scale"]), int(out["output_dim_rotat"]),
            int(out["layers_scale"][-1][2]), int(out["layers_rotat"][-1][2]))


def load_pca_from_checkpoint(ckpt_path, hparams=None):
    """-> (compT_scale [6*n_tris, Ks], means_scale, compT_rotat [3*n_tris, Kr], means_rotat) float32, ready for
    ``set_pca``, from a reference checkpoint (``torch.load`` -> ``["state"]``; a bare state dict is accepted too).
    Legacy key names are renamed like api.py:170-197 does.  With ``hparams`` (dict, file or log dir) the shapes are
    checked against the model description the checkpoint was trained with."""
    import torch
    ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    state = ckpt["state"] if isinstance(ckpt, dict) and "state" in ckpt else ckpt
    renamed = {}
    for k, v in state.items():
        for old, new in _CKPT_LEGACY_KEYS:
            k = k.replace(old, new)
        renamed[k] = v
    missing = [k for k in _PCA_KEYS.values() if k not in renamed]
    if missing:
        raise KeyError(f"{ckpt_path}: no PCA buffers in the checkpoint (missing {missing[0]}); "
                       "was the model trained with using_pca=True?")
    got = {n: renamed[k].detach().to(torch.float32).cpu().numpy() for n, k in _PCA_KEYS.items()}
    cs, ms, cr, mr = got["compT_scale"], got["means_scale"].reshape(-1), got["compT_rotat"], got["means_rotat"].reshape(-1)
    if cs.ndim != 2 or cr.ndim != 2 or len(ms) != cs.shape[0] or len(mr) != cr.shape[0] or cs.shape[0] != 2 * cr.shape[0]:
        raise ValueError(f"{ckpt_path}: inconsistent PCA buffer shapes {cs.shape} {ms.shape} {cr.shape} {mr.shape}")
    if hparams is not None:
        hp = hparams if isinstance(hparams, dict) else load_hparams(hparams)
        want = pca_dims_from_hparams(hp)
        have = (cs.shape[0], cr.shape[0], cs.shape[1], cr.shape[1])
        if want != have:
            raise ValueError(f"{ckpt_path}: PCA buffers {have} do not match hparams {want}")
    return (np.ascontiguousarray(cs), np.ascontiguousarray(ms), np.ascontiguousarray(cr), np.ascontiguousarray(mr))


# ---------------------------------------------------------------------------------------------------------------
# Foreign-topology inputs of ``set_template_mesh`` (viewer/frame.py:48-96; evaluate.sh --mesh_constraints / --mesh_tricorres)

def read_constraints(path):
    """frame.py:55-58: whitespace-separated vertex indices over any number of lines."""
    with open(path) as fp:
        return np.asarray([int(x) for x in " ".join(l.strip() for l in fp.readlines()).split()], dtype=np.uint32)


def read_tricorres(path, n_faces):
    """frame.py:59-90: first line = number of records, then ``src_tri,dst_tri,<anything>`` per line; the source triangles
    of a target triangle keep file order.  -> dict(corr_count [n_faces], corr_faces [sum(max(count,1))]) exactly as the
    reference builds them (a target triangle without sources gets count 0 and one placeholder entry 0)."""
    by_dst = {}
    with open(path) as fp:
        count = 0
        for i, line in enumerate(fp):
            if i == 0:
                count = int(line.strip())
                continue
            if count == 0:
                break
            src, dst, _ = line.strip().split(",")
            by_dst.setdefault(int(dst), []).append(int(src))
            count -= 1
    corr_count, corr_faces = [], []
    for i in range(n_faces):
        srcs = by_dst.get(i)
        if not srcs:
            corr_count.append(0)
            corr_faces.append(0)
        else:
            corr_count.append(len(srcs))
            corr_faces += srcs
    return dict(corr_count=np.asarray(corr_count, dtype=np.uint32), corr_faces=np.asarray(corr_faces, dtype=np.uint32))


def load_template(template_path, constraints_path=None, corres_path=None):
    """The inputs ``set_template_mesh`` (frame.py:48-96) hands to ``deformation.set_target``:
    -> (verts [n,3] f32, faces [m,3] u32, c_indices or None, corres dict or None)."""
    verts, faces = read_mesh(template_path, dtype=np.float32)
    c = read_constraints(constraints_path) if constraints_path is not None else None
    corres = read_tricorres(corres_path, len(faces)) if corres_path is not None else None
    return verts, faces, c, corres
