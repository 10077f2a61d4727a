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
