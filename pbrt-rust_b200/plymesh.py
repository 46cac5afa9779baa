"""PLY triangle meshes for `Shape "plymesh"` (src/shapes/plymesh.rs:17-167, which reads through the `ply-rs`
crate).  Host side only: the result is the argument list of `create_trianglemesh` (P, indices, N, uv).

Handles `ascii`, `binary_little_endian` and `binary_big_endian` files, any scalar property types, and the
list property `vertex_indices` with 3 or 4 entries per face (quads split 0-1-2 / 3-0-2 as plymesh.rs:98-114;
other face sizes are skipped with a warning).  Elements are read in file order.
"""
from __future__ import annotations

import warnings

import numpy as np

from .host import B200Error

f32 = np.float32
_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2", "int": "i4",
          "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}
_UV_NAMES = (("u", "v"), ("s", "t"), ("texture_u", "texture_v"), ("texture_s", "texture_t"))  # plymesh.rs:51-56


def _read_header(f, filename):
    if f.readline().strip() != b"ply":
        raise B200Error(f'PLY file "{filename}": missing "ply" magic')
    fmt, elements = None, []
    while True:
        line = f.readline()
        if not line:
            raise B200Error(f'PLY file "{filename}": unterminated header')
        tok = line.decode("ascii", "replace").split()
        if not tok or tok[0] in ("comment", "obj_info"):
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append({"name": tok[1], "count": int(tok[2]), "props": []})
        elif tok[0] == "property":
            if not elements:
                raise B200Error(f'PLY file "{filename}": property before any element')
            if tok[1] == "list":
                elements[-1]["props"].append((tok[4], "list", _TYPES[tok[2]], _TYPES[tok[3]]))
            else:
                elements[-1]["props"].append((tok[2], "scalar", _TYPES[tok[1]], None))
        elif tok[0] == "end_header":
            break
    if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
        raise B200Error(f'PLY file "{filename}": unknown format {fmt}')
    return fmt, elements


def _read_binary_element(f, el, order):
    props, n = el["props"], el["count"]
    if all(p[1] == "scalar" for p in props):
        dt = np.dtype([(p[0], order + p[2]) for p in props])
        raw = f.read(dt.itemsize * n)
        if len(raw) != dt.itemsize * n:
            raise B200Error("PLY payload truncated")
        return np.frombuffer(raw, dt), None
    if len(props) == 1:  # the usual face element: one list property
        name, _, ct, it = props[0]
        cdt, idt = np.dtype(order + ct), np.dtype(order + it)
        pos = f.tell()
        first = np.frombuffer(f.read(cdt.itemsize), cdt)
        if len(first):
            k = int(first[0])
            rec = np.dtype([("n", cdt), ("v", idt, (k,))])
            f.seek(pos)
            raw = f.read(rec.itemsize * n)
            if len(raw) == rec.itemsize * n:
                a = np.frombuffer(raw, rec)
                if np.all(a["n"] == k):  # fast path: every face has the same vertex count
                    return None, {name: (np.full(n, k, np.int64), a["v"].astype(np.int64).reshape(-1))}
            f.seek(pos)
    # general path: row by row
    scal = {p[0]: np.zeros(n, order + p[2]) for p in props if p[1] == "scalar"}
    lists = {p[0]: ([], []) for p in props if p[1] == "list"}
    for r in range(n):
        for name, kind, t, it in props:
            if kind == "scalar":
                d = np.dtype(order + t)
                scal[name][r] = np.frombuffer(f.read(d.itemsize), d)[0]
            else:
                cd, idd = np.dtype(order + t), np.dtype(order + it)
                k = int(np.frombuffer(f.read(cd.itemsize), cd)[0])
                lists[name][0].append(k)
                lists[name][1].append(np.frombuffer(f.read(idd.itemsize * k), idd).astype(np.int64))
    lists = {k: (np.array(c, np.int64), np.concatenate(v) if v else np.zeros(0, np.int64)) for k, (c, v) in lists.items()}
    rec = None
    if scal:
        rec = np.zeros(n, np.dtype([(k, v.dtype) for k, v in scal.items()]))
        for k, v in scal.items():
            rec[k] = v
    return rec, lists


def _read_ascii_element(f, el):
    props, n = el["props"], el["count"]
    scal = {p[0]: np.zeros(n, np.float64) for p in props if p[1] == "scalar"}
    lists = {p[0]: ([], []) for p in props if p[1] == "list"}
    for r in range(n):
        tok = f.readline().split()
        c = 0
        for name, kind, t, it in props:
            if kind == "scalar":
                scal[name][r] = float(tok[c])
                c += 1
            else:
                k = int(tok[c])
                lists[name][0].append(k)
                lists[name][1].append(np.array([int(float(v)) for v in tok[c + 1:c + 1 + k]], np.int64))
                c += 1 + k
    lists = {k: (np.array(c, np.int64), np.concatenate(v) if v else np.zeros(0, np.int64)) for k, (c, v) in lists.items()}
    rec = None
    if scal:
        rec = np.zeros(n, np.dtype([(k, "f8") for k in scal]))
        for k, v in scal.items():
            rec[k] = v
    return rec, lists


def read_ply(filename):
    """-> dict(P, indices, N, S, uv) for SceneBuilder.shape("trianglemesh", ...), or None when the reference
    would log an error and create no shape."""
    try:
        f = open(filename, "rb")
    except OSError as e:
        raise B200Error(f'plymesh: cannot open "{filename}": {e}')  # File::open(..).unwrap() panics
    with f:
        fmt, elements = _read_header(f, filename)
        names = [e["name"] for e in elements]
        vcount = next((e["count"] for e in elements if e["name"] == "vertex"), 0)
        fcount = next((e["count"] for e in elements if e["name"] == "face"), 0)
        if "vertex" in names:
            vp = [p[0] for p in elements[names.index("vertex")]["props"]]
            if not all(k in vp for k in "xyz"):
                warnings.warn(f'PLY file "{filename}": Vertex coordinate property not found')
                return None
        if vcount == 0 or fcount == 0:
            warnings.warn(f'PLY file "{filename}" is invalid! No face/vertex elements found')
            return None
        order = {"ascii": "", "binary_little_endian": "<", "binary_big_endian": ">"}[fmt]
        vert = faces = None
        for el in elements:
            rec, lists = _read_ascii_element(f, el) if fmt == "ascii" else _read_binary_element(f, el, order)
            if el["name"] == "vertex":
                vert = rec
            elif el["name"] == "face":
                if lists is None or "vertex_indices" not in lists:
                    raise B200Error(f'PLY file "{filename}": face element has no "vertex_indices" list')
                faces = lists["vertex_indices"]
    vp = vert.dtype.names
    P = np.stack([vert["x"], vert["y"], vert["z"]], axis=1).astype(f32)
    N = uv = None
    if "nx" in vp:  # plymesh.rs:45-49 tests "nx" three times; ny / nz default to 0 when absent
        N = np.stack([vert[k] if k in vp else np.zeros(len(vert)) for k in ("nx", "ny", "nz")], axis=1).astype(f32)
    if any(a in vp and b in vp for a, b in _UV_NAMES):
        u = next((vert[k] for k in ("u", "texture_u", "s", "texture_s") if k in vp), np.zeros(len(vert)))  # set_property, plymesh.rs:190-199
        v = next((vert[k] for k in ("v", "texture_v", "t", "texture_t") if k in vp), np.zeros(len(vert)))
        uv = np.stack([u, v], axis=1).astype(f32)
    counts, flat = faces
    if np.all(counts == 3):
        idx = flat.reshape(-1, 3)
    else:
        off = np.concatenate([[0], np.cumsum(counts)[:-1]])
        bad = (counts != 3) & (counts != 4)
        if bad.any():
            warnings.warn(f"plymesh: Ignoring {int(bad.sum())} faces that are neither triangles nor quads")
        rows = []
        for o, c in zip(off[~bad], counts[~bad]):  # file order is kept: a quad becomes 0-1-2, 3-0-2
            rows.append(flat[o:o + 3])
            if c == 4:
                rows.append(np.array([flat[o + 3], flat[o], flat[o + 2]], np.int64))
        idx = np.array(rows, np.int64).reshape(-1, 3)
    if len(idx) == 0:
        return None
    if idx.min() < 0 or idx.max() >= len(P):
        raise B200Error(f'PLY file "{filename}": vertex index out of range')
    return {"P": P, "indices": idx.astype(np.uint32), "N": N, "S": None, "uv": uv}


def write_ply(filename, P, indices, N=None, uv=None, fmt="binary_little_endian"):
    """Small PLY writer (tests and scene generators): float x y z [nx ny nz] [u v], uint8-counted int faces."""
    P, idx = np.asarray(P, f32).reshape(-1, 3), np.asarray(indices, np.int32).reshape(-1, 3)
    cols, names = [P], ["x", "y", "z"]
    if N is not None:
        cols.append(np.asarray(N, f32).reshape(-1, 3)); names += ["nx", "ny", "nz"]
    if uv is not None:
        cols.append(np.asarray(uv, f32).reshape(-1, 2)); names += ["u", "v"]
    V = np.concatenate(cols, axis=1)
    hdr = ["ply", f"format {fmt} 1.0", f"element vertex {len(P)}"] + [f"property float {n}" for n in names]
    hdr += [f"element face {len(idx)}", "property list uint8 int vertex_indices", "end_header"]
    with open(filename, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        if fmt == "ascii":
            for row in V:
                f.write((" ".join(repr(float(x)) for x in row) + "\n").encode("ascii"))
            for t in idx:
                f.write(("3 %d %d %d\n" % tuple(t)).encode("ascii"))
        else:
            o = "<" if fmt == "binary_little_endian" else ">"
            f.write(V.astype(o + "f4").tobytes())
            rec = np.zeros(len(idx), np.dtype([("n", "u1"), ("v", o + "i4", (3,))]))
            rec["n"], rec["v"] = 3, idx
            f.write(rec.tobytes())
