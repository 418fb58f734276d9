"""OBJ reader / writer of the reference's drivers, host side, NumPy only.

`load_mesh` follows reference Code/utils.py:476-641 as every caller uses it (`K = 0`, `bGetAdj = False`:
dataClasses.py:30,491-523, computeMetrics.py:41,66) and `write_mesh` Code/utils.py:659-699; same return
tuple, same dtypes, same text.  Pinned by `tests/golden/obj_io.npz`, produced by the reference's own two
functions (oracle/make_golden.py:obj_cases).
"""
import os

import numpy as np

__all__ = ["load_mesh", "write_mesh", "compute_vertex_normals"]


def _unit_rows_twice(a):
    """utils.py:26-35: a / (|a| + 1e-8), applied two times."""
    for _ in range(2):
        a = a * (1 / (np.sqrt((a * a).sum(1))[:, np.newaxis] + 0.00000001))
    return a


def compute_vertex_normals(verts, faces):
    """`computeNormals` (utils.py:44-59), including its indexed `+=`: when a vertex occurs several times in
    one corner column only the last face of that column contributes."""
    tri = verts[faces]
    fn = _unit_rows_twice(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]))
    normals = np.zeros(verts.shape, dtype=np.float32)
    for corner in range(3):
        col = faces[:, corner]
        normals[col] = normals[col] + fn  # buffered: one write per vertex, the last one wins
    return _unit_rows_twice(normals)


def _parse_plain(lines):
    """Fast path for files made of `v x y z` and `f a b c` records only (what `write_mesh` and most scanners
    emit): the numbers of all records go through one C-level parse.  None when anything else is present --
    extra vertex columns, `v/vt/vn` corners, polygons, indented records -- and the general loop takes over."""
    vrec = [l for l in lines if l.startswith("v ")]
    frec = [l for l in lines if l.startswith("f ")]
    rest = [l for l in lines if l and not l.startswith(("v ", "f ", "#"))]
    if not vrec or any(l.split() and l.split()[0] in ("v", "f") for l in rest):
        return None
    ftext = " ".join(l[2:] for l in frec)
    if "/" in ftext:
        return None
    try:
        v = np.fromstring(" ".join(l[2:] for l in vrec), dtype=np.float64, sep=" ")
        c = np.fromstring(ftext, dtype=np.int64, sep=" ") if frec else np.zeros(0, np.int64)
    except (ValueError, DeprecationWarning):
        return None
    if v.size != 3 * len(vrec) or c.size != 3 * len(frec):
        return None
    return v.reshape(-1, 3), (c - 1)


def load_mesh(path, filename, K=0, bGetAdj=False):
    """Returns (vertices float32 [V,3], adj, free_ind, faces uint16|uint32 [F,3] zero-based, vertex normals).
    Polygons are fan-triangulated around their first vertex; `v/vt/vn` corners keep the vertex index; `vn`,
    `vt`, `mtllib`, `usemtl`, comments and blank lines are skipped.  The per-vertex ring adjacency
    (`bGetAdj`) is not built: no driver of the reference asks for it."""
    if bGetAdj:
        raise NotImplementedError("load_mesh(bGetAdj=True): the vertex-ring adjacency is never requested by "
                                  "the reference's drivers; use build_faces_adj for the facet graph")
    with open(os.path.join(path, filename), "r") as f:
        lines = f.read().split("\n")
    parsed = _parse_plain(lines)
    if parsed is not None:
        verts, corners = parsed
    else:
        verts = []
        corners = []
        for line in lines:
            tok = line.split()
            if not tok or line.startswith("#"):
                continue
            if tok[0] == "v":
                verts.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif tok[0] == "f":
                ids = [int(t.split("/")[0]) - 1 for t in tok[1:]]
                for t in range(len(ids) - 2):
                    corners += (ids[0], ids[t + 1], ids[t + 2])
    vertices = np.array(verts).astype(np.float32)
    itype = np.uint16 if vertices.shape[0] < 65536 else np.uint32
    corners = np.asarray(corners)
    faces = corners.reshape(corners.size // 3, 3).astype(itype)
    return vertices, [], [], faces, compute_vertex_normals(vertices, faces)


def write_mesh(vl, fl, strFileName):
    """One `v` line per vertex with `%.6f` coordinates (every column of `vl`), then one `f` line per face,
    one-based; every token is followed by a blank.  Face rows (-1, -1, .) are skipped and the first row
    (0, 0, .) ends the list (the padding conventions of the patch pipeline)."""
    vl = np.asarray(vl)
    fl = np.asarray(fl)
    text = ""
    if vl.size:
        row = "v " + "%.6f " * vl.shape[1] + "\n"
        text = (row * vl.shape[0]) % tuple(vl.ravel().tolist())
    if fl.size:
        one = fl.astype(np.int64) + 1
        end = np.flatnonzero((one[:, 0] == 1) & (one[:, 1] == 1))
        if end.size:
            one = one[: end[0]]
        one = one[~((one[:, 0] == 0) & (one[:, 1] == 0))]
        row = "f " + "%d " * one.shape[1] + "\n"
        text += (row * one.shape[0]) % tuple(one.ravel().tolist())
    with open(strFileName, "w") as f:
        f.write(text)
