"""engine/readfile.py of the reference on the B200 build: TetGen readers for the pad / ball meshes (read_node / read_ele / read_smesh,
code/engine/readfile.py:1-51) and save_cloth_mesh (:117-128; a plain ASCII PLY writer -- the reference goes through open3d).

File lookup: the reference opens "../data/<name>" relative to the current directory (its scripts run from code/).  Here a path is
tried as given, then under $TSL_DATA_DIR, then under the assets shipped with the package (thinshelllab_b200/data/assets: copies of the
reference's data/tactile.* and data/ball.* mesh files -- assets, not code)."""
import os

import numpy as np

_ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "assets")


def resolve(filename):
    cands = [filename]
    if os.environ.get("TSL_DATA_DIR"):
        cands.append(os.path.join(os.environ["TSL_DATA_DIR"], os.path.basename(filename)))
    cands.append(os.path.join(_ASSETS, os.path.basename(filename)))
    for c in cands:
        if os.path.exists(c):
            return c
    raise FileNotFoundError(f"{filename}: not found (also tried $TSL_DATA_DIR and {_ASSETS})")


def _table(filename, dtype):
    """TetGen text table: a header line whose first number is the row count, then `index v1 v2 ...` rows"""
    with open(resolve(filename), encoding="utf-8") as fh:
        n = int(fh.readline().split()[0])
        rows = [fh.readline().split() for _ in range(n)]
    return n, [np.array(r[1:], dtype=dtype) for r in rows]


def read_node(filename="../data/tactile.node"):
    """-> (count, positions [count][3])"""
    n, rows = _table(filename, np.float64)
    return n, [r[:3].tolist() for r in rows]


def read_ele(filename="../data/tactile.ele"):
    """-> (count, cells [count][4])"""
    n, rows = _table(filename, np.int64)
    return n, [r[:4].astype(int).tolist() for r in rows]


def read_smesh(filename="../data/tactile.face"):
    """-> (count, surface triangles [count][3]); the trailing boundary marker of a .face row is dropped"""
    n, rows = _table(filename, np.int64)
    return n, [r[:3].astype(int).tolist() for r in rows]


def save_cloth_mesh(cloth, path):
    v = np.asarray(cloth.pos.to_numpy(), np.float64)
    f = np.asarray(cloth.f2v.to_numpy(), np.int64)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as fh:
        fh.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty double x\nproperty double y\nproperty double z\n" % len(v))
        fh.write("element face %d\nproperty list uchar int vertex_indices\nend_header\n" % len(f))
        np.savetxt(fh, v, fmt="%.17g")
        np.savetxt(fh, np.concatenate([np.full((len(f), 1), 3), f], 1), fmt="%d")
