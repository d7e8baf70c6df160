"""engine/readfile.py of the reference, the one function its trajectory drivers call on the hot loop's results:
save_cloth_mesh (code/engine/readfile.py:117-128) -- here a plain ASCII PLY writer (positions + triangles; the reference goes through
open3d and also stores vertex normals).  Mesh loading (read_node / read_ele / read_smesh) is not part of this build."""
import os

import numpy as np


def save_cloth_mesh(cloth, path):
    v = np.asarray(cloth.pos.to_numpy(), np.float64)
    f = np.asarray(cloth.f2v.to_numpy(), np.int64)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as fh:
        fh.write("ply\nformat ascii 1.0\nelement vertex %d\nproperty double x\nproperty double y\nproperty double z\n" % len(v))
        fh.write("element face %d\nproperty list uchar int vertex_indices\nend_header\n" % len(f))
        np.savetxt(fh, v, fmt="%.17g")
        np.savetxt(fh, np.concatenate([np.full((len(f), 1), 3), f], 1), fmt="%d")
