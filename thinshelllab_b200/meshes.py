"""Host-side mesh builders (cold path, numpy): the structured box body ("table") of the reference scenes.

Follows Elastic.get_vertices / init_pos / get_surface_indices of the reference
(code/engine/model_elastic_offset.py:240-245, 279-312, 346-376): 5 tets per cube with alternating parity,
lumped vertex mass = sum W/4 * density, surface triangles in the serial loop order of the reference (the order
matters: the contact query breaks distance ties by candidate order)."""
import numpy as np


def box_body(Len, Nx, Ny, Nz, offset, density=2000.0, arch=0.0):
    """returns pos [nv,3] f64, tets [nc,4] i32, faces [nf,3] i32 (body-local vertex ids), mass [nv] f64.
    arch != 0: Elastic.init_pos_arch (model_elastic_offset.py:253-270) -- the rest shape is bent upwards by arch * sin(pi x / Lx)
    (3.1415926 as the reference has it) before the rest matrices, volumes and masses are taken"""
    n = np.array([Nx, Ny, Nz], np.int64)
    dx = Len / (n.max() - 1)
    gx, gy, gz = np.meshgrid(np.arange(Nx), np.arange(Ny), np.arange(Nz), indexing="ij")
    rest = np.stack([gx, gy, gz], -1).reshape(-1, 3).astype(np.float64) * dx      # vertex id = (x*Ny + y)*Nz + z
    if arch != 0.0:
        rest[:, 2] += arch * np.sin(gx.reshape(-1).astype(np.float64) / float(Nx - 1) * 3.1415926)
    cx, cy, cz = np.meshgrid(np.arange(Nx - 1), np.arange(Ny - 1), np.arange(Nz - 1), indexing="ij")
    cube = np.stack([cx, cy, cz], -1).reshape(-1, 3)                               # cube id = (x*(Ny-1) + y)*(Nz-1) + z
    codes = np.array([[j, j ^ 1, j ^ 2, j ^ 4] for j in (0, 3, 5, 6)] + [[1, 2, 4, 7]])   # [5,4] corner codes
    bits = (codes[..., None] >> np.arange(3)) & 1                                  # [5,4,3]
    corner = cube[:, None, None, :] + ((bits[None] ^ cube[:, None, None, :]) & 1)  # parity flip per cube
    tets = ((corner[..., 0] * Ny + corner[..., 1]) * Nz + corner[..., 2]).reshape(-1, 4).astype(np.int32)
    Ds = rest[tets[:, :3]] - rest[tets[:, 3:4]]
    W = np.abs(np.linalg.det(Ds)) / 6.0
    mass = np.zeros(rest.shape[0])
    np.add.at(mass, tets.reshape(-1), np.repeat(W / 4.0 * density, 4))
    pos = rest + np.asarray(offset, np.float64)
    # surface: faces (i, i+1, i+2) of the four corner tets of each cube for i in (0, 2, 3), kept when all three
    # vertices share a boundary plane; orientation flipped so the normal points away from the 4th vertex
    t4 = tets.reshape(-1, 5, 4)[:, :4, :]
    order = np.array([[(i + j) % 4 for j in range(4)] for i in (0, 2, 3)])          # [3,4]: v0,v1,v2,opposite
    cand = t4[:, :, order].reshape(-1, 4)                                           # (cube, tet, i) order
    vx = cand // (Ny * Nz); vy = (cand // Nz) % Ny; vz = cand % Nz
    code = ((vz == 0) * 1 + (vz == Nz - 1) * 2 + (vy == 0) * 4 + (vy == Ny - 1) * 8 + (vx == 0) * 16 + (vx == Nx - 1) * 32)
    keep = (code[:, 0] & code[:, 1] & code[:, 2]) != 0
    cand = cand[keep]
    nrm = np.cross(pos[cand[:, 1]] - pos[cand[:, 0]], pos[cand[:, 2]] - pos[cand[:, 0]])
    flip = np.einsum("ij,ij->i", nrm, pos[cand[:, 3]] - pos[cand[:, 0]]) > 0
    faces = cand[:, :3].copy()
    faces[flip, 1], faces[flip, 2] = cand[flip, 2], cand[flip, 1]
    return pos, tets, faces.astype(np.int32), mass
