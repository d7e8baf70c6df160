"""Synthetic sheet scenes of BASELINE.json's configs: an N x N cloth over a frozen box table, Scene_bouncing physics.

Geometry is chosen so that the reference's contact query is meaningful at every size: cloth grid step `dx` (default
2 mm), table triangles no larger than the 3 mm contact grid cell, and the contact grid extent (`grid_n`, the
reference hard-codes 132 cells = +-0.1965 m, geometry.py:8-10) widened to cover the sheet.

`sheet_spec` is pure numpy (host, no GPU): the same description feeds the CUDA scene, the parity tests' oracle and
bench.py's CPU baseline arm."""
import numpy as np

from .meshes import box_body


def sheet_spec(N, dx=0.002, dt=5e-3, seed=0, z0=0.0003, bump=0.25, noise=0.01, k_contact=40000.0, mu=0.5):
    size = N * dx
    table_size = size + 0.02
    tn = int(np.ceil(table_size / 0.003)) + 1
    grid_n = max(132, 2 * int(np.ceil((0.5 * table_size + 0.01) / 0.003)) + 2)
    NV = (N + 1) ** 2
    tdx = table_size / (tn - 1)
    tpos, ttets, tfaces, tmass = box_body(table_size, tn, tn, 2, (-0.5 * table_size, -0.5 * table_size, -tdx))
    # deterministic start: flat sheet + smooth bump + small noise
    rng = np.random.default_rng(seed)
    i, j = np.meshgrid(np.arange(N + 1), np.arange(N + 1), indexing="ij")
    cpos = np.stack([i * dx - 0.5 * size, j * dx - 0.5 * size, np.full(i.shape, z0, np.float64)], -1).reshape(-1, 3)
    cpos[:, 2] += (bump * dx * (1 + np.sin(2 * np.pi * i / 32.0) * np.cos(2 * np.pi * j / 32.0))).reshape(-1)
    cpos += rng.uniform(-noise * dx, noise * dx, (NV, 3))
    return dict(N=N, dx=dx, dt=dt, size=size, table_size=table_size, table_N=(tn, tn, 2), table_offset=(-0.5 * table_size, -0.5 * table_size, -tdx),
                table_pos=tpos, table_faces=tfaces, table_mass=tmass, cloth_pos=cpos, grid_n=grid_n, k_contact=k_contact, mu=mu,
                eps_contact=0.0004, eps_v=0.01, max_n_constraints=NV + 16, n_tris=2 * N * N, n_verts=NV + tpos.shape[0])


def sheet_scene(N, device="cuda:0", pinned_vertices=(), **kw):
    import torch

    from .task_scene.Scene_bouncing import Scene
    sp = sheet_spec(N, **kw)
    size = sp["size"]
    s = Scene(cloth_size=size, cloth_N=N, dt=sp["dt"], table_size=sp["table_size"], table_N=sp["table_N"], table_offset=sp["table_offset"],
              cloth_offset=(-0.5 * size, -0.5 * size, 0.0), reset_offset=(-0.5 * size, -0.5 * size, 0.0),
              k_contact=sp["k_contact"], max_n_constraints=sp["max_n_constraints"], grid_n=sp["grid_n"], pinned_vertices=pinned_vertices, device=device)
    s.mu_cloth_elastic[None] = sp["mu"]
    s.init_all()
    NV = (N + 1) ** 2
    s.engine.pos[:NV] = torch.from_numpy(sp["cloth_pos"]).to(s.engine.device)
    s.engine.prev_pos.copy_(s.engine.pos)
    s.engine.cloth_ref_angle[0].zero_()     # a flat sheet: no pre-creased rows (Scene_bouncing's init_ref_angle_bridge is scene specific)
    s.spec = sp
    return s
