"""Synthetic sheet scenes of BASELINE.json's configs: an N x N cloth over a frozen box table, Scene_bouncing physics.

Geometry is chosen so that the reference's contact query is meaningful at every size: cloth grid step `dx` (default
2 mm), table triangles no larger than the 3 mm contact grid cell, and the contact grid extent (`grid_n`, the
reference hard-codes 132 cells = +-0.1965 m, geometry.py:8-10) widened to cover the sheet."""
import numpy as np
import torch

from .task_scene.Scene_bouncing import Scene


def sheet_scene(N, dx=0.002, dt=5e-3, seed=0, z0=0.0006, bump=0.25, noise=0.01, k_contact=40000.0, mu=0.5,
                max_n_constraints=None, device="cuda:0"):
    size = N * dx
    table_size = size + 0.02
    tn = int(np.ceil(table_size / 0.003)) + 1
    grid_n = max(132, 2 * int(np.ceil((0.5 * table_size + 0.01) / 0.003)) + 2)
    NV = (N + 1) ** 2
    s = Scene(cloth_size=size, cloth_N=N, dt=dt, table_size=table_size, table_N=(tn, tn, 2),
              table_offset=(-0.5 * table_size, -0.5 * table_size, -table_size / (tn - 1)),
              cloth_offset=(-0.5 * size, -0.5 * size, z0), reset_offset=(-0.5 * size, -0.5 * size, z0),
              k_contact=k_contact, max_n_constraints=max_n_constraints or (NV + 16), grid_n=grid_n, device=device)
    s.mu_cloth_elastic[None] = mu
    s.init_all()
    # deterministic start: flat sheet + smooth bump + small noise
    rng = np.random.default_rng(seed)
    i, j = np.meshgrid(np.arange(N + 1), np.arange(N + 1), indexing="ij")
    pos = s.engine.pos.cpu().numpy()
    pos[:NV, 2] += (bump * dx * (1 + np.sin(2 * np.pi * i / 32.0) * np.cos(2 * np.pi * j / 32.0))).reshape(-1)
    pos[:NV] += rng.uniform(-noise * dx, noise * dx, (NV, 3))
    s.engine.pos.copy_(torch.from_numpy(pos))
    s.engine.prev_pos.copy_(s.engine.pos)
    return s
