#!/usr/bin/env python
"""Diagnostic: the adjoint linear solve of a golden Scene_folding / Scene_forming state with the library's BiCGStab under different
preconditioners, against the reference's own solution.  python tools/adjoint_probe.py forming|folding"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thinshelllab_b200 import _lib  # noqa: E402
from thinshelllab_b200.task_scene.Scene_folding import Scene  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "forming"
g = np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
for precond, safety in ((1, 1.2), (0, 1.2)):
    s = Scene(g)
    e = s.engine
    e.set_option(_lib.OPT_PRECOND, precond)
    e.set_option(_lib.OPT_MG_SAFETY, safety)
    for j in (2, 1):
        x_t, x_p = (torch.from_numpy(g["pos_buffer"][k]).to(e.device) for k in (j, j - 1))
        e.pos.copy_(x_p); e.prev_pos.copy_(x_p)
        nc = e.contact_detect()
        e.pos.copy_(x_t)
        e.cloth_ref_angle[0].copy_(torch.from_numpy(g["ref_angle_buffer"][j - 1, 0]).to(e.device))
        e.assemble(_lib.ASM_HESSIAN | _lib.ASM_F64)
        rhs = torch.from_numpy(g[f"b{j}_rhs"]).to(e.device)
        x, (it, flags, rr) = e.solve(rhs, rel_tol=1e-10, max_iters=20000)
        ref = g[f"b{j}_z"]
        H = e.matrix()
        print(f"{name} precond={precond} safety={safety} step {j}: nc {nc} (ref {int(g[f'b{j}_nc'])}) iters {it} flags {flags} rel_res {rr:.2e} "
              f"|x-ref|/|ref| {np.abs(x.cpu().numpy() - ref).max() / np.abs(ref).max():.2e} true res {np.abs(H @ x.cpu().numpy() - g[f'b{j}_rhs']).max():.2e}", flush=True)
