#!/usr/bin/env python
"""Diagnostic run on a GPU box: Newton traces (TSL_TRACE=1) of the synthetic sheet, position error against the CPU oracle at
small sizes, per-kernel-class timings.  Usage: python tools/gpu_trace.py N STEPS [--oracle]"""
import os
import sys
import time

os.environ.setdefault("TSL_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from thinshelllab_b200 import _lib
from thinshelllab_b200.synthetic import sheet_scene

N = int(sys.argv[1]); steps = int(sys.argv[2]); use_oracle = "--oracle" in sys.argv
max_newton = int(sys.argv[sys.argv.index("--max-newton") + 1]) if "--max-newton" in sys.argv else 1000
s = sheet_scene(N)
e = s.engine
if "--mode" in sys.argv:
    e.set_option(_lib.OPT_NEWTON_MODE, int(sys.argv[sys.argv.index("--mode") + 1]))
o = None
if use_oracle:
    from oracle import tsl_oracle as orc
    c = s.cloths[0]
    tpos, tfaces, tmass = s._table
    o = orc.OracleScene(c.N, c.M, c.dx, s.dt, tpos, tfaces, tmass, Kb=100.0, k_angle=3.14, k_contact=s.k_contact, eps_contact=s.eps_contact,
                        eps_v=s.eps_v, mu=0.5, max_n_constraints=s.max_n_constraints, grid_n=e.cfg.grid_n)
    o.pos[:] = e.pos.cpu().numpy(); o.prev_pos[:] = o.pos; o.vel[:] = 0
    o.ref_angle[:] = e.cloth_ref_angle[0].cpu().numpy()
for k in range(steps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    st = s.time_step(max_newton=max_newton)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    msg = f"step {k}: {1e3 * dt:.1f} ms newton={st.newton_iters} pcg={st.linear_iters} ls={st.linesearch_evals} nc={st.n_contacts} conv={st.converged} flags={st.flags} " \
          f"ms(contact/asm/solve/ls)={st.ms_contact:.1f}/{st.ms_assembly:.1f}/{st.ms_solve:.1f}/{st.ms_linesearch:.1f}"
    if o is not None:
        it = o.time_step()
        msg += f" | oracle newton={it} nc={o.nc} max|dx|={np.abs(e.pos.cpu().numpy() - o.pos).max():.3e}"
    print(msg, flush=True)
e.contact_detect()
e.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_NEWTON | _lib.ASM_SPD)
if "--solves" in sys.argv:
    F = torch.from_numpy(e.residual()).to(e.device)
    g = torch.Generator(device="cpu").manual_seed(0)
    R = torch.randn(F.shape[0], generator=g, dtype=torch.float64).to(e.device) * (e.frozen == 0)
    for name, rhs in (("F(converged)", F), ("random", R)):
        for tol in (1e-2, 1e-4, 1e-6):
            x, (it, fl, rr) = e.solve(rhs, rel_tol=tol, max_iters=300)
            print(f"  solve {name} |rhs|={float(rhs.norm()):.3e} tol={tol:g}: its={it} flags={fl} rr={rr:.3e}")
sz = e.sizes()
print("sizes", sz)
for name, what in (("pcg_iter", 0), ("spmv", 1), ("energy", 2), ("residual", 3), ("hessian", 4), ("vcycle", 5), ("mg_setup", 6)):
    e.bench_kernel(what, 5)
    print(f"  {name}: {1e3 * e.bench_kernel(what, 50):.1f} us")
for lev in range(e.mg_level(0, values=False)[2]):
    n0, n1, nl, lmax, _ = e.mg_level(lev, values=False)
    print(f"  level {lev}: {n0}x{n1} lmax {lmax:.3f}")
