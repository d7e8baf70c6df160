#!/usr/bin/env python
"""Golden fixtures at the BASELINE.json sizes from the tier-2 CPU oracle (oracle/tsl_oracle.py: the C restatement of the reference,
itself pinned to the tier-1 goldens by tests/test_oracle_golden.py).  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_sheet_goldens.py 158        # configs[1] / [2]: T = 5 rollout of the landing + 4 adjoint steps   (~15-40 min here)

Writes tests/golden/sheet<N>_landing.npz: per frame the cloth positions, constraint count and a hash of the sorted constraint index
set; the adjoint sweep's pos_grad[0], grad_kb and |z| per step.  The scenario is bench.py's (thinshelllab_b200.synthetic.LANDING)."""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tsl_oracle as orc  # noqa: E402
from thinshelllab_b200.synthetic import LANDING, sheet_spec  # noqa: E402


def idx_hash(idx):
    a = np.ascontiguousarray(np.asarray(sorted(map(tuple, idx)), np.int32).reshape(-1, 4))
    return hashlib.sha256(a.tobytes()).hexdigest()


def main(N, T=5):
    sp = sheet_spec(N, **LANDING)
    o = orc.OracleScene(N, N, sp["dx"], sp["dt"], sp["table_pos"], sp["table_faces"], sp["table_mass"], Kb=100.0, k_angle=3.14,
                        k_contact=sp["k_contact"], eps_contact=sp["eps_contact"], eps_v=sp["eps_v"], mu=sp["mu"],
                        max_n_constraints=sp["max_n_constraints"], grid_n=sp["grid_n"])
    NVc = o.NVc
    o.pos[:NVc] = sp["cloth_pos"]; o.prev_pos[:] = o.pos
    g = orc.OracleGrad(o, T)
    g.copy_pos(0)
    out = dict(N=N, T=T, dx=sp["dx"], dt=sp["dt"], z0=LANDING["z0"], pos_f0=o.pos[:NVc].copy())
    for f in range(1, T):
        t0 = time.time()
        log = []
        it = o.time_step(log=log)
        g.copy_pos(f)
        out[f"pos_f{f}"] = o.pos[:NVc].copy()
        out[f"vel_f{f}"] = o.vel[:NVc].copy()
        out[f"nc_f{f}"] = o.nc
        out[f"idx_hash_f{f}"] = idx_hash(o.c_idx[:o.nc])
        out[f"newton_f{f}"] = it
        out[f"E_f{f}"] = log[-1][3]
        out[f"delta_f{f}"] = log[-1][1]
        print(f"frame {f}: newton {it} nc {o.nc} delta {log[-1][1]:.2e} E {log[-1][3]:.12e} ({time.time() - t0:.0f} s)", flush=True)
    assert np.abs(o.ref_angle).max() == 0.0          # k_angle = 3.14: no plastic flow on a sheet, buffers stay zero
    g.pos_grad[T - 1, :NVc, 2] = 1.0                  # loss seed of SURVEY.md section 8d config 2
    for j in range(T - 1, 0, -1):
        t0 = time.time()
        g.transfer_grad(j)
        out[f"z_norm_b{j}"] = float(np.linalg.norm(g.z))
        out[f"z_b{j}_sample"] = g.z.reshape(-1, 3)[:NVc:97].copy()
        out[f"nc_b{j}"] = o.nc
        out[f"grad_kb_b{j}"] = g.grad_kb
        print(f"adjoint {j}: |z| {out[f'z_norm_b{j}']:.6e} grad_kb {g.grad_kb:.12e} ({time.time() - t0:.0f} s)", flush=True)
    out["pos_grad0"] = g.pos_grad[0, :NVc].copy()
    out["grad_kb"] = g.grad_kb
    path = os.path.join(ROOT, "tests", "golden", f"sheet{N}_landing.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 158, int(sys.argv[2]) if len(sys.argv) > 2 else 5)
