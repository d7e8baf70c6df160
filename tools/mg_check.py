#!/usr/bin/env python
"""Diagnostic (GPU box): the library's V-cycle against a numpy/scipy restatement of the same cycle on the same matrices."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, scipy.sparse as sp, torch
from thinshelllab_b200 import _lib
from thinshelllab_b200.synthetic import sheet_scene
from test_gpu_multigrid import _prolongation, _stencil_to_csr

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
s = sheet_scene(N); e = s.engine
for _ in range(steps): s.time_step()
e.contact_detect()
e.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_NEWTON | _lib.ASM_SPD)
NVc = s.cloths[0].NV; nc3 = 3 * NVc
A0 = e.matrix()[:nc3, :nc3].tocsr()
n0, n1, nlev, _, _ = e.mg_level(0, values=False)
levels = []
for l in range(nlev):
    a, b, _, lmax, val = e.mg_level(l)
    A = A0 if l == 0 else _stencil_to_csr(a, b, val)
    Ab = A.tobsr((3, 3)); nv = a * b
    D = np.zeros((nv, 3, 3)); r = np.repeat(np.arange(nv), np.diff(Ab.indptr)); D[r[r == Ab.indices]] = Ab.data[r == Ab.indices]
    levels.append(dict(A=A, Dinv=np.linalg.inv(D), n=(a, b), lmax=float(lmax)))
    if l + 1 < nlev: levels[-1]["P"] = _prolongation(a, b)[0]
print("levels", [(L["n"], round(L["lmax"], 3)) for L in levels])
def apply_D(Dinv, r): return np.einsum("nij,nj->ni", Dinv, r.reshape(-1, 3)).reshape(-1)
def cheb(L, x, b, deg, ratio, zero):
    lmax = L["lmax"]; lmin = lmax / ratio
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin); sigma = theta / delta; rho = 1 / sigma
    r = b if zero else b - L["A"] @ x
    d = apply_D(L["Dinv"], r) / theta
    x = d.copy() if zero else x + d
    for k in range(1, deg):
        rn = 1 / (2 * sigma - rho); r = b - L["A"] @ x
        d = rn * rho * d + (2 * rn / delta) * apply_D(L["Dinv"], r); x = x + d; rho = rn
    return x
def vcycle(k, b, trace=None):
    L = levels[k]
    if k == nlev - 1: return cheb(L, None, b, 8, 200.0, True)
    x = cheb(L, None, b, 2, 8.0, True)
    r = b - L["A"] @ x
    x = x + L["P"] @ vcycle(k + 1, L["P"].T @ r)
    return cheb(L, x, b, 2, 8.0, False)
rng = np.random.default_rng(0)
b = rng.standard_normal(nc3)
bf = np.zeros(3 * e.n_verts); bf[:nc3] = b
z_gpu = e.precond_apply(torch.from_numpy(bf).to(e.device)).cpu().numpy()[:nc3]
z_py = vcycle(0, b)
print("V-cycle: |z_gpu - z_py| / |z_py| =", np.linalg.norm(z_gpu - z_py) / np.linalg.norm(z_py), " b.z gpu/py", b @ z_gpu, b @ z_py)
# smoother only (no coarse correction) and one-level checks
x2 = cheb(levels[0], None, b, 2, 8.0, True)
print("   pre-smoothed fine iterate: rel diff of z_gpu vs smoother-only", np.linalg.norm(z_gpu - cheb(levels[0], x2, b, 2, 8.0, False)) / np.linalg.norm(z_py))
def pcg(A, b, M, tol, maxit):
    x = np.zeros_like(b); r = b.copy(); z = M(r); p = z.copy(); rz = r @ z; rr0 = r @ r
    for it in range(1, maxit + 1):
        q = A @ p; a = rz / (p @ q); x += a * p; r -= a * q
        if r @ r <= tol * tol * rr0: return it
        z = M(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return maxit
def M_gpu(r):
    rf = np.zeros(3 * e.n_verts); rf[:nc3] = r
    return e.precond_apply(torch.from_numpy(rf).to(e.device)).cpu().numpy()[:nc3]
for tol in (1e-2, 1e-4):
    print(f"PCG tol {tol:g}: python V-cycle {pcg(A0, b, lambda r: vcycle(0, r), tol, 400)} its, GPU V-cycle {pcg(A0, b, M_gpu, tol, 400)} its")
F = torch.from_numpy(bf).to(e.device)
for tol in (1e-2, 1e-4):
    x, (it, fl, rr) = e.solve(F, rel_tol=tol, max_iters=400)
    print(f"library PCG tol {tol:g}: {it} its flags {fl}")
