"""GPU tests of the multigrid preconditioner (tsl_mg.cu) through the C ABI: the Galerkin hierarchy against P^T A P formed
with scipy from the exported fine matrix, symmetry / definiteness of the V-cycle, and its effect on PCG iteration counts.
Tolerances: fp32 storage -> 2e-5 of the largest entry per level; V-cycle symmetry 1e-4 relative."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg  # noqa: F401
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from thinshelllab_b200 import _lib
    from thinshelllab_b200.synthetic import sheet_scene


def _interp1d(nf):
    nc = (nf - 1) // 2 + 1
    rows, cols, vals = [], [], []
    for i in range(nf):
        if i % 2 == 0:
            rows.append(i); cols.append(i // 2); vals.append(1.0)
        else:
            a, b = (i - 1) // 2, (i + 1) // 2
            if b < nc:
                rows += [i, i]; cols += [a, b]; vals += [0.5, 0.5]
            else:
                rows.append(i); cols.append(a); vals.append(1.0)
    return sp.csr_matrix((vals, (rows, cols)), shape=(nf, nc)), nc


def _prolongation(n0, n1):
    P0, c0 = _interp1d(n0)
    P1, c1 = _interp1d(n1)
    return sp.kron(sp.kron(P0, P1), sp.identity(3)).tocsr(), c0, c1


def _stencil_to_csr(n0, n1, val):
    """val [25, 3, 3, n0*n1] -> scipy CSR (3 n0 n1)^2"""
    rows, cols, data = [], [], []
    I, J = np.meshgrid(np.arange(n0), np.arange(n1), indexing="ij")
    I, J = I.reshape(-1), J.reshape(-1)
    v = I * n1 + J
    for slot in range(25):
        dI, dJ = slot // 5 - 2, slot % 5 - 2
        ok = (I + dI >= 0) & (I + dI < n0) & (J + dJ >= 0) & (J + dJ < n1)
        u = (I + dI) * n1 + (J + dJ)
        for a in range(3):
            for b in range(3):
                rows.append(3 * v[ok] + a); cols.append(3 * u[ok] + b); data.append(val[slot, a, b, ok])
        # entries that point outside the grid must be zero
        assert np.abs(val[slot][:, :, ~ok]).max(initial=0.0) == 0.0
    A = sp.csr_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))), shape=(3 * n0 * n1, 3 * n0 * n1))
    A.eliminate_zeros()
    return A


@pytest.mark.parametrize("N,pinned", [(40, False), (37, False), (24, True), (200, False)])   # 200: element-major (large-level) layout
def test_galerkin_hierarchy_matches_scipy(N, pinned):
    kw = {}
    s = sheet_scene(N)
    e = s.engine
    if pinned:                                  # frozen cloth DOFs are left out of the coarse spaces
        e.frozen[0:3 * (N + 1)] = 1            # the whole first grid row
    e.contact_detect()
    e.assemble(_lib.ASM_HESSIAN | _lib.ASM_NEWTON | _lib.ASM_SPD)
    NVc = s.cloths[0].NV
    A = e.matrix()[:3 * NVc, :3 * NVc].tocsr()
    n0, n1, nlev, lmax, val = e.mg_level(0)
    assert (n0, n1) == (N + 1, N + 1) and nlev >= 3
    A0 = _stencil_to_csr(n0, n1, val)
    assert abs(A0 - A).max() <= 1e-6 * abs(A).max()          # level 0: the stencil copy is the fine cloth block
    free = 1.0 - e.frozen.cpu().numpy()[:3 * NVc].astype(np.float64)
    Al = A
    for lev in range(1, nlev):
        P, c0, c1 = _prolongation(n0, n1)
        if lev == 1:
            P = sp.diags(free) @ P
        Al = (P.T @ Al @ P).tocsr()
        n0, n1, _, lmax, val = e.mg_level(lev)
        assert (n0, n1) == (c0, c1)
        G = _stencil_to_csr(n0, n1, val)
        assert abs(G - Al).max() <= 2e-5 * abs(Al).max(), lev
        assert abs(G - G.T).max() <= 2e-5 * abs(Al).max()
        if n0 * n1 > 3000:
            continue
        if lev == nlev - 1 and 3 * n0 * n1 <= 120:
            continue            # the coarsest level is solved exactly (dense inverse, tsl_mg.cu k_coarse_inverse): no smoother, no estimate
        # power-iteration estimate (times the safety factor) must not be below the true lambda_max(D^-1 A)
        Gb = G.tobsr((3, 3))
        D = np.zeros((n0 * n1, 3, 3))
        r = np.repeat(np.arange(n0 * n1), np.diff(Gb.indptr))
        D[r[r == Gb.indices]] = Gb.data[r == Gb.indices]
        ok = np.abs(np.linalg.det(D)) > 0
        Dinv = np.zeros_like(D); Dinv[ok] = np.linalg.inv(D[ok])
        DA = sp.block_diag([sp.csr_matrix(b) for b in Dinv]).tocsr() @ G
        true = np.abs(np.linalg.eigvals(DA.toarray())).max() if DA.shape[0] <= 1500 else abs(sp.linalg.eigs(DA, k=1, which="LM", return_eigenvectors=False, tol=1e-4)[0])
        assert lmax >= 0.97 * true, (lev, lmax, true)
        assert lmax <= 1.6 * true, (lev, lmax, true)


def test_vcycle_is_symmetric_positive_and_cuts_pcg_iterations():
    s = sheet_scene(64)
    e = s.engine
    for _ in range(2):
        s.time_step()                         # a state with contacts and in-plane stress
    e.contact_detect()
    e.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_NEWTON | _lib.ASM_SPD)
    n = 3 * e.n_verts
    g = torch.Generator(device="cpu").manual_seed(0)
    u = torch.randn(n, generator=g, dtype=torch.float64).to(e.device)
    v = torch.randn(n, generator=g, dtype=torch.float64).to(e.device)
    Mu, Mv = e.precond_apply(u), e.precond_apply(v)
    a, b = float(v @ Mu), float(u @ Mv)
    assert abs(a - b) <= 1e-4 * max(abs(a), abs(b))
    assert float(u @ Mu) > 0 and float(v @ Mv) > 0
    F = torch.from_numpy(e.residual()).to(e.device)
    x_mg, (it_mg, fl_mg, rr_mg) = e.solve(F, rel_tol=1e-4, max_iters=500)
    e.set_option(_lib.OPT_PRECOND, 0)
    x_bj, (it_bj, fl_bj, rr_bj) = e.solve(F, rel_tol=1e-4, max_iters=20000)
    e.set_option(_lib.OPT_PRECOND, 1)
    assert fl_mg == 0, (it_mg, fl_mg, rr_mg)
    assert fl_bj == 0, (it_bj, fl_bj, rr_bj)
    assert it_mg <= 40 and it_bj >= 4 * it_mg, (it_mg, it_bj)
    H = e.matrix()
    Fh = F.cpu().numpy()
    for x in (x_mg, x_bj):
        assert np.linalg.norm(H @ x.cpu().numpy() - Fh) <= 2e-4 * np.linalg.norm(Fh)
