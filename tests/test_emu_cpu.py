"""CPU runs of libtsl kernels through the CUDA-on-CPU shim tests/csrc/cuda_emu.h (one std::thread per CUDA thread): the kernels with
real thread cooperation -- pivoted LU panels, shared-memory tiles, warp shuffles -- are exercised here, where there is no GPU, with
the same launch sequences the library uses.  The GPU tests repeat the comparisons on the real device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "csrc", "_build")


def _build(name):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "csrc", name + ".cpp")
    out = os.path.join(BUILD, "lib" + name + ".so")
    deps = [src, os.path.join(HERE, "csrc", "cuda_emu.h")]
    csrc = os.path.join(os.path.dirname(HERE), "thinshelllab_b200", "csrc")
    deps += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", src, "-o", out])
    return C.CDLL(out)


@pytest.mark.parametrize("n", [1, 7, 32, 33, 70, 129])
def test_dense_lu_kernels_match_numpy(n):
    """blocked right-looking LU with partial pivoting (tsl_dense_kernels.cuh) + triangular solves against numpy.linalg.solve"""
    L = _build("emu_dense")
    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    if n > 1:
        A[0, 0] = 0.0                                 # forces a row exchange in the first panel
    if n > 40:
        A[35, :36] = 0.0; A[35, 35] = 1e-9            # a tiny pivot candidate inside the second panel
    b = rng.standard_normal(n)
    x = np.zeros(n)
    Af = np.asfortranarray(A)
    info = L.emu_lu_solve(n, Af.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), 128)
    assert info == 0
    ref = np.linalg.solve(A, b)
    assert np.abs(x - ref).max() <= 1e-9 * np.abs(ref).max() * max(1.0, np.linalg.cond(A) * 1e-6)


def test_dense_lu_flags_singular_matrix():
    L = _build("emu_dense")
    n = 40
    A = np.random.default_rng(0).standard_normal((n, n))
    A[:, 5] = 0.0
    b = np.ones(n); x = np.zeros(n)
    Af = np.asfortranarray(A)
    assert L.emu_lu_solve(n, Af.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), 64) == 1


@pytest.mark.parametrize("N,M,offset,seed", [(5, 4, 0, 0), (9, 37, 0, 1), (6, 66, 40, 2), (15, 3, 0, 3), (7, 2, 0, 4), (4, 1, 32, 5)])
def test_owner_computes_hessian_rows_match_the_oracle_twin(N, M, offset, seed):
    """k_hessian_rows (tsl_assembly_kernels.cuh: one thread per matrix block, tiles of 4 x 32 grid vertices, hinge gradients staged in
    shared memory, contribution lists from the tables read off the reference's mesher) against the oracle's CPU twin of the forward
    Newton matrices (exact and clamped), on grids that span several tiles, with a frozen DOF and a non-zero row offset"""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import tsl_oracle as orc
    L = _build("emu_assembly")
    rng = np.random.default_rng(seed)
    dx, dt = 0.002, 5e-3
    NV = (N + 1) * (M + 1)
    tpos = np.array([[1., 1., -1.], [1.1, 1., -1.], [1., 1.1, -1.]])
    o = orc.OracleScene(N, M, dx, dt, tpos, np.array([[0, 1, 2]], np.int32), np.ones(3))
    i, j = np.meshgrid(np.arange(N + 1), np.arange(M + 1), indexing="ij")
    p = np.stack([i * dx, j * dx, 0.0 * i], -1).reshape(-1, 3).astype(float)
    p += rng.uniform(-0.3 * dx, 0.3 * dx, p.shape)                 # strong in-plane noise: compressed and stretched elements
    p[:, 2] += 0.5 * dx * np.sin(i.ravel() * 0.9) * np.cos(j.ravel() * 0.7)
    o.pos[:NV] = p; o.prev_pos[:] = o.pos
    o.frozen[3 * 5 + 1] = 1
    o.nc = 0
    o._build_pattern()
    ref = {}
    for name, pf in (("e", 0), ("c", 3)):
        o.hessian_mode = "psd"; orc.lib().orc_set_psd_flags(pf)
        o.compute_residual_and_hessian(spd=True)
        ref[name] = o.matrix().toarray()[:3 * NV, :3 * NV]
    orc.lib().orc_set_psd_flags(3)
    nv = offset + NV + 3
    pos = np.zeros((nv, 3)); pos[offset:offset + NV] = p; pos[:offset] = 7.0
    frozen = np.zeros(3 * nv, np.int32); frozen[3 * offset:3 * (offset + NV)] = o.frozen[:3 * NV]
    oe = np.zeros((3 * nv, 3 * nv)); oc = np.zeros((3 * nv, 3 * nv))
    L.emu_hessian_rows.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p] + [C.c_double] * 5 + [C.c_void_p, C.c_void_p]
    bad = L.emu_hessian_rows(N, M, offset, 3, pos.ctypes.data, frozen.ctypes.data, 1000.0, 1000.0, 100.0, dx, o.cloth_mass / dt ** 2,
                             oe.ctypes.data, oc.ctypes.data)
    assert bad == 0                                                   # padding slots untouched
    a, b = 3 * offset, 3 * (offset + NV)
    for name, out in (("e", oe), ("c", oc)):
        assert np.abs(out[a:b, a:b] - ref[name]).max() <= 2e-6 * np.abs(ref[name]).max(), name
        out[a:b, a:b] = 0
        assert not out.any()                                          # nothing written outside the cloth rows


@pytest.mark.parametrize("N,M,offset,seed", [(5, 4, 0, 0), (9, 37, 0, 1), (6, 66, 40, 2)])
def test_owner_computes_residual_and_energy_match_the_oracle(N, M, offset, seed):
    """k_residual_rows / k_energy_rows (fp64 tile kernels) against the oracle's Cloth.compute_residual / compute_energy restatement:
    wavy sheets with non-zero rest angles, velocities and a previous position"""
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import tsl_oracle as orc
    L = _build("emu_assembly")
    rng = np.random.default_rng(seed)
    dx, dt = 0.002, 5e-3
    NV, NF = (N + 1) * (M + 1), 2 * N * M
    tpos = np.array([[1., 1., -1.], [1.1, 1., -1.], [1., 1.1, -1.]])
    o = orc.OracleScene(N, M, dx, dt, tpos, np.array([[0, 1, 2]], np.int32), np.ones(3))
    i, j = np.meshgrid(np.arange(N + 1), np.arange(M + 1), indexing="ij")
    p = np.stack([i * dx, j * dx, 0.0 * i], -1).reshape(-1, 3).astype(float)
    p += rng.uniform(-0.2 * dx, 0.2 * dx, p.shape)
    p[:, 2] += 0.7 * dx * np.sin(i.ravel() * 0.9) * np.cos(j.ravel() * 0.7)       # both signs of the dihedral angles
    o.pos[:NV] = p
    o.prev_pos[:NV] = p + rng.uniform(-1e-5, 1e-5, p.shape)
    o.vel[:NV] = rng.uniform(-1e-2, 1e-2, p.shape)
    o.ref_angle[:] = rng.uniform(-0.3, 0.3, (NF, 3))
    o.nc = 0
    o._bind()
    Lo = orc.lib()
    Lo.orc_cloth_normals(o.cloth); Lo.orc_cloth_prepare_bending(o.cloth)
    Fb = np.zeros((NV, 3))
    Lo.orc_cloth_residual(o.cloth, orc._d(Fb), 15)
    E_ref = Lo.orc_cloth_energy(o.cloth, None)
    nv = offset + NV + 3
    def emb(a, fill=0.0):
        out = np.full((nv, 3), fill); out[offset:offset + NV] = a[:NV]; return out
    pos, prev, vel = emb(o.pos, 7.0), emb(o.prev_pos, 7.0), emb(o.vel)
    F = np.full(3 * nv, 123.0)
    L.emu_residual_energy_rows.restype = C.c_double
    L.emu_residual_energy_rows.argtypes = [C.c_int] * 4 + [C.c_void_p] * 4 + [C.c_double] * 6 + [C.c_void_p, C.c_void_p]
    grav = np.array([0.0, 0.0, -9.8])
    ra = np.ascontiguousarray(o.ref_angle)
    E = L.emu_residual_energy_rows(N, M, offset, nv, pos.ctypes.data, prev.ctypes.data, vel.ctypes.data, ra.ctypes.data, 1000.0, 1000.0, 100.0,
                                   dx, dt, o.cloth_mass, grav.ctypes.data, F.ctypes.data)
    assert abs(E - E_ref) <= 1e-12 * abs(E_ref), (E, E_ref)
    Fc = F.reshape(-1, 3)[offset:offset + NV]
    assert np.abs(Fc - Fb).max() <= 1e-11 * np.abs(Fb).max()
    F.reshape(-1, 3)[offset:offset + NV] = 123.0
    assert np.all(F == 123.0)                                         # only the cloth rows are written


@pytest.mark.parametrize("n0f,n1f,emf,emc", [(13, 21, 0, 0), (12, 35, 1, 1), (7, 9, 1, 0)])
def test_tiled_galerkin_matches_the_entrywise_kernel(n0f, n1f, emf, emc):
    """k_galerkin_tiled (fine stencil rows staged in shared memory per tile of 2 x 8 coarse vertices) against k_galerkin (one thread
    per coarse entry), odd and even grid sizes, row-major and element-major levels"""
    L = _build("emu_mg")
    rng = np.random.default_rng(n0f * 100 + n1f)
    nvf = n0f * n1f
    vf = rng.standard_normal((nvf, 25, 9)).astype(np.float32)
    # entries that point outside the grid hold zeros (as the fine levels are built)
    for v in range(nvf):
        i, j = divmod(v, n1f)
        for s in range(25):
            ii, jj = i + s // 5 - 2, j + s % 5 - 2
            if not (0 <= ii < n0f and 0 <= jj < n1f):
                vf[v, s] = 0
    nvc = ((n0f - 1) // 2 + 1) * ((n1f - 1) // 2 + 1)
    ref = np.zeros((nvc, 225), np.float32); out = np.zeros((nvc, 225), np.float32)
    L.emu_galerkin_pair(n0f, n1f, emf, emc, vf.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert np.abs(ref).max() > 1 and np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("N,M,off", [(12, 20, 0), (9, 33, 40)])
def test_tiled_galerkin_from_the_sliced_ell_matrix(N, M, off):
    """k_galerkin_sell_tiled (level 0 -> 1 straight from the sliced-ELL matrix, frozen DOFs masked out) against k_sell_to_stencil +
    k_galerkin<MASK>, on the cloth's own block pattern with extra rows before / after the cloth"""
    import scipy.sparse as sp
    import sys
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import tsl_oracle as orc
    L = _build("emu_mg")
    rng = np.random.default_rng(N)
    n0f, n1f = N + 1, M + 1
    NV = n0f * n1f
    f2v, cf, cp = orc.cloth_mesh(N, M)
    hi, hl = np.nonzero(cf > np.arange(2 * N * M)[:, None])
    hinge = np.stack([f2v[hi, hl], f2v[hi, (hl + 1) % 3], f2v[hi, (hl + 2) % 3], f2v[cf[hi, hl], cp[hi, hl]]], 1)
    r = np.concatenate([st[:, a] for st in (f2v, hinge) for a in range(st.shape[1]) for b in range(st.shape[1])])
    c = np.concatenate([st[:, b] for st in (f2v, hinge) for a in range(st.shape[1]) for b in range(st.shape[1])])
    nv = off + NV + 5
    A = sp.csr_matrix((np.ones(r.size), (r + off, c + off)), shape=(nv, nv)) + sp.identity(nv, format="csr")
    A.sum_duplicates(); A.sort_indices()
    rowptr, colidx = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    blocks = rng.standard_normal((colidx.size, 9)).astype(np.float32)
    mask = np.zeros(3 * NV, np.int32)
    mask[rng.integers(0, 3 * NV, 25)] = 1
    nvc = ((n0f - 1) // 2 + 1) * ((n1f - 1) // 2 + 1)
    ref = np.zeros((nvc, 225), np.float32); out = np.zeros((nvc, 225), np.float32)
    L.emu_galerkin_sell(off, n0f, n1f, nv, rowptr.ctypes.data_as(C.c_void_p), colidx.ctypes.data_as(C.c_void_p), blocks.ctypes.data_as(C.c_void_p),
                        mask.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert np.abs(ref).max() > 1 and np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max()
