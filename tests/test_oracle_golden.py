"""Pins the CPU oracle (oracle/csrc/tsl_oracle.c + oracle/tsl_oracle.py) against golden vectors that
were produced by executing the reference's own sources (oracle/gen_goldens.py, tier 1).
Tolerances are fp64 round-off only (different summation order / algebraically equivalent forms)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import tsl_oracle as orc

CLOTH_CASES = ["6x4_wavy", "6x4_flat", "8x8_wavy", "15x3_fold"]


def _rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_spd_projector_known_answers(golden_dir):
    g = np.load(os.path.join(golden_dir, "spd_projector.npz"))
    L = orc.lib()
    for D, K in ((3, 10), (9, 20)):
        a = g[f"in{D}"].copy()
        for t in range(a.shape[0]):
            m = np.ascontiguousarray(a[t])
            L.orc_spd_project(orc._d(m), D, K)
            ref = g[f"out{D}"][t]
            assert np.abs(m - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max()), (D, t)
    a2 = g["in2"]
    for t in range(a2.shape[0]):
        m = np.ascontiguousarray(a2[t])
        L.orc_spd_project_2d(orc._d(m))
        assert np.abs(m - g["out2"][t]).max() <= 1e-12 * max(1.0, np.abs(a2[t]).max())


def noise_sign_override(f2v, cf, pos, norm_dir):
    """Signs that the emulated reference produced for the topologically degenerate side tests
    (see cloth_neg() in oracle/csrc/tsl_oracle.c): recomputed with the same numpy expression the
    taichi stand-in evaluates, so they are bit-identical to the golden run."""
    import importlib.util
    import sys
    # the emulation module by PATH (another test may have registered an inert `taichi` stand-in under that name)
    ti_emu = sys.modules.get("_tsl_ti_emu")
    if ti_emu is None:
        path = os.path.join(os.path.dirname(orc.__file__), "ti_emu", "taichi", "__init__.py")
        spec = importlib.util.spec_from_file_location("_tsl_ti_emu", path, submodule_search_locations=[os.path.dirname(path)])
        ti_emu = importlib.util.module_from_spec(spec)
        sys.modules["_tsl_ti_emu"] = ti_emu
        spec.loader.exec_module(ti_emu)
    NF = f2v.shape[0]
    ov = -np.ones((NF, 3), np.int8)
    for i in range(NF):
        for l in range(3):
            i2 = cf[i, l]
            if i2 == -1:
                continue
            va, vb = f2v[i, (l + 1) % 2], f2v[i, l]
            if va in f2v[i2] and vb in f2v[i2]:
                n2 = norm_dir[i2].view(ti_emu._Arr)
                e = pos[va].view(ti_emu._Arr) - pos[vb].view(ti_emu._Arr)
                ov[i, l] = 1 if n2.dot(e) < 0 else 0
    return ov


def _dense_mat(nv):
    rowptr = (np.arange(nv + 1) * nv).astype(np.int32)
    colidx = np.tile(np.arange(nv, dtype=np.int32), nv)
    val = np.zeros((nv * nv, 3, 3))
    frozen = np.zeros(3 * nv, np.int32)
    mat = C.c_void_p(orc.lib().orc_mat_create(nv, orc._i(rowptr), orc._i(colidx), orc._d(val), orc._i(frozen)))
    keep = (rowptr, colidx, val, frozen)

    def dense():
        return val.reshape(nv, nv, 3, 3).transpose(0, 2, 1, 3).reshape(3 * nv, 3 * nv).copy()
    return mat, val, dense, keep


@pytest.mark.parametrize("name", CLOTH_CASES)
def test_cloth_terms_match_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"cloth_{name}.npz"))
    L = orc.lib()
    N, M = int(g["N"]), int(g["M"])
    f2v, cf, cp = orc.cloth_mesh(N, M)
    assert np.array_equal(f2v, g["f2v"])
    assert np.array_equal(cf, g["counter_face"])          # incl. the never-written entries (Q2)
    assert np.array_equal(cp, g["counter_point"])
    NV, NF = (N + 1) * (M + 1), 2 * N * M
    pos = np.ascontiguousarray(g["pos"]); prev = np.ascontiguousarray(g["prev_pos"]); vel = np.ascontiguousarray(g["vel"])
    ref = np.ascontiguousarray(g["ref_angle"]); grav = np.array([0, 0, -9.8])
    c = C.c_void_p(L.orc_cloth_create(N, M, 0, orc._f(g["dx"]), orc._f(g["dt"]), orc._f(g["mass"]), orc._i(f2v), orc._i(cf), orc._i(cp)))
    L.orc_cloth_bind(c, orc._d(pos), orc._d(prev), orc._d(vel), orc._d(ref), orc._d(grav), orc._f(g["Kl"]), orc._f(g["Ka"]), orc._f(g["Kb"]))
    ov = noise_sign_override(f2v, cf, pos, g["norm_dir"])
    assert (ov >= 0).sum() > 0
    L.orc_cloth_set_neg_override(c, ov.ctypes.data_as(C.c_char_p))
    L.orc_cloth_normals(c)
    L.orc_cloth_prepare_bending(c)
    nd = np.zeros((NF, 3)); mM = np.zeros((NF * 3, 3, 3)); mN = np.zeros((NF * 3, 3, 3))
    an = np.zeros((NF, 3)); he = np.zeros((NF, 3)); ci = np.zeros((NF, 3)); di = np.zeros((NF, 3))
    L.orc_cloth_get_derived(c, *[orc._d(x) for x in (nd, mM, mN, an, he, ci, di)])
    for got, key in ((nd, "norm_dir"), (mM, "mat_M"), (mN, "mat_N"), (an, "angle"), (he, "heights"), (ci, "c_i"), (di, "d_i")):
        assert _rel(got, g[key]) < 1e-11, key
    parts = np.zeros(4)
    U = L.orc_cloth_energy(c, orc._d(parts))
    assert abs(U - g["U"]) <= 1e-12 * abs(g["U"])
    assert abs(parts[2] - g["U_ma"]) <= 1e-12 * abs(g["U_ma"]) + 1e-18
    assert abs(parts[3] - g["U_bending"]) <= 1e-11 * abs(g["U_bending"]) + 1e-18
    assert abs(parts[0] + parts[1] - g["U_me"]) <= 1e-12 * abs(g["U_me"])
    F = np.zeros((NV, 3))
    for mask, key in ((15, "F_b"), (3, "F_me"), (4, "F_ma"), (8, "F_bending")):
        L.orc_cloth_residual(c, orc._d(F), mask)
        assert _rel(F, g[key]) < 1e-11, key
    mass = np.full(NV, float(g["mass"]))
    for spd in (0, 1):
        mat, val, dense, keep = _dense_mat(NV)
        L.orc_add_mass_diag(mat, orc._d(mass), orc._f(g["dt"]))
        L.orc_cloth_hessian_me(c, mat, spd)
        assert _rel(dense(), g[f"H_me_spd{spd}"]) < 1e-10, spd
    mat, val, dense, keep = _dense_mat(NV)
    L.orc_cloth_hessian_ma(c, mat)
    assert _rel(dense(), g["H_ma"]) < 1e-11
    mat, val, dense, keep = _dense_mat(NV)
    L.orc_cloth_hessian_bending(c, mat)
    assert _rel(dense(), g["H_bending"]) < 1e-10
    assert L.orc_mat_missing(mat) == 0
    dkl = np.zeros((NV, 3)); dka = np.zeros((NV, 3)); dkb = np.zeros((NV, 3))
    L.orc_cloth_compute_deri(c, orc._d(dkl), orc._d(dka), orc._d(dkb))
    assert _rel(dkl, g["d_kl"]) < 1e-11 and _rel(dka, g["d_ka"]) < 1e-11 and _rel(dkb, g["d_kb"]) < 1e-11
    ref2 = ref.copy()
    L.orc_cloth_update_ref_angle(c, orc._d(ref2), orc._f(0.02))
    assert _rel(ref2, g["ref_angle_after_k0p02"]) < 1e-12
    assert np.abs(ref2 - ref).max() > 1e-3          # the plastic branch was exercised
    L.orc_cloth_destroy(c)


# ----------------------------------------------------------------------------- scene level
def _scene_from_golden(g):
    N, M = int(g["cloth_N"]), int(g["cloth_M"])
    off = int(g["table_offset"])
    NFc = 2 * N * M
    s = orc.OracleScene(N, M, float(g["cloth_dx"]), float(g["dt"]), g["pos0"][off:], g["faces"][NFc:] - off, g["mass"][off:],
                        Kb=float(g["Kb"]), k_angle=float(g["k_angle"]), k_contact=float(g["k_contact"]),
                        eps_contact=float(g["eps_contact"]), eps_v=float(g["eps_v"]), mu=float(g["mu"]))
    s.pos[:] = g["pos0"]; s.prev_pos[:] = g["pos0"]; s.vel[:] = g["vel0"]; s.ref_angle[:] = g["ref_angle0"]
    return s


def _scene_override(s, pos):
    """noise-sign override for the current cloth state (see noise_sign_override)"""
    import importlib.util
    import sys
    # the emulation module by PATH (another test may have registered an inert `taichi` stand-in under that name)
    ti_emu = sys.modules.get("_tsl_ti_emu")
    if ti_emu is None:
        path = os.path.join(os.path.dirname(orc.__file__), "ti_emu", "taichi", "__init__.py")
        spec = importlib.util.spec_from_file_location("_tsl_ti_emu", path, submodule_search_locations=[os.path.dirname(path)])
        ti_emu = importlib.util.module_from_spec(spec)
        sys.modules["_tsl_ti_emu"] = ti_emu
        spec.loader.exec_module(ti_emu)
    p = pos[:s.NVc]
    nd = np.zeros((s.NFc, 3))
    for i in range(s.NFc):
        a, b, c = (p[s.f2v[i, k]].view(ti_emu._Arr) for k in range(3))
        nd[i] = (b - a).cross(c - b).normalized()      # Cloth.compute_normal_dir as the stand-in evaluates it
    return noise_sign_override(s.f2v, s.cf, p, nd)


def test_box_mesher_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "bouncing.npz"))
    off = int(g["table_offset"]); NFc = 2 * int(g["cloth_N"]) * int(g["cloth_M"])
    pos, tets, faces, mass = orc.box_mesh(0.07, 9, 9, 2, (-0.035, -0.035, -0.00875))
    assert np.abs(pos - g["pos0"][off:]).max() < 1e-15
    assert np.array_equal(tets, g["tet_vertices"])
    assert np.array_equal(faces + off, g["faces"][NFc:])
    assert _rel(mass, g["mass"][off:]) < 1e-12


def test_scene_contact_query_and_first_newton_iterations(golden_dir):
    g = np.load(os.path.join(golden_dir, "bouncing.npz"))
    s = _scene_from_golden(g)
    L = orc.lib()
    frame = 1
    s.prev_pos[:] = s.pos
    s.calc_vn(); s.projection_query(); s.contact_analysis()
    tv = slice(int(g["table_offset"]), None)
    # vertex normals of surface vertices (interior tet vertices are 0/0 = NaN in both)
    assert np.allclose(s.vn, g[f"f{frame}_vn"], rtol=0, atol=1e-12, equal_nan=True)
    assert np.array_equal(s.proj_flag, g[f"f{frame}_proj_flag"])
    assert np.array_equal(s.proj_dir, g[f"f{frame}_proj_dir"])
    assert np.array_equal(s.proj_idx, g[f"f{frame}_proj_idx"])
    assert np.abs(s.proj_w - g[f"f{frame}_proj_w"]).max() < 1e-12
    nc = int(g[f"f{frame}_nc"])
    assert s.nc == nc and nc > 0
    # constraint SET is bit-exact; the reference's append order is nondeterministic, ours is by vertex
    key = lambda a: sorted(map(tuple, a))
    assert key(s.c_idx[:nc]) == key(g[f"f{frame}_const_idx"])
    o1 = np.argsort(s.c_idx[:nc, 3]); o2 = np.argsort(g[f"f{frame}_const_idx"][:, 3])
    for mine, k in ((s.c_w, "const_w"), (s.c_k, "const_k"), (s.c_mu, "const_mu"), (s.c_dx0, "const_dx0"), (s.c_T, "const_T"), (s.c_n, "const_n")):
        assert _rel(mine[:nc][o1], g[f"f{frame}_{k}"][o2]) < 1e-10, k
    s._build_pattern()
    for it in (1, 2):
        s.pos[:] = g[f"f{frame}_it{it}_pos"]
        ov = _scene_override(s, s.pos)
        L.orc_cloth_set_neg_override(s.cloth, ov.ctypes.data_as(C.c_char_p))
        E0 = s.compute_energy()
        assert abs(E0 - g[f"f{frame}_newton_log"][it - 1, 0]) <= 1e-12 * abs(E0)
        s.compute_residual_and_hessian(spd=True)
        assert _rel(s.F, g[f"f{frame}_it{it}_F"]) < 1e-10
        import scipy.sparse as sp
        Hg = sp.csr_matrix((g[f"f{frame}_it{it}_H_data"], g[f"f{frame}_it{it}_H_indices"], g[f"f{frame}_it{it}_H_indptr"]),
                           shape=(3 * s.NV, 3 * s.NV))
        D = (s.matrix() - Hg).tocoo()
        assert np.abs(D.data).max() <= 1e-9 * np.abs(Hg.data).max()
        p = s.solve(s.F)
        assert _rel(p, g[f"f{frame}_it{it}_p"]) < 1e-7
    L.orc_cloth_set_neg_override(s.cloth, None)


def test_scene_rollout_positions_match_reference(golden_dir):
    """Forward rollout with the oracle's own Newton path (canonical side-test rule, SuperLU): the converged
    positions must agree with the reference's to the Newton tolerance (1e-7 * dt on the last step size)."""
    g = np.load(os.path.join(golden_dir, "bouncing.npz"))
    s = _scene_from_golden(g)
    T = int(g["T"])
    for frame in range(1, T):
        log = []
        s.time_step(log=log)
        assert s.nc == int(g[f"f{frame}_nc"])
        assert sorted(map(tuple, s.c_idx[:s.nc])) == sorted(map(tuple, g[f"f{frame}_const_idx"]))
        err = np.abs(s.pos - g[f"f{frame}_pos"]).max()
        assert err < 2e-9, (frame, err)           # both stop at |p|_inf < 5e-10 m with a linear rate < 0.5
        assert np.abs(s.vel - g[f"f{frame}_vel"]).max() < 2e-9 / float(g["dt"]) * 1.01
        assert np.abs(s.ref_angle - g[f"f{frame}_ref_angle"]).max() < 1e-6
    assert abs(s.reward() - float(g["reward"])) < 1e-7


def test_scene_adjoint_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "bouncing.npz"))
    s = _scene_from_golden(g)
    T = int(g["T"])
    L = orc.lib()
    gr = orc.OracleGrad(s, T)
    gr.pos_buffer[:] = g["pos_buffer"]                       # reference trajectory: isolates the backward pass
    gr.ref_angle_buffer[0] = g["ref_angle0"]
    for f in range(1, T):
        gr.ref_angle_buffer[f] = g[f"f{f}_ref_angle"]
    gr.get_loss_table()
    assert np.array_equal(gr.pos_grad, g["pos_grad_seed"])
    for j in range(T - 1, 0, -1):
        ov = _scene_override(s, gr.pos_buffer[j])
        L.orc_cloth_set_neg_override(s.cloth, ov.ctypes.data_as(C.c_char_p))
        gr.transfer_grad(j)
        assert s.nc == int(g[f"b{j}_nc"])
        assert sorted(map(tuple, s.c_idx[:s.nc])) == sorted(map(tuple, g[f"b{j}_const_idx"]))
        assert _rel(gr.d_kb, g[f"b{j}_d_kb"]) < 1e-10
        assert _rel(gr.z, g[f"b{j}_z"]) < 1e-8, j
        assert _rel(gr.pos_grad, g[f"b{j}_pos_grad"]) < 1e-8, j
        assert _rel(s.tmp_z_frozen, g[f"b{j}_tmp_z_frozen"]) < 1e-8
        assert abs(gr.grad_kb - float(g[f"b{j}_grad_kb"])) <= 1e-8 * abs(float(g[f"b{j}_grad_kb"]))
    L.orc_cloth_set_neg_override(s.cloth, None)
    assert abs(gr.grad_kb - float(g["grad_kb"])) <= 1e-8 * abs(float(g["grad_kb"]))


@pytest.mark.parametrize("name", ["box_4x3x3", "tactile"])
def test_tet_terms_match_reference(golden_dir, name):
    """Elastic.compute_energy / get_force / compute_residual / compute_Hessian / compute_deri of both tet models
    (engine/model_elastic_offset.py, engine/model_elastic_tactile.py) against the emulated reference run"""
    g = np.load(os.path.join(golden_dir, f"tet_{name}.npz"))
    kind = orc.TET_TACTILE if str(g["kind"]) == "tactile" else orc.TET_BOX
    t = orc.Tets(kind, g["rest"], g["tets"], float(g["density"]), g["mu"], g["lam"], g["alpha"], g["dt"], g["gravity"], g["ext_force"])
    assert _rel(t.B, g["F_B"]) < 1e-12 and _rel(t.W, g["F_W"]) < 1e-12 and _rel(t.m, g["F_m"]) < 1e-12
    pos = np.ascontiguousarray(g["pos"]); prev = np.ascontiguousarray(g["prev_pos"]); vel = np.ascontiguousarray(g["vel"])
    U = t.energy(pos, prev, vel)
    assert abs(U - g["U"]) <= 1e-12 * abs(g["U"])
    assert _rel(t.force(pos), g["F_f"]) < 1e-12
    assert _rel(t.residual(pos, prev, vel), g["F_b"]) < 1e-12
    for spd in (0, 1):
        mat, val, dense, keep = _dense_mat(t.nv)
        t.hessian(pos, 0, mat, spd)
        assert orc.lib().orc_mat_missing(mat) == 0
        ref = g[f"H_spd{spd}"]
        assert np.abs(dense() - ref).max() <= 1e-9 * np.abs(ref).max(), spd
    if kind == orc.TET_BOX:
        # the reference's box model has no projection: spd is ignored
        assert np.array_equal(g["H_spd0"], g["H_spd1"])
    d_mu, d_lam = t.deri(pos)
    assert _rel(d_mu, g["d_mu"]) < 1e-12 and _rel(d_lam, g["d_lam"]) < 1e-12
