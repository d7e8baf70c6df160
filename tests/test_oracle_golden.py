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
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(orc.__file__), "ti_emu"))
    import taichi as ti_emu
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
