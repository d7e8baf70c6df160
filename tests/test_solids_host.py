"""Host build of thinshelllab_b200/csrc/tsl_solids.cuh -- the __host__ __device__ element functions the CUDA kernels of the
tetrahedral bodies and of the general (moving-triangle) contact call -- against the reference goldens and the oracle.
No GPU needed: the header is compiled with g++ by this test (tests/csrc/solids_host.cpp)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tsl_oracle as orc  # noqa: E402

_dp = C.POINTER(C.c_double)
# forward (projected) matrices: |library's exact PSD clamp - reference's thresholded SPD_Projector| / largest entry (quirk Q8)
PROJ_TOL = 5e-3


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("solids") / "solids_host.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "csrc", "solids_host.cpp")])
    return C.CDLL(out)


def _d(a):
    return a.ctypes.data_as(_dp)


def _expand(tets, blocks, nv):
    H = np.zeros((3 * nv, 3 * nv))
    for c, tv in enumerate(tets):
        for a in range(4):
            for b in range(4):
                H[3 * tv[a]:3 * tv[a] + 3, 3 * tv[b]:3 * tv[b] + 3] += blocks[c, a, b]
    return H


@pytest.mark.parametrize("name", ["box_4x3x3", "tactile"])
def test_tet_element_functions_match_reference(host, golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"tet_{name}.npz"))
    kind = 1 if str(g["kind"]) == "tactile" else 0
    tets = np.ascontiguousarray(g["tets"], np.int32)
    B = np.ascontiguousarray(g["F_B"]); W = np.ascontiguousarray(g["F_W"]); pos = np.ascontiguousarray(g["pos"])
    nc, nv = tets.shape[0], pos.shape[0]
    m = g["F_m"]; dt = float(g["dt"])
    for spd in (0, 1):
        E = np.zeros(nc); G = np.zeros((nc, 4, 3)); H9 = np.zeros((nc, 81)); blocks = np.zeros((nc, 4, 4, 3, 3))
        host.host_tets(kind, C.c_double(g["mu"]), C.c_double(g["lam"]), C.c_double(g["alpha"]), nc, tets.ctypes.data_as(C.c_void_p),
                       _d(B), _d(W), _d(pos), int(spd and kind == 1), _d(E), _d(G), _d(H9), _d(blocks))
        H = _expand(tets, blocks, nv) + np.kron(np.diag(m / dt ** 2), np.eye(3))
        ref = g[f"H_spd{spd}"]
        # spd = 1 (tactile): the library clamps eigenvalues exactly (cyclic Jacobi), the reference's SPD_Projector stops at its
        # sweep / threshold limits (quirk Q8): forward-matrix entries agree to the reference's own approximation error
        assert np.abs(H - ref).max() <= (PROJ_TOL if (spd and kind == 1) else 1e-9) * np.abs(ref).max(), spd
    # energy: the golden's U includes the vertex terms (gravity, external force, inertia)
    X = pos - g["prev_pos"] - g["vel"] * dt
    U = E.sum() - (m[:, None] * pos * g["gravity"][None]).sum() - (g["ext_force"] * pos).sum() + 0.5 * (m * (X * X).sum(1)).sum() / dt ** 2
    assert abs(U - g["U"]) <= 1e-11 * abs(g["U"])
    # force: F_f = -dE/dx + m g + ext
    F = np.zeros((nv, 3))
    for c, tv in enumerate(tets):
        for q in range(4):
            F[tv[q]] -= G[c, q]
    F += m[:, None] * g["gravity"][None] + g["ext_force"]
    assert np.abs(F - g["F_f"]).max() <= 1e-11 * np.abs(g["F_f"]).max()


@pytest.mark.parametrize("name", ["box_4x3x3", "tactile"])
def test_tet_parameter_derivatives_match_reference(host, golden_dir, name):
    """Elastic.compute_deri (model_elastic_offset.py:423-438 / model_elastic_tactile.py:329-347)"""
    g = np.load(os.path.join(golden_dir, f"tet_{name}.npz"))
    kind = 1 if str(g["kind"]) == "tactile" else 0
    tets = np.ascontiguousarray(g["tets"], np.int32)
    B = np.ascontiguousarray(g["F_B"]); W = np.ascontiguousarray(g["F_W"]); pos = np.ascontiguousarray(g["pos"])
    nc, nv = tets.shape[0], pos.shape[0]
    gm = np.zeros((nc, 4, 3)); gl = np.zeros((nc, 4, 3))
    host.host_tets_deri(kind, C.c_double(g["mu"]), C.c_double(g["lam"]), C.c_double(g["alpha"]), nc, tets.ctypes.data_as(C.c_void_p),
                        _d(B), _d(W), _d(pos), _d(gm), _d(gl))
    dm = np.zeros((nv, 3)); dl = np.zeros((nv, 3))
    np.add.at(dm, tets.ravel(), gm.reshape(-1, 3)); np.add.at(dl, tets.ravel(), gl.reshape(-1, 3))
    assert np.abs(dm - g["d_mu"]).max() <= 1e-11 * np.abs(g["d_mu"]).max()
    assert np.abs(dl - g["d_lam"]).max() <= 1e-11 * max(np.abs(g["d_lam"]).max(), 1e-300)


def test_psd_clamp_is_the_exact_projection(host, golden_dir):
    """psd_clamp<9> / psd_project_3x3 (own cyclic-Jacobi eigen-clamp) against numpy.linalg.eigh, and the distance to the reference's
    SPD_Projector outputs (golden, generated from code/engine/linalg.py:15-148): that one is approximate (quirk Q8)"""
    g = np.load(os.path.join(golden_dir, "spd_projector.npz"))
    for n, fn in ((9, host.host_spd9), (3, host.host_spd3)):
        devs = []
        for A, R in zip(g[f"in{n}"], g[f"out{n}"]):
            M = np.ascontiguousarray(A.copy())
            fn(_d(M), 20)
            S = 0.5 * (A + A.T)
            w, V = np.linalg.eigh(S)
            exact = (V * np.maximum(w, 0)) @ V.T
            assert np.abs(M - exact).max() <= 1e-12 * np.abs(S).max(), n
            devs.append(np.abs(M - R).max() / max(np.abs(R).max(), 1e-30))
        assert np.median(devs) < 1e-7 and max(devs) < PROJ_TOL, (n, np.median(devs), max(devs))
    rng = np.random.default_rng(3)
    for _ in range(50):
        S = rng.normal(size=(9, 9)); S = S + S.T
        a = np.ascontiguousarray(S.copy())
        host.host_spd9(_d(a), 20)
        w, V = np.linalg.eigh(S)
        assert np.abs(a - (V * np.maximum(w, 0)) @ V.T).max() <= 1e-12 * np.abs(S).max()
    # already PSD: input bits are kept
    P = rng.normal(size=(9, 9)); P = P @ P.T
    a = np.ascontiguousarray(P.copy())
    host.host_spd9(_d(a), 20)
    assert np.array_equal(a, P)


def test_general_contact_matches_oracle(host):
    """normal part of BaseScene.contact_energy over (f0, f1, f2, v): gradient-free check of the 12x12 against the oracle's
    restatement of contact_diff.det / cross (oracle/csrc/tsl_oracle.c), with and without SPD_Projector(9, K=20)"""
    rng = np.random.default_rng(11)
    L = orc.lib()
    L.orc_contacts_create.restype = C.c_void_p
    L.orc_mat_create.restype = C.c_void_p
    for trial in range(30):
        x = np.zeros((4, 3))
        x[0] = rng.normal(size=3) * 0.01
        x[1] = x[0] + [0.004, 0, 0] + rng.normal(size=3) * 2e-4
        x[2] = x[0] + [0, 0.004, 0] + rng.normal(size=3) * 2e-4
        x[3] = x[0] + [0.001, 0.001, 1e-4] + rng.normal(size=3) * 5e-5
        idx = np.array([[0, 1, 2, 3]], np.int32)
        w = np.array([[0.5, 0.25, 0.25]]); k = np.zeros(1); mu = np.ones(1); dx0 = np.zeros((1, 3)); T = np.zeros((1, 6)); n = np.zeros((1, 3))
        T[0, 0] = 1; T[0, 4] = 1
        for spd in (0, 1):
            cs = C.c_void_p(L.orc_contacts_create(1, orc._i(idx), orc._d(w), orc._d(k), orc._d(mu), orc._d(dx0), orc._d(T), orc._d(n),
                                                  orc._f(1e4), orc._f(4e-4), orc._f(0.01), orc._f(5e-3)))
            rowptr = (np.arange(5) * 4).astype(np.int32); colidx = np.tile(np.arange(4, dtype=np.int32), 4)
            val = np.zeros((16, 3, 3)); frozen = np.zeros(12, np.int32)
            mat = C.c_void_p(L.orc_mat_create(4, orc._i(rowptr), orc._i(colidx), orc._d(val), orc._i(frozen)))
            F = np.zeros(12)
            L.orc_contact_grad_hess(cs, orc._d(x), orc._i(frozen), orc._d(F), mat, spd)
            ref = val.reshape(4, 4, 3, 3)
            G = np.zeros(9); H = np.zeros(81); blocks = np.zeros((4, 4, 3, 3))
            act = host.host_contact(_d(x), C.c_double(1e4), C.c_double(4e-4), spd, _d(G), _d(H), _d(blocks))
            assert act == int(np.abs(F).max() > 0)
            if not act:
                continue
            # friction with k = 0 contributes f1-terms times 0: the oracle matrix is the normal part only
            assert np.abs(blocks - ref).max() <= (PROJ_TOL if spd else 1e-9) * np.abs(ref).max(), (trial, spd)
            Fh = np.zeros((4, 3))
            Fh[1:] = G.reshape(3, 3); Fh[0] = -G.reshape(3, 3).sum(0)
            assert np.abs(Fh.ravel() - F).max() <= 1e-10 * np.abs(F).max()
            L.orc_contacts_destroy(cs); L.orc_mat_destroy(mat)
