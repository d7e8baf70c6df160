"""Host driver of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module; the product package (thinshelllab_b200) never does.

It restates, on top of oracle/csrc/tsl_oracle.c, the reference's scene-level control flow for the
cloth + frozen-table scene family (Scene_bouncing):
  * BaseScene.time_step / newton_step / line search  (engine/BaseScene.py:1159-1230, 1327-1370)
  * Scene_bouncing.contact_analysis / timestep_finish  (task_scene/Scene_bouncing.py:91-121)
  * geometry.projection_query                          (engine/geometry.py:223-229)
  * analytic_grad_system.Grad.transfer_grad            (engine/analytic_grad_system.py:115-160)
  * Elastic box mesher / surface / lumped mass         (engine/model_elastic_offset.py:240-245,293-376)
The linear solve is SciPy SuperLU (fp64 direct), standing in for the reference's cupyx spsolve
(engine/sparse_solver.py:103).  Parity pinning: tests/test_oracle_golden.py.
"""
import ctypes as C
import os
import subprocess
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build_lib(force=False):
    so = os.path.join(_HERE, "_build", "libtsl_oracle.so")
    src = os.path.join(_HERE, "csrc", "tsl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build_lib())
        L.orc_mat_create.restype = C.c_void_p
        L.orc_cloth_create.restype = C.c_void_p
        L.orc_contacts_create.restype = C.c_void_p
        L.orc_cloth_energy.restype = C.c_double
        L.orc_contact_energy.restype = C.c_double
        L.orc_vertex_energy.restype = C.c_double
        L.orc_tets_create.restype = C.c_void_p
        L.orc_tets_energy.restype = C.c_double
        _LIB = L
    return _LIB


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def _i(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_ip)


def _f(x):
    return C.c_double(float(x))


# ----------------------------------------------------------------------------- tetrahedral bodies
TET_BOX, TET_TACTILE = 0, 1


class Tets:
    """one Elastic body of the reference: kind TET_BOX (engine/model_elastic_offset.py) or TET_TACTILE
    (engine/model_elastic_tactile.py); arrays are body-local"""

    def __init__(self, kind, rest, tets, density, mu, lam, alpha, dt, gravity, ext=None):
        L = lib()
        self.kind = kind
        self.rest = np.ascontiguousarray(rest, np.float64)
        self.tets = np.ascontiguousarray(tets, np.int32)
        self.nv, self.nc = self.rest.shape[0], self.tets.shape[0]
        self.B = np.zeros((self.nc, 3, 3)); self.W = np.zeros(self.nc); self.m = np.zeros(self.nv)
        L.orc_tets_rest(self.nv, self.nc, _i(self.tets), _d(self.rest), _f(density), _d(self.B), _d(self.W), _d(self.m))
        self.gravity = np.ascontiguousarray(gravity, np.float64)
        self.ext = None if ext is None else np.ascontiguousarray(ext, np.float64)
        self.mu, self.lam, self.alpha, self.dt = float(mu), float(lam), float(alpha), float(dt)
        self._make()

    def _make(self):
        L = lib()
        self.h = C.c_void_p(L.orc_tets_create(self.kind, self.nv, self.nc, _i(self.tets), _d(self.B), _d(self.W), _d(self.m),
                                              _f(self.mu), _f(self.lam), _f(self.alpha), _f(self.dt), _d(self.gravity),
                                              _d(self.ext) if self.ext is not None else None))

    def set_params(self, mu, lam):
        lib().orc_tets_destroy(self.h)
        self.mu, self.lam = float(mu), float(lam)
        self._make()

    def energy(self, pos, prev, vel):
        return lib().orc_tets_energy(self.h, _d(pos), _d(prev), _d(vel))

    def force(self, pos):
        F = np.zeros((self.nv, 3)); lib().orc_tets_force(self.h, _d(pos), _d(F)); return F

    def residual(self, pos, prev, vel):
        F = np.zeros((self.nv, 3)); lib().orc_tets_residual(self.h, _d(pos), _d(prev), _d(vel), _d(F)); return F

    def hessian(self, pos, offset, mat, spd):
        lib().orc_tets_hessian(self.h, _d(pos), int(offset), mat, int(spd))

    def deri(self, pos):
        a = np.zeros((self.nv, 3)); b = np.zeros((self.nv, 3)); lib().orc_tets_deri(self.h, _d(pos), _d(a), _d(b)); return a, b


# ----------------------------------------------------------------------------- meshers
def cloth_mesh(N, M):
    """Cloth.init_mesh (engine/model_fold_offset.py:929-1018)"""
    NF = 2 * N * M
    f2v = np.zeros((NF, 3), np.int32)
    cf = np.zeros((NF, 3), np.int32)
    cp = np.zeros((NF, 3), np.int32)
    lib().orc_cloth_init_mesh(N, M, _i(f2v), _i(cf), _i(cp))
    return f2v, cf, cp


def box_mesh(Len, Nx, Ny, Nz, off, density=2000.0):
    """Elastic box: get_vertices / init_pos / get_surface_indices
    (engine/model_elastic_offset.py:240-245, 279-312, 346-376).  Returns pos, tets, faces, mass."""
    n = np.array([Nx, Ny, Nz])
    dx = Len / (n.max() - 1)

    def i2p(I):
        return (I[0] * n[1] + I[1]) * n[2] + I[2]

    nv = int(n.prod())
    ox = np.zeros((nv, 3))
    for x in range(Nx):
        for y in range(Ny):
            for z in range(Nz):
                ox[i2p((x, y, z))] = np.array([x, y, z]) * dx
    tets = np.zeros((5 * int((n - 1).prod()), 4), np.int32)

    def set_element(e, I, verts):
        for i in range(4):
            c = [I[k] + (((verts[i] >> k) ^ I[k]) & 1) for k in range(3)]
            tets[e, i] = i2p(c)

    for x in range(Nx - 1):
        for y in range(Ny - 1):
            for z in range(Nz - 1):
                I = (x, y, z)
                e = ((x * (Ny - 1) + y) * (Nz - 1) + z) * 5
                for i, j in enumerate([0, 3, 5, 6]):
                    set_element(e + i, I, (j, j ^ 1, j ^ 2, j ^ 4))
                set_element(e + 4, I, (1, 2, 4, 7))
    mass = np.zeros(nv)
    for c in range(tets.shape[0]):
        v = tets[c]
        Ds = np.stack([ox[v[i]] - ox[v[3]] for i in range(3)], axis=1)
        W = abs(np.linalg.det(Ds)) / 6
        for i in range(4):
            mass[v[i]] += W / 4 * density
    pos = ox + np.asarray(off, dtype=np.float64)

    def check(u):
        ans, rest = 0, int(u)
        for i in range(3):
            k = rest % n[2 - i]
            rest //= n[2 - i]
            if k == 0:
                ans |= 1 << (i * 2)
            if k == n[2 - i] - 1:
                ans |= 1 << (i * 2 + 1)
        return ans

    su = sum((n[i] - 1) * (n[(i + 1) % 3] - 1) for i in range(3))
    faces = np.zeros((int(2 * su * 2), 3), np.int32)
    cnt = 0
    for c in range(tets.shape[0]):
        if c % 5 != 4:
            for i in (0, 2, 3):
                verts = [int(tets[c][(i + j) % 4]) for j in range(3)]
                if check(verts[0]) & check(verts[1]) & check(verts[2]):
                    v3 = int(tets[c][(i + 3) % 4])
                    nrm = np.cross(pos[verts[1]] - pos[verts[0]], pos[verts[2]] - pos[verts[0]])
                    if nrm.dot(pos[v3] - pos[verts[0]]) > 0:
                        verts[1], verts[2] = verts[2], verts[1]
                    faces[cnt] = verts
                    cnt += 1
    return pos, tets, faces, mass


# ----------------------------------------------------------------------------- scene
class OracleScene:
    """cloth (vertex offset 0) + one frozen box body ("table"), Scene_bouncing style."""

    def __init__(self, N, M, dx, dt, table_pos, table_faces, table_mass, *, rho=40.0, Kl=1000.0, Ka=1000.0,
                 Kb=100.0, k_angle=3.14, k_contact=40000.0, eps_contact=4e-4, eps_v=0.01, mu=0.5,
                 damping=1.0, gravity=(0.0, 0.0, -9.8), max_n_constraints=10000, extra_frozen_vertices=(), grid_h=0.003, grid_n=0):
        L = lib()
        self.grid = (grid_h, grid_n)
        self.hessian_mode = "reference"
        self.N, self.M, self.dx, self.dt = N, M, dx, dt
        self.NVc = (N + 1) * (M + 1)
        self.NFc = 2 * N * M
        self.f2v, self.cf, self.cp = cloth_mesh(N, M)
        self.cloth_mass = rho * dx * dx
        nt = table_pos.shape[0]
        self.NV = self.NVc + nt
        self.table_offset = self.NVc
        self.pos = np.zeros((self.NV, 3))
        self.pos[self.NVc:] = table_pos
        self.prev_pos = self.pos.copy()
        self.vel = np.zeros((self.NV, 3))
        self.mass = np.concatenate([np.full(self.NVc, self.cloth_mass), np.asarray(table_mass, np.float64)])
        self.frozen = np.zeros(3 * self.NV, np.int32)
        self.frozen[3 * self.NVc:] = 1
        for v in extra_frozen_vertices:
            self.frozen[3 * v:3 * v + 3] = 1
        self.faces = np.ascontiguousarray(np.concatenate([self.f2v, np.asarray(table_faces, np.int32) + self.NVc]), np.int32)
        self.body_v = [(0, self.NVc), (self.NVc, self.NV)]
        self.body_f = [(0, self.NFc), (self.NFc, self.faces.shape[0])]
        self.Kl, self.Ka, self.Kb, self.k_angle = Kl, Ka, Kb, k_angle
        self.k_contact, self.eps_contact, self.eps_v, self.mu = k_contact, eps_contact, eps_v, mu
        self.damping = damping
        self.gravity = np.array(gravity, np.float64)
        self.ref_angle = np.zeros((self.NFc, 3))
        self.border_flag = np.zeros(self.NV, np.int32)
        nb = 2
        self.proj_flag = np.zeros((nb, self.NV), np.int32)
        self.proj_dir = np.zeros((nb, self.NV), np.int32)
        self.proj_idx = np.zeros((nb, self.NV, 3), np.int32)
        self.proj_w = np.zeros((nb, self.NV, 3))
        self.vn = np.zeros((self.NV, 3))
        self.max_nc = max_n_constraints
        self.nc = 0
        self.c_idx = np.zeros((self.max_nc, 4), np.int32)
        self.c_w = np.zeros((self.max_nc, 3)); self.c_k = np.zeros(self.max_nc); self.c_mu = np.zeros(self.max_nc)
        self.c_dx0 = np.zeros((self.max_nc, 3)); self.c_T = np.zeros((self.max_nc, 2, 3)); self.c_n = np.zeros((self.max_nc, 3))
        self.cloth = C.c_void_p(L.orc_cloth_create(N, M, 0, _f(dx), _f(dt), _f(self.cloth_mass), _i(self.f2v), _i(self.cf), _i(self.cp)))
        self.F = np.zeros(3 * self.NV)
        self.tmp_z_frozen = np.zeros(3 * self.NV)
        self.timing = {}
        # static stencil pairs of the block pattern: triangles + hinges + diagonal
        hi, hl = np.nonzero(self.cf > np.arange(self.NFc)[:, None])
        opp = self.f2v[self.cf[hi, hl], self.cp[hi, hl]]
        hinge = np.stack([self.f2v[hi, hl], self.f2v[hi, (hl + 1) % 3], self.f2v[hi, (hl + 2) % 3], opp], 1)
        self.hinges = hinge
        self._static_pairs = self._pairs(self.f2v) + self._pairs(hinge) + [(np.arange(self.NV), np.arange(self.NV))]
        self._pattern_nc = -1

    # ---- helpers
    @staticmethod
    def _pairs(st):
        k = st.shape[1]
        return [(st[:, a].astype(np.int64), st[:, b].astype(np.int64)) for a in range(k) for b in range(k)]

    def _tick(self, key, t0):
        self.timing[key] = self.timing.get(key, 0.0) + time.perf_counter() - t0

    def _bind(self):
        lib().orc_cloth_bind(self.cloth, _d(self.pos), _d(self.prev_pos), _d(self.vel), _d(self.ref_angle),
                             _d(self.gravity), _f(self.Kl), _f(self.Ka), _f(self.Kb))

    def _contacts(self):
        return C.c_void_p(lib().orc_contacts_create(
            self.nc, _i(self.c_idx), _d(self.c_w), _d(self.c_k), _d(self.c_mu), _d(self.c_dx0), _d(self.c_T), _d(self.c_n),
            _f(self.k_contact), _f(self.eps_contact), _f(self.eps_v), _f(self.dt)))

    def _build_pattern(self):
        pr = list(self._static_pairs)
        if self.nc:
            pr += self._pairs(self.c_idx[:self.nc])
        r = np.concatenate([p[0] for p in pr]); c = np.concatenate([p[1] for p in pr])
        A = sp.csr_matrix((np.ones(r.size, np.int8), (r, c)), shape=(self.NV, self.NV))
        A.sum_duplicates(); A.sort_indices()
        self.rowptr = A.indptr.astype(np.int32); self.colidx = A.indices.astype(np.int32)
        self.val = np.zeros((self.colidx.size, 3, 3))
        self.mat = C.c_void_p(lib().orc_mat_create(self.NV, _i(self.rowptr), _i(self.colidx), _d(self.val), _i(self.frozen)))

    def matrix(self):
        """assembled H as scipy CSR (3NV x 3NV)"""
        return sp.bsr_matrix((self.val, self.colidx, self.rowptr), shape=(3 * self.NV, 3 * self.NV)).tocsr()

    # ---- geometry.projection_query + Scene_bouncing.contact_analysis
    def calc_vn(self):
        lib().orc_calc_vn(self.NV, self.faces.shape[0], _i(self.faces), _d(self.pos), _d(self.vn))

    def projection_query(self):
        L = lib()
        L.orc_set_grid(_f(self.grid[0]), int(self.grid[1]))
        for b, (fs, fe) in enumerate(self.body_f):
            for b2, (vs, ve) in enumerate(self.body_v):
                if b2 != b:
                    L.orc_project_pair(self.NV, _d(self.pos), _d(self.vn), _i(self.faces), fs, fe, vs, ve, _i(self.border_flag),
                                       _i(self.proj_flag[b]), _i(self.proj_dir[b]), _i(self.proj_idx[b]), _d(self.proj_w[b]))

    def contact_analysis(self):
        # Scene_bouncing.contact_analysis: cloth vertices against the table surface only
        b = 1
        nc = lib().orc_contact_pair_analysis(
            _d(self.pos), _d(self.prev_pos), 0, self.NVc, _f(self.mu), _f(self.k_contact), _f(self.eps_contact),
            _i(self.proj_flag[b]), _i(self.proj_dir[b]), _i(self.proj_idx[b]), _d(self.proj_w[b]),
            0, self.max_nc, _i(self.c_idx), _d(self.c_w), _d(self.c_k), _d(self.c_mu), _d(self.c_dx0), _d(self.c_T), _d(self.c_n))
        if nc < 0:
            raise RuntimeError("max_n_constraints exceeded")
        self.nc = nc

    # ---- energy / residual / Hessian
    def compute_energy(self):
        L = lib()
        self._bind()
        L.orc_cloth_normals(self.cloth)
        cs = self._contacts()
        E = L.orc_contact_energy(cs, _d(self.pos))
        L.orc_contacts_destroy(cs)
        E += L.orc_cloth_energy(self.cloth, None)
        E += L.orc_vertex_energy(self.NVc, self.NV, _d(self.pos), _d(self.prev_pos), _d(self.vel), _d(self.mass), _d(self.gravity), _f(self.dt))
        return E

    def compute_residual_and_hessian(self, spd=True):
        """BaseScene.compute_residual_and_Hessian (engine/BaseScene.py:976-1040)"""
        L = lib()
        self._bind()
        t0 = time.perf_counter()
        L.orc_cloth_normals(self.cloth)
        L.orc_cloth_prepare_bending(self.cloth)
        Fb = np.zeros((self.NVc, 3))
        L.orc_cloth_residual(self.cloth, _d(Fb), 15)
        self.F[:] = 0
        self.F[:3 * self.NVc] = Fb.reshape(-1)
        # table: F_b = m (x - x_prev - v dt)/dt^2 - m g  -> frozen, zeroed by apply_frozen
        self.F[self.frozen != 0] = 0
        self._tick("residual", t0)
        t0 = time.perf_counter()
        self.val[:] = 0
        cs = self._contacts()
        if self.hessian_mode == "psd":      # CPU twin of the product's forward Newton matrix (not a reference restatement)
            L.orc_contact_grad_hess(cs, _d(self.pos), _i(self.frozen), _d(self.F), None, 0)
            L.orc_contact_hessian_psd_vertex(cs, _d(self.pos), self.mat)
            L.orc_contacts_destroy(cs)
            L.orc_add_mass_diag(self.mat, _d(self.mass), _f(self.dt))
            L.orc_cloth_hessian_psd(self.cloth, self.mat)
            assert L.orc_mat_missing(self.mat) == 0
            self._tick("hessian", t0)
            return
        L.orc_contact_grad_hess(cs, _d(self.pos), _i(self.frozen), _d(self.F), self.mat, int(spd))
        L.orc_contacts_destroy(cs)
        L.orc_add_mass_diag(self.mat, _d(self.mass), _f(self.dt))
        L.orc_cloth_hessian_me(self.cloth, self.mat, int(spd))
        L.orc_cloth_hessian_ma(self.cloth, self.mat)
        L.orc_cloth_hessian_bending(self.cloth, self.mat)
        assert L.orc_mat_missing(self.mat) == 0
        self._tick("hessian", t0)

    def compute_hessian(self, spd):
        """BaseScene.compute_Hessian (engine/BaseScene.py:1042-1052): adds on top of the current values"""
        L = lib()
        self._bind()
        L.orc_add_mass_diag(self.mat, _d(self.mass), _f(self.dt))
        L.orc_cloth_normals(self.cloth)
        L.orc_cloth_prepare_bending(self.cloth)
        L.orc_cloth_hessian_me(self.cloth, self.mat, int(spd))
        L.orc_cloth_hessian_ma(self.cloth, self.mat)
        L.orc_cloth_hessian_bending(self.cloth, self.mat)
        cs = self._contacts()
        L.orc_contact_grad_hess(cs, _d(self.pos), _i(self.frozen), None, self.mat, int(spd))
        L.orc_contacts_destroy(cs)
        assert L.orc_mat_missing(self.mat) == 0

    def solve(self, b):
        t0 = time.perf_counter()
        x = spla.spsolve(self.matrix().tocsc(), b)
        self._tick("solve", t0)
        return x

    # ---- BaseScene.time_step
    def time_step(self, max_newton=1000, tol=1e-7, log=None):
        t0 = time.perf_counter()
        self.prev_pos[:] = self.pos
        self.calc_vn()
        self.projection_query()
        self.contact_analysis()
        self._tick("contact", t0)
        self._build_pattern()
        it = 0
        while it < max_newton:
            it += 1
            t0 = time.perf_counter()
            E0 = self.compute_energy()
            self._tick("energy", t0)
            self.compute_residual_and_hessian(spd=True)
            p = self.solve(self.F)
            p_norm = np.abs(p).max()
            x1 = self.pos.copy()
            alpha = 1.0
            t0 = time.perf_counter()
            while alpha > 1e-8:
                self.pos[:] = x1 - alpha * p.reshape(-1, 3)
                E = self.compute_energy()
                if E < E0:
                    break
                alpha /= 2
            self._tick("linesearch", t0)
            delta = p_norm / self.dt
            if log is not None:
                log.append((E0, delta, alpha, E))
            if delta < tol:
                break
        # Scene_bouncing.timestep_finish
        self.vel[:] = (self.pos - self.prev_pos) * self.damping / self.dt
        self._bind()
        L = lib()
        L.orc_cloth_normals(self.cloth)
        L.orc_cloth_update_ref_angle(self.cloth, _d(self.ref_angle), _f(self.k_angle))
        return it

    def reward(self):
        """Scene_bouncing.compute_reward (task_scene/Scene_bouncing.py:107-113)"""
        row = np.arange(self.NVc) // (self.M + 1)
        return self.pos[:self.NVc][(row == 5) | (row == 10), 2].sum()


class OracleGrad:
    """analytic_grad_system.Grad (engine/analytic_grad_system.py) for OracleScene"""

    def __init__(self, s, T):
        self.s, self.T = s, T
        self.pos_buffer = np.zeros((T, s.NV, 3))
        self.ref_angle_buffer = np.zeros((T, s.NFc, 3))
        self.pos_grad = np.zeros((T, s.NV, 3))
        self.angleref_grad = np.zeros((T, s.NFc, 3))
        self.grad_kb = 0.0
        self.damping = 1.0
        self.clamp = 1.0

    def copy_pos(self, step):
        self.pos_buffer[step] = self.s.pos
        self.ref_angle_buffer[step] = self.s.ref_angle

    def get_loss_table(self):
        """analytic_grad_system.py:176-180 (uses cloth.N + 1 where M + 1 is meant; square sheets only)"""
        s = self.s
        row = (np.arange(s.NVc) / (s.N + 1)).astype(np.int64)
        sel = np.nonzero((row == 5) | (row == 10))[0]
        self.pos_grad[1:, sel, 2] = -1

    def transfer_grad(self, step):
        s, L = self.s, lib()
        np.clip(self.pos_grad[step], -self.clamp, self.clamp, out=self.pos_grad[step])
        # sys.copy_pos_only(pos_buffer, step-1): pos <- x_{t-1}, prev_pos <- x_{t-1}
        s.pos[:] = self.pos_buffer[step - 1]; s.prev_pos[:] = self.pos_buffer[step - 1]
        s.calc_vn(); s.projection_query(); s.contact_analysis()
        # copy_pos_and_refangle(step)
        s.pos[:] = self.pos_buffer[step]; s.prev_pos[:] = self.pos_buffer[step - 1]
        s.ref_angle[:] = self.ref_angle_buffer[step - 1]
        s._bind()
        L.orc_cloth_normals(s.cloth); L.orc_cloth_prepare_bending(s.cloth)
        pg = np.ascontiguousarray(self.pos_grad[step])
        L.orc_cloth_refangle_a2ax(s.cloth, _d(np.ascontiguousarray(self.angleref_grad[step])), _d(self.angleref_grad[step - 1]),
                                  _d(pg), _f(s.k_angle))
        self.pos_grad[step] = pg
        # get_paramters_grad -> d_kb
        d_kb = np.zeros((s.NV, 3))
        dk = np.zeros((s.NVc, 3))
        L.orc_cloth_compute_deri(s.cloth, None, None, _d(dk))
        d_kb[:s.NVc] = dk
        self.d_kb = d_kb
        s._build_pattern()
        L.orc_mat_set_counting(s.mat, 0, None, None)
        s.compute_hessian(False)
        rhs = self.pos_grad[step].reshape(-1).copy()
        self.H = s.matrix()
        p = s.solve(rhs)
        self.z = p
        s.tmp_z_frozen[:] = 0
        zc = np.ascontiguousarray(p)
        L.orc_mat_set_counting(s.mat, 1, _d(zc), _d(s.tmp_z_frozen))
        s.compute_hessian(False)
        L.orc_mat_set_counting(s.mat, 0, None, None)
        x_hat_grad = p.reshape(-1, 3) * s.mass[:, None] / (s.dt ** 2)
        cs = s._contacts()
        L.orc_contact_backprop(cs, _d(s.pos), _d(zc), _d(self.pos_grad[step - 1]))
        L.orc_contacts_destroy(cs)
        L.orc_cloth_refangle_x2a(s.cloth, _d(self.angleref_grad[step - 1]), _d(zc))
        free = (s.frozen == 0)
        self.grad_kb += float((p[free] * d_kb.reshape(-1)[free]).sum())
        fm = free.reshape(-1, 3)
        if step > 0:
            self.pos_grad[step - 1][fm] += (x_hat_grad * (1 + self.damping))[fm]
        if step > 1:
            self.pos_grad[step - 2][fm] -= (x_hat_grad * self.damping)[fm]
