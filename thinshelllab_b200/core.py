"""ShellEngine: torch-owned device state + the C-ABI calls of libtsl.so.

Mirrors the parts of BaseScene (code/engine/BaseScene.py) that hold global state -- pos / prev_pos / vel /
mass / frozen / faces / body_list / contact pair table -- and forwards the hot path to CUDA.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import TslError


@dataclass
class StepStats:
    newton_iters: int
    linear_iters: int
    linesearch_evals: int
    n_contacts: int
    converged: bool
    flags: int
    delta: float
    energy: float
    ms_contact: float
    ms_assembly: float
    ms_solve: float
    ms_linesearch: float


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


class ShellEngine:
    """One context per GPU.  All tensors live on `device` (a CUDA device); nothing here computes on the CPU."""

    def __init__(self, n_verts, dt, *, k_contact, eps_contact, eps_v=0.01, damping=1.0, gravity=(0.0, 0.0, -9.8),
                 max_n_constraints=10000, grid_h=0.003, grid_n=132, device="cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("thinshelllab_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.L = _lib.lib()
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        self.n_verts = int(n_verts)
        self.dt = float(dt)
        cfg = _lib.Config()
        cfg.struct_size = C.sizeof(_lib.Config)
        cfg.n_verts = self.n_verts
        cfg.dt, cfg.k_contact, cfg.eps_contact, cfg.eps_v, cfg.damping = dt, k_contact, eps_contact, eps_v, damping
        cfg.gravity[0], cfg.gravity[1], cfg.gravity[2] = gravity
        cfg.max_n_constraints, cfg.grid_h, cfg.grid_n = int(max_n_constraints), grid_h, int(grid_n)
        self.cfg = cfg
        self.ctx = C.c_void_p()
        rc = self.L.tsl_create(C.byref(cfg), C.byref(self.ctx))
        if rc != 0:
            raise TslError(rc, "tsl_create failed (no CUDA device?)")
        f64 = dict(dtype=torch.float64, device=self.device)
        self.pos = torch.zeros((self.n_verts, 3), **f64)
        self.prev_pos = torch.zeros((self.n_verts, 3), **f64)
        self.vel = torch.zeros((self.n_verts, 3), **f64)
        self.mass = torch.zeros((self.n_verts,), **f64)
        self.frozen = torch.zeros((3 * self.n_verts,), dtype=torch.int32, device=self.device)
        self.border_flag = torch.zeros((self.n_verts,), dtype=torch.int32, device=self.device)
        self.tet_bodies = []
        self.cloth_ref_angle = []
        self.cloth_shape = []
        self.n_bodies = 0
        self.finalized = False

    # ---- error handling
    def _ck(self, rc):
        if rc < 0:
            raise TslError(rc, self.L.tsl_last_error(self.ctx).decode())
        return rc

    def __del__(self):
        try:
            if getattr(self, "ctx", None) and self.ctx.value:
                self.L.tsl_destroy(self.ctx)
                self.ctx = C.c_void_p()
        except Exception:
            pass

    # ---- scene description
    def add_cloth(self, N, M, v_offset, dx, rho=40.0, Kl=1000.0, Ka=1000.0, Kb=100.0, k_angle=3.14):
        ref = torch.zeros((2 * N * M, 3), dtype=torch.float64, device=self.device)
        cid = self._ck(self.L.tsl_add_cloth(self.ctx, N, M, v_offset, dx, rho, Kl, Ka, Kb, k_angle, _ptr(ref)))
        self.cloth_ref_angle.append(ref)
        self.cloth_shape.append((N, M, v_offset, dx, rho))
        nv = (N + 1) * (M + 1)
        self.mass[v_offset:v_offset + nv] = rho * dx * dx
        return cid

    def set_cloth_params(self, cloth, Kl, Ka, Kb, k_angle):
        self._ck(self.L.tsl_set_cloth_params(self.ctx, cloth, Kl, Ka, Kb, k_angle))

    def update_ref_angle(self, cloth=0):
        """Cloth.update_ref_angle / init_ref_angle at the bound positions (plastic flow beyond k_angle)"""
        self._sync_stream()
        self._ck(self.L.tsl_cloth_update_ref_angle(self.ctx, int(cloth)))

    def set_side_test_override(self, cloth, ov):
        """test hook: ov [NF, 3] int8 with 1 = negative, or None for the canonical rule"""
        if ov is None:
            self._ck(self.L.tsl_set_side_test_override(self.ctx, cloth, C.c_void_p(0)))
        else:
            ov = np.ascontiguousarray(ov, np.int8)
            self._ck(self.L.tsl_set_side_test_override(self.ctx, cloth, _np_ptr(ov)))

    def cloth_topology(self, cloth=0):
        N, M = self.cloth_shape[cloth][:2]
        nf = 2 * N * M
        f2v, cf, cp = (np.zeros((nf, 3), np.int32) for _ in range(3))
        self._ck(self.L.tsl_get_cloth_topology(self.ctx, cloth, _np_ptr(f2v), _np_ptr(cf), _np_ptr(cp)))
        return f2v, cf, cp

    def add_tets(self, kind, v_offset, n_verts, tets, B, W, mu, lam, alpha=0.0, gravity=None):
        """Elastic body: kind 0 = neo-Hookean box (model_elastic_offset.py), 1 = tactile pad / ball (model_elastic_tactile.py);
        tets [nc, 4] body-local ids, B [nc, 3, 3] inverse rest Ds, W [nc] rest volumes.  Masses: fill self.mass."""
        tets = np.ascontiguousarray(tets, np.int32); B = np.ascontiguousarray(B, np.float64); W = np.ascontiguousarray(W, np.float64)
        g = None if gravity is None else np.ascontiguousarray(gravity, np.float64)
        bid = self._ck(self.L.tsl_add_tets(self.ctx, int(kind), int(v_offset), int(n_verts), tets.shape[0], _np_ptr(tets), _np_ptr(B),
                                           _np_ptr(W), float(mu), float(lam), float(alpha), _np_ptr(g) if g is not None else C.c_void_p(0)))
        self.tet_bodies.append((int(kind), int(v_offset), int(n_verts)))
        return bid

    def set_tet_params(self, body, mu, lam):
        self._ck(self.L.tsl_set_tet_params(self.ctx, body, float(mu), float(lam)))

    def set_surfaces(self, faces, bodies):
        faces = np.ascontiguousarray(faces, np.int32)
        bodies = np.ascontiguousarray(bodies, np.int32)
        self.n_bodies = bodies.shape[0]
        self._ck(self.L.tsl_set_surfaces(self.ctx, _np_ptr(faces), faces.shape[0], _np_ptr(bodies), bodies.shape[0]))

    def add_contact_pair(self, surface_body, v_start, v_end, mu):
        pid = self._ck(self.L.tsl_add_contact_pair(self.ctx, surface_body, v_start, v_end, mu))
        self.__dict__.setdefault("_pairs_registered", []).append(pid)
        return pid

    def set_contact_mu(self, pair, mu):
        self._ck(self.L.tsl_set_contact_mu(self.ctx, pair, mu))

    def finalize(self):
        self._ck(self.L.tsl_bind_state(self.ctx, _ptr(self.pos), _ptr(self.prev_pos), _ptr(self.vel), _ptr(self.mass),
                                       _ptr(self.frozen), _ptr(self.border_flag)))
        self._ck(self.L.tsl_finalize(self.ctx))
        self.finalized = True

    def dist_init(self, rank, world, ghost_lo_rows, ghost_hi_rows, first_row_global=0):
        """strip partition (include/tsl.h: tsl_dist_init).  The ncclUniqueId is made by rank 0 and broadcast through the
        torch.distributed process group the caller has already initialised (backend nccl)."""
        import torch.distributed as dist
        buf = (C.c_ubyte * 128)()
        if world > 1:
            t = torch.zeros(128, dtype=torch.uint8, device=self.device)
            if rank == 0:
                self._ck(self.L.tsl_dist_unique_id(buf))
                t.copy_(torch.frombuffer(bytearray(buf), dtype=torch.uint8))
            dist.broadcast(t, 0)
            raw = bytes(t.cpu().numpy().tobytes())
            buf = (C.c_ubyte * 128).from_buffer_copy(raw)
        self._ck(self.L.tsl_dist_init(self.ctx, buf, int(rank), int(world), int(ghost_lo_rows), int(ghost_hi_rows), int(first_row_global)))
        self.dist = (int(rank), int(world), int(ghost_lo_rows), int(ghost_hi_rows))

    def dist_stats(self):
        a, b = C.c_longlong(), C.c_longlong()
        self._ck(self.L.tsl_dist_stats(self.ctx, C.byref(a), C.byref(b)))
        return {"halo_exchanges": a.value, "allreduces": b.value}

    def reset_contact_state(self):
        self._ck(self.L.tsl_reset_contact_state(self.ctx))

    # ---- hot path
    def _sync_stream(self):
        self._ck(self.L.tsl_set_stream(self.ctx, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def contact_detect(self):
        self._sync_stream()
        n = C.c_int()
        self._ck(self.L.tsl_contact_detect(self.ctx, C.byref(n)))
        return n.value

    def energy(self):
        self._sync_stream()
        e = C.c_double()
        self._ck(self.L.tsl_energy(self.ctx, C.byref(e)))
        return e.value

    def assemble(self, flags):
        self._sync_stream()
        self._ck(self.L.tsl_assemble(self.ctx, flags))

    def solve(self, rhs, rel_tol=1e-6, max_iters=10000):
        self._sync_stream()
        x = torch.zeros_like(rhs)
        st = _lib.SolveStatsC()
        self._ck(self.L.tsl_solve(self.ctx, _ptr(rhs), _ptr(x), rel_tol, max_iters, C.byref(st)))
        return x, (st.iters, st.flags, st.rel_residual)

    @staticmethod
    def _stats(st):
        return StepStats(st.newton_iters, st.linear_iters, st.linesearch_evals, st.n_contacts, bool(st.converged), st.flags,
                         st.delta, st.energy, st.ms_contact, st.ms_assembly, st.ms_solve, st.ms_linesearch)

    def step_forward(self, max_newton=1000, tol=1e-7):
        self._sync_stream()
        st = _lib.StepStatsC()
        self._ck(self.L.tsl_step_forward(self.ctx, max_newton, tol, C.byref(st)))
        return self._stats(st)

    def step_forward_host(self, pos_host, vel_host, max_newton=1000, tol=1e-7):
        """pos_host / vel_host: pinned CPU float64 tensors [n_verts, 3], updated in place"""
        self._sync_stream()
        st = _lib.StepStatsC()
        self._ck(self.L.tsl_step_forward_host(self.ctx, _ptr(pos_host), _ptr(vel_host), max_newton, tol, C.byref(st)))
        return self._stats(st)

    def step_backward(self, x_t, x_tm1, ref_angle_tm1, pg_t, pg_tm1, pg_tm2, ag_t, ag_tm1, grad_kb, z_out=None, clamp=1.0,
                      rel_tol=1e-10, max_iters=20000):
        self._sync_stream()
        st = _lib.SolveStatsC()
        self._ck(self.L.tsl_step_backward(self.ctx, _ptr(x_t), _ptr(x_tm1), _ptr(ref_angle_tm1), _ptr(pg_t), _ptr(pg_tm1),
                                          _ptr(pg_tm2), _ptr(ag_t), _ptr(ag_tm1), _ptr(grad_kb), _ptr(z_out), clamp, rel_tol,
                                          max_iters, C.byref(st)))
        return st.iters, st.flags, st.rel_residual

    def step_backward_ex(self, x_t, x_tm1, ref_angle_tm1, pg_t, pg_tm1, pg_tm2, ag_t, ag_tm1, grad_kb=None, z_out=None, z_frozen=None,
                         clamp=1000.0, clamp_angleref=1000.0, rel_tol=1e-10, max_iters=20000):
        """Grad.transfer_grad of the trajectory optimiser (analytic_grad_single.py:217-255); z_frozen receives tmp_z_frozen"""
        self._sync_stream()
        st = _lib.SolveStatsC()
        self._ck(self.L.tsl_step_backward_ex(self.ctx, _ptr(x_t), _ptr(x_tm1), _ptr(ref_angle_tm1), _ptr(pg_t), _ptr(pg_tm1),
                                             _ptr(pg_tm2), _ptr(ag_t), _ptr(ag_tm1), _ptr(grad_kb), _ptr(z_out), _ptr(z_frozen), clamp,
                                             clamp_angleref, rel_tol, max_iters, C.byref(st)))
        return st.iters, st.flags, st.rel_residual

    def elastic_param_grad(self, z=None):
        """(d_mu, d_lam [n_verts, 3] CUDA tensors, (sum z d_mu, sum z d_lam) over the free DOFs or None) at the bound positions"""
        self._sync_stream()
        d_mu = torch.empty((self.n_verts, 3), dtype=torch.float64, device=self.device); d_lam = torch.empty_like(d_mu)
        out = (C.c_double * 2)()
        self._ck(self.L.tsl_elastic_param_grad(self.ctx, _ptr(z), _ptr(d_mu), _ptr(d_lam), out))
        return d_mu, d_lam, ((out[0], out[1]) if z is not None else None)

    def cloth_param_deri(self, cloth=0, kl=True, ka=True, kb=True):
        """(d_kl, d_ka, d_kb) [n_verts, 3] CUDA tensors (None where not requested): Cloth.compute_deri at the bound positions"""
        self._sync_stream()
        out = [torch.empty((self.n_verts, 3), dtype=torch.float64, device=self.device) if want else None for want in (kl, ka, kb)]
        self._ck(self.L.tsl_cloth_param_deri(self.ctx, int(cloth), _ptr(out[0]), _ptr(out[1]), _ptr(out[2])))
        return tuple(out)

    def friction_coef_grad(self, z, pair_begin=0, pair_end=None):
        """one adjoint step's contribution to Grad.grad_friction_coef (Scene.contact_energy_backprop_friction)"""
        self._sync_stream()
        out = C.c_double()
        n_pairs = len(getattr(self, "_pairs_registered", [])) if pair_end is None else pair_end
        self._ck(self.L.tsl_friction_coef_grad(self.ctx, _ptr(z), int(pair_begin), int(n_pairs), C.byref(out)))
        return out.value

    def elastic_force(self, body):
        """Elastic.get_force of tetrahedral body `body`: F_f [body verts, 3] CUDA tensor"""
        self._sync_stream()
        nv = self.tet_bodies[body][2]
        F = torch.empty((nv, 3), dtype=torch.float64, device=self.device)
        self._ck(self.L.tsl_elastic_force(self.ctx, int(body), _ptr(F)))
        return F

    def gripper_apply(self, v_offset, bound_idx, F_x, pos3, rotmat32):
        """pos[v_offset + bound_idx] = pos3 + R F_x[bound_idx] (gripper.get_vert_pos + update_bound, gripper_single.py:79-83, 157-161)"""
        self._sync_stream()
        p = np.ascontiguousarray(pos3, np.float64); R = np.ascontiguousarray(rotmat32, np.float32)
        self._ck(self.L.tsl_gripper_apply(self.ctx, int(v_offset), int(bound_idx.numel()), _ptr(bound_idx), _ptr(F_x),
                                          p.ctypes.data_as(C.POINTER(C.c_double)), R.ctypes.data_as(C.POINTER(C.c_float))))

    def gripper_gather(self, z_frozen, v_offset, bound_idx, F_x, rotmat32, clamp_pos=10.0, clamp_angle=100.0):
        """(d_pos, d_angle) of gripper.gather_grad (gripper_single.py:133-150)"""
        self._sync_stream()
        R = np.ascontiguousarray(rotmat32, np.float32); out = np.zeros(6)
        self._ck(self.L.tsl_gripper_gather(self.ctx, _ptr(z_frozen), int(v_offset), int(bound_idx.numel()), _ptr(bound_idx), _ptr(F_x),
                                           R.ctypes.data_as(C.POINTER(C.c_float)), clamp_pos, clamp_angle,
                                           out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    # ---- introspection
    def residual(self):
        F = np.zeros(3 * self.n_verts)
        self._ck(self.L.tsl_get_residual(self.ctx, _np_ptr(F)))
        return F

    def matrix(self):
        """last assembled Hessian as scipy BSR -> CSR"""
        import scipy.sparse as sp
        n = C.c_int()
        self._ck(self.L.tsl_get_matrix_nnzb(self.ctx, C.byref(n)))
        rowptr = np.zeros(self.n_verts + 1, np.int32); colidx = np.zeros(n.value, np.int32); val = np.zeros((n.value, 3, 3))
        self._ck(self.L.tsl_get_matrix(self.ctx, _np_ptr(rowptr), _np_ptr(colidx), _np_ptr(val)))
        H = sp.bsr_matrix((val, colidx, rowptr), shape=(3 * self.n_verts, 3 * self.n_verts)).tocsr()
        rows, cols, blocks = self.contact_blocks()
        if rows.size:
            # blocks outside the static pattern (contact against moving triangles)
            r = (3 * rows[:, None, None] + np.arange(3)[None, :, None]) + np.zeros((1, 1, 3), np.int64)
            c = (3 * cols[:, None, None] + np.arange(3)[None, None, :]) + np.zeros((1, 3, 1), np.int64)
            H = H + sp.coo_matrix((blocks.ravel(), (r.ravel(), c.ravel())), shape=H.shape).tocsr()
        return H

    def contact_blocks(self):
        n = C.c_int()
        self._ck(self.L.tsl_get_contact_blocks(self.ctx, C.byref(n), C.c_void_p(0), C.c_void_p(0), C.c_void_p(0)))
        rows = np.zeros(n.value, np.int32); cols = np.zeros(n.value, np.int32); val = np.zeros((n.value, 3, 3))
        if n.value:
            self._ck(self.L.tsl_get_contact_blocks(self.ctx, C.byref(n), _np_ptr(rows), _np_ptr(cols), _np_ptr(val)))
        return rows, cols, val

    def projection(self, body):
        nv = self.n_verts
        flag = np.zeros(nv, np.int32); d = np.zeros(nv, np.int32); idx = np.zeros((nv, 3), np.int32); w = np.zeros((nv, 3))
        self._ck(self.L.tsl_get_projection(self.ctx, body, _np_ptr(flag), _np_ptr(d), _np_ptr(idx), _np_ptr(w)))
        return flag, d, idx, w

    def constraints(self):
        m = self.cfg.max_n_constraints
        n = C.c_int()
        idx = np.zeros((m, 4), np.int32); w = np.zeros((m, 3)); k = np.zeros(m); dx0 = np.zeros((m, 3)); T = np.zeros((m, 2, 3)); nn = np.zeros((m, 3))
        self._ck(self.L.tsl_get_constraints(self.ctx, C.byref(n), _np_ptr(idx), _np_ptr(w), _np_ptr(k), _np_ptr(dx0), _np_ptr(T), _np_ptr(nn)))
        c = n.value
        return dict(nc=c, idx=idx[:c], w=w[:c], k=k[:c], dx0=dx0[:c], T=T[:c], n=nn[:c])

    def sizes(self):
        s = _lib.SizesC()
        self._ck(self.L.tsl_get_sizes(self.ctx, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def bench_kernel(self, what, iters):
        self._sync_stream()
        ms = C.c_float()
        self._ck(self.L.tsl_bench_kernel(self.ctx, what, iters, C.byref(ms)))
        return ms.value

    def set_option(self, key, value):
        self._ck(self.L.tsl_set_option(self.ctx, int(key), float(value)))

    def mg_level(self, level, values=True):
        """multigrid level read-back: (n0, n1, n_levels, lmax, val[25, 3, 3, n0*n1] or None)"""
        dims = (C.c_int * 3)()
        lmax = C.c_float()
        self._ck(self.L.tsl_mg_get_level(self.ctx, level, dims, C.byref(lmax), C.c_void_p(0)))
        val = None
        if values:
            val = np.zeros((25, 3, 3, dims[0] * dims[1]), np.float32)
            self._ck(self.L.tsl_mg_get_level(self.ctx, level, dims, C.byref(lmax), _np_ptr(val)))
        return dims[0], dims[1], dims[2], lmax.value, val

    def precond_apply(self, b):
        """z = M^-1 b with the current preconditioner (b: [3 n_verts] float64 CUDA tensor)"""
        self._sync_stream()
        z = torch.zeros_like(b)
        self._ck(self.L.tsl_precond_apply(self.ctx, _ptr(b), _ptr(z)))
        return z

    def dense_solve(self, A, b):
        """test hook of the dense LU behind the direct adjoint solve: x = A^-1 b (host arrays, factorised on the GPU)"""
        A = np.asfortranarray(A, np.float64); b = np.ascontiguousarray(b, np.float64)
        x = np.zeros_like(b)
        self._ck(self.L.tsl_dense_solve_host(self.ctx, int(b.shape[0]), _np_ptr(A), _np_ptr(b), _np_ptr(x)))
        return x

    def launch_count(self):
        return int(self.L.tsl_launch_count(self.ctx))
