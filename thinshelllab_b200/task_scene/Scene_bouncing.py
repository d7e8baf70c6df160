"""Scene_bouncing on the B200 engine: same constructor, attributes and methods the reference's driver uses
(code/task_scene/Scene_bouncing.py, code/training/trajopt_bouncing.py:43-48,57-58,66-79,121), state in torch CUDA
tensors, hot path in libtsl.  One cloth (N x M grid) over a frozen box ("table"); cloth vertices are projected on
the table surface (contact_analysis :91-96)."""
import numpy as np
import torch

from ..core import ShellEngine
from ..engine.BaseScene import SceneCommon
from ..fields import Scalar, TensorField
from ..meshes import box_body


class Body:
    """BaseScene.Body (code/engine/BaseScene.py:23-27): vertex / face range of one body in the scene-global arrays"""

    def __init__(self, v_start=0, v_end=0, f_start=0, f_end=0):
        self.v_start, self.v_end, self.f_start, self.f_end = v_start, v_end, f_start, f_end


class _ClothView:
    """the attributes of engine.model_fold_offset.Cloth that drivers touch"""

    def __init__(self, scene, cid, N, M, dx, offset, rho, Kl=1000.0, Ka=1000.0, Kb=100.0, k_angle=3.14):
        self._s, self._cid = scene, cid
        self.N, self.M, self.dx, self.offset = N, M, dx, offset
        self.NV, self.NF = (N + 1) * (M + 1), 2 * N * M
        self.mass = rho * dx * dx
        # the Scalars read back exactly what the engine simulates with
        self._p = dict(Kl=Kl, Ka=Ka, Kb=Kb, k_angle=k_angle)
        self.Kl = Scalar(Kl, lambda v: self._set("Kl", v))
        self.Ka = Scalar(Ka, lambda v: self._set("Ka", v))
        self.Kb = Scalar(Kb, lambda v: self._set("Kb", v))
        self.k_angle = Scalar(k_angle, lambda v: self._set("k_angle", v))

    def _set(self, k, v):
        self._p[k] = v
        if self._s.engine.finalized or self._s.engine.cloth_shape:
            self._s.engine.set_cloth_params(self._cid, self._p["Kl"], self._p["Ka"], self._p["Kb"], self._p["k_angle"])

    @property
    def pos(self):
        return TensorField(self._s.engine.pos[self.offset:self.offset + self.NV])

    @property
    def vel(self):
        return TensorField(self._s.engine.vel[self.offset:self.offset + self.NV])

    @property
    def ref_angle(self):
        return TensorField(self._s.engine.cloth_ref_angle[self._cid])

    @property
    def f2v(self):
        return TensorField(torch.from_numpy(self._s.engine.cloth_topology(self._cid)[0]))


class Scene(SceneCommon):
    """reference: Scene(cloth_size=0.06); extra keyword arguments expose the constants the reference hard-codes in
    init_scene_parameters so the same class serves the synthetic sheet sizes of BASELINE.json."""

    def __init__(self, cloth_size=0.06, *, cloth_N=15, cloth_M=None, dt=2e-3, table_size=0.07, table_N=(9, 9, 2),
                 table_offset=(-0.035, -0.035, -0.00875), cloth_offset=(-0.03, -0.03, 0.00039), reset_offset=(-0.03, -0.03, 0.0039),
                 k_contact=40000.0, eps_contact=0.0004, eps_v=0.01, max_n_constraints=10000, rho=40.0, pinned_vertices=(),
                 grid_h=0.003, grid_n=132, device="cuda:0"):
        cloth_M = cloth_N if cloth_M is None else cloth_M
        self.dt = self.h = dt
        self.cloth_cnt, self.elastic_cnt, self.effector_cnt = 1, 1, 1
        self.cloth_N, self.cloth_M, self.cloth_size = cloth_N, cloth_M, cloth_size
        self.k_contact, self.eps_contact, self.eps_v = k_contact, eps_contact, eps_v
        self.max_n_constraints, self.damping = max_n_constraints, 1.0
        self._cloth_offset, self._reset_offset = cloth_offset, reset_offset
        self._pinned = tuple(pinned_vertices)
        dx = cloth_size / cloth_N
        NVc = (cloth_N + 1) * (cloth_M + 1)
        tpos, ttets, tfaces, tmass = box_body(table_size, *table_N, table_offset)
        self._table = (tpos, tfaces, tmass)
        self.tot_NV = NVc + tpos.shape[0]
        self.engine = ShellEngine(self.tot_NV, dt, k_contact=k_contact, eps_contact=eps_contact, eps_v=eps_v, damping=self.damping,
                                  max_n_constraints=max_n_constraints, grid_h=grid_h, grid_n=grid_n, device=device)
        cid = self.engine.add_cloth(cloth_N, cloth_M, 0, dx, rho)
        self.cloths = [_ClothView(self, cid, cloth_N, cloth_M, dx, 0, rho)]
        self.mu_cloth_elastic = Scalar(1.0, self._set_mu)
        self.gravity = np.array([0.0, 0.0, -9.8])
        f2v = self.engine.cloth_topology(cid)[0]
        NFc = f2v.shape[0]
        faces = np.concatenate([f2v, tfaces + NVc]).astype(np.int32)
        self.tot_NF = faces.shape[0]
        self.faces = faces
        self.engine.set_surfaces(faces, [[0, NVc, 0, NFc], [NVc, self.tot_NV, NFc, self.tot_NF]])
        # Scene_bouncing.contact_analysis: cloth vertices against the table surface (body 1)
        self._pair = self.engine.add_contact_pair(1, 0, NVc, self.mu_cloth_elastic[None])
        self.engine.mass[NVc:] = torch.from_numpy(tmass).to(self.engine.device)
        self.elastic_offset = NVc
        self._initialised = False

    # ---- reference API
    def _set_mu(self, v):
        self.engine.set_contact_mu(self._pair, v)

    def _place(self, off):
        e, c = self.engine, self.cloths[0]
        i, j = np.meshgrid(np.arange(c.N + 1), np.arange(c.M + 1), indexing="ij")
        p = np.stack([i * c.dx + off[0], j * c.dx + off[1], np.full(i.shape, off[2], np.float64)], -1).reshape(-1, 3)   # Cloth.init_pos_offset
        pos = np.concatenate([p, self._table[0]])
        e.pos.copy_(torch.from_numpy(pos))
        e.prev_pos.copy_(e.pos)
        e.vel.zero_()
        e.cloth_ref_angle[0].zero_()
        self._init_ref_angle_bridge()

    def _init_ref_angle_bridge(self):
        """Cloth.init_ref_angle_bridge (model_fold_offset.py:811-822): two pre-creased rows, only if the sheet has them"""
        c = self.cloths[0]
        f2v, cf, cp = self.engine.cloth_topology(0)
        hi, hl = np.nonzero(cf > np.arange(c.NF)[:, None])
        own = f2v[hi, hl] // (c.M + 1)
        opp = f2v[cf[hi, hl], cp[hi, hl]] // (c.M + 1)
        sel = ((own == 4) & (opp == 6)) | ((own == 9) & (opp == 11))
        if sel.any():
            ra = self.engine.cloth_ref_angle[0].cpu().numpy()
            ra[hi[sel], hl[sel]] = 1.7
            self.engine.cloth_ref_angle[0].copy_(torch.from_numpy(ra))

    def init_all(self):
        e = self.engine
        if not e.finalized:
            e.frozen[3 * self.elastic_offset:] = 1                       # set_frozen_kernel: the whole table
            for v in self._pinned:
                e.frozen[3 * v:3 * v + 3] = 1
            e.finalize()
        self._place(self._cloth_offset)
        self._initialised = True

    def reset(self):
        self._place(self._reset_offset)
        self.engine.reset_contact_state()

    def time_step(self, f_contact=None, frame_idx=0, max_newton=1000, tol=1e-7):
        """BaseScene.time_step; f_contact is accepted for signature compatibility (see engine.geometry)"""
        self.last_stats = self.engine.step_forward(max_newton, tol)
        return self.last_stats

    def compute_reward(self):
        c = self.cloths[0]
        row = torch.arange(c.NV, device=self.engine.device) // (c.M + 1)
        z = self.engine.pos[:c.NV, 2]
        return float(z[(row == 5) | (row == 10)].sum().item())

    @property
    def pos(self):
        return TensorField(self.engine.pos)

    @property
    def vel(self):
        return TensorField(self.engine.vel)

    @property
    def mass(self):
        return TensorField(self.engine.mass)

    @property
    def frozen(self):
        return TensorField(self.engine.frozen)
