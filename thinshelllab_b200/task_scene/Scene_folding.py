"""Scene_folding on the B200 engine (code/task_scene/Scene_folding.py): one cloth strip (N x M grid, pinned along row N), the
frozen neo-Hookean table and one tactile pad whose bottom / inner-circle vertices follow a gripper pose.  Contacts go both ways
(cloth faces against pad and table vertices, pad and table faces against cloth vertices, contact_analysis :99-108), so constraint
triangles move and the general contact path of libtsl is exercised.

Construction (cold path): engine/scene_builder.py builds what `Scene.init_all(); Scene.reset()` leave in the reference's fields --
the folded strip (Cloth.init_fold), the box table, the tactile pad from the TetGen assets (data/tactile.*), masses, frozen flags,
surface triangles, the gripper frame -- for any cloth_size; tests/test_scene_builder_cpu.py pins it to the arrays the reference itself
produced (tests/golden/folding.npz, forming.npz).  A `state` mapping with the same keys can be passed instead (synthetic scenes)."""
import os

import numpy as np
import torch

from ..core import ShellEngine
from ..engine.BaseScene import SceneCommon
from ..engine.gripper_single import gripper
from ..fields import Scalar, TensorField
from .Scene_bouncing import Body, _ClothView


def tet_rest(rest, tets):
    """Elastic.init_pos (model_elastic_offset.py:240-253): B = Ds^-1, W = |det Ds| / 6 of the rest positions"""
    D = np.stack([rest[tets[:, i]] - rest[tets[:, 3]] for i in range(3)], -1)
    return np.linalg.inv(D), np.abs(np.linalg.det(D)) / 6


class _ElasticView:
    def __init__(self, scene, bid, offset, n_verts, mu, lam):
        self._s, self._bid, self.offset, self.n_verts = scene, bid, offset, n_verts
        self._p = dict(mu=mu, lam=lam)
        self.mu = Scalar(mu, lambda v: self._set("mu", v))
        self.lam = Scalar(lam, lambda v: self._set("lam", v))

    def _set(self, k, v):
        self._p[k] = v
        if self._bid >= 0:
            self._s.engine.set_tet_params(self._bid, self._p["mu"], self._p["lam"])

    @property
    def F_x(self):
        return TensorField(self._s.engine.pos[self.offset:self.offset + self.n_verts])


class Scene(SceneCommon):
    FORMING = False                            # Scene_forming subclasses with True (15 x 7 strip, k_contact 20000)

    def __init__(self, cloth_size=0.06, device="cuda:0", *, state=None, max_newton=50):
        """reference signature: Scene(cloth_size=0.06, device="cuda:0") (code/task_scene/Scene_folding.py:27).  The scene is built by
        engine/scene_builder.folding_state(cloth_size); a ready `state` mapping (keys of tests/golden/folding.npz, see
        oracle/gen_goldens.py:gen_folding) may be passed instead, also as the first positional argument (synthetic scenes, tests)."""
        if state is None and hasattr(cloth_size, "keys"):
            state, cloth_size = cloth_size, None
        if state is None:
            from ..engine.scene_builder import folding_state
            state = folding_state(cloth_size=float(cloth_size), forming=self.FORMING)
        g = state
        self.dt = self.h = float(g["dt"])
        self.cloth_cnt, self.elastic_cnt, self.effector_cnt = 1, 2, 2
        self.k_contact, self.eps_contact, self.eps_v = float(g["k_contact"]), float(g["eps_contact"]), float(g["eps_v"])
        self.max_n_constraints, self.damping = int(g["max_n_constraints"]) if "max_n_constraints" in g else 10000, 1.0
        self.max_newton = max_newton
        N, M, dx = int(g["cloth_N"]), int(g["cloth_M"]), float(g["cloth_dx"])
        self.cloth_N, self.cloth_M = N, M
        self.tot_NV = int(g["pos0"].shape[0])
        gravity = tuple(float(v) for v in g["gravity"])
        e = self.engine = ShellEngine(self.tot_NV, self.dt, k_contact=self.k_contact, eps_contact=self.eps_contact, eps_v=self.eps_v,
                                      damping=self.damping, gravity=gravity, max_n_constraints=self.max_n_constraints,
                                      grid_n=int(g["grid_n"]) if "grid_n" in g else 132, device=device)
        rho = float(g["cloth_mass"]) / (dx * dx)
        cid = e.add_cloth(N, M, 0, dx, rho, Kb=float(g["Kb"]), k_angle=float(g["k_angle"]))
        self.cloths = [_ClothView(self, cid, N, M, dx, 0, rho, Kb=float(g["Kb"]), k_angle=float(g["k_angle"]))]
        self.cloths[0].body_idx = 0
        pos0 = np.asarray(g["pos0"], np.float64)
        to, tn = int(g["table_offset"]), int(g["table_nverts"])
        po, pn = int(g["pad_offset"]), int(g["pad_nverts"])
        b0 = -1
        if "table_tets" in g:                  # a frozen table only adds a constant to the energy: its cells are optional
            tB, tW = tet_rest(pos0[to:to + tn], np.asarray(g["table_tets"]))
            b0 = e.add_tets(0, to, tn, g["table_tets"], tB, tW, float(g["table_mu"]), float(g["table_lam"]), 0.0, g["table_gravity"])
        b1 = e.add_tets(1, po, pn, g["pad_tets"], g["pad_F_B"], g["pad_F_W"], float(g["pad_mu"]), float(g["pad_lam"]), float(g["pad_alpha"]),
                        g["pad_gravity"])
        self.elastics = [_ElasticView(self, b0, to, tn, float(g["table_mu"]) if b0 >= 0 else 0.0, float(g["table_lam"]) if b0 >= 0 else 0.0),
                         _ElasticView(self, b1, po, pn, float(g["pad_mu"]), float(g["pad_lam"]))]
        self.elastics[0].body_idx, self.elastics[1].body_idx = 1, 2
        if "faces" in g:
            self.faces = np.ascontiguousarray(g["faces"], np.int32)
            bv, bf = np.asarray(g["body_v"]), np.asarray(g["body_f"])
        else:                                  # synthetic scenes: cloth faces from the library's own topology + table + pad pieces
            cf_, tf_, pf_ = e.cloth_topology(cid)[0], np.asarray(g["_table_faces"]), np.asarray(g["_pad_faces"])
            self.faces = np.ascontiguousarray(np.concatenate([cf_, tf_, pf_]), np.int32)
            n0, n1 = cf_.shape[0], cf_.shape[0] + tf_.shape[0]
            bv = np.array([[0, to], [to, to + tn], [po, po + pn]]); bf = np.array([[0, n0], [n0, n1], [n1, self.faces.shape[0]]])
        self.tot_NF = self.faces.shape[0]
        self.body_list = [Body(int(bv[i, 0]), int(bv[i, 1]), int(bf[i, 0]), int(bf[i, 1])) for i in range(bv.shape[0])]
        e.set_surfaces(self.faces, [[b.v_start, b.v_end, b.f_start, b.f_end] for b in self.body_list])
        # Scene_folding.contact_analysis (:99-108): for every elastic j: cloth surface vs its vertices, its surface vs cloth vertices
        self.mu_cloth_elastic = Scalar(float(g["mu"]), self._set_mu)
        self._pairs = []
        for el in self.elastics:
            self._pairs.append(e.add_contact_pair(0, el.offset, el.offset + el.n_verts, self.mu_cloth_elastic[None]))
            self._pairs.append(e.add_contact_pair(el.body_idx, 0, self.cloths[0].NV, self.mu_cloth_elastic[None]))
        e.mass.copy_(torch.from_numpy(np.asarray(g["mass"], np.float64)))
        e.frozen.copy_(torch.from_numpy(np.asarray(g["frozen"], np.int32)))
        e.border_flag.copy_(torch.from_numpy(np.asarray(g["border_flag"], np.int32)))
        self._pos0 = pos0
        self._vel0 = np.asarray(g["vel0"], np.float64)
        self._ref0 = np.asarray(g["ref_angle0"], np.float64)
        self._gpos0 = np.asarray(g["gripper_pos0"], np.float64)
        self.gripper = gripper(self, [po], g["gripper_F_x"], g["gripper_bound_idx"], self._gpos0)
        self.gravity = np.array(gravity)
        e.finalize()
        if "init_ref_angle" in g and bool(g["init_ref_angle"]):
            # Cloth.init_fold -> init_ref_angle (model_fold_offset.py:1053-1057, 788-797): plastic flow of the folded strip's rest angles
            e.pos.copy_(torch.from_numpy(self._pos0)); e.prev_pos.copy_(e.pos)
            e.cloth_ref_angle[0].zero_()
            e.update_ref_angle(0)
            self._ref0 = e.cloth_ref_angle[0].cpu().numpy().copy()
        self.reset()

    def _set_mu(self, v):
        for p in getattr(self, "_pairs", []):
            self.engine.set_contact_mu(p, v)

    # ---- reference API
    def init_all(self):
        pass                                   # the state was given

    def reset(self):
        e = self.engine
        e.pos.copy_(torch.from_numpy(self._pos0)); e.prev_pos.copy_(e.pos)
        e.vel.copy_(torch.from_numpy(self._vel0))
        e.cloth_ref_angle[0].copy_(torch.from_numpy(self._ref0))
        self.gripper.init(self, self._gpos0)
        e.reset_contact_state()

    def action(self, step, delta_pos, delta_rot):
        """:213-224: move the gripper, then the driven vertices of the pad"""
        self.gripper.step_simple(delta_pos, delta_rot)
        self.gripper.update_bound(self)

    def time_step(self, f_contact=None, frame_idx=0, force_stick=True, tol=1e-7):
        """:275-321 (at most 50 Newton iterations; timestep_finish includes update_ref_angle)"""
        self.last_stats = self.engine.step_forward(self.max_newton, tol)
        return self.last_stats

    def _crease_hinges(self):
        c = self.cloths[0]
        f2v, cf, cp = self.engine.cloth_topology(0)
        hi, hl = np.nonzero(cf > np.arange(c.NF)[:, None])
        own = f2v[hi, hl] // (c.M + 1)
        opp = f2v[cf[hi, hl], cp[hi, hl]] // (c.M + 1)
        s7 = (own == 6) & (opp == 8)
        s8 = (own == 7) & (opp == 9)
        return (hi[s7], hl[s7]), (hi[s8], hl[s8])

    def compute_reward(self, curve7=1.0, curve8=-1.0):
        """:159-177"""
        ra = self.engine.cloth_ref_angle[0].cpu().numpy()
        s7, s8 = self._crease_hinges()
        return float(-(ra[s7[0], s7[1]] * curve7).sum() - (ra[s8[0], s8[1]] * curve8).sum())

    pos = property(lambda self: TensorField(self.engine.pos))
    vel = property(lambda self: TensorField(self.engine.vel))
    mass = property(lambda self: TensorField(self.engine.mass))
    frozen = property(lambda self: TensorField(self.engine.frozen))
