"""One cloth + a list of elastic bodies (frozen or free boxes, tactile pads on gripper parts): the layout shared by the reference's
Scene_lifting / Scene_pick / Scene_balancing / Scene_interact (code/task_scene/Scene_*.py).  The scene arrays come from
engine/scene_builder.multi_body_state; contacts go both ways between the cloth and every elastic body (contact_analysis of those
scenes), so constraint triangles move and the general contact path of libtsl is used."""
import numpy as np
import torch

from ..core import ShellEngine
from ..engine.BaseScene import SceneCommon
from ..engine.gripper_single import gripper
from ..engine.gripper_tactile import gripper as gripper_tactile
from ..fields import Scalar, TensorField
from .Scene_bouncing import Body, _ClothView
from .Scene_folding import _ElasticView


class MultiBodyScene(SceneCommon):
    max_newton = 50

    def _build(self, g, device="cuda:0"):
        self.dt = self.h = float(g["dt"])
        els = g["elastics"]
        pads = [r for r in els if r["kind"] == 1]
        n_cloths = int(g["n_cloths"]) if "n_cloths" in g else 1
        self.cloth_cnt, self.elastic_cnt = n_cloths, len(els)
        self.effector_cnt = self.elastic_cnt                       # BaseScene.__init__: effector_cnt defaults to elastic_cnt
        self.k_contact, self.eps_contact, self.eps_v = float(g["k_contact"]), float(g["eps_contact"]), float(g["eps_v"])
        self.max_n_constraints, self.damping = int(g["max_n_constraints"]), float(g["damping"]) if "damping" in g else 1.0
        N, M, dx = int(g["cloth_N"]), int(g["cloth_M"]), float(g["cloth_dx"])
        self.cloth_N, self.cloth_M = N, M
        pos0 = np.asarray(g["pos0"], np.float64)
        self.tot_NV = pos0.shape[0]
        gravity = tuple(float(v) for v in g["gravity"])            # the cloths' gravity; bodies carry their own
        e = self.engine = ShellEngine(self.tot_NV, self.dt, k_contact=self.k_contact, eps_contact=self.eps_contact, eps_v=self.eps_v,
                                      damping=self.damping, gravity=gravity, max_n_constraints=self.max_n_constraints,
                                      grid_n=int(g["grid_n"]) if "grid_n" in g else 132, device=device)
        rho = float(g["cloth_mass"]) / (dx * dx)
        NVc = (N + 1) * (M + 1)
        self.cloths = []
        for k in range(n_cloths):
            cid = e.add_cloth(N, M, k * NVc, dx, rho, Kb=float(g["Kb"]), k_angle=float(g["k_angle"]))
            self.cloths.append(_ClothView(self, cid, N, M, dx, k * NVc, rho, Kb=float(g["Kb"]), k_angle=float(g["k_angle"])))
            self.cloths[k].body_idx = k
        self.elastics = []
        for j, r in enumerate(els):
            bid = e.add_tets(int(r["kind"]), int(r["offset"]), int(r["nverts"]), r["tets"], r["F_B"], r["F_W"], float(r["mu"]), float(r["lam"]),
                             float(r["alpha"]), r["gravity"])
            v = _ElasticView(self, bid, int(r["offset"]), int(r["nverts"]), float(r["mu"]), float(r["lam"]))
            v.body_idx = n_cloths + j
            v.rest = r.get("rest")
            self.elastics.append(v)
        f2v = e.cloth_topology(0)[0]
        self.faces = np.ascontiguousarray(np.concatenate([f2v + k * NVc for k in range(n_cloths)] + list(g["elastic_faces"])), np.int32)
        self.tot_NF = self.faces.shape[0]
        NFc = f2v.shape[0]
        self.body_list = [Body(k * NVc, (k + 1) * NVc, k * NFc, (k + 1) * NFc) for k in range(n_cloths)]
        f0 = n_cloths * NFc
        for r, fa in zip(els, g["elastic_faces"]):
            self.body_list.append(Body(int(r["offset"]), int(r["offset"]) + int(r["nverts"]), f0, f0 + fa.shape[0]))
            f0 += fa.shape[0]
        e.set_surfaces(self.faces, [[b.v_start, b.v_end, b.f_start, b.f_end] for b in self.body_list])
        # contact_analysis as a list of (surface body, vertex body, mu): bodies are numbered cloths first, then elastics.  Default: every
        # elastic j against cloth 0 both ways with mu_cloth_elastic (or a fixed per-body value); scenes with several cloths or
        # elastic-elastic contact give the list (or its additions) explicitly, in the reference's order.
        self.mu_cloth_elastic = Scalar(float(g["mu"]), lambda v: self._set_mu("elastic", v))
        self.mu_cloth_cloth = Scalar(float(g["mu"]), lambda v: self._set_mu("cloth", v))
        mu_fixed = g.get("mu_per_elastic") if hasattr(g, "get") else None
        if "pairs" in g and g["pairs"] is not None:
            spec = list(g["pairs"])
        else:
            spec = []
            for j, el in enumerate(self.elastics):
                fixed = None if mu_fixed is None else mu_fixed[j]
                mu = "elastic" if fixed is None else fixed
                spec += [(0, el.body_idx, mu), (el.body_idx, 0, mu)]
            spec += list(g["extra_pairs"]) if "extra_pairs" in g else []
        self._pairs = []
        for (a, b, mu) in spec:
            vb = self.body_list[int(b)]
            follow, factor = (mu, 1.0) if isinstance(mu, str) else (tuple(mu) if isinstance(mu, (tuple, list)) else (None, 1.0))
            value = float(mu) if follow is None else (self.mu_cloth_elastic[None] if follow == "elastic" else self.mu_cloth_cloth[None]) * factor
            self._pairs.append((e.add_contact_pair(int(a), vb.v_start, vb.v_end, value), follow, factor))
        e.mass.copy_(torch.from_numpy(np.asarray(g["mass"], np.float64)))
        e.frozen.copy_(torch.from_numpy(np.asarray(g["frozen"], np.int32)))
        self._pos0, self._vel0 = pos0, np.asarray(g["vel0"], np.float64)
        self._ref0 = np.asarray(g["ref_angle0"], np.float64).reshape(n_cloths, -1, 3)
        self._gpos0 = np.asarray(g["gripper_pos0"], np.float64)
        self._grot0 = np.asarray(g["gripper_rot0"], np.float64) if "gripper_rot0" in g and np.abs(np.asarray(g["gripper_rot0"])[:, 1:]).max() > 0 else None
        pad_part = [int(v) for v in g["pad_part"]] if "pad_part" in g else list(range(len(pads)))
        offs = [int(r["offset"]) for r in pads]
        if len(set(pad_part)) == len(pad_part):
            self.enable_gripper = False
            self.gripper = gripper(self, offs, g["gripper_F_x"], g["gripper_bound_idx"], self._gpos0)
        else:
            # BaseScene(enable_gripper=True): the two-finger gripper, pads 2 j (upper) and 2 j + 1 (lower) on part j
            assert pad_part == [k // 2 for k in range(len(pads))], "two-finger gripper: pads must come as (upper, lower) per part"
            self.enable_gripper = True
            Fx = np.asarray(g["gripper_F_x"])
            self.gripper = gripper_tactile(self, offs[0::2], offs[1::2], Fx[0::2], Fx[1::2], g["gripper_bound_idx"], self._gpos0)
        self.gravity = np.array([0.0, 0.0, -9.8])
        e.finalize()
        if bool(g["init_ref_angle"]):
            e.pos.copy_(torch.from_numpy(self._pos0)); e.prev_pos.copy_(e.pos)
            for k in range(n_cloths):
                e.cloth_ref_angle[k].zero_()
                e.update_ref_angle(k)
            self._ref0 = np.stack([e.cloth_ref_angle[k].cpu().numpy() for k in range(n_cloths)])
        self.reset()

    def _set_mu(self, which, v):
        for p, follow, factor in getattr(self, "_pairs", []):
            if follow == which:
                self.engine.set_contact_mu(p, v * factor)

    def _effectors(self):
        pads = [el for el in self.elastics if el._bid >= 0 and self.engine.tet_bodies[el._bid][0] == 1]
        return [(el._bid, self.gripper._bound_idx) for el in pads]

    # ---- reference API
    def init_all(self):
        pass

    def reset(self):
        e = self.engine
        e.pos.copy_(torch.from_numpy(self._pos0)); e.prev_pos.copy_(e.pos)
        e.vel.copy_(torch.from_numpy(self._vel0))
        for k in range(self.cloth_cnt):
            e.cloth_ref_angle[k].copy_(torch.from_numpy(self._ref0[k]))
        self.gripper.init(self, self._gpos0)
        if self._grot0 is not None:                                  # Scene_card.init: pads turned before the first step
            self.gripper._rot[:] = self._grot0
            self.gripper.get_rotmat()
        e.reset_contact_state()

    def action(self, step, delta_pos, delta_rot):
        self.gripper.step_simple(delta_pos, delta_rot)
        self.gripper.update_bound(self)

    def time_step(self, f_contact=None, frame_idx=0, force_stick=True, tol=1e-7):
        self.last_stats = self.engine.step_forward(self.max_newton, tol)
        return self.last_stats

    def _hinges_between_rows(self, r1, r2):
        """(face, slot) of the hinges (counter_face[i][l] > i) whose own vertex f2v[i][l] lies in grid row r1 and whose opposite vertex in r2"""
        c = self.cloths[0]
        f2v, cf, cp = self.engine.cloth_topology(0)
        hi, hl = np.nonzero(cf > np.arange(c.NF)[:, None])
        own = f2v[hi, hl] // (c.M + 1)
        opp = f2v[cf[hi, hl], cp[hi, hl]] // (c.M + 1)
        sel = (own == r1) & (opp == r2)
        return hi[sel], hl[sel]

    def _crease_hinges(self):
        return self._hinges_between_rows(6, 8), self._hinges_between_rows(7, 9)

    def _hinge_angles(self, hi, hl):
        """Cloth.compute_angle (code/engine/model_fold_offset.py:126-138) for a handful of hinges, on the host (rewards only)"""
        c = self.cloths[0]
        f2v, cf, cp = self.engine.cloth_topology(0)
        x = self.engine.pos[c.offset:c.offset + c.NV].cpu().numpy()

        def normal(f):
            a, b, cc = x[f2v[f, 0]], x[f2v[f, 1]], x[f2v[f, 2]]
            n = np.cross(b - a, cc - b)
            return n / np.linalg.norm(n, axis=-1, keepdims=True)
        n1, n2 = normal(hi), normal(cf[hi, hl])
        ct = np.einsum("ij,ij->i", n1, n2)
        th = np.where(ct < 0.999999, np.arccos(np.clip(ct, -1, 1)), 2 * np.sqrt(np.abs(1 - ct)) / np.sqrt(1 + ct))
        e = x[f2v[hi, (hl + 1) % 2]] - x[f2v[hi, hl]]                  # `% 2` is the reference's (quirk Q3)
        return np.where(np.einsum("ij,ij->i", n2, e) < 0, -th, th)

    pos = property(lambda self: TensorField(self.engine.pos))
    vel = property(lambda self: TensorField(self.engine.vel))
    mass = property(lambda self: TensorField(self.engine.mass))
    frozen = property(lambda self: TensorField(self.engine.frozen))
