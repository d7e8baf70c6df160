"""The parts of the reference's BaseScene (code/engine/BaseScene.py) that sit on top of the forward step and are shared by every task
scene of this package: state files, the early-stop test of the CMA-ES / RL drivers, force gathering on the effector pads, the
observation vector of the Gym wrappers, and the per-vertex parameter sensitivities.  Host-side orchestration over torch tensors; the
numerics (elastic forces, dF/dK) come from libtsl (tsl_elastic_force, tsl_cloth_param_deri, tsl_elastic_param_grad)."""
import numpy as np
import torch

from ..fields import TensorField


class SceneCommon:
    # ---- BaseScene.save_state / load_state (:1376-1392): {'pos', 'vel'} as CPU tensors, torch.save format
    def save_state(self, save_path):
        e = self.engine
        torch.save({"pos": e.pos.detach().cpu(), "vel": e.vel.detach().cpu()}, save_path)

    def load_state(self, save_path):
        e = self.engine
        data = torch.load(save_path)
        e.pos.copy_(torch.as_tensor(data["pos"], dtype=torch.float64).reshape(e.pos.shape))
        e.vel.copy_(torch.as_tensor(data["vel"], dtype=torch.float64).reshape(e.vel.shape))
        e.prev_pos.copy_(e.pos)
        for cid in range(len(self.cloths)):          # the reference calls update_ref_angle() after loading (:1385)
            e.update_ref_angle(cid)

    # ---- BaseScene.check_pos_nan / gather_force / check_early_stop (:1542-1585)
    def check_pos_nan(self):
        return bool(torch.isnan(self.engine.pos).any().item())

    def _effectors(self):
        """(tet body id, driven-vertex ids) of the effector pads elastics[1 .. effector_cnt)"""
        out = []
        for j in range(1, getattr(self, "effector_cnt", 1)):
            el = self.elastics[j]
            out.append((el._bid, self.gripper._bound_idx if hasattr(self, "gripper") else None))
        return out

    def gather_force(self):
        """tot_force[j - 1] = sum of Elastic.get_force over the bottom / inner-circle vertices of pad j"""
        eff = self._effectors()
        self.tot_force = np.zeros((max(len(eff), 1), 3))
        for k, (bid, bound) in enumerate(eff):
            Ff = self.engine.elastic_force(bid)
            self.tot_force[k] = Ff[bound.long()].sum(0).cpu().numpy()
        return self.tot_force

    def check_early_stop(self, frame, ifprint=False, RL=False):
        if self.check_pos_nan():
            if ifprint:
                print("exist nan")
            return True
        self.gather_force()
        for i in range(len(self._effectors())):
            if (np.abs(self.tot_force[i]) > 10).any():
                if ifprint:
                    print("too much force")
                return True
            if np.linalg.norm(self.tot_force[i]) < 0.2 and frame > 10 and not RL:
                if ifprint:
                    print("no contact")
                return True
        return False

    # ---- BaseScene.get_observation_kernel (:1586-1619); the sampling constants of BaseScene.__init__ (:182-191)
    n_obs_cloth, n_obs_elastic = 4, 16

    def get_observation_kernel(self):
        e = self.engine
        c0 = self.cloths[0]
        ns, ms = c0.N // 4, c0.M // 4
        obs = []
        for c in self.cloths:
            for j in range(self.n_obs_cloth):
                for k in range(self.n_obs_cloth):
                    # (the reference indexes with cloth_N where M + 1 is meant: kept, it only selects which vertices are observed)
                    v = c.offset + (ns // 2 + j * ns) * self.cloth_N + (ms // 2 + k * ms)
                    obs.append(torch.cat([e.pos[v], e.vel[v]]))
        for el in getattr(self, "elastics", []):
            for j in range(self.n_obs_elastic):
                ii = ((el.n_verts // self.n_obs_elastic) * j - 1) % el.n_verts      # (-1 for j = 0: wraps to the last vertex)
                obs.append(torch.cat([e.pos[el.offset + ii], e.vel[el.offset + ii]]))
        o = torch.stack(obs).reshape(-1)
        if hasattr(self, "gripper"):
            g = np.concatenate([np.concatenate([self.gripper._pos[j], self.gripper._rot[j]]) for j in range(self.gripper.n_part)])
            o = torch.cat([o, torch.as_tensor(g, dtype=torch.float64, device=o.device)])
        self._observation = o
        return o

    observation = property(lambda self: TensorField(self._observation))

    # ---- BaseScene.get_paramters_grad (:1513-1525): per-vertex dF/dK of every body at the current positions
    def get_paramters_grad(self):
        e = self.engine
        self._d_kl, self._d_ka, self._d_kb = e.cloth_param_deri(0)
        if e.tet_bodies:
            self._d_mu, self._d_lam, _ = e.elastic_param_grad(None)
        else:
            self._d_mu = torch.zeros_like(self._d_kb); self._d_lam = torch.zeros_like(self._d_kb)

    d_kl = property(lambda self: TensorField(self._d_kl))
    d_ka = property(lambda self: TensorField(self._d_ka))
    d_kb = property(lambda self: TensorField(self._d_kb))
    d_mu = property(lambda self: TensorField(self._d_mu))
    d_lam = property(lambda self: TensorField(self._d_lam))
