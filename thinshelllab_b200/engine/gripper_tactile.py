"""Two-finger gripper (code/engine/gripper_tactile.py) on the B200 engine: every part carries an UPPER and a LOWER tactile pad
(elastics[2 j + 1] / elastics[2 j + 2]) and a half opening; the pose bookkeeping is host arithmetic in float64 as in gripper_single,
the passes over vertices are tsl_gripper_apply / tsl_gripper_gather, once per pad."""
import os

import numpy as np
import torch

from ..fields import TensorField
from .gripper_single import pose_step, quat_to_rotmat32


class gripper:
    def __init__(self, sys, upper_offsets, lower_offsets, F_x_upper, F_x_lower, bound_idx, pos0):
        """upper_offsets / lower_offsets: scene-global vertex offsets of the two pads of every part; F_x_upper / F_x_lower
        [parts, n_verts, 3]: pad vertices relative to the part's position (init_kernel :41-50); bound_idx: the driven vertices, the same
        body-local ids in every pad"""
        e = sys.engine
        self._sys = sys
        self.n_part = len(upper_offsets)
        self.upper_offsets, self.lower_offsets = [int(o) for o in upper_offsets], [int(o) for o in lower_offsets]
        dev = e.device
        self._F_x_upper = torch.as_tensor(np.ascontiguousarray(F_x_upper), dtype=torch.float64, device=dev).contiguous()
        self._F_x_lower = torch.as_tensor(np.ascontiguousarray(F_x_lower), dtype=torch.float64, device=dev).contiguous()
        self._F0 = (self._F_x_upper.clone(), self._F_x_lower.clone())
        self._bound_idx = torch.as_tensor(np.ascontiguousarray(bound_idx), dtype=torch.int32, device=dev).contiguous()
        self._all_idx = torch.arange(self._F_x_upper.shape[1], dtype=torch.int32, device=dev)
        self.n_verts, self.n_bound = self._F_x_upper.shape[1], int(self._bound_idx.numel())
        self._pos = np.array(pos0, np.float64).reshape(self.n_part, 3).copy()
        self._rot = np.tile(np.array([1.0, 0.0, 0.0, 0.0]), (self.n_part, 1))
        self._rotmat = np.stack([quat_to_rotmat32(q) for q in self._rot])
        self._half = np.zeros(self.n_part)
        self._d_pos = np.zeros((self.n_part, 3)); self._d_angle = np.zeros((self.n_part, 3)); self._d_dist = np.zeros(self.n_part)

    pos = property(lambda self: TensorField(torch.from_numpy(self._pos)))
    rot = property(lambda self: TensorField(torch.from_numpy(self._rot)))
    rotmat = property(lambda self: TensorField(torch.from_numpy(self._rotmat)))
    d_pos = property(lambda self: TensorField(torch.from_numpy(self._d_pos)))
    d_angle = property(lambda self: TensorField(torch.from_numpy(self._d_angle)))
    d_dist = property(lambda self: TensorField(torch.from_numpy(self._d_dist)))
    half_gripper_dist = property(lambda self: TensorField(torch.from_numpy(self._half)))
    F_x_upper = property(lambda self: TensorField(self._F_x_upper))
    F_x_lower = property(lambda self: TensorField(self._F_x_lower))
    bound_idx = property(lambda self: TensorField(self._bound_idx))

    def init(self, sys, pos_array):
        """init_kernel (:39-61): pose, zero opening, pads relative to the pose as first built"""
        self._pos[:] = np.asarray(pos_array, np.float64).reshape(self.n_part, 3)
        self._rot[:] = [1.0, 0.0, 0.0, 0.0]
        self._half[:] = 0.0
        self._F_x_upper.copy_(self._F0[0]); self._F_x_lower.copy_(self._F0[1])
        self.get_rotmat()

    def set(self, pos, rot, step):
        p = pos.to_numpy() if hasattr(pos, "to_numpy") else np.asarray(pos)
        r = rot.to_numpy() if hasattr(rot, "to_numpy") else np.asarray(rot)
        self._pos[:] = p[step]
        self._rot[:] = r[step]

    def get_rotmat(self):
        for j in range(self.n_part):
            self._rotmat[j] = quat_to_rotmat32(self._rot[j])

    def get_vert_pos(self):
        pass                                            # world positions are produced where they are consumed (update_bound / update_all)

    def step_simple(self, delta_pos, delta_rot):
        dp = delta_pos.to_numpy() if hasattr(delta_pos, "to_numpy") else np.asarray(delta_pos)
        dr = delta_rot.to_numpy() if hasattr(delta_rot, "to_numpy") else np.asarray(delta_rot)
        for j in range(self.n_part):
            self._pos[j], self._rot[j] = pose_step(self._pos[j], self._rot[j], dp[j], dr[j])
        self.get_rotmat()

    def step(self, delta_pos, delta_rot, delta_dis):
        """:131-148: step_simple plus the opening: the upper pad moves by +delta, the lower by -delta along the part's own z axis"""
        dd = delta_dis.to_numpy() if hasattr(delta_dis, "to_numpy") else np.asarray(delta_dis)
        self.step_simple(delta_pos, delta_rot)
        for j in range(self.n_part):
            self._half[j] += float(dd[j])
            self.open_gripper(float(dd[j]), j)

    def open_gripper(self, delta_dis, j):
        self._F_x_upper[j, :, 2] += delta_dis
        self._F_x_lower[j, :, 2] -= delta_dis

    def _apply(self, idx):
        e = self._sys.engine
        for j in range(self.n_part):
            e.gripper_apply(self.upper_offsets[j], idx, self._F_x_upper[j], self._pos[j], self._rotmat[j])
            e.gripper_apply(self.lower_offsets[j], idx, self._F_x_lower[j], self._pos[j], self._rotmat[j])

    def update_bound(self, sys=None):
        self._apply(self._bound_idx)

    def update_all(self, sys=None):
        self._apply(self._all_idx)

    def gather_grad(self, grad, sys=None):
        """:150-172: mean over the 2 n_bound driven vertices of the two pads, then clamped to +-10"""
        e = self._sys.engine
        g = grad.t if isinstance(grad, TensorField) else grad
        for j in range(self.n_part):
            up = e.gripper_gather(g, self.upper_offsets[j], self._bound_idx, self._F_x_upper[j], self._rotmat[j], clamp_pos=1e300, clamp_angle=1e300)
            lo = e.gripper_gather(g, self.lower_offsets[j], self._bound_idx, self._F_x_lower[j], self._rotmat[j], clamp_pos=1e300, clamp_angle=1e300)
            m = np.clip(0.5 * (np.asarray(up) + np.asarray(lo)), -10.0, 10.0)
            self._d_pos[j], self._d_angle[j] = m[:3], m[3:]

    # ---- state files (:174-222): one .npy per field, the names of the reference
    def _world(self, F):
        R = torch.from_numpy(self._rotmat.astype(np.float64)).to(F.device)
        return torch.from_numpy(self._pos).to(F.device)[:, None, :] + torch.einsum("jab,jvb->jva", R, F)

    def save_all(self, path):
        os.makedirs(path, exist_ok=True)
        out = {"F_x_upper": self._F_x_upper.cpu().numpy(), "F_x_upper_world": self._world(self._F_x_upper).cpu().numpy(),
               "F_x_lower": self._F_x_lower.cpu().numpy(), "F_x_lower_world": self._world(self._F_x_lower).cpu().numpy(),
               "pos": self._pos, "rot": self._rot, "rotmat": self._rotmat, "half_gripper_dist": self._half}
        for k, v in out.items():
            np.save(os.path.join(path, k + ".npy"), v)

    def load_all(self, path):
        self._F_x_upper.copy_(torch.from_numpy(np.load(os.path.join(path, "F_x_upper.npy"))))
        self._F_x_lower.copy_(torch.from_numpy(np.load(os.path.join(path, "F_x_lower.npy"))))
        self._pos[:] = np.load(os.path.join(path, "pos.npy"))
        self._rot[:] = np.load(os.path.join(path, "rot.npy"))
        self._rotmat[:] = np.load(os.path.join(path, "rotmat.npy"))
        self._half[:] = np.load(os.path.join(path, "half_gripper_dist.npy"))
