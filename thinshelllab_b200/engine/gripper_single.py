"""gripper (code/engine/gripper_single.py) on the B200 engine: the rigid pose that drives the bound vertices of a tactile pad.
Pose bookkeeping (a position and a quaternion per part) is host arithmetic in float64, exactly as the reference does it in
Python scope; the two passes over vertices -- moving the bound vertices and gathering the adjoint -- are CUDA
(tsl_gripper_apply / tsl_gripper_gather)."""
import numpy as np
import torch

from ..fields import TensorField


def quat_to_rotmat32(q):
    """gripper.get_rotmat (:85-94): evaluated in float64, stored in a float32 matrix field"""
    s, x, y, z = (float(v) for v in q)
    R = np.array([[s * s + x * x - y * y - z * z, 2 * (x * y - s * z), 2 * (x * z + s * y)],
                  [2 * (x * y + s * z), s * s - x * x + y * y - z * z, 2 * (y * z - s * x)],
                  [2 * (x * z - s * y), 2 * (y * z + s * x), s * s - x * x - y * y + z * z]])
    return R.astype(np.float32)


def pose_step(pos, rot, dpos, drot):
    """one part of gripper.step_simple (:111-128): translate, then q += (-(w . v), s w + w x v) and renormalise
    (w = delta_rot, q = (s, v))"""
    w = np.asarray(drot, np.float64)
    s, v = float(rot[0]), np.asarray(rot[1:], np.float64)
    real = -float(np.dot(w, v))
    res = s * w + np.cross(w, v)
    q = np.array([s + real, v[0] + res[0], v[1] + res[1], v[2] + res[2]])
    return np.asarray(pos, np.float64) + np.asarray(dpos, np.float64), q / np.sqrt(np.dot(q, q))


class gripper:
    def __init__(self, sys, part_offsets, F_x, bound_idx, pos0):
        """part_offsets: scene-global vertex offset of every driven body; F_x [parts, n_verts, 3] body vertices relative to the pose
        (init_kernel :51-59); bound_idx [n_bound] body-local ids of the driven vertices (the same for every part)"""
        e = sys.engine
        self._sys = sys
        self.n_part = len(part_offsets)
        self.part_offsets = [int(o) for o in part_offsets]
        self._F_x = torch.as_tensor(np.ascontiguousarray(F_x), dtype=torch.float64, device=e.device).contiguous()
        self._bound_idx = torch.as_tensor(np.ascontiguousarray(bound_idx), dtype=torch.int32, device=e.device).contiguous()
        self.n_verts, self.n_bound = self._F_x.shape[1], int(self._bound_idx.numel())
        self._pos = np.array(pos0, np.float64).reshape(self.n_part, 3).copy()
        self._rot = np.tile(np.array([1.0, 0.0, 0.0, 0.0]), (self.n_part, 1))
        self._rotmat = np.stack([quat_to_rotmat32(q) for q in self._rot])
        self._d_pos = np.zeros((self.n_part, 3)); self._d_angle = np.zeros((self.n_part, 3))

    pos = property(lambda self: TensorField(torch.from_numpy(self._pos)))
    rot = property(lambda self: TensorField(torch.from_numpy(self._rot)))
    rotmat = property(lambda self: TensorField(torch.from_numpy(self._rotmat)))
    d_pos = property(lambda self: TensorField(torch.from_numpy(self._d_pos)))
    d_angle = property(lambda self: TensorField(torch.from_numpy(self._d_angle)))
    F_x = property(lambda self: TensorField(self._F_x))
    bound_idx = property(lambda self: TensorField(self._bound_idx))

    def init(self, sys, pos_array):
        self._pos[:] = np.asarray(pos_array, np.float64).reshape(self.n_part, 3)
        self._rot[:] = [1.0, 0.0, 0.0, 0.0]
        self.get_rotmat()

    def set(self, pos, rot, step):
        """pose of a stored frame (Grad.gripper_pos_buffer / gripper_rot_buffer)"""
        p = pos.to_numpy() if hasattr(pos, "to_numpy") else np.asarray(pos)
        r = rot.to_numpy() if hasattr(rot, "to_numpy") else np.asarray(rot)
        self._pos[:] = p[step]
        self._rot[:] = r[step]

    def get_rotmat(self):
        for j in range(self.n_part):
            self._rotmat[j] = quat_to_rotmat32(self._rot[j])

    def step_simple(self, delta_pos, delta_rot):
        dp = delta_pos.to_numpy() if hasattr(delta_pos, "to_numpy") else np.asarray(delta_pos)
        dr = delta_rot.to_numpy() if hasattr(delta_rot, "to_numpy") else np.asarray(delta_rot)
        for j in range(self.n_part):
            self._pos[j], self._rot[j] = pose_step(self._pos[j], self._rot[j], dp[j], dr[j])
        self.get_rotmat()

    def update_bound(self, sys=None):
        """get_vert_pos + update_bound + pushup for the driven vertices only (:79-83, 157-161)"""
        e = self._sys.engine
        for j in range(self.n_part):
            e.gripper_apply(self.part_offsets[j], self._bound_idx, self._F_x[j], self._pos[j], self._rotmat[j])

    def gather_grad(self, grad, sys=None):
        """:133-150; grad = tmp_z_frozen [3 tot_NV] (CUDA tensor)"""
        e = self._sys.engine
        g = grad.t if isinstance(grad, TensorField) else grad
        for j in range(self.n_part):
            out = e.gripper_gather(g, self.part_offsets[j], self._bound_idx, self._F_x[j], self._rotmat[j])
            self._d_pos[j], self._d_angle[j] = out[:3], out[3:]
