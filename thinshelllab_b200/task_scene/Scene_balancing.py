"""Scene_balancing on the B200 engine (code/task_scene/Scene_balancing.py, training/trajopt_balancing.py): a 15 x 7 cloth strip held at
both ends by a two-finger gripper (two parts, an upper and a lower tactile pad each: engine/gripper_tactile.py) with the TetGen ball of
data/ball.* lying on it; the task keeps the ball over the centre of the strip (or throws it: compute_reward_throwing)."""
import os

import numpy as np
import torch

from ..engine.scene_builder import balancing_state
from ._multi_body import MultiBodyScene
from .Scene_bouncing import Body  # noqa: F401  (the reference module exports it)


class Scene(MultiBodyScene):
    def __init__(self, cloth_size=0.06, device="cuda:0", *, state=None, max_newton=50):
        self.max_newton = max_newton
        self.cloth_size = cloth_size
        self._build(state if state is not None else balancing_state(cloth_size=float(cloth_size)), device=device)

    def _centre(self):
        return (self.cloth_N + 1) // 2 * (self.cloth_M + 1) + (self.cloth_M + 1) // 2

    def _ball_minus_centre(self, x):
        b = self.elastics[0]
        return x[..., b.offset:b.offset + b.n_verts, :2] - x[..., self._centre():self._centre() + 1, :2]

    def compute_reward(self):
        """:137-144: squared horizontal distance of every ball vertex from the centre vertex of the strip"""
        return -float((self._ball_minus_centre(self.engine.pos) ** 2).sum().item())

    def compute_reward_all(self, analy_grad):
        """:146-153: the same, summed over every stored frame"""
        return -float((self._ball_minus_centre(analy_grad._pos_buffer) ** 2).sum().item())

    def _end_rows_z(self):
        z = self.engine.pos[:self.cloths[0].NV, 2]
        M1 = self.cloth_M + 1
        return torch.cat([z[:M1], z[self.cloth_N * M1:self.cloth_N * M1 + M1]])

    def compute_reward_throwing(self, analy_grad):
        """:155-166: height of the ball in the last stored frame, the two held ends of the strip kept at z = 0"""
        b = self.elastics[0]
        up = analy_grad._pos_buffer[analy_grad.tot_timestep - 1, b.offset:b.offset + b.n_verts, 2].sum()
        return float((up - 10.0 * (self._end_rows_z() ** 2).sum()).item())

    def compute_reward_throwing_RL(self):
        b = self.elastics[0]
        return float((self.engine.pos[b.offset:b.offset + b.n_verts, 2].sum() - 10.0 * (self._end_rows_z() ** 2).sum()).item())

    # ---- save_all / load_all (:187-209): gripper files, the state file, and the contact-projection flags.  The projection of a vertex
    # onto the surfaces is recomputed at the start of every time step here (and by f_contact in the reference's time_step :231-235), so
    # proj_flag / proj_dir carry no state across a save: they are written as zeros of the reference's shapes and ignored on load.
    def save_all(self, path):
        os.makedirs(path, exist_ok=True)
        self.gripper.save_all(path)
        self.save_state(os.path.join(path, "state"))
        np.save(os.path.join(path, "proj_flag.npy"), np.zeros(self.tot_NV, np.int32))
        np.save(os.path.join(path, "proj_dir.npy"), np.zeros((self.tot_NV, 3)))
        np.save(os.path.join(path, "border_flag.npy"), np.zeros(self.tot_NV, np.int32))

    def load_all(self, path):
        self.gripper.load_all(path)
        self.load_state(os.path.join(path, "state"))
        self.engine.reset_contact_state()
