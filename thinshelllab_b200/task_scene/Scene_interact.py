"""Scene_interact on the B200 engine (code/task_scene/Scene_interact.py, training/trajopt_interact.py): a cloth half on a frozen table
with a free box lying on it; ONE two-finger gripper part (engine/gripper_tactile.py) closes on the free end of the cloth during the first
frames (the opening shrinks by 0.6 mm per frame) and then pulls: the task separates the box from the cloth (compute_reward) or drags it
along (compute_reward_1)."""
import numpy as np

from ..engine.scene_builder import interact_state
from ._multi_body import MultiBodyScene
from .Scene_bouncing import Body  # noqa: F401  (the reference module exports it)


class Scene(MultiBodyScene):
    def __init__(self, cloth_size=0.06, device="cuda:0", soft=False, dense=10000.0, *, state=None, max_newton=50):
        self.max_newton = max_newton
        self.cloth_size, self.soft, self.dense, self.extra_obj = cloth_size, soft, dense, True
        self._build(state if state is not None else interact_state(cloth_size=float(cloth_size), dense=float(dense)), device=device)
        self.effector_cnt = 3                                   # init_scene_parameters :48: elastics[1], elastics[2] are the effector pads

    def compute_reward(self):
        """:149-156"""
        e, c, b = self.engine, self.cloths[0], self.elastics[3]
        return float((-e.pos[c.offset:c.offset + c.NV, 0].sum() + e.pos[b.offset:b.offset + b.n_verts, 0].sum() * 256.0 / 144.0).item())

    def compute_reward_1(self):
        """:158-163"""
        b = self.elastics[3]
        return -float(self.engine.pos[b.offset:b.offset + b.n_verts, 0].sum().item())

    def action(self, step, delta_pos, delta_rot):
        """:165-172: the gripper closes over frames 1..4, then only moves"""
        if step < 5:
            self.gripper.step(delta_pos, delta_rot, np.array([-0.0006]))
        else:
            self.gripper.step_simple(delta_pos, delta_rot)
        self.gripper.update_bound(self)
