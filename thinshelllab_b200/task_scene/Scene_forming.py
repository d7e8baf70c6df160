"""Scene_forming on the B200 engine (code/task_scene/Scene_forming.py): the Scene_folding layout -- cloth strip pinned along its last row,
frozen table, one tactile pad on a gripper, contacts both ways -- with a 15 x 7 strip, k_contact = 20000 and a position reward
(:127-133).  Everything on the hot path is shared with Scene_folding; only the scene arrays and the reward differ."""
import numpy as np
import torch

from .Scene_bouncing import Body  # noqa: F401  (the reference module exports it)
from .Scene_folding import Scene as _FoldingScene


class Scene(_FoldingScene):
    FORMING = True                             # engine/scene_builder.folding_state(forming=True): 15 x 7 strip, half_curve_num 3, k_contact 20000

    def compute_reward(self, target_pos):
        """:127-133: minus the squared distance of the cloth to the target shape"""
        c = self.cloths[0]
        t = torch.as_tensor(np.asarray(target_pos, np.float64), device=self.engine.device)
        d = self.engine.pos[c.offset:c.offset + c.NV] - t
        return float(-(d * d).sum().item())
