"""Scene_pick on the B200 engine (code/task_scene/Scene_pick.py): a flat 16 x 16 cloth on an arched frozen table (friction 0.1), two
tactile pads above on two gripper parts that pinch the cloth and crease it (training/trajopt_pick_fold.py: k_angle 0.5, reward on the
rest angles of the crease between grid rows 7 and 9)."""
import numpy as np

from ..engine.scene_builder import pick_state
from ._multi_body import MultiBodyScene
from .Scene_bouncing import Body  # noqa: F401  (the reference module exports it)


class Scene(MultiBodyScene):
    def __init__(self, cloth_size=0.06, device="cuda:0", *, state=None, max_newton=50):
        self.max_newton = max_newton
        self.cloth_size = cloth_size
        self._build(state if state is not None else pick_state(cloth_size=float(cloth_size)), device=device)

    def compute_reward(self):
        """:121-127: height of grid row 8"""
        c = self.cloths[0]
        z = self.engine.pos[c.offset:c.offset + c.NV, 2].cpu().numpy()
        return float(z[(np.arange(c.NV) // (c.M + 1)) == 8].sum())

    def compute_reward_pick_fold(self):
        """:139-152: rest angle + 0.01 x current dihedral angle over the hinges between grid rows 7 and 9"""
        hi, hl = self._hinges_between_rows(7, 9)
        ra = self.engine.cloth_ref_angle[0].cpu().numpy()
        return float(ra[hi, hl].sum() + 0.01 * self._hinge_angles(hi, hl).sum())
