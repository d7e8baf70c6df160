"""Scene_card on the B200 engine (code/task_scene/Scene_card.py, training/trajopt_card.py): three stacked cards (12 x 8 cloths) on a
frozen table, two pads at the ends turned to face each other and one above; neighbouring cards touch each other (cloth-cloth contact,
friction 0.1), the pads only push cloth vertices.  The driver identifies the bending stiffness of the cards (analytic_grad_system)."""
from ..engine.scene_builder import card_state
from ._multi_body import MultiBodyScene
from .Scene_bouncing import Body  # noqa: F401  (the reference module exports it)


class Scene(MultiBodyScene):
    def __init__(self, cloth_size=0.06, device="cuda:0", *, state=None, max_newton=50):
        self.max_newton = max_newton
        self.cloth_size = cloth_size
        self._build(state if state is not None else card_state(cloth_size=float(cloth_size)), device=device)

    def compute_reward(self):
        """:166-171"""
        c = self.cloths[0]
        return -float(self.engine.pos[c.offset:c.offset + c.NV, 0].sum().item())
