"""Scene_sliding on the B200 engine (code/task_scene/Scene_sliding.py, training/trajopt_silding.py): three stacked 15 x 15 cloths on a
frozen table, one pad pressing on the stack and dragging it; the cloth-cloth friction coefficient mu_cloth_cloth is the parameter the
driver identifies (Grad.count_friction_grad -> contact_energy_backprop_friction over the first nc1 constraints, the cloth-cloth ones:
tsl_friction_coef_grad over the first n_cloth_cloth_pairs contact pairs)."""
from ..engine.scene_builder import sliding_state
from ._multi_body import MultiBodyScene
from .Scene_bouncing import Body  # noqa: F401  (the reference module exports it)


class Scene(MultiBodyScene):
    def __init__(self, cloth_size=0.06, device="cuda:0", *, state=None, max_newton=50):
        self.max_newton = max_newton
        self.cloth_size = cloth_size
        g = state if state is not None else sliding_state(cloth_size=float(cloth_size))
        self.n_cloth_cloth_pairs = int(g["n_cloth_cloth_pairs"])
        self._build(g, device=device)

    def compute_reward(self):
        """:127-132"""
        c = self.cloths[0]
        return -float(self.engine.pos[c.offset:c.offset + c.NV, 0].sum().item())
