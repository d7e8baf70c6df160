"""Scene_lifting on the B200 engine (code/task_scene/Scene_lifting.py): a flat 15 x 15 cloth carrying a small heavy neo-Hookean box, one
tactile pad above and two below on three gripper parts; the task lifts / carries the box by moving the pads (training/trajopt_lifting.py).

The reference caps Newton at 15 iterations per step (:203) -- its projected Newton is usually not converged by then; this engine
iterates to the same stopping test (delta < 1e-7) with its own Newton matrix, so `max_newton` only bounds the worst case."""
import numpy as np
import torch

from ..engine.scene_builder import lifting_state
from ._multi_body import MultiBodyScene
from .Scene_bouncing import Body  # noqa: F401  (the reference module exports it)


class Scene(MultiBodyScene):
    def __init__(self, cloth_size=0.06, device="cuda:0", *, state=None, max_newton=50):
        self.max_newton = max_newton
        self.cloth_size = cloth_size
        self._build(state if state is not None else lifting_state(cloth_size=float(cloth_size)), device=device)

    def compute_reward(self):
        """:153-159: minus the squared distance of the box from its rest shape shifted by (-0.012, -0.012, 0) -- F_x - F_ox + offsets"""
        b = self.elastics[0]
        x = self.engine.pos[b.offset:b.offset + b.n_verts]
        ox = torch.as_tensor(b.rest - np.array([-0.025, -0.005, 0.0003]), device=x.device)       # F_ox: rest shape before init's offset
        d = x - ox + torch.tensor([0.025 + 0.012, 0.005 + 0.012, -0.0003], dtype=torch.float64, device=x.device)
        return float(-(d * d).sum().item())
