#!/usr/bin/env python
"""One forward + trajectory-adjoint frame of the configs[3]-style scene (N x N sheet on the table, tactile pad pressed into it) -- a
short, self-contained command for `ncu` launch lists:  python tools/run_pad_sheet.py [N] [frames]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thinshelllab_b200.agent.traj_opt_single import agent_trajopt  # noqa: E402
from thinshelllab_b200.engine.analytic_grad_single import Grad  # noqa: E402
from thinshelllab_b200.synthetic import pad_sheet_scene  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 316
T = 1 + (int(sys.argv[2]) if len(sys.argv) > 2 else 2)
g = None   # pad arrays from engine/scene_builder.folding_state
s = pad_sheet_scene(N, g)
agent = agent_trajopt(T, 1, max_moving_dist=0.001)
traj = np.zeros((T, 1, 6)); traj[:, 0, 2] = -1.5e-4 * np.arange(T)
agent.traj.from_numpy(traj)
grad = Grad(s, T, 1)
grad.copy_pos(s, 0)
for frame in range(1, T):
    agent.get_action(frame)
    s.action(frame, agent.delta_pos, agent.delta_rot)
    st = s.time_step()
    print(f"frame {frame}: newton {st.newton_iters} pcg {st.linear_iters} contacts {st.n_contacts} converged {st.converged}")
    grad.copy_pos(s, frame)
grad._pos_grad[T - 1, :s.cloths[0].NV, 2] = 1.0
for j in range(T - 1, 0, -1):
    print("backward", j, grad.transfer_grad(j, s, rel_tol=1e-8), grad._gripper_grad[j])
