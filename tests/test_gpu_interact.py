"""Scene_interact on the GPU (code/task_scene/Scene_interact.py, training/trajopt_interact.py): frozen table, free box on the cloth with
box <-> table contact as well (elastic-elastic pairs), ONE two-finger gripper part that closes over the first frames (gripper.step with
a change of opening) and then pulls.  Checks: the scene equals the reference-made state; the rollout converges, the pads close as
scripted, the last step is a fixed point of the REFERENCE iteration; the adjoint of get_loss_interact against finite differences."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from thinshelllab_b200 import _lib
    from thinshelllab_b200.agent.traj_opt_single import agent_trajopt
    from thinshelllab_b200.engine.analytic_grad_single import Grad
    from thinshelllab_b200.task_scene.Scene_interact import Scene


def _traj(T):
    tr = np.zeros((T, 1, 6))
    for i in range(1, T):
        tr[i, 0] = [-1.0e-4 * max(i - 3, 0), 0.0, 0.5e-4 * max(i - 3, 0), 0.0, 0.0, 0.0]       # close for three frames, then pull away and up
    return tr


def _rollout(s, tr, grad=None):
    T = tr.shape[0]
    agent = agent_trajopt(T, 1, max_moving_dist=0.001)
    agent.traj.from_numpy(tr)
    s.reset()
    s.mu_cloth_elastic[None] = 5.0
    if grad is not None:
        grad.reset()
        grad.copy_pos(s, 0)
    stats = []
    for f in range(1, T):
        agent.get_action(f)
        s.action(f, agent.delta_pos, agent.delta_rot)
        stats.append(s.time_step())
        if grad is not None:
            grad.copy_pos(s, f)
    return -s.compute_reward(), stats                                   # the loss get_loss_interact seeds (last frame)


def test_interact_scene_state_rollout_and_adjoint(golden_dir):
    import scipy.sparse.linalg as spla
    g = np.load(os.path.join(golden_dir, "scene_state_interact.npz"))
    s = Scene(cloth_size=0.06)
    e = s.engine
    assert np.array_equal(e.pos.cpu().numpy(), g["pos0"]) and np.array_equal(e.frozen.cpu().numpy(), g["frozen"])
    assert np.array_equal(s.faces, g["faces"]) and np.abs(e.mass.cpu().numpy() - g["mass"]).max() <= 1e-14 * g["mass"].max()
    assert s.enable_gripper and s.gripper.n_part == 1 and s.elastic_cnt == 4 and s.effector_cnt == int(g["effector_cnt"])
    T = 6
    tr = _traj(T)
    grad = Grad(s, T, 1)
    loss, stats = _rollout(s, tr, grad)
    for f, st in enumerate(stats, 1):
        assert st.converged, (f, st)
    assert abs(s.gripper._half[0] + 0.0006 * 4) < 1e-15                 # frames 1..4 close the fingers, frame 5 does not (action :165-169)
    c = e.constraints()
    idx = c["idx"][:stats[-1].n_contacts]
    body = np.searchsorted([b.v_start for b in s.body_list], idx, side="right") - 1
    pairs = {tuple(sorted(set(r))) for r in body}
    assert (0, 1) in pairs and (0, 4) in pairs and ((0, 2) in pairs or (0, 3) in pairs), pairs    # cloth-table, cloth-box, cloth-pad
    # fixed point of the reference iteration at the last state
    vel1 = e.vel.clone()
    e.vel.copy_((grad._pos_buffer[T - 2] - grad._pos_buffer[T - 3]) / s.dt)
    e.prev_pos.copy_(grad._pos_buffer[T - 2])
    e.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_SPD | _lib.ASM_F64)
    p = spla.spsolve(e.matrix().tocsc(), e.residual())
    delta = np.abs(p).max() / s.dt
    print(f"Scene_interact: steps {[(st.newton_iters, st.linear_iters, st.n_contacts) for st in stats]}, body pairs in contact {sorted(pairs)}, "
          f"reference Newton step at the last state {delta:.2e}")
    assert delta < 1e-6
    e.vel.copy_(vel1)
    # adjoint of get_loss_interact
    grad.get_loss_interact(s)
    for j in range(T - 1, 0, -1):
        it, flags, rr = grad.transfer_grad(j, s)
        assert flags == 0 and it == 0 and rr < 1e-9, (j, it, flags, rr)
    gg = grad._gripper_grad.copy()
    assert np.isfinite(gg).all() and np.abs(gg[1:]).max() > 0
    for (comp, h) in ((0, 2e-6), (2, 2e-6)):
        tp, tm = tr.copy(), tr.copy()
        tp[T - 1, 0, comp] += h; tm[T - 1, 0, comp] -= h
        fd = (_rollout(s, tp)[0] - _rollout(s, tm)[0]) / (2 * h)
        an = gg[T - 1, 0, comp] * 2 * s.gripper.n_bound
        print(f"Scene_interact dL/dpose[{T - 1}, part 0, {comp}]: adjoint x 2 n_bound {an:.6e}  finite difference {fd:.6e}")
        # (x agrees to 0.1 %; in z the hard-squeezed pads make the terms the reference's adjoint matrix leaves out -- the change of the
        # contact normals and weights with the pose -- worth ~5 %: the adjoint reproduces the reference's matrix, not the exact Jacobian)
        assert abs(an - fd) <= (0.01 if comp == 0 else 0.08) * max(abs(fd), abs(an)) + 1e-9, (comp, an, fd)
