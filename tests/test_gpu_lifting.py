"""Scene_lifting on the GPU (code/task_scene/Scene_lifting.py, training/trajopt_lifting.py): a scene with a FREE neo-Hookean box, three
tactile pads on three gripper parts and no frozen table.  Checks: the scene the product builds equals the reference-made state; every
step of a short rollout is a fixed point of the REFERENCE iteration (its projected matrix assembled by libtsl with the reference's
formulas -- verified entry by entry against the reference in test_gpu_folding -- and a SciPy direct solve); the trajectory adjoint of
get_loss_lift runs through the dense-LU path and its gripper gradient agrees with finite differences of the rollout."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from thinshelllab_b200 import _lib
    from thinshelllab_b200.agent.traj_opt_single import agent_trajopt
    from thinshelllab_b200.engine.analytic_grad_single import Grad
    from thinshelllab_b200.task_scene.Scene_lifting import Scene


def _traj(T):
    tr = np.zeros((T, 3, 6))
    for i in range(1, T):
        tr[i, 0] = [0.0, 0.0, -2.0e-4 * i, 0.0, 0.0, 0.0]            # the pad above presses down
        tr[i, 1] = [1.0e-4 * i, 0.0, 1.5e-4 * i, 0.0, 1e-3 * i, 0.0]   # the two below push up and sideways
        tr[i, 2] = [1.0e-4 * i, 0.0, 1.5e-4 * i, 0.0, 0.0, -1e-3 * i]
    return tr


def _rollout(s, tr, grad=None):
    T = tr.shape[0]
    agent = agent_trajopt(T, 3, max_moving_dist=0.001)
    agent.traj.from_numpy(tr)
    s.reset()
    if grad is not None:
        grad.reset()
        grad.copy_pos(s, 0)
    x0 = s.engine.pos.clone()
    stats = []
    for f in range(1, T):
        agent.get_action(f)
        s.action(f, agent.delta_pos, agent.delta_rot)
        stats.append(s.time_step())
        if grad is not None:
            grad.copy_pos(s, f)
    b = s.elastics[0]
    d = s.engine.pos[b.offset:b.offset + b.n_verts] - x0[b.offset:b.offset + b.n_verts]
    d[:, 0] += 0.012; d[:, 1] += 0.012
    return 0.5 * float((d * d).sum().item()), stats                       # the loss whose gradient get_loss_lift seeds


def test_lifting_scene_state_rollout_and_adjoint(golden_dir):
    import scipy.sparse.linalg as spla
    g = np.load(os.path.join(golden_dir, "scene_state_lifting.npz"))
    s = Scene(cloth_size=0.06)
    e = s.engine
    assert np.array_equal(e.pos.cpu().numpy(), g["pos0"]) and np.array_equal(e.frozen.cpu().numpy(), g["frozen"])
    assert np.array_equal(s.faces, g["faces"]) and np.abs(e.mass.cpu().numpy() - g["mass"]).max() <= 1e-14 * g["mass"].max()
    assert s.gripper.n_part == 3 and abs(s.compute_reward() + 3 * 0.0 + 125 * 2 * 0.012 ** 2) < 1e-12      # the box at rest: -(2 x 0.012^2) per vertex
    T = 3
    tr = _traj(T)
    grad = Grad(s, T, 3)
    loss, stats = _rollout(s, tr, grad)
    for f, st in enumerate(stats, 1):
        assert st.converged, (f, st)
    # the last step as a fixed point of the reference iteration (velocity of the start of the step)
    vel1 = e.vel.clone()
    e.vel.copy_((grad._pos_buffer[T - 2] - grad._pos_buffer[T - 3]) / s.dt if T > 2 else torch.zeros_like(e.vel))
    e.prev_pos.copy_(grad._pos_buffer[T - 2])
    e.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_SPD | _lib.ASM_F64)
    p = spla.spsolve(e.matrix().tocsc(), e.residual())
    delta = np.abs(p).max() / s.dt
    print(f"Scene_lifting: steps {[(st.newton_iters, st.linear_iters, st.n_contacts) for st in stats]}, reference Newton step at the last state {delta:.2e}")
    assert delta < 1e-6
    e.vel.copy_(vel1)
    assert stats[-1].n_contacts > 0
    # adjoint of get_loss_lift
    grad.get_loss_lift(s)
    for j in range(T - 1, 0, -1):
        it, flags, rr = grad.transfer_grad(j, s)
        assert flags == 0 and it == 0 and rr < 1e-9, (j, it, flags, rr)              # 3.6 k unknowns: the dense-LU path
    gg = grad._gripper_grad.copy()
    assert np.isfinite(gg).all() and np.abs(gg[1:]).max() > 0
    # finite differences of the rollout with respect to the pose of frame T-1 (pad 0 z, pad 1 x)
    for (part, comp) in ((0, 2), (1, 0)):
        h = 2e-6
        tp, tm = tr.copy(), tr.copy()
        tp[T - 1, part, comp] += h; tm[T - 1, part, comp] -= h
        fd = (_rollout(s, tp)[0] - _rollout(s, tm)[0]) / (2 * h)
        # gripper.gather_grad returns the MEAN over the driven vertices (d_pos /= n_bound, gripper_single.py:146-147): dL/dpose = n_bound x that
        an = gg[T - 1, part, comp] * s.gripper.n_bound
        print(f"Scene_lifting dL/dpose[{T - 1}, part {part}, {comp}]: adjoint x n_bound {an:.6e}  finite difference {fd:.6e}")
        assert abs(an - fd) <= 0.01 * max(abs(fd), abs(an)) + 1e-9, (part, comp, an, fd)
