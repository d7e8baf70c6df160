"""Scene_pick on the GPU (code/task_scene/Scene_pick.py, training/trajopt_pick_fold.py): a cloth under gravity on an arched frozen table
whose friction is pinned to 0.1 while the two pads use mu_cloth_elastic, with bending plasticity (k_angle 0.5).  Checks: the scene the
product builds equals the reference-made state; the agent's init_traj_pick_fold trajectory; every step of the press-down converges and
the last one is a fixed point of the REFERENCE iteration; the rewards against the engine's own dihedral angles; the adjoint of
get_loss_pick runs through the dense-LU path and agrees with finite differences of the rollout."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from thinshelllab_b200 import _lib
    from thinshelllab_b200.agent.traj_opt_single import agent_trajopt
    from thinshelllab_b200.engine.analytic_grad_single import Grad
    from thinshelllab_b200.task_scene.Scene_pick import Scene


def _rollout(s, tr, grad=None):
    T = tr.shape[0]
    agent = agent_trajopt(T, 2, max_moving_dist=0.001)
    agent.traj.from_numpy(tr)
    s.reset()
    s.mu_cloth_elastic[None] = 10.0
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
    return -s.compute_reward(), stats                                   # the loss get_loss_pick seeds (last frame)


def test_pick_scene_state_rollout_rewards_and_adjoint(golden_dir):
    import scipy.sparse.linalg as spla
    g = np.load(os.path.join(golden_dir, "scene_state_pick.npz"))
    s = Scene(cloth_size=0.06)
    e = s.engine
    assert np.array_equal(e.pos.cpu().numpy(), g["pos0"]) and np.array_equal(e.frozen.cpu().numpy(), g["frozen"])
    assert np.array_equal(s.faces, g["faces"]) and np.abs(e.mass.cpu().numpy() - g["mass"]).max() <= 1e-14 * g["mass"].max()
    assert s.gripper.n_part == 2 and s.elastic_cnt == 3
    s.cloths[0].Kb[None] = 200.0
    T = 6
    agent = agent_trajopt(T, 2, max_moving_dist=0.001)
    agent.init_traj_pick_fold()
    tr = agent.traj.to_numpy()
    assert np.allclose(tr[:, :2, 2], -0.0006 * np.arange(T)[:, None]) and not tr[:, :, [0, 1, 3, 4, 5]].any()
    grad = Grad(s, T, 2)
    loss, stats = _rollout(s, tr, grad)
    for f, st in enumerate(stats, 1):
        assert st.converged, (f, st)
    assert stats[-1].n_contacts > 50                                    # table below, pads above
    # rewards: row 8 height; the crease hinges between rows 7 and 9 (16 of them on the 17 x 17 grid) against the engine's angles
    c = s.cloths[0]
    z = e.pos[:c.NV, 2].cpu().numpy().reshape(17, 17)
    assert abs(s.compute_reward() - z[8].sum()) < 1e-15
    hi, hl = s._hinges_between_rows(7, 9)
    assert len(hi) == 16
    th = s._hinge_angles(hi, hl)
    ra = e.cloth_ref_angle[0].cpu().numpy()
    assert abs(s.compute_reward_pick_fold() - (ra[hi, hl].sum() + 0.01 * th.sum())) < 1e-15
    # the last step as a fixed point of the reference iteration
    vel1 = e.vel.clone()
    e.vel.copy_((grad._pos_buffer[T - 2] - grad._pos_buffer[T - 3]) / s.dt)
    e.prev_pos.copy_(grad._pos_buffer[T - 2])
    e.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_SPD | _lib.ASM_F64)
    p = spla.spsolve(e.matrix().tocsc(), e.residual())
    delta = np.abs(p).max() / s.dt
    print(f"Scene_pick: steps {[(st.newton_iters, st.linear_iters, st.n_contacts) for st in stats]}, reference Newton step at the last state {delta:.2e}")
    assert delta < 1e-6
    e.vel.copy_(vel1)
    # adjoint of get_loss_pick
    grad.get_loss_pick(s)
    for j in range(T - 1, 0, -1):
        it, flags, rr = grad.transfer_grad(j, s)
        assert flags == 0 and it == 0 and rr < 1e-9, (j, it, flags, rr)
    gg = grad._gripper_grad.copy()
    assert np.isfinite(gg).all() and np.abs(gg[1:]).max() > 0
    # finite differences of the rollout with respect to the pose of the last frame of part 0 (x, z, rotation about y).  The quaternion
    # update of gripper.step_simple (gripper_single.py:99-110) adds (-d.v, s d + d x v) without the factor 1/2, so a delta_rot of d turns
    # the pad by 2 d while gather_grad (:133-150) returns dL/d(angle): the reference's rotation gradient is half the derivative with
    # respect to its own action, and so is ours.
    for (part, comp, h, scale) in ((0, 2, 2e-6, 1.0), (0, 0, 2e-5, 1.0), (0, 4, 2e-4, 2.0)):
        tp, tm = tr.copy(), tr.copy()
        tp[T - 1, part, comp] += h; tm[T - 1, part, comp] -= h
        fd = (_rollout(s, tp)[0] - _rollout(s, tm)[0]) / (2 * h)
        an = gg[T - 1, part, comp] * s.gripper.n_bound * scale          # gather_grad returns the mean over the driven vertices
        print(f"Scene_pick dL/dpose[{T - 1}, part {part}, {comp}]: adjoint x n_bound x {scale:g} {an:.6e}  finite difference {fd:.6e}")
        assert abs(an - fd) <= 0.03 * max(abs(fd), abs(an)) + 1e-9, (part, comp, an, fd)
