"""Scene_balancing on the GPU (code/task_scene/Scene_balancing.py, training/trajopt_balancing.py): the TetGen ball (data/ball.*: BASELINE
configs[3]'s second volumetric body) on a cloth strip between the pads of the two-finger gripper (engine/gripper_tactile.py: two parts,
an upper and a lower pad each).  Checks: the scene the product builds equals the reference-made state; a short rollout converges and its
last step is a fixed point of the REFERENCE iteration; save_all / load_all round trip; the adjoint of get_loss_balance (dense-LU path)
against finite differences of the rollout through gripper_tactile.gather_grad."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from thinshelllab_b200 import _lib
    from thinshelllab_b200.agent.traj_opt_single import agent_trajopt
    from thinshelllab_b200.engine.analytic_grad_single import Grad
    from thinshelllab_b200.task_scene.Scene_balancing import Scene


def _traj(T):
    tr = np.zeros((T, 2, 6))
    for i in range(1, T):
        tr[i, 0] = [0.0, 0.0, 1.5e-4 * i, 0.0, 1e-3 * i, 0.0]        # one end lifted and tilted: the ball starts rolling
        tr[i, 1] = [1.0e-4 * i, 0.0, -1.0e-4 * i, 0.0, 0.0, 0.0]      # the other lowered and pushed inwards
    return tr


def _grip(s, path):
    """what the reference's `trajopt_balancing.py --save` run leaves in ../data/balance_state: the two grippers closed on the ends of the
    strip (here: 0.2 mm per side, two frames to settle), written with save_all"""
    s.reset()
    s.mu_cloth_elastic[None] = 5.0
    z = np.zeros((2, 3))
    s.gripper.step(z, z, np.array([-2e-4, -2e-4]))
    s.gripper.update_bound(s)
    stats = [s.time_step() for _ in range(2)]
    s.save_all(path)
    return stats


def _rollout(s, tr, state, grad=None):
    T = tr.shape[0]
    agent = agent_trajopt(T, 2, max_moving_dist=0.001)
    agent.traj.from_numpy(tr)
    s.reset()
    s.mu_cloth_elastic[None] = 5.0
    s.load_all(state)                                                  # (the driver: sys.load_all(state_path) at the start of every rollout)
    buf = [s.engine.pos.clone()]
    if grad is not None:
        grad.reset()
        grad.copy_pos(s, 0)
    stats = []
    for f in range(1, T):
        agent.get_action(f)
        s.action(f, agent.delta_pos, agent.delta_rot)
        stats.append(s.time_step(tol=1e-9))        # (tight: the finite differences below subtract two rollouts)
        buf.append(s.engine.pos.clone())
        if grad is not None:
            grad.copy_pos(s, f)
    return torch.stack(buf), stats


def test_balancing_scene_state_rollout_state_files_and_adjoint(golden_dir, tmp_path):
    import scipy.sparse.linalg as spla
    g = np.load(os.path.join(golden_dir, "scene_state_balancing.npz"))
    s = Scene(cloth_size=0.06)
    e = s.engine
    assert np.array_equal(e.pos.cpu().numpy(), g["pos0"]) and np.array_equal(e.frozen.cpu().numpy(), g["frozen"])
    assert np.array_equal(s.faces, g["faces"]) and np.abs(e.mass.cpu().numpy() - g["mass"]).max() <= 1e-14 * g["mass"].max()
    assert s.enable_gripper and s.gripper.n_part == 2 and s.elastic_cnt == 5 and s.elastics[0].n_verts == 100
    state = str(tmp_path / "balance_state")
    gstats = _grip(s, state)
    assert all(st.converged for st in gstats) and gstats[-1].n_contacts >= 8, gstats     # both ends pinched between two pads, ball on top
    assert np.array_equal(s.gripper._half, [-2e-4, -2e-4])
    T = 4
    tr = _traj(T)
    grad = Grad(s, T, 2)
    buf, stats = _rollout(s, tr, state, grad)
    loss = float((s._ball_minus_centre(buf) ** 2).sum().item())
    for f, st in enumerate(stats, 1):
        assert st.converged, (f, st)
    assert stats[-1].n_contacts >= 8
    assert abs(s.compute_reward_all(grad) + loss) < 1e-12 * max(1.0, loss) and abs(s.compute_reward() + float((s._ball_minus_centre(buf[-1]) ** 2).sum())) < 1e-15
    # fixed point of the reference iteration at the last state
    vel1 = e.vel.clone()
    e.vel.copy_((grad._pos_buffer[T - 2] - grad._pos_buffer[T - 3]) / s.dt)
    e.prev_pos.copy_(grad._pos_buffer[T - 2])
    e.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_SPD | _lib.ASM_F64)
    p = spla.spsolve(e.matrix().tocsc(), e.residual())
    delta = np.abs(p).max() / s.dt
    print(f"Scene_balancing: steps {[(st.newton_iters, st.linear_iters, st.n_contacts) for st in stats]}, reference Newton step at the last state {delta:.2e}")
    assert delta < 1e-6
    e.vel.copy_(vel1)
    # save_all / load_all: the files of the reference (gripper fields, state, flags), round trip
    x_end, v_end, gp, gr = e.pos.clone(), e.vel.clone(), s.gripper._pos.copy(), s.gripper._rot.copy()
    d = str(tmp_path / "balance_state")
    s.save_all(d)
    for name in ("F_x_upper", "F_x_upper_world", "F_x_lower", "F_x_lower_world", "pos", "rot", "rotmat", "half_gripper_dist", "proj_flag",
                 "proj_dir", "border_flag"):
        assert os.path.exists(os.path.join(d, name + ".npy")), name
    assert np.load(os.path.join(d, "F_x_upper.npy")).shape == (2, 276, 3)
    # (F_x_*_world of the driven vertices are where the engine holds them)
    w = np.load(os.path.join(d, "F_x_lower_world.npy"))
    bi = s.gripper._bound_idx.cpu().numpy()
    assert np.abs(w[1][bi] - x_end[s.gripper.lower_offsets[1] + torch.from_numpy(bi).long().to(x_end.device)].cpu().numpy()).max() < 1e-9
    assert np.array_equal(np.load(os.path.join(d, "half_gripper_dist.npy")), [-2e-4, -2e-4])
    s.reset()
    assert np.array_equal(s.gripper._half, [0.0, 0.0])
    s.load_all(d)
    assert torch.equal(e.pos, x_end) and torch.equal(e.vel, v_end) and np.array_equal(s.gripper._pos, gp) and np.array_equal(s.gripper._rot, gr)
    # adjoint of get_loss_balance
    _rollout(s, tr, state, grad)
    grad.get_loss_balance(s)
    # The seeds are NOT the gradient of the reward: the entries of the centre vertex are plain assignments inside a loop over the ball's
    # vertices (analytic_grad_single.py:428-444), so only the last vertex's term survives.  What the adjoint differentiates is the
    # linear functional <seeds, x(pose)>; the finite differences below take exactly that.
    seeds = grad._pos_grad.clone()
    b = s.elastics[0]
    assert torch.equal(seeds[1:, b.offset:b.offset + b.n_verts, :2], 2 * s._ball_minus_centre(grad._pos_buffer)[1:]) and not seeds[0].any()
    assert torch.equal(seeds[1:, s._centre(), :2], -seeds[1:, b.offset + b.n_verts - 1, :2])
    lin = lambda x: float((seeds * x).sum().item())
    # (1) exact: the gripper gradient of the last step is -z^T dF/db on the contact set of the step, z the solution of the adjoint system
    #     -- evaluated here with SciPy and differences of the engine's RESIDUAL under a common shift of the driven vertices of BOTH pads
    #     of a part (this is what pins gripper_tactile.gather_grad: the mean over the 2 n_bound driven vertices)
    import scipy.sparse.linalg as spla
    x = grad._pos_buffer[T - 1].clone()
    free = ~e.frozen.cpu().numpy().astype(bool)
    it, flags, rr = grad.transfer_grad(T - 1, s)
    assert flags == 0 and it == 0 and rr < 1e-9, (it, flags, rr)                     # 4 k unknowns: the dense-LU path
    gg1 = grad._gripper_grad[T - 1].copy()
    e.pos.copy_(x); e.prev_pos.copy_(grad._pos_buffer[T - 2])
    e.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_F64)
    H = e.matrix().tocsr()
    z = np.zeros(3 * s.tot_NV)
    z[free] = spla.spsolve(H[free][:, free].T.tocsc(), seeds[T - 1].cpu().numpy().reshape(-1)[free])
    assert np.abs(grad._z.cpu().numpy().reshape(-1)[free] - z[free]).max() <= 1e-4 * np.abs(z).max()
    bi = s.gripper._bound_idx.cpu().numpy()
    for part in range(2):
        bound = np.concatenate([bi + s.gripper.upper_offsets[part], bi + s.gripper.lower_offsets[part]])
        for comp in (0, 2):
            F = []
            for sgn in (1.0, -1.0):
                d = np.zeros((s.tot_NV, 3)); d[bound, comp] = sgn * 1e-7
                e.pos.copy_(x + torch.from_numpy(d).to(e.device))
                e.assemble(_lib.ASM_RESIDUAL)
                F.append(e.residual())
            ref = -(z[free] * ((F[0] - F[1]) / 2e-7)[free]).sum()
            an = gg1[part, comp] * 2 * s.gripper.n_bound
            print(f"Scene_balancing dL/dpose[{T - 1}, part {part}, {comp}]: engine x 2 n_bound {an:.8e}   -z^T dF/db (SciPy + residual differences) {ref:.8e}")
            assert abs(an - ref) <= 1e-4 * max(abs(ref), np.abs(gg1[part, :3]).max() * 2 * s.gripper.n_bound), (part, comp, an, ref)
    e.pos.copy_(x)
    for j in range(T - 2, 0, -1):
        it, flags, rr = grad.transfer_grad(j, s)
        assert flags == 0 and it == 0 and rr < 1e-9, (j, it, flags, rr)
    gg = grad._gripper_grad.copy()
    assert np.isfinite(gg).all() and np.abs(gg[1:]).max() > 0
    # (2) against finite differences of the rollout: same sign and size.  Not digit by digit: the adjoint matrix is the reference's, which
    #     leaves out how the contact normals turn with the surface they belong to -- exactly the effect that makes the ball roll when the
    #     strip tilts (measured: -2.74e-3 against -2.24e-3).
    for (part, comp, h) in ((0, 2, 1e-5),):
        tp, tm = tr.copy(), tr.copy()
        tp[T - 1, part, comp] += h; tm[T - 1, part, comp] -= h
        fd = (lin(_rollout(s, tp, state)[0]) - lin(_rollout(s, tm, state)[0])) / (2 * h)
        an = gg[T - 1, part, comp] * 2 * s.gripper.n_bound
        print(f"Scene_balancing dL/dpose[{T - 1}, part {part}, {comp}]: adjoint x 2 n_bound {an:.6e}  finite difference of the rollout {fd:.6e}")
        assert np.sign(an) == np.sign(fd) and 0.5 < an / fd < 2.0, (part, comp, an, fd)
