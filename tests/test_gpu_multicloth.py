"""Several cloths in one system (Scene_card, Scene_sliding: code/task_scene/Scene_card.py, Scene_sliding.py) on the GPU.

* Engine level: two separated cloths in one context give, cloth by cloth, the energy, residual, fp64 reference matrix, Newton matrices
  and dF/dKb of the same cloths in contexts of their own (additivity: the property that does not need an oracle).
* Scene_card: the product-built scene equals the reference-made state (three cards, two turned pads); rollout of the driver's opening
  trajectory (init_traj_card) with cloth-cloth contacts; last step a fixed point of the REFERENCE iteration; trajectory adjoint across
  the three cloths against finite differences; the Kb identification recurrence of trajopt_card.py runs.
* Scene_sliding: reference-made state; press-and-drag rollout; the friction-coefficient gradient of trajopt_silding.py
  (count_friction_grad, cloth-cloth constraints only) against a finite difference in mu_cloth_cloth."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from thinshelllab_b200 import _lib
    from thinshelllab_b200.agent.traj_opt_single import agent_trajopt
    from thinshelllab_b200.core import ShellEngine
    from thinshelllab_b200.engine.analytic_grad_single import Grad
    from thinshelllab_b200.engine.analytic_grad_system import Grad as GradSystem
    from thinshelllab_b200.task_scene.Scene_card import Scene as SceneCard
    from thinshelllab_b200.task_scene.Scene_sliding import Scene as SceneSliding


def _bent(N, M, dx, origin, seed):
    rng = np.random.default_rng(seed)
    i, j = np.meshgrid(np.arange(N + 1), np.arange(M + 1), indexing="ij")
    x = np.stack([i * dx, j * dx, 0.3 * dx * np.sin(0.9 * i) * np.cos(0.7 * j)], -1).reshape(-1, 3) + np.asarray(origin)
    return x + 0.02 * dx * rng.standard_normal(x.shape)


def _engine(shapes, xs, kbs):
    nv = sum((N + 1) * (M + 1) for N, M in shapes)
    e = ShellEngine(nv, 5e-3, k_contact=10000.0, eps_contact=4e-4, gravity=(0.0, 0.0, -9.8))
    off = 0
    for (N, M), kb in zip(shapes, kbs):
        e.add_cloth(N, M, off, 0.004, 40.0, Kb=kb, k_angle=3.14)
        off += (N + 1) * (M + 1)
    e.finalize()
    x = torch.from_numpy(np.concatenate(xs)).to(e.device)
    e.pos.copy_(x); e.prev_pos.copy_(x - 1e-5); e.vel.zero_()
    return e


def test_two_cloths_are_the_sum_of_their_parts():
    shapes = [(6, 4), (5, 5)]
    xs = [_bent(6, 4, 0.004, (0.0, 0.0, 0.0), 1), _bent(5, 5, 0.004, (0.2, 0.1, 0.05), 2)]
    kbs = [100.0, 250.0]
    both = _engine(shapes, xs, kbs)
    parts = [_engine([s], [x], [kb]) for s, x, kb in zip(shapes, xs, kbs)]
    assert abs(both.energy() - sum(p.energy() for p in parts)) <= 1e-13 * abs(both.energy())
    nv0 = 35
    for flags, tol in ((_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_F64, 1e-12), (_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_SPD | _lib.ASM_F64, 1e-12)):
        both.assemble(flags)
        F, A = both.residual(), both.matrix().toarray()
        for k, p in enumerate(parts):
            p.assemble(flags)
            sl = slice(0, 3 * nv0) if k == 0 else slice(3 * nv0, None)
            Fp, Ap = p.residual(), p.matrix().toarray()
            assert np.abs(F[sl] - Fp).max() <= tol * np.abs(Fp).max()
            assert np.abs(A[sl, sl] - Ap).max() <= tol * np.abs(Ap).max()
        assert not A[:3 * nv0, 3 * nv0:].any() and not A[3 * nv0:, :3 * nv0].any()
    for k, p in enumerate(parts):
        d2 = both.cloth_param_deri(k, kl=False, ka=False)[2].cpu().numpy()
        d1 = p.cloth_param_deri(0, kl=False, ka=False)[2].cpu().numpy()
        sl = slice(0, nv0) if k == 0 else slice(nv0, None)
        assert np.abs(d2[sl] - d1).max() <= 1e-12 * np.abs(d1).max()
        other = slice(nv0, None) if k == 0 else slice(0, nv0)
        assert not d2[other].any()
    # one implicit step of both cloths together: every cloth ends at a stationary point of ITS OWN system (crumpled sheets relaxing: the
    # minimisation is not convex, so the joint and the separate Newton paths may pick different minima -- compare residuals, not positions)
    import scipy.sparse.linalg as spla
    x0, v0 = both.prev_pos.clone(), both.vel.clone()
    st = both.step_forward(200, 1e-8)
    assert st.converged
    for k, p in enumerate(parts):
        sl = slice(0, nv0) if k == 0 else slice(nv0, None)
        p.pos.copy_(both.pos[sl]); p.prev_pos.copy_(x0[sl] + 1e-5); p.vel.copy_(v0[sl])
        p.prev_pos.copy_(both.prev_pos[sl])
        p.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_SPD | _lib.ASM_F64)
        step = spla.spsolve(p.matrix().tocsc(), p.residual())
        assert np.abs(step).max() / 5e-3 < 1e-6, (k, np.abs(step).max())


def _fixed_point_delta(s, grad, T):
    import scipy.sparse.linalg as spla
    e = s.engine
    vel1 = e.vel.clone()
    e.vel.copy_((grad._pos_buffer[T - 2] - grad._pos_buffer[T - 3]) / s.dt * s.damping)
    e.prev_pos.copy_(grad._pos_buffer[T - 2])
    e.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_SPD | _lib.ASM_F64)
    p = spla.spsolve(e.matrix().tocsc(), e.residual())
    e.vel.copy_(vel1)
    return np.abs(p).max() / s.dt


def _pairs_in_contact(s, n):
    idx = s.engine.constraints()["idx"][:n]
    body = np.searchsorted([b.v_start for b in s.body_list], idx, side="right") - 1
    return {tuple(sorted(set(int(v) for v in r))) for r in body}


def _card_rollout(s, tr, grad=None):
    T = tr.shape[0]
    agent = agent_trajopt(T, 3, max_moving_dist=0.001)
    agent.traj.from_numpy(tr)
    s.reset()
    s.mu_cloth_elastic[None] = 1.0
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
    c = s.cloths[0]
    return float(s.engine.pos[c.offset:c.offset + c.NV, 0].sum().item()), stats      # = -compute_reward(): what get_loss_slide_simple seeds


def test_card_scene_three_cloths(golden_dir):
    g = np.load(os.path.join(golden_dir, "scene_state_card.npz"))
    s = SceneCard(cloth_size=0.06)
    e = s.engine
    assert np.array_equal(e.pos.cpu().numpy(), g["pos0"]) and np.array_equal(e.frozen.cpu().numpy(), g["frozen"])
    assert np.array_equal(s.faces, g["faces"]) and np.abs(e.mass.cpu().numpy() - g["mass"]).max() <= 1e-14 * g["mass"].max()
    assert s.cloth_cnt == 3 and s.gripper.n_part == 3 and s.damping == 0.95
    s.cloths[0].Kb[None] = 1000.0
    T = 7
    agent = agent_trajopt(T, 3, max_moving_dist=0.001)
    agent.init_traj_card()
    tr = agent.traj.to_numpy()
    assert np.allclose(tr[:5, 0, 0], 0.0003 * np.arange(1, 6)) and np.allclose(tr[:5, 1, 0], -0.0003 * np.arange(1, 6)) and np.allclose(tr[5:, 0, 2], [0.0003, 0.0006])
    grad = Grad(s, T, 3)
    loss, stats = _card_rollout(s, tr, grad)
    for f, st in enumerate(stats, 1):
        assert st.converged, (f, st)
    pairs = _pairs_in_contact(s, stats[-1].n_contacts)
    delta = _fixed_point_delta(s, grad, T)
    print(f"Scene_card: steps {[(st.newton_iters, st.linear_iters, st.n_contacts) for st in stats]}, body pairs in contact {sorted(pairs)}, "
          f"reference Newton step at the last state {delta:.2e}")
    assert (0, 1) in pairs and (1, 2) in pairs                           # card on card
    assert any(p in pairs for p in ((0, 4), (0, 5), (1, 4), (1, 5), (2, 4), (2, 5)))     # an end pad on a card
    assert delta < 1e-6
    # ---- trajectory adjoint across three cloths (dense-LU path), seeds of get_loss_slide_simple.
    # The cards start exactly eps_contact apart and float without gravity: card-card constraints appear and vanish with rounding, the
    # rollout is not a smooth function of the trajectory at the 1e-6 m level and finite differences of it are noise (measured: +29, -21,
    # +4.9 for h = 2e-7, 2e-6, 2e-5 where the adjoint says 6.6).  What can be checked exactly is the adjoint's own definition on the
    # contact set of the step: (a) its matrix is the Jacobian of the engine's residual, (b) z solves it, (c) the gripper gradient is
    # -z^T dF/db with dF/db by finite differences of the RESIDUAL with respect to the driven vertices.
    import scipy.sparse.linalg as spla
    x = grad._pos_buffer[T - 1].clone()
    fz = e.frozen.cpu().numpy().astype(bool)
    free = ~fz

    def residual_at(dx=None):
        e.pos.copy_(x if dx is None else x + torch.from_numpy(dx).to(e.device))
        e.prev_pos.copy_(grad._pos_buffer[T - 2])
        e.assemble(_lib.ASM_RESIDUAL)
        return e.residual()
    e.pos.copy_(x); e.prev_pos.copy_(grad._pos_buffer[T - 2])
    e.vel.copy_((grad._pos_buffer[T - 2] - grad._pos_buffer[T - 3]) / s.dt * s.damping)
    e.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_F64)
    H = e.matrix().tocsr()
    rng = np.random.default_rng(0)
    for k in range(3):                                                     # (a) on the three cloths (contacts between them included)
        b = s.body_list[k]
        d = np.zeros((s.tot_NV, 3)); d[b.v_start:b.v_end] = 1e-7 * rng.standard_normal((b.v_end - b.v_start, 3))
        jd = (residual_at(d) - residual_at(-d)) / 2
        hd = H @ d.reshape(-1)
        assert np.abs(hd[free] - jd[free]).max() <= 0.01 * np.abs(jd[free]).max(), (k, np.abs(hd[free] - jd[free]).max(), np.abs(jd[free]).max())
    grad.get_loss_slide_simple(s)
    seeds = grad._pos_grad[T - 1].cpu().numpy().reshape(-1).copy()
    it, flags, rr = grad.transfer_grad(T - 1, s)
    assert flags == 0 and it == 0 and rr < 1e-9, (it, flags, rr)
    z = np.zeros(3 * s.tot_NV)
    z[free] = spla.spsolve(H[free][:, free].T.tocsc(), seeds[free])
    z_eng = grad._z.cpu().numpy().reshape(-1)
    assert np.abs(z_eng[free] - z[free]).max() <= 1e-4 * np.abs(z).max()      # (b)
    gg1 = grad._gripper_grad[T - 1].copy()
    for part in range(3):                                                  # (c): every part (two of them turned by 90 degrees), x and z
        bound = s.gripper._bound_idx.cpu().numpy() + s.gripper.part_offsets[part]
        for comp in (0, 2):
            d = np.zeros((s.tot_NV, 3)); d[bound, comp] = 1e-7
            dFdb = (residual_at(d) - residual_at(-d)) / 2e-7
            ref = -(z[free] * dFdb[free]).sum()
            an = gg1[part, comp] * s.gripper.n_bound
            print(f"Scene_card dL/dpose[{T - 1}, part {part}, {comp}]: engine x n_bound {an:.8e}   -z^T dF/db (SciPy + residual differences) {ref:.8e}")
            assert abs(an - ref) <= 1e-4 * max(abs(ref), abs(gg1[part, :3]).max() * s.gripper.n_bound), (part, comp, an, ref)
    e.pos.copy_(x)
    for j in range(T - 2, 0, -1):
        it, flags, rr = grad.transfer_grad(j, s)
        assert flags == 0 and it == 0 and rr < 1e-9, (j, it, flags, rr)
    gg = grad._gripper_grad.copy()
    assert np.isfinite(gg).all() and np.abs(gg[1:]).max() > 0
    # the Kb recurrence of trajopt_card.py (system-identification Grad over a scene with several cloths)
    gs = GradSystem(s, T, 3)
    s.reset(); s.mu_cloth_elastic[None] = 1.0
    gs.copy_pos(s, 0)
    for f in range(1, T):
        agent.get_action(f)
        s.action(f, agent.delta_pos, agent.delta_rot)
        s.time_step()
        gs.copy_pos(s, f)
    gs.get_loss_card(s)
    for j in range(T - 1, 0, -1):
        it, flags, rr = gs.transfer_grad(j, s)
        assert flags == 0 and rr < 1e-9, (j, it, flags, rr)
    print(f"Scene_card grad_kb {gs.grad_kb[None]:.6e}")
    assert np.isfinite(gs.grad_kb[None])


def _slide_traj(T):
    tr = np.zeros((T, 1, 6))
    for i in range(1, T):
        tr[i, 0, 2] = -0.0005 * min(i, 5)                                 # reach and press the stack (the pad starts 1.9 mm above it)
        tr[i, 0, 0] = -0.0004 * max(i - 5, 0)                             # then drag along -x
    return tr


def _slide_rollout(s, tr, mu_cc, grad=None):
    T = tr.shape[0]
    agent = agent_trajopt(T, 1, max_moving_dist=0.001)
    agent.traj.from_numpy(tr)
    s.reset()
    s.mu_cloth_elastic[None] = 1.0
    s.mu_cloth_cloth[None] = mu_cc
    if grad is not None:
        grad.reset()
        grad.copy_pos(s, 0)
    stats, loss = [], 0.0
    c = s.cloths[0]
    for f in range(1, T):
        agent.get_action(f)
        s.action(f, agent.delta_pos, agent.delta_rot)
        stats.append(s.time_step())
        loss += float(s.engine.pos[c.offset:c.offset + c.NV, 0].sum().item())          # get_loss_slide: +1 on cloth 0's x, frames 1..
        if grad is not None:
            grad.copy_pos(s, f)
    return loss, stats


def test_sliding_scene_friction_coefficient_gradient(golden_dir):
    g = np.load(os.path.join(golden_dir, "scene_state_sliding.npz"))
    s = SceneSliding(cloth_size=0.06)
    e = s.engine
    assert np.array_equal(e.pos.cpu().numpy(), g["pos0"]) and np.array_equal(e.frozen.cpu().numpy(), g["frozen"])
    assert np.array_equal(s.faces, g["faces"]) and np.abs(e.mass.cpu().numpy() - g["mass"]).max() <= 1e-14 * g["mass"].max()
    assert s.cloth_cnt == 3 and s.n_cloth_cloth_pairs == 8 and s.elastics[1].mu[None] == float(g["el1_mu"])
    s.cloths[0].Kb[None] = 1000.0
    T = 9
    tr = _slide_traj(T)
    gs = GradSystem(s, T, 1)
    gs.count_friction_grad, gs.count_kb_grad = True, False
    mu0 = 0.5
    loss, stats = _slide_rollout(s, tr, mu0, gs)
    for f, st in enumerate(stats, 1):
        assert st.converged, (f, st)
    pairs = _pairs_in_contact(s, stats[-1].n_contacts)
    delta = _fixed_point_delta(s, gs, T)
    print(f"Scene_sliding: steps {[(st.newton_iters, st.linear_iters, st.n_contacts) for st in stats]}, body pairs in contact {sorted(pairs)}, "
          f"reference Newton step at the last state {delta:.2e}")
    assert (0, 1) in pairs and (1, 2) in pairs and (0, 3) in pairs and (2, 4) in pairs    # cloth/cloth, bottom cloth/table, top cloth/pad
    assert delta < 1e-6
    # ---- the friction-coefficient gradient.  As in Scene_card the stacked cloths start exactly eps_contact apart, so finite differences of
    # the ROLLOUT in mu are dominated by constraints flickering on and off; the step-wise definition is checked instead:
    # contribution of step t = -z^T dF/dmu on the contact set of the step (Scene_sliding.contact_energy_backprop_friction: the friction
    # stiffness k = mu x pressure is linear in mu), with dF/dmu from differences of the engine's RESIDUAL after re-detecting with mu +- h.
    import scipy.sparse.linalg as spla
    gs.get_loss_slide(s)
    x_t, x_tm1 = gs._pos_buffer[T - 1].clone(), gs._pos_buffer[T - 2].clone()
    it, flags, rr = gs.transfer_grad(T - 1, s)
    assert flags == 0 and rr < 1e-9, (it, flags, rr)
    an1 = gs.grad_friction_coef[None]
    assert gs.grad_kb[None] == 0.0                                        # count_friction_grad replaces the stiffness gradients
    z = gs._z.cpu().numpy().reshape(-1)
    free = ~e.frozen.cpu().numpy().astype(bool)

    def residual_with(mu):
        s.mu_cloth_cloth[None] = mu
        e.pos.copy_(x_tm1); e.prev_pos.copy_(x_tm1)
        e.contact_detect()                                                # where transfer_grad detects (copy_pos_only(step - 1): quirk Q7)
        e.pos.copy_(x_t)
        e.assemble(_lib.ASM_RESIDUAL)
        return e.residual()
    h = 1e-4
    dFdmu = (residual_with(mu0 + h) - residual_with(mu0 - h)) / (2 * h)
    s.mu_cloth_cloth[None] = mu0
    ref = -(z[free] * dFdmu[free]).sum()
    print(f"Scene_sliding d/dmu_cloth_cloth, step {T - 1}: engine {an1:.8e}   -z^T dF/dmu (residual differences) {ref:.8e}")
    assert an1 != 0.0 and abs(an1 - ref) <= 1e-5 * abs(ref), (an1, ref)
    for j in range(T - 2, 0, -1):
        it, flags, rr = gs.transfer_grad(j, s)
        assert flags == 0 and rr < 1e-9, (j, it, flags, rr)
    an = gs.grad_friction_coef[None]
    assert np.isfinite(an) and an != an1
    fd = (_slide_rollout(s, tr, mu0 + 1e-2)[0] - _slide_rollout(s, tr, mu0 - 1e-2)[0]) / 2e-2
    print(f"Scene_sliding dL/dmu_cloth_cloth over the rollout: adjoint {an:.6e}  finite difference of the rollout (h = 0.01, informational) {fd:.6e}")
