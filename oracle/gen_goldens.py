#!/usr/bin/env python
"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN SOURCES under the taichi
stand-in (oracle tier 1).  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):

    python oracle/gen_goldens.py cloth      # term-by-term cloth goldens      (~1 min)
    python oracle/gen_goldens.py spd        # SPD_Projector known answers     (seconds)
    python oracle/gen_goldens.py bouncing   # Scene_bouncing rollout + adjoint (~30-60 min)

Outputs: tests/golden/*.npz (committed; the GPU box has no /root/reference).
Ground truth statement: these are outputs of the reference's Python source executed by
a serial Taichi emulation with SciPy SuperLU in place of CuPy/cuSOLVER spsolve; the
Taichi JIT itself cannot be installed in this image.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "ti_emu"))
import hook  # noqa: E402

hook.install()
os.chdir(hook.REF_CODE)  # the reference opens ../data/* relative to code/
import taichi as ti  # noqa: E402

ti.init(ti.cpu, default_fp=ti.f64, default_ip=ti.i32, fast_math=False)
OUT = os.path.join(HERE, "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)


class _Rec:
    """stands in for BaseScene in Cloth.compute_Hessian_*: records H.add / add_H into a dense matrix"""

    def __init__(self, n):
        self.M = np.zeros((n, n))
        self.H = self

    def add(self, i, j, v):
        self.M[int(i), int(j)] += v

    def add_H(self, i, j, v):
        self.M[int(i), int(j)] += v


def cloth_case(name, N, M, square, Len, mode, seed, kb=100.0, k_angle=3.14):
    from thinshelllab.engine.model_fold_offset import Cloth
    rng = np.random.default_rng(seed)
    dt = 5e-3
    if square:
        c = Cloth(N, dt, Len, 0, 40.0, 0)
    else:
        c = Cloth(N, dt, Len, 0, 40.0, 0, False, M)
    c.Kb[None] = kb
    c.k_angle[None] = k_angle
    if mode == "fold":
        c.init_fold(-0.07, -0.01, 0.0004, 2)
    else:
        c.init(-0.5 * Len, -0.5 * Len * c.M / c.N, 0.002)
    pos0 = c.pos.to_numpy()
    dx = c.dx
    ii = np.arange(c.NV) // (c.M + 1)
    jj = np.arange(c.NV) % (c.M + 1)
    pos = pos0.copy()
    if mode == "flat":
        pos += rng.uniform(-1e-5 * dx, 1e-5 * dx, pos.shape)
    elif mode == "wavy":
        pos[:, 2] += 0.6 * dx * np.sin(2 * np.pi * ii / 5.0) * np.cos(2 * np.pi * jj / 4.0)
        pos += rng.uniform(-0.04 * dx, 0.04 * dx, pos.shape)
    elif mode == "fold":
        pos += rng.uniform(-0.02 * dx, 0.02 * dx, pos.shape)
    c.pos.from_numpy(pos)
    prev = pos + rng.uniform(-0.01 * dx, 0.01 * dx, pos.shape)
    vel = rng.uniform(-0.05, 0.05, pos.shape)
    c.prev_pos.from_numpy(prev)
    c.vel.from_numpy(vel)
    ref0 = c.ref_angle.to_numpy()
    ref = ref0 + rng.uniform(-0.05, 0.05, ref0.shape)
    c.ref_angle.from_numpy(ref)
    out = dict(N=c.N, M=c.M, dx=dx, dt=dt, rho=40.0, Kl=c.Kl[None], Ka=c.Ka[None], Kb=c.Kb[None],
               k_angle=c.k_angle[None], mass=c.mass, pos=pos, prev_pos=prev, vel=vel, ref_angle=ref,
               f2v=c.f2v.to_numpy(), counter_face=c.counter_face.to_numpy(),
               counter_point=c.counter_point.to_numpy(), V=c.V.to_numpy(), l_i=c.l_i.to_numpy())
    c.compute_normal_dir()
    c.prepare_bending()
    for k in ("norm_dir", "mat_M", "mat_N", "angle", "heights", "c_i", "d_i"):
        out[k] = getattr(c, k).to_numpy()
    c.compute_energy(); out["U"] = c.U[None]
    c.compute_energy_me(); out["U_me"] = c.U[None]
    c.compute_energy_ma(); out["U_ma"] = c.U[None]
    c.compute_energy_bending(); out["U_bending"] = c.U[None]
    c.compute_residual(); out["F_b"] = c.F_b.to_numpy()
    c.compute_residual_me(); out["F_me"] = c.F_b.to_numpy()
    c.compute_residual_ma(); out["F_ma"] = c.F_b.to_numpy()
    c.compute_residual_bending(); out["F_bending"] = c.F_b.to_numpy()
    n = 3 * c.NV
    for spd in (0, 1):
        r = _Rec(n); c.compute_Hessian_me(r, spd); out[f"H_me_spd{spd}"] = r.M
    r = _Rec(n); c.compute_Hessian_ma(r); out["H_ma"] = r.M
    r = _Rec(n); c.compute_Hessian_bending(r); out["H_bending"] = r.M
    c.compute_deri()
    out["d_kl"] = c.d_kl.to_numpy(); out["d_ka"] = c.d_ka.to_numpy(); out["d_kb"] = c.d_kb.to_numpy()
    c.k_angle[None] = 0.02  # force plastic flow on some hinges
    c.update_ref_angle()
    out["ref_angle_after_k0p02"] = c.ref_angle.to_numpy()
    np.savez_compressed(os.path.join(OUT, f"cloth_{name}.npz"), **out)
    print("wrote", name, "U", out["U"])


def gen_cloth():
    cloth_case("6x4_wavy", 6, 4, False, 0.03, "wavy", 1)
    cloth_case("6x4_flat", 6, 4, False, 0.03, "flat", 2)
    cloth_case("8x8_wavy", 8, 8, True, 0.04, "wavy", 3, kb=400.0)
    cloth_case("15x3_fold", 15, 3, False, 0.1, "fold", 4, kb=400.0, k_angle=0.5)


def gen_spd():
    from thinshelllab.engine import linalg
    rng = np.random.default_rng(7)
    out = {}
    for D, K, cnt in ((3, 10, 40), (9, 20, 24)):
        p = linalg.SPD_Projector(cnt, D, K)
        A = ti.field(ti.f64, (cnt, D, D))
        a = rng.standard_normal((cnt, D, D))
        a = a + a.transpose(0, 2, 1)
        # vary scale / rank structure like the physics blocks do
        for t in range(cnt):
            if t % 4 == 1:
                v = rng.standard_normal(D); a[t] = 5e5 * np.outer(v, v) + 1e2 * a[t]
            if t % 4 == 2:
                a[t] *= 1e-3
            if t % 4 == 3:
                a[t] = a[t] @ a[t].T * 100.0 - 50.0 * np.eye(D)
        A.from_numpy(a)
        for t in range(cnt):
            p.project(A, t, D)
        out[f"in{D}"] = a
        out[f"out{D}"] = A.to_numpy()
    # 2x2 friction projector
    a2 = rng.standard_normal((32, 2, 2)); a2 = a2 + a2.transpose(0, 2, 1)
    o2 = np.stack([np.asarray(linalg.SPD_project_2d(ti.Matrix(a2[t]))) for t in range(32)])
    out["in2"], out["out2"] = a2, o2
    np.savez_compressed(os.path.join(OUT, "spd_projector.npz"), **out)
    print("wrote spd_projector")


def tet_case(name, kind, seed):
    """term-by-term goldens of a tet body: model_elastic_tactile.Elastic (kind 'tactile') or
    model_elastic_offset.Elastic (kind 'box')"""
    rng = np.random.default_rng(seed)
    dt = 5e-3
    if kind == "tactile":
        from thinshelllab.engine.model_elastic_tactile import Elastic
        e = Elastic(dt, 0, 0.5)
        e.init(0.01, -0.002, 0.03, True)
        par = dict(mu=e.mu[None], lam=e.lam[None], alpha=e.alpha[None])
        amp = 0.03
    else:
        from thinshelllab.engine.model_elastic_offset import Elastic
        e = Elastic(dt, 0.03, 0, 4, 3, 3)
        e.init(-0.01, 0.0, 0.002)
        par = dict(mu=e.mu[None], lam=e.lam[None], alpha=0.0)
        e.lam[None] = 1.5e5    # the reference default nu = 0 gives lam = 0: exercise the lam terms too
        par["lam"] = 1.5e5
        amp = 0.06
    e.gravity[None] = ti.Vector([0.3, -0.2, -9.8])
    rest = e.F_x.to_numpy()
    tets = e.F_vertices.to_numpy()
    edge = np.linalg.norm(rest[tets[:, 0]] - rest[tets[:, 1]], axis=1).mean()
    pos = rest + rng.uniform(-amp * edge, amp * edge, rest.shape)
    prev = pos + rng.uniform(-0.01 * edge, 0.01 * edge, rest.shape)
    vel = rng.uniform(-0.05, 0.05, rest.shape)
    ext = rng.uniform(-1e-3, 1e-3, rest.shape)
    e.F_x.from_numpy(pos); e.F_x_prev.from_numpy(prev); e.F_v.from_numpy(vel); e.ext_force.from_numpy(ext)
    out = dict(kind=kind, dt=dt, rest=rest, tets=tets, f2v=e.f2v.to_numpy(), F_B=e.F_B.to_numpy(), F_W=e.F_W.to_numpy(),
               F_m=e.F_m.to_numpy(), pos=pos, prev_pos=prev, vel=vel, ext_force=ext, gravity=e.gravity.to_numpy(),
               density=e.density, **par)
    if kind == "tactile":
        out["F_ox"] = e.F_ox.to_numpy(); out["ratio"] = e.ratio; out["is_surface"] = e.is_surface.to_numpy()
    e.compute_energy(); out["U"] = e.U[None]
    e.get_force(); out["F_f"] = e.F_f.to_numpy()
    e.compute_residual(); out["F_b"] = e.F_b.to_numpy()
    n = 3 * e.n_verts
    for spd in (0, 1):
        r = _Rec(n); e.compute_Hessian(r, spd); out[f"H_spd{spd}"] = r.M
    e.d_mu.fill(0); e.d_lam.fill(0)
    e.compute_deri()
    out["d_mu"] = e.d_mu.to_numpy(); out["d_lam"] = e.d_lam.to_numpy()
    np.savez_compressed(os.path.join(OUT, f"tet_{name}.npz"), **out)
    print("wrote tet", name, "U", out["U"], "nverts", e.n_verts, "ncells", e.n_cells, flush=True)


def gen_tets():
    tet_case("box_4x3x3", "box", 21)
    tet_case("tactile", "tactile", 22)


def _csr_of(sysm):
    """dense-backed SparseMatrix -> scipy CSR (code/engine/sparse_solver.py:13-17)"""
    import scipy.sparse as sp
    n = sysm.n
    val = sysm.value.to_numpy()
    rc = sysm.row_cnt.to_numpy()
    ri = sysm.row_idx.to_numpy()
    rows, cols = [], []
    for i in range(n):
        for j in ri[i, :rc[i]]:
            rows.append(i); cols.append(int(j))
    rows = np.array(rows); cols = np.array(cols)
    return sp.csr_matrix((val[rows, cols], (rows, cols)), shape=(n, n))


def gen_bouncing(T=4, use_reset=False, tag="bouncing"):
    import scipy.sparse as sp
    from thinshelllab.task_scene.Scene_bouncing import Scene
    from thinshelllab.engine.geometry import projection_query
    from thinshelllab.engine.analytic_grad_system import Grad
    from cupyx.scipy.sparse import linalg as fake_linalg
    s = Scene(cloth_size=0.06)
    s.device = "cpu"; s.H.device = "cpu"
    s.cloths[0].Kb[None] = 120.0
    g = Grad(s, T, s.elastic_cnt - 1)
    s.init_all()
    g.init_mass(s)
    if use_reset:
        s.reset()
    s.mu_cloth_elastic[None] = 0.5
    # deterministic perturbation so that every term is active from frame 1
    rng = np.random.default_rng(11)
    NVc = s.cloths[0].NV
    pos = s.pos.to_numpy()
    dx = s.cloths[0].dx
    ii = np.arange(NVc) // (s.cloths[0].M + 1)
    jj = np.arange(NVc) % (s.cloths[0].M + 1)
    pos[:NVc, 2] += 0.05 * dx * (1 + np.sin(2 * np.pi * ii / 8.0) * np.cos(2 * np.pi * jj / 8.0))
    pos[:NVc] += rng.uniform(-0.01 * dx, 0.01 * dx, (NVc, 3))
    s.pos.from_numpy(pos)
    s.push_down_pos()
    out = dict(T=T, dt=s.dt, k_contact=s.k_contact, eps_contact=s.eps_contact, eps_v=s.eps_v, mu=0.5,
               Kb=120.0, k_angle=s.cloths[0].k_angle[None], cloth_N=s.cloths[0].N, cloth_M=s.cloths[0].M,
               cloth_dx=dx, cloth_mass=s.cloths[0].mass, pos0=s.pos.to_numpy(), vel0=s.vel.to_numpy(),
               mass=s.mass.to_numpy(), frozen=s.frozen.to_numpy(), faces=s.faces.to_numpy(),
               ref_angle0=s.cloths[0].ref_angle.to_numpy(), border_flag=s.border_flag.to_numpy(),
               gravity=s.gravity.to_numpy(), tet_vertices=s.elastics[0].F_vertices.to_numpy(),
               table_offset=s.elastics[0].offset, table_nverts=s.elastics[0].n_verts,
               body_v=np.array([[b.v_start, b.v_end] for b in s.body_list]),
               body_f=np.array([[b.f_start, b.f_end] for b in s.body_list]))
    g.copy_pos(s, 0)
    t0 = time.time()
    for frame in range(1, T):
        # instrumented copy of BaseScene.time_step (code/engine/BaseScene.py:1327-1370): same calls, same order
        s.timestep_init()
        s.calc_vn()
        projection_query(s)
        s.contact_analysis()
        nc = s.nc[None]
        out[f"f{frame}_nc"] = nc
        out[f"f{frame}_vn"] = s.vn.to_numpy()
        out[f"f{frame}_proj_flag"] = s.proj_flag.to_numpy()
        out[f"f{frame}_proj_dir"] = s.proj_dir.to_numpy()
        out[f"f{frame}_proj_idx"] = s.proj_idx.to_numpy()
        out[f"f{frame}_proj_w"] = s.proj_w.to_numpy()
        for k in ("const_idx", "const_w", "const_k", "const_mu", "const_dx0", "const_T", "const_n"):
            out[f"f{frame}_{k}"] = getattr(s, k).to_numpy()[:nc]
        it = 0
        log = []
        while it < 1000:
            it += 1
            s.newton_step_init()
            s.compute_energy()
            E0 = s.E[None]
            s.compute_residual_and_Hessian(False, it, spd=True)
            if it <= 2:
                H = _csr_of(s.H)
                out[f"f{frame}_it{it}_H_data"] = H.data
                out[f"f{frame}_it{it}_H_indices"] = H.indices
                out[f"f{frame}_it{it}_H_indptr"] = H.indptr
                out[f"f{frame}_it{it}_F"] = s.F.to_numpy()
                out[f"f{frame}_it{it}_pos"] = s.pos.to_numpy()
            delta, alpha = s.newton_step(it)
            if it <= 2:
                out[f"f{frame}_it{it}_p"] = fake_linalg.LAST_SOLVE["x"].copy()
            log.append((E0, delta, alpha, s.E[None]))
            print(f"frame {frame} it {it} E0 {E0:.10e} delta {delta:.3e} alpha {alpha} t {time.time()-t0:.0f}s", flush=True)
            if delta < 1e-7:
                break
        s.timestep_finish()
        out[f"f{frame}_newton_log"] = np.array(log)
        out[f"f{frame}_pos"] = s.pos.to_numpy()
        out[f"f{frame}_vel"] = s.vel.to_numpy()
        out[f"f{frame}_ref_angle"] = s.cloths[0].ref_angle.to_numpy()
        g.copy_pos(s, frame)
        np.savez_compressed(os.path.join(OUT, f"{tag}_partial.npz"), **out)
    out["reward"] = s.compute_reward()
    # backward: trajopt_bouncing.py:106-110
    g.get_loss_table(s)
    out["pos_grad_seed"] = g.pos_grad.to_numpy()
    for j in range(T - 1, 0, -1):
        g.transfer_grad(j, s, projection_query)
        out[f"b{j}_z"] = fake_linalg.LAST_SOLVE["x"].copy()
        out[f"b{j}_rhs"] = fake_linalg.LAST_SOLVE["b"].copy()
        Hb = sp.csr_matrix(fake_linalg.LAST_SOLVE["H"])
        out[f"b{j}_H_data"], out[f"b{j}_H_indices"], out[f"b{j}_H_indptr"] = Hb.data, Hb.indices, Hb.indptr
        out[f"b{j}_nc"] = s.nc[None]
        out[f"b{j}_const_idx"] = s.const_idx.to_numpy()[:s.nc[None]]
        out[f"b{j}_pos_grad"] = g.pos_grad.to_numpy()
        out[f"b{j}_angleref_grad"] = g.angleref_grad.to_numpy()
        out[f"b{j}_grad_kb"] = g.grad_kb[None]
        out[f"b{j}_tmp_z_frozen"] = s.tmp_z_frozen.to_numpy()
        out[f"b{j}_d_kb"] = s.d_kb.to_numpy()
        print(f"backward {j} grad_kb {g.grad_kb[None]:.10e} t {time.time()-t0:.0f}s", flush=True)
    out["grad_kb"] = g.grad_kb[None]
    out["pos_buffer"] = g.pos_buffer.to_numpy()
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    os.remove(os.path.join(OUT, f"{tag}_partial.npz"))
    print("wrote", tag)


def gen_folding(T=3, tag="folding", forming=False):
    """config 0: Scene_folding forward rollout + trajectory adjoint, following training/trajopt_folding.py:50-133
    (cloth 15x3 + frozen table + tactile pad on a kinematic gripper).  forming=True: the same for Scene_forming
    (training/trajopt_forming.py:45-141: cloth 15x7, k_contact 20000, Kb 200, position loss get_loss_push)"""
    import scipy.sparse as sp
    if forming:
        from thinshelllab.task_scene.Scene_forming import Scene
    else:
        from thinshelllab.task_scene.Scene_folding import Scene
    from thinshelllab.engine.geometry import projection_query
    from thinshelllab.engine.analytic_grad_single import Grad
    from thinshelllab.agent.traj_opt_single import agent_trajopt
    from cupyx.scipy.sparse import linalg as fake_linalg
    s = Scene(cloth_size=0.1)
    s.device = "cpu"; s.H.device = "cpu"
    Kb = 200.0 if forming else 400.0
    s.cloths[0].Kb[None] = Kb
    g = Grad(s, T, s.elastic_cnt - 1)
    agent = agent_trajopt(T, s.elastic_cnt - 1, max_moving_dist=0.001)
    s.init_all()
    g.init_mass(s)
    s.reset()
    s.mu_cloth_elastic[None] = 5.0
    traj = np.zeros((T, s.elastic_cnt - 1, 6))
    for i in range(1, T):
        traj[i, 0] = [2e-4 * i, -1e-4 * i, -4e-4 * i, 2e-3 * i, 1e-2 * i, -3e-3 * i]
    agent.traj.from_numpy(traj)
    pad = s.elastics[1]
    out = dict(T=T, dt=s.dt, k_contact=s.k_contact, eps_contact=s.eps_contact, eps_v=s.eps_v, mu=5.0, Kb=Kb,
               k_angle=s.cloths[0].k_angle[None], cloth_N=s.cloths[0].N, cloth_M=s.cloths[0].M, cloth_dx=s.cloths[0].dx,
               cloth_mass=s.cloths[0].mass, cloth_size=0.1, traj=traj,
               pos0=s.pos.to_numpy(), vel0=s.vel.to_numpy(), mass=s.mass.to_numpy(), frozen=s.frozen.to_numpy(),
               faces=s.faces.to_numpy(), ref_angle0=s.cloths[0].ref_angle.to_numpy(), n_cloths=len(s.cloths), damping=s.damping, border_flag=s.border_flag.to_numpy(),
               gravity=s.gravity.to_numpy(),
               table_tets=s.elastics[0].F_vertices.to_numpy(), table_offset=s.elastics[0].offset, table_nverts=s.elastics[0].n_verts,
               pad_tets=pad.F_vertices.to_numpy(), pad_offset=pad.offset, pad_nverts=pad.n_verts, pad_F_ox=pad.F_ox.to_numpy(),
               pad_ratio=pad.ratio, pad_f2v=pad.f2v.to_numpy(), pad_is_surface=pad.is_surface.to_numpy(),
               pad_F_B=pad.F_B.to_numpy(), pad_F_W=pad.F_W.to_numpy(), pad_mu=pad.mu[None], pad_lam=pad.lam[None],
               pad_alpha=pad.alpha[None], pad_gravity=pad.gravity.to_numpy(), table_gravity=s.elastics[0].gravity.to_numpy(),
               table_mu=s.elastics[0].mu[None], table_lam=s.elastics[0].lam[None],
               gripper_pos0=s.gripper.pos.to_numpy(), gripper_rot0=s.gripper.rot.to_numpy(), gripper_F_x=s.gripper.F_x.to_numpy(),
               gripper_bound_idx=s.gripper.bound_idx.to_numpy(),
               body_v=np.array([[b.v_start, b.v_end] for b in s.body_list]),
               body_f=np.array([[b.f_start, b.f_end] for b in s.body_list]))
    g.copy_pos(s, 0)
    t0 = time.time()
    for frame in range(1, T):
        agent.get_action(frame)
        s.action(frame, agent.delta_pos, agent.delta_rot)
        out[f"f{frame}_pos_after_action"] = s.pos.to_numpy()
        out[f"f{frame}_gripper_pos"] = s.gripper.pos.to_numpy()
        out[f"f{frame}_gripper_rot"] = s.gripper.rot.to_numpy()
        out[f"f{frame}_gripper_rotmat"] = s.gripper.rotmat.to_numpy()
        # instrumented copy of Scene_folding.time_step (code/task_scene/Scene_folding.py:275-321): same calls, same order
        s.timestep_init()
        s.calc_vn()
        projection_query(s)
        s.contact_analysis()
        nc = s.nc[None]
        out[f"f{frame}_nc"] = nc
        out[f"f{frame}_proj_flag"] = s.proj_flag.to_numpy()
        out[f"f{frame}_proj_dir"] = s.proj_dir.to_numpy()
        out[f"f{frame}_proj_idx"] = s.proj_idx.to_numpy()
        out[f"f{frame}_proj_w"] = s.proj_w.to_numpy()
        for k in ("const_idx", "const_w", "const_k", "const_mu", "const_dx0", "const_T", "const_n"):
            out[f"f{frame}_{k}"] = getattr(s, k).to_numpy()[:nc]
        it = 0
        log = []
        while it < 50:
            it += 1
            s.newton_step_init()
            s.compute_energy()
            E0 = s.E[None]
            s.compute_residual_and_Hessian(False, it, spd=True)
            if it <= 1:
                H = _csr_of(s.H)
                out[f"f{frame}_it{it}_H_data"] = H.data
                out[f"f{frame}_it{it}_H_indices"] = H.indices
                out[f"f{frame}_it{it}_H_indptr"] = H.indptr
                out[f"f{frame}_it{it}_F"] = s.F.to_numpy()
                out[f"f{frame}_it{it}_pos"] = s.pos.to_numpy()
                out[f"f{frame}_it{it}_E0"] = E0
            delta, alpha = s.newton_step(it)
            log.append((E0, delta, alpha, s.E[None]))
            print(f"frame {frame} it {it} E0 {E0:.10e} delta {delta:.3e} alpha {alpha} nc {nc} t {time.time()-t0:.0f}s", flush=True)
            if delta < 1e-7:
                break
        s.timestep_finish()
        out[f"f{frame}_newton_log"] = np.array(log)
        out[f"f{frame}_pos"] = s.pos.to_numpy()
        out[f"f{frame}_vel"] = s.vel.to_numpy()
        out[f"f{frame}_ref_angle"] = s.cloths[0].ref_angle.to_numpy()
        g.copy_pos(s, frame)
        np.savez_compressed(os.path.join(OUT, f"{tag}_partial.npz"), **out)
    NVc = s.cloths[0].NV
    if forming:
        target = out["pos0"][:NVc].copy()
        target[:, 2] -= 1e-3
        out["target_pos"] = target
        out["reward"] = s.compute_reward(target)
        g.get_loss_push(s, target)
    else:
        out["reward"] = s.compute_reward(1.0, -1.0)
        g.get_loss_fold(s, 1.0, -1.0)
        # a position loss on the last frame as well, so that the adjoint right-hand side is not almost empty
        pg = g.pos_grad.to_numpy()
        pg[T - 1, :NVc, 2] = 1.0
        g.pos_grad.from_numpy(pg)
    out["pos_grad_seed"] = g.pos_grad.to_numpy()
    out["angleref_grad_seed"] = g.angleref_grad.to_numpy()
    for j in range(T - 1, 0, -1):
        g.transfer_grad(j, s, projection_query)
        out[f"b{j}_z"] = fake_linalg.LAST_SOLVE["x"].copy()
        out[f"b{j}_rhs"] = fake_linalg.LAST_SOLVE["b"].copy()
        Hb = sp.csr_matrix(fake_linalg.LAST_SOLVE["H"])
        out[f"b{j}_H_data"], out[f"b{j}_H_indices"], out[f"b{j}_H_indptr"] = Hb.data, Hb.indices, Hb.indptr
        out[f"b{j}_nc"] = s.nc[None]
        out[f"b{j}_const_idx"] = s.const_idx.to_numpy()[:s.nc[None]]
        out[f"b{j}_pos_grad"] = g.pos_grad.to_numpy()
        out[f"b{j}_angleref_grad"] = g.angleref_grad.to_numpy()
        out[f"b{j}_tmp_z_frozen"] = s.tmp_z_frozen.to_numpy()
        out[f"b{j}_gripper_grad"] = g.gripper_grad.to_numpy()
        print(f"backward {j} gripper_grad {g.gripper_grad.to_numpy()[j]} t {time.time()-t0:.0f}s", flush=True)
    out["gripper_grad"] = g.gripper_grad.to_numpy()
    out["pos_buffer"] = g.pos_buffer.to_numpy()
    out["ref_angle_buffer"] = g.ref_angle_buffer.to_numpy()
    out["gripper_pos_buffer"] = g.gripper_pos_buffer.to_numpy()
    out["gripper_rot_buffer"] = g.gripper_rot_buffer.to_numpy()
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    os.remove(os.path.join(OUT, f"{tag}_partial.npz"))
    print("wrote", tag)


def gen_trajopt_folding(T=3, iters=2, lr=0.001):
    """the optimisation loop of training/trajopt_folding.py:50-140, statement by statement, with the reference's own engine, agent and
    optimiser under the emulation; records what the script does not print (gripper gradient, trajectory after every Adam step)"""
    import json
    from thinshelllab.task_scene.Scene_folding import Scene
    from thinshelllab.engine.geometry import projection_query
    from thinshelllab.engine.analytic_grad_single import Grad
    from thinshelllab.agent.traj_opt_single import agent_trajopt
    from thinshelllab.optimizer.optim import Adam_single
    sys_ = Scene(cloth_size=0.1)
    sys_.device = "cpu"; sys_.H.device = "cpu"
    sys_.cloths[0].Kb[None] = 400.0
    analy_grad = Grad(sys_, T, sys_.elastic_cnt - 1)
    adam = Adam_single((T, sys_.elastic_cnt - 1, 6), lr, 0.9, 0.9999, 1e-8)
    agent = agent_trajopt(T, sys_.elastic_cnt - 1, max_moving_dist=0.001)
    sys_.init_all()
    analy_grad.init_mass(sys_)
    sys_.reset()
    sys_.mu_cloth_elastic[None] = 5.0
    adam.reset()
    out = dict(what="optimisation loop of the reference's training/trajopt_folding.py (statements :73-140, --tot_step 3 --iter 2 --lr 0.001) run with "
                    "the reference's own Scene_folding / Grad / agent_trajopt / Adam_single under the Taichi emulation shim (SuperLU in place of CuPy spsolve)",
               generated_by="python oracle/gen_goldens.py trajopt_folding", args=dict(tot_step=T, iter=iters, lr=lr, Kb=400.0, mu_cloth_elastic=5.0),
               total_reward=[], gripper_grad=[], traj_after_step=[])
    t0 = time.time()
    for i in range(iters):
        analy_grad.copy_pos(sys_, 0)
        for frame in range(1, T):
            agent.get_action(frame)
            sys_.action(frame, agent.delta_pos, agent.delta_rot)
            sys_.time_step(projection_query, frame)
            analy_grad.copy_pos(sys_, frame)
        out["total_reward"].append(float(sys_.compute_reward(1.0, -1.0)))
        analy_grad.get_loss_fold(sys_, 1.0, -1.0)
        for j in range(T - 1, 0, -1):
            analy_grad.transfer_grad(j, sys_, projection_query)
        out["gripper_grad"].append(analy_grad.gripper_grad.to_numpy().tolist())
        sys_.reset()
        adam.step(agent.traj, analy_grad.gripper_grad)
        agent.fix_action(0.015)
        analy_grad.reset()
        out["traj_after_step"].append(agent.traj.to_numpy().tolist())
        print(f"iter {i} reward {out['total_reward'][-1]} t {time.time() - t0:.0f}s", flush=True)
        out["final_traj"] = out["traj_after_step"][-1]
        with open(os.path.join(OUT, "trajopt_folding_T3.json"), "w") as fh:
            json.dump(out, fh, indent=1)
    print("wrote trajopt_folding_T3.json")


if __name__ == "__main__":
    what = sys.argv[1:] or ["cloth", "spd"]
    if "cloth" in what:
        gen_cloth()
    if "spd" in what:
        gen_spd()
    if "tets" in what:
        gen_tets()
    if "bouncing" in what:
        gen_bouncing()
    if "folding" in what:
        gen_folding()
    if "forming" in what:
        gen_folding(tag="forming", forming=True)
    if "trajopt_folding" in what:
        gen_trajopt_folding()


def gen_scene_states():
    """initial states of the multi-pad task scenes (Scene_lifting, Scene_pick, Scene_balancing, Scene_interact, Scene_card, Scene_sliding) as the reference builds them: Scene(); init_all(); reset().
    No time stepping (seconds).  Pins thinshelllab_b200/engine/scene_builder.py."""
    import importlib
    ALL = ("lifting", "pick", "balancing", "interact", "card", "sliding")
    tags = [t for t in ALL if t in sys.argv[2:]] or ALL
    for tag in tags:
        mod = importlib.import_module(f"thinshelllab.task_scene.Scene_{tag}")
        s = mod.Scene(cloth_size=0.06)
        s.device = "cpu"; s.H.device = "cpu"
        s.init_all()
        s.reset()
        out = dict(dt=s.dt, k_contact=s.k_contact, eps_contact=s.eps_contact, eps_v=s.eps_v, max_n_constraints=s.max_n_constraints,
                   cloth_N=s.cloths[0].N, cloth_M=s.cloths[0].M, cloth_dx=s.cloths[0].dx, cloth_mass=s.cloths[0].mass, k_angle=s.cloths[0].k_angle[None],
                   Kb=s.cloths[0].Kb[None], pos0=s.pos.to_numpy(), vel0=s.vel.to_numpy(), mass=s.mass.to_numpy(), frozen=s.frozen.to_numpy(),
                   faces=s.faces.to_numpy(), ref_angle0=s.cloths[0].ref_angle.to_numpy(), n_cloths=len(s.cloths), damping=s.damping, border_flag=s.border_flag.to_numpy(), gravity=s.gravity.to_numpy(),
                   cloth_gravity=s.cloths[0].gravity.to_numpy(), n_elastics=len(s.elastics), effector_cnt=s.effector_cnt,
                   body_v=np.array([[b.v_start, b.v_end] for b in s.body_list]), body_f=np.array([[b.f_start, b.f_end] for b in s.body_list]),
                   gripper_pos0=s.gripper.pos.to_numpy(), gripper_rot0=s.gripper.rot.to_numpy(), gripper_bound_idx=s.gripper.bound_idx.to_numpy())
        if hasattr(s.gripper, "F_x_upper"):       # gripper_tactile: two pads per part
            out.update(gripper_F_x_upper=s.gripper.F_x_upper.to_numpy(), gripper_F_x_lower=s.gripper.F_x_lower.to_numpy(),
                       gripper_half_dist=s.gripper.half_gripper_dist.to_numpy(), gripper_surface_idx=s.gripper.surface_idx.to_numpy())
        else:
            out.update(gripper_F_x=s.gripper.F_x.to_numpy())
        for j, el in enumerate(s.elastics):
            out[f"el{j}_offset"], out[f"el{j}_nverts"] = el.offset, el.n_verts
            out[f"el{j}_tets"] = el.F_vertices.to_numpy()
            out[f"el{j}_F_B"], out[f"el{j}_F_W"] = el.F_B.to_numpy(), el.F_W.to_numpy()
            out[f"el{j}_mu"], out[f"el{j}_lam"] = el.mu[None], el.lam[None]
            out[f"el{j}_gravity"] = el.gravity.to_numpy()
            out[f"el{j}_tactile"] = int(hasattr(el, "alpha"))
            if hasattr(el, "alpha"):
                out[f"el{j}_alpha"] = el.alpha[None]
        np.savez_compressed(os.path.join(OUT, f"scene_state_{tag}.npz"), **out)
        print("wrote scene_state_" + tag, out["pos0"].shape, flush=True)


if __name__ == "__main__" and "scene_states" in sys.argv[1:]:
    gen_scene_states()

