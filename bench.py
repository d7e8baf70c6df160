#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: element-updates/sec per implicit fwd+bwd step; HBM GB/s vs peak.

Workload (both arms, every N): the LANDING of a square sheet on a frozen table (thinshelllab_b200.synthetic.LANDING: released 0.3 mm above
the table, inside the 0.4 mm contact gap; Scene_bouncing physics, dt 5 ms).  One "step" = one implicit forward time step (contact
query, Newton with multigrid-preconditioned PCG, line search) plus one adjoint step for it (contact re-detection, un-projected fp64
Hessian, adjoint solve, dL/dKb).  The timed window is FIXED: steps 0 .. K-1 of the landing from the initial state (warm-up steps run
the same steps first, then the state is restored), so the number does not depend on where a landing is cut; Newton / PCG totals and
ms per Newton iteration / per PCG iteration are first-class fields of the line.

`value` = triangles x steps / device time with state resident in HBM; `e2e` = the same window through the host-buffer C-ABI entry
point (tsl_step_forward_host) + host-side loss seed / gradient read-back, copies inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--sheet-n 707] [--impl reference]

N > 1 (torchrun): N replicas of the same sheet (same seed), weak scaling, no data-path collective (DESIGN.md section 6 says why the
strip partition of one sheet is opt-in).  --impl reference: the CPU oracle (restatement of the reference, fp64, SuperLU) on the host
cores, same config object, each step a bounded sample (a sub-sheet of the same landing, see cpu_baseline.sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION (set in this image) and at every higher level: drop the
# variable (level NONE) before torch is imported, so that stdout carries the one JSON line only
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    del os.environ["NCCL_DEBUG"]

METRIC = "tri_steps_per_s (implicit fwd+bwd step)"
UNIT = "tri-steps/s"

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel from the committed ncu --set full capture
# (profiles/), keyed by sheet size; None where no capture exists
TRAFFIC = {707: 283.47e6}     # profiles/r1_ncu_full_k_spmv_mixed.csv: 273.57 MB read + 9.90 MB written


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(N):
    """the `config` object of BOTH arms (the driver compares them)"""
    n_tris = 2 * N * N
    tag = {158: "BASELINE configs[1] / [2]", 316: "BASELINE configs[3] sheet size", 707: "the 1 M-triangle sheet of BASELINE configs[4] / north_star"}.get(N, "custom size")
    return {"workload": f"sheet {N}x{N} ({n_tris} tris, dx 2 mm) released 0.3 mm above a frozen table (landing), Scene_bouncing physics, dt 5 ms; "
                        f"window = steps 0..K-1 of the landing, forward + adjoint (dL/dKb) per step; {tag}",
            "sheet_n": N, "n_tris": n_tris, "window": "steps 0..K-1 from the initial state (state restored after warm-up)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self._stop_ev = gpu_index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=5)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def oracle_window(sample_n, steps, budget_s, threads=None):
    """steps 0..steps-1 of the landing on a sample_n x sample_n sheet with the CPU oracle, forward + adjoint per step, stopping early when
    `budget_s` seconds of wall clock are spent.  Returns (tris, [seconds per completed step], threads, [newton iterations per step])."""
    from oracle import tsl_oracle as orc
    from thinshelllab_b200.synthetic import LANDING, sheet_spec
    if threads:
        orc.lib().orc_set_num_threads(int(threads))
    sp = sheet_spec(sample_n, **LANDING)
    o = orc.OracleScene(sample_n, sample_n, sp["dx"], sp["dt"], sp["table_pos"], sp["table_faces"], sp["table_mass"], k_contact=sp["k_contact"],
                        mu=sp["mu"], max_n_constraints=sp["max_n_constraints"], grid_n=sp["grid_n"])
    o.pos[:o.NVc] = sp["cloth_pos"]; o.prev_pos[:] = o.pos
    times, newton = [], []
    t_start = time.perf_counter()
    for _ in range(steps):
        g = orc.OracleGrad(o, 2)
        t0 = time.perf_counter()
        g.copy_pos(0)
        newton.append(o.time_step())
        g.copy_pos(1)
        g.pos_grad[1, :o.NVc, 2] = 1.0
        g.transfer_grad(1)
        # transfer_grad leaves the scene at (x_t, x_{t-1}); the next forward step starts from x_t with its own velocity
        o.pos[:] = g.pos_buffer[1]
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    return sp["n_tris"], times, orc.lib().orc_num_threads(), newton


def _cpu_sample_text(n, tris, times, newton, kind="window"):
    return (f"{n}x{n} sub-sheet ({tris} tris) of the same landing, steps 0..{len(times) - 1} forward + adjoint "
            f"({int(np.sum(newton))} Newton iterations), CPU oracle (fp64 restatement of the reference, SuperLU direct solves, OpenMP assembly)")


def run_reference(args, rank):
    if rank != 0:
        return
    n = args.cpu_sample_n
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())       # torchrun pins it to 1: the CPU arm uses every host core
    want = args.warmup + args.steps
    # no warm-up state exists on the CPU: the W + K launched steps are all timed; when the wall-clock bound cuts the window short the
    # throughput is that of the steps completed (steps_run)
    tris, times, threads, newton = oracle_window(n, want, args.cpu_budget_s, threads=os.cpu_count())
    T = float(np.sum(times))
    val = tris * len(times) / T
    sample = _cpu_sample_text(n, tris, times, newton)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "steps_run": len(times), "ms_per_step": 1e3 * T / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.sheet_n),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "newton_iters": [int(x) for x in newton],
                         "note": "CPU oracle = restatement of the reference sources pinned to goldens made by the reference itself under "
                                 "a Taichi emulation; Taichi / CuPy cannot be installed in this image, the reference's dense-backed "
                                 "SparseMatrix needs 92 GB at 50 k triangles"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ GPU arm
def _fixed_window(s, g, e, NVc, steps, adjoint_tol, snap):
    """steps 0..steps-1 from the snapshot `snap`, device resident; returns (ms, per-step stats)"""
    import torch
    e.pos.copy_(snap[0]); e.prev_pos.copy_(snap[0]); e.vel.copy_(snap[1]); e.cloth_ref_angle[0].copy_(snap[2]); e.reset_contact_state()
    stats = []
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        g.copy_pos(s, 0)
        st = s.time_step()
        g.copy_pos(s, 1)
        g._pos_grad.zero_(); g._angleref_grad.zero_()
        g._pos_grad[1, :NVc, 2] = 1.0                      # loss seed: dL/dz = 1 on the cloth (cf. Grad.get_loss_slide_simple)
        its, flags, rr = g.transfer_grad(1, s, rel_tol=adjoint_tol)
        stats.append((st.newton_iters, st.linear_iters, st.linesearch_evals, st.n_contacts, its, st.flags, flags, int(st.converged)))
    ev1.record()
    torch.cuda.synchronize()
    return ev0.elapsed_time(ev1), stats


def secondary_50k(args, dev):
    """BASELINE.json configs[1] / configs[2] (50 k-triangle sheet, forward only and forward + adjoint), same drop window, one GPU"""
    import torch
    from thinshelllab_b200.engine.analytic_grad_system import Grad
    from thinshelllab_b200.synthetic import LANDING, sheet_scene
    N, K = 158, 10
    s = sheet_scene(N, device=dev, **LANDING)
    e, NVc, g = s.engine, s.cloths[0].NV, Grad(s, 2, 0)
    snap = (e.pos.clone(), e.vel.clone(), e.cloth_ref_angle[0].clone())
    _fixed_window(s, g, e, NVc, 3, args.adjoint_tol, snap)              # warm-up
    out = {}
    ms, stats = _fixed_window(s, g, e, NVc, K, args.adjoint_tol, snap)
    st = np.array(stats, dtype=np.float64)
    out["configs[2] forward + adjoint"] = {"sheet": "158x158 (49928 tris)", "window": f"steps 0..{K - 1} of the landing", "ms_per_step": ms / K,
                                           "tri_steps_per_s": 2 * N * N * K / (ms * 1e-3), "newton_iters": int(st[:, 0].sum()),
                                           "pcg_iters": int(st[:, 1].sum()), "adjoint_iters": int(st[:, 4].sum())}
    # forward only
    e.pos.copy_(snap[0]); e.prev_pos.copy_(snap[0]); e.vel.copy_(snap[1]); e.cloth_ref_angle[0].copy_(snap[2]); e.reset_contact_state()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    nw = 0
    for _ in range(K):
        nw += s.time_step().newton_iters
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    out["configs[1] forward only"] = {"sheet": "158x158 (49928 tris)", "window": f"steps 0..{K - 1} of the landing", "ms_per_step": ms / K,
                                      "tri_steps_per_s": 2 * N * N * K / (ms * 1e-3), "newton_iters": int(nw)}
    del s, g
    torch.cuda.empty_cache()
    return out


def secondary_solids(args, dev):
    """BASELINE.json configs[0] (Scene_folding: cloth strip + table + tactile pad on a gripper, T = 3 rollout + trajectory adjoint) and a
    configs[3] (316 x 316 = 200 k-triangle sheet on the table, the free TetGen ball lying on it, its overhanging edge between the two
    tactile pads of a two-finger gripper that closes and lifts: data/tactile.*, data/ball.*; contacts against moving triangles), one
    GPU, state resident in HBM"""
    import torch
    from thinshelllab_b200.agent.traj_opt_single import agent_trajopt
    from thinshelllab_b200.engine.analytic_grad_single import Grad
    from thinshelllab_b200.synthetic import config3_scene
    from thinshelllab_b200.task_scene.Scene_folding import Scene
    T0 = 3
    traj0 = np.zeros((T0, 1, 6))
    for i in range(1, T0):                       # the press-and-tilt trajectory of the golden run (oracle/gen_goldens.py:gen_folding)
        traj0[i, 0] = [2e-4 * i, -1e-4 * i, -4e-4 * i, 2e-3 * i, 1e-2 * i, -3e-3 * i]
    out = {}

    def rollout(s, T, traj, reps):
        NVc = s.cloths[0].NV
        agent = agent_trajopt(T, 1, max_moving_dist=0.001)
        agent.traj.from_numpy(traj)
        grad = Grad(s, T, 1)
        res = None
        for rep in range(reps + 1):              # first repetition = warm-up (graph capture, allocations)
            s.reset()
            grad.reset()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            grad.copy_pos(s, 0)
            newton = krylov = 0
            for frame in range(1, T):
                agent.get_action(frame)
                s.action(frame, agent.delta_pos, agent.delta_rot)
                st = s.time_step()
                newton += st.newton_iters; krylov += st.linear_iters
                grad.copy_pos(s, frame)
            grad._pos_grad[T - 1, :NVc, 2] = 1.0
            bi = 0
            for j in range(T - 1, 0, -1):
                bi += grad.transfer_grad(j, s, rel_tol=args.adjoint_tol)[0]
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            res = {"steps": T - 1, "ms_per_step": ms / (T - 1), "tri_steps_per_s": s.cloths[0].NF * (T - 1) / (ms * 1e-3), "newton_iters": newton,
                   "pcg_iters": krylov, "adjoint_iters (0 = dense LU)": bi, "contacts_last_step": int(st.n_contacts), "converged_last_step": bool(st.converged)}
        return res

    s = Scene(cloth_size=0.1, device=dev)
    s.cloths[0].Kb[None] = 400.0; s.mu_cloth_elastic[None] = 5.0           # training/trajopt_folding.py:52-56
    out["configs[0] Scene_folding fwd + trajectory adjoint"] = dict(scene="cloth 15x3 (90 tris) + frozen table + tactile pad (1365 tets) on a gripper",
                                                                    **rollout(s, T0, traj0, 2))
    del s
    N, T = 316, 6
    s = config3_scene(N, device=dev)
    traj = np.zeros((T, 1, 6)); traj[:, 0, 2] = 1.5e-4 * np.maximum(np.arange(T) - 2, 0)
    out["configs[3] 200k-tri sheet + ball + two pads gripping an edge, fwd + trajectory adjoint"] = dict(
        scene=f"sheet {N}x{N} ({2 * N * N} tris) resting on a frozen table, free TetGen ball (295 tets) lying on it, the overhanging edge "
              "between the two tactile pads (1365 tets each) of a two-finger gripper that closes 0.15 mm per frame and lifts", **rollout(s, T, traj, 1))
    del s
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    from thinshelllab_b200.engine.analytic_grad_system import Grad
    from thinshelllab_b200.synthetic import LANDING, sheet_scene
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    N = args.sheet_n
    s = sheet_scene(N, device=dev, **LANDING)                    # the same sheet (seed 0) on every rank
    e = s.engine
    if args.newton_mode >= 0:
        from thinshelllab_b200 import _lib as _l
        e.set_option(_l.OPT_NEWTON_MODE, args.newton_mode)
    NVc = s.cloths[0].NV
    n_tris = 2 * N * N
    g = Grad(s, 2, 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    snap = (e.pos.clone(), e.vel.clone(), e.cloth_ref_angle[0].clone())
    # host buffers of the e2e path
    pos_h = torch.empty((e.n_verts, 3), dtype=torch.float64).pin_memory()
    vel_h = torch.empty((e.n_verts, 3), dtype=torch.float64).pin_memory()
    seed_h = torch.zeros((e.n_verts, 3), dtype=torch.float64).pin_memory(); seed_h[:NVc, 2] = 1.0
    pg_h = torch.empty((e.n_verts, 3), dtype=torch.float64).pin_memory()

    def fwd_bwd_host():
        g._pos_buffer[0].copy_(pos_h, non_blocking=True)    # x_{t-1} travels with the step's inputs
        g._ref_angle_buffer[0, 0].copy_(e.cloth_ref_angle[0])
        e.step_forward_host(pos_h, vel_h)                   # H2D pos, vel -> step -> D2H pos, vel
        g.copy_pos(s, 1)
        g._pos_grad.zero_(); g._angleref_grad.zero_()
        g._pos_grad[1].copy_(seed_h, non_blocking=True)     # H2D loss seed
        g.transfer_grad(1, s, rel_tol=args.adjoint_tol)
        pg_h.copy_(g._pos_grad[0], non_blocking=True)       # D2H dL/dx_{t-1}
        return g.grad_kb[None]                              # D2H dL/dKb (8 bytes, syncs)

    # warm-up: the first W steps of the same window (graph capture, allocations, clocks), then the state is restored
    if args.warmup > 0:
        _fixed_window(s, g, e, NVc, args.warmup, args.adjoint_tol, snap)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = e.launch_count()
    ms, stats = _fixed_window(s, g, e, NVc, args.steps, args.adjoint_tol, snap)
    barrier()
    launches = e.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # e2e: the same window through host buffers
    e.pos.copy_(snap[0]); e.prev_pos.copy_(snap[0]); e.vel.copy_(snap[1]); e.cloth_ref_angle[0].copy_(snap[2]); e.reset_contact_state()
    pos_h.copy_(e.pos); vel_h.copy_(e.vel)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fwd_bwd_host()
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    from thinshelllab_b200 import dist as tdist
    units, (ms, ms_e2e) = tdist.aggregate(n_tris * args.steps, [ms, ms_e2e], device=dev)   # SUM of units, MAX of times
    # ---- rooflines, timed live with CUDA events on the launching stream inside libtsl (tsl_bench_kernel), at the state the window left
    from thinshelllab_b200 import _lib
    e.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_NEWTON | _lib.ASM_SPD)
    sz = e.sizes()
    e.bench_kernel(1, 20)
    ms_spmv = e.bench_kernel(1, 200)
    e.bench_kernel(0, 10)
    ms_pcg = e.bench_kernel(0, 100)
    e.bench_kernel(5, 10)
    ms_vcycle = e.bench_kernel(5, 100)
    ms_setup = e.bench_kernel(6, 20)
    ms_hess = e.bench_kernel(4, 20)
    ms_resid = e.bench_kernel(3, 20)
    ms_energy = e.bench_kernel(2, 20)
    V = sz["n_verts"]
    Vs, Bs = sz["n_solve"], sz["nnzb_solve"]                         # rows / blocks the forward solve touches (the frozen table is skipped)
    spmv_bytes = 40.0 * Bs + (4 + 24 + 24) * Vs                    # 36 B + 4 B per block, row ptr 4V, fp64 x 24V, fp64 y 24V
    # one multigrid-PCG iteration: 5 fine-level matrix passes (SpMV, 3 smoother steps, 1 residual) + 4 passes over every
    # coarse level (900 B per coarse vertex, ~1/3 V in total) + ~25 fine vector passes
    n_coarse, ng = 0, N + 1
    while ng > 6:
        ng = (ng - 1) // 2 + 1
        n_coarse += ng * ng
    pcg_bytes = 5 * 40.0 * Bs + 4 * 900.0 * n_coarse + 25 * 12.0 * Vs
    peak, peak_src = _peaks()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    st = np.array(stats, dtype=np.float64)
    value = units / (ms * 1e-3)
    e2e = units / (ms_e2e * 1e-3)
    nb = e.n_verts * 24
    tot_newton, tot_pcg = float(st[:, 0].sum()), float(st[:, 1].sum())

    def rl(us, bytes_):
        return {"us": us, "algorithmic_bytes": bytes_, "achieved_GBps": bytes_ / (us * 1e-6) / 1e9, "frac": bytes_ / (us * 1e-6) / 1e9 / peak}

    cfg = workload_config(N)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (state, energy, residual, contact, adjoint matrix and solve) + f32 (forward Newton matrix, multigrid)", "data": "synthetic",
        "config": cfg,
        "work": {"newton_iters": int(tot_newton), "pcg_iters": int(tot_pcg), "linesearch_evals": int(st[:, 2].sum()), "adjoint_iters": int(st[:, 4].sum()),
                 "ms_per_newton_iter": ms / max(tot_newton, 1), "ms_per_pcg_iter_incl_everything": ms / max(tot_pcg, 1),
                 "unconverged_steps": int((st[:, 7] == 0).sum()),
                 "per_step": [{"newton": int(r[0]), "pcg": int(r[1]), "adjoint": int(r[4]), "contacts": int(r[3])} for r in st],
                 "flags": {"pcg_negative_curvature_steps": int((st[:, 5].astype(int) & 1).sum()), "krylov_cap_hit": int(((st[:, 5].astype(int)) & 2).sum() // 2),
                           "adjoint_fallbacks": int(((st[:, 6].astype(int)) & 8).sum() // 8)}},
        "system": {"n_verts": V, "nnzb": sz["nnzb"], "nnzb_padded": sz["nnzb_padded"], "n_solve": Vs, "nnzb_solve": Bs,
                   "parallelism": "single GPU" if world == 1 else f"{world} replicas of the same sheet (no data-path collective; the strip partition of one sheet is opt-in: DESIGN.md section 6)",
                   "l2": "matrix %.0f MB > 126 MB L2" % (sz["bytes_matrix_f32"] / 1e6) if sz["bytes_matrix_f32"] > 126e6 else
                         "working set %.0f MB fits the 126 MB L2: roofline fractions can exceed 1" % (sz["bytes_matrix_f32"] / 1e6),
                   "solver": "Newton (exact / clamped / blended matrix, line search) + multigrid-preconditioned PCG; adjoint: dense LU below 12288 unknowns, multigrid-FGMRES(50) fp64 above",
                   "newton_mode": ("library default" if args.newton_mode < 0 else args.newton_mode)},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": 4 * nb, "d2h_bytes_per_step": 3 * nb + 8},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_spmv_mixed (fine-level sliced-ELL matrix pass of PCG, fp32 matrix x fp64 vector + fused dot; k_cheb_step_sell / k_mg_residual_sell stream the same matrix)",
                     "achieved": spmv_bytes / (ms_spmv * 1e-3) / 1e9, "peak": peak,
                     "peak_source": peak_src, "unit": "GB/s", "frac": spmv_bytes / (ms_spmv * 1e-3) / 1e9 / peak, "traffic": TRAFFIC.get(N),
                     "us_per_launch": 1e3 * ms_spmv, "algorithmic_bytes_per_launch": spmv_bytes,
                     "pcg_iteration": dict(rl(1e3 * ms_pcg, pcg_bytes), note="one captured CUDA graph: SpMV + update + V-cycle + direction"),
                     # SURVEY 8d algorithmic bytes: Hessian 280 B/tri per matrix, residual 82 B/tri, energy 76 B/tri.  A Newton iteration
                     # builds TWO matrices (exact + clamped) in one owner-computes pass plus the contact / mass kernels of each: that
                     # pair is what is timed (at the end of the window: ~500 k active constraints)
                     "assembly": {"hessian pair (exact + clamped Newton matrices)": rl(1e3 * ms_hess, 2 * 280.0 * n_tris), "residual": rl(1e3 * ms_resid, 82.0 * n_tris),
                                  "energy": rl(1e3 * ms_energy, 76.0 * n_tris)},
                     "other_us": {"vcycle": 1e3 * ms_vcycle, "mg_setup": 1e3 * ms_setup}},
    }
    if world == 1 and N != 158 and not args.no_secondary:
        out["secondary"] = {}
        try:
            out["secondary"]["baseline_configs_50k"] = secondary_50k(args, dev)
        except Exception as ex:                  # a secondary measurement must not take the headline line down with it
            out["secondary"]["baseline_configs_50k"] = {"error": repr(ex)}
        try:
            out["secondary"]["baseline_configs_solids"] = secondary_solids(args, dev)
        except Exception as ex:
            out["secondary"]["baseline_configs_solids"] = {"error": repr(ex)}
    if world == 1 and not args.no_cpu_baseline:
        tris, times, threads, newton = oracle_window(args.cpu_sample_n, 2, 25.0, threads=os.cpu_count())
        out["cpu_baseline"] = {"value": tris * len(times) / float(np.sum(times)), "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": _cpu_sample_text(args.cpu_sample_n, tris, times, newton)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sheet-n", type=int, default=707, help="sheet is N x N quads: 158 -> 50k tris, 316 -> 200k, 707 -> 1M")
    ap.add_argument("--cpu-sample-n", type=int, default=32)
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="wall-clock bound of the CPU arm's window")
    ap.add_argument("--adjoint-tol", type=float, default=1e-8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the 50 k-triangle / solids secondary measurements")
    ap.add_argument("--newton-mode", type=int, default=-1, help="TSL_OPT_NEWTON_MODE of the forward solve (-1: library default)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world)


if __name__ == "__main__":
    main()
