#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: element-updates/sec per implicit fwd+bwd step; HBM GB/s vs peak.

One "step" = one implicit forward time step of the sheet (contact query, Newton with multigrid-preconditioned PCG, line search) plus
one adjoint step for it (contact re-detection, un-projected fp64 Hessian, multigrid-preconditioned BiCGStab solve, parameter gradient dL/dKb).
`value` = triangles x steps / device time with state resident in HBM; `e2e` = the same through the host-buffer C-ABI entry
point (tsl_step_forward_host) + host-side loss seed / gradient read-back, copies inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--sheet-n 707] [--impl reference]

N > 1 (torchrun): one independent sheet per rank (replicas, weak scaling, no data-path collective).  The strip partition of ONE sheet
over the GPUs (tsl_dist_init, NCCL halo exchange + all-reduced Krylov scalars) exists for the forward step and is
exact but, lacking a coarse space across strips, slower than one GPU (tools/bench_partition.py,
profiles/r1_partition_2gpu.md): the benchmark keeps replicas and says so in config.parallelism.
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


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel from the committed ncu --set full capture
# (profiles/), keyed by sheet size; None where no capture exists
TRAFFIC = {707: 283.47e6}     # profiles/r1_ncu_full_k_spmv_mixed.csv: 273.57 MB read + 9.90 MB written


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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
def oracle_fwd_bwd(sample_n, steps, threads=None):
    """times `steps` fwd+bwd steps of the CPU oracle on a sample_n x sample_n sheet (each from the same initial state)."""
    from oracle import tsl_oracle as orc
    from thinshelllab_b200.synthetic import sheet_spec
    if threads:
        orc.lib().orc_set_num_threads(int(threads))
    sp = sheet_spec(sample_n)
    times = []
    for _ in range(steps):
        o = orc.OracleScene(sample_n, sample_n, sp["dx"], sp["dt"], sp["table_pos"], sp["table_faces"], sp["table_mass"], k_contact=sp["k_contact"],
                            mu=sp["mu"], max_n_constraints=sp["max_n_constraints"], grid_n=sp["grid_n"])
        o.pos[:o.NVc] = sp["cloth_pos"]; o.prev_pos[:] = o.pos
        g = orc.OracleGrad(o, 2)
        t0 = time.perf_counter()
        g.copy_pos(0)
        o.time_step()
        g.copy_pos(1)
        g.pos_grad[1, :o.NVc, 2] = 1.0
        g.transfer_grad(1)
        times.append(time.perf_counter() - t0)
    return sp["n_tris"], times, orc.lib().orc_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    n = args.cpu_sample_n
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())       # torchrun pins it to 1: the CPU arm uses every host core
    for _ in range(args.warmup):
        pass                                    # the CPU arm has no warm-up state worth paying ~25 s per step for
    tris, times, threads = oracle_fwd_bwd(n, max(1, args.steps), threads=os.cpu_count())
    T = float(np.sum(times))
    val = tris * len(times) / T
    sample = f"{n}x{n} sheet ({tris} tris) over the table, {len(times)} fwd+bwd step(s) from the bench's initial-state generator; SuperLU direct solves"
    print(json.dumps({
        "impl": "reference", "metric": "tri_steps_per_s (implicit fwd+bwd step)", "value": val, "unit": "tri-steps/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": 0, "ms_per_step": 1e3 * T / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"sheet {args.sheet_n}x{args.sheet_n} fwd+bwd (CPU arm runs the bounded sample below)", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "tri-steps/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "CPU oracle (restatement of the reference sources, pinned to goldens); Taichi/CuPy cannot be installed here"},
        "e2e": {"value": val, "unit": "tri-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------------------ GPU arm
def secondary_50k(args, dev):
    """BASELINE.json configs[1] / configs[2] (50 k-triangle sheet, forward only and forward + adjoint) on one GPU, state resident in
    HBM: reported inside config, next to the 1 M-triangle headline"""
    import torch
    from thinshelllab_b200.engine.analytic_grad_system import Grad
    from thinshelllab_b200.synthetic import sheet_scene
    N = 158
    s = sheet_scene(N, device=dev)
    if args.newton_mode >= 0:
        from thinshelllab_b200 import _lib as _l
        s.engine.set_option(_l.OPT_NEWTON_MODE, args.newton_mode)
    e, NVc, g = s.engine, s.cloths[0].NV, Grad(s, 2, 0)

    def fwd():
        g.copy_pos(s, 0)
        s.time_step()
        g.copy_pos(s, 1)

    def bwd():
        g._pos_grad.zero_(); g._angleref_grad.zero_()
        g._pos_grad[1, :NVc, 2] = 1.0
        g.transfer_grad(1, s, rel_tol=args.adjoint_tol)

    for _ in range(3):
        fwd(); bwd()
    snap = (e.pos.clone(), e.vel.clone(), e.cloth_ref_angle[0].clone())
    out = {}
    for name, with_bwd in (("configs[1] forward only", False), ("configs[2] forward + adjoint", True)):
        e.pos.copy_(snap[0]); e.prev_pos.copy_(snap[0]); e.vel.copy_(snap[1]); e.cloth_ref_angle[0].copy_(snap[2]); e.reset_contact_state()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(3):
            fwd()
            if with_bwd:
                bwd()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        out[name] = {"sheet": "158x158 (49928 tris)", "steps": 3, "ms_per_step": ms / 3, "tri_steps_per_s": 2 * N * N * 3 / (ms * 1e-3)}
    del s, g
    torch.cuda.empty_cache()
    return out


def secondary_solids(args, dev):
    """BASELINE.json configs[0] (Scene_folding: cloth strip + table + tactile pad on a gripper, T = 3 rollout + trajectory adjoint, the
    scene state of thinshelllab_b200/data/scene_folding_cloth0p1.npz) and a configs[3]-style scene (316 x 316 = 200 k-triangle sheet on the table with the volumetric
    tactile pad pressed into it; contacts against moving triangles), one GPU, state resident in HBM"""
    import torch
    from thinshelllab_b200.agent.traj_opt_single import agent_trajopt
    from thinshelllab_b200.engine.analytic_grad_single import Grad
    from thinshelllab_b200.synthetic import pad_sheet_scene
    from thinshelllab_b200.task_scene.Scene_folding import Scene
    g = np.load(os.path.join(ROOT, "thinshelllab_b200", "data", "scene_folding_cloth0p1.npz"))
    T0 = 3
    traj0 = np.zeros((T0, 1, 6))
    for i in range(1, T0):                       # the press-and-tilt trajectory of the golden run (oracle/gen_goldens.py:gen_folding)
        traj0[i, 0] = [2e-4 * i, -1e-4 * i, -4e-4 * i, 2e-3 * i, 1e-2 * i, -3e-3 * i]
    out = {}

    def rollout(s, T, traj, reps):
        e = s.engine
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
                   "pcg_iters": krylov, "bicgstab_iters": bi, "contacts_last_step": int(st.n_contacts), "converged_last_step": bool(st.converged)}
        return res

    s = Scene(g, device=dev)
    out["configs[0] Scene_folding fwd + trajectory adjoint"] = dict(scene="cloth 15x3 (90 tris) + frozen table + tactile pad (1365 tets) on a gripper",
                                                                    **rollout(s, T0, traj0, 2))
    del s
    N, T = 316, 4
    s = pad_sheet_scene(N, g, device=dev)
    traj = np.zeros((T, 1, 6)); traj[:, 0, 2] = -1.5e-4 * np.arange(T)
    out["configs[3] 200k-tri sheet + tactile pad, fwd + trajectory adjoint"] = dict(
        scene=f"sheet {N}x{N} ({2 * N * N} tris) resting on a frozen table, tactile pad (1365 tets) pressed 0.15 mm per step into it", **rollout(s, T, traj, 1))
    del s
    torch.cuda.empty_cache()
    return out


def run_ours(args, rank, world):
    import torch
    import torch.distributed as dist
    from thinshelllab_b200.engine.analytic_grad_system import Grad
    from thinshelllab_b200.synthetic import sheet_scene
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    N = args.sheet_n
    s = sheet_scene(N, device=dev, seed=rank)
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

    stats = []

    def fwd_bwd_device():
        g.copy_pos(s, 0)
        st = s.time_step()
        g.copy_pos(s, 1)
        g._pos_grad.zero_(); g._angleref_grad.zero_()
        g._pos_grad[1, :NVc, 2] = 1.0                      # loss seed: Grad.get_loss_slide-style dL/dz = 1 on the cloth
        its, flags, rr = g.transfer_grad(1, s, rel_tol=args.adjoint_tol)
        stats.append((st.newton_iters, st.linear_iters, st.linesearch_evals, st.n_contacts, its, st.flags, flags))

    # host buffers of the e2e path
    pos_h = torch.empty((e.n_verts, 3), dtype=torch.float64).pin_memory()
    vel_h = torch.empty((e.n_verts, 3), dtype=torch.float64).pin_memory()
    seed_h = torch.zeros((e.n_verts, 3), dtype=torch.float64).pin_memory(); seed_h[:NVc, 2] = 1.0
    pg_h = torch.empty((e.n_verts, 3), dtype=torch.float64).pin_memory()

    def fwd_bwd_host():
        g._pos_buffer[0].copy_(pos_h, non_blocking=True)    # x_{t-1} travels with the step's inputs
        g._ref_angle_buffer[0, 0].copy_(e.cloth_ref_angle[0])
        st = e.step_forward_host(pos_h, vel_h)             # H2D pos, vel -> step -> D2H pos, vel
        g.copy_pos(s, 1)
        g._pos_grad.zero_(); g._angleref_grad.zero_()
        g._pos_grad[1].copy_(seed_h, non_blocking=True)     # H2D loss seed
        g.transfer_grad(1, s, rel_tol=args.adjoint_tol)
        pg_h.copy_(g._pos_grad[0], non_blocking=True)       # D2H dL/dx_{t-1}
        kb = g.grad_kb[None]                                # D2H dL/dKb (8 bytes, syncs)
        return kb

    for _ in range(args.warmup):
        fwd_bwd_device()
    stats.clear()
    # both timed regions run the SAME physical steps: snapshot the state after warm-up
    snap = (e.pos.clone(), e.vel.clone(), e.cloth_ref_angle[0].clone())

    def restore():
        e.pos.copy_(snap[0]); e.prev_pos.copy_(snap[0]); e.vel.copy_(snap[1]); e.cloth_ref_angle[0].copy_(snap[2])
        e.reset_contact_state()
    restore()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = e.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        fwd_bwd_device()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = e.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # e2e
    restore()
    pos_h.copy_(e.pos); vel_h.copy_(e.vel)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fwd_bwd_host()
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    from thinshelllab_b200 import dist as tdist
    units, (ms, ms_e2e) = tdist.aggregate(n_tris * args.steps, [ms, ms_e2e], device=dev)   # SUM of units, MAX of times
    # ---- roofline of the dominant kernel class (fine-level block-sparse matrix pass: PCG SpMV and the V-cycle's fine
    # smoother / residual kernels stream the same bytes), timed live with CUDA events on the launching stream inside libtsl
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
    out = {
        "metric": "tri_steps_per_s (implicit fwd+bwd step)", "value": value, "unit": "tri-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (state, energy, residual, contact, adjoint matrix and solve) + f32 (forward Newton matrix, PCG vectors, multigrid)", "data": "synthetic",
        "config": {"workload": f"sheet {N}x{N} ({n_tris} tris, dx 2 mm) landing on a frozen table, Scene_bouncing physics, fwd + adjoint (dL/dKb) per step"
                               + (" -- the 1 M-triangle sheet of BASELINE configs[4] / north_star on ONE GPU (largest single-GPU configuration; configs[1] and [2] are in baseline_configs_50k)" if N == 707 else ""),
                   "sheet_n": N, "n_tris": n_tris, "n_verts": V, "nnzb": sz["nnzb"], "nnzb_padded": sz["nnzb_padded"], "n_solve": Vs, "nnzb_solve": Bs,
                   "parallelism": "single GPU" if world == 1 else f"{world} independent replicas (the strip partition of one sheet is exact but not yet faster than one GPU: DESIGN.md section 6)",
                   "l2": "matrix %.0f MB > 126 MB L2" % (sz["bytes_matrix_f32"] / 1e6) if sz["bytes_matrix_f32"] > 126e6 else
                         "working set %.0f MB fits the 126 MB L2: roofline fraction can exceed 1" % (sz["bytes_matrix_f32"] / 1e6),
                   "solver": "Newton (exact / clamped / blended matrix, line search) + multigrid-preconditioned PCG; adjoint: multigrid-preconditioned BiCGStab fp64",
                   "newton_mode": ("library default" if args.newton_mode < 0 else args.newton_mode),
                   "per_step_mean": {"newton_iters": st[:, 0].mean(), "pcg_iters": st[:, 1].mean(), "linesearch_evals": st[:, 2].mean(),
                                     "contacts": st[:, 3].mean(), "bicgstab_iters": st[:, 4].mean()},
                   "per_step": [{"newton": int(r[0]), "pcg": int(r[1]), "bicgstab": int(r[4]), "contacts": int(r[3])} for r in st],
                   "flags": {"pcg_negative_curvature_steps": int((st[:, 5].astype(int) & 1).sum()), "krylov_cap_hit": int(((st[:, 5].astype(int) | st[:, 6].astype(int)) & 2).sum() // 2)}},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "tri-steps/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": 4 * nb, "d2h_bytes_per_step": 3 * nb + 8},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_spmv_mixed (fine-level sliced-ELL matrix pass of PCG, fp32 matrix x fp64 vector + fused dot; k_cheb_step_sell / k_mg_residual_sell stream the same matrix)",
                     "achieved": spmv_bytes / (ms_spmv * 1e-3) / 1e9, "peak": peak,
                     "peak_source": peak_src, "unit": "GB/s", "frac": spmv_bytes / (ms_spmv * 1e-3) / 1e9 / peak, "traffic": TRAFFIC.get(N),
                     "us_per_launch": 1e3 * ms_spmv, "algorithmic_bytes_per_launch": spmv_bytes,
                     "pcg_iteration": {"us": 1e3 * ms_pcg, "algorithmic_bytes": pcg_bytes, "achieved": pcg_bytes / (ms_pcg * 1e-3) / 1e9,
                                       "frac": pcg_bytes / (ms_pcg * 1e-3) / 1e9 / peak, "note": "one captured CUDA graph: SpMV + update + V-cycle + direction"},
                     "other_us": {"vcycle": 1e3 * ms_vcycle, "mg_setup": 1e3 * ms_setup, "hessian": 1e3 * ms_hess, "residual": 1e3 * ms_resid, "energy": 1e3 * ms_energy}},
    }
    if world == 1 and N != 158 and not args.no_secondary:
        out["config"]["baseline_configs_50k"] = secondary_50k(args, dev)
        try:
            out["config"]["baseline_configs_solids"] = secondary_solids(args, dev)
        except Exception as ex:                  # a secondary measurement must not take the headline line down with it
            out["config"]["baseline_configs_solids"] = {"error": repr(ex)}
    if world == 1 and not args.no_cpu_baseline:
        tris, times, threads = oracle_fwd_bwd(args.cpu_sample_n, 1, threads=os.cpu_count())
        out["cpu_baseline"] = {"value": tris / times[0], "unit": "tri-steps/s", "cores": threads, "kind": "port",
                               "sample": f"{args.cpu_sample_n}x{args.cpu_sample_n} sheet ({tris} tris), 1 fwd+bwd step, CPU oracle (fp64, SuperLU direct solves; not Taichi)"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sheet-n", type=int, default=707, help="sheet is N x N quads: 158 -> 50k tris, 316 -> 200k, 707 -> 1M")
    ap.add_argument("--cpu-sample-n", type=int, default=32)
    ap.add_argument("--adjoint-tol", type=float, default=1e-8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the 50 k-triangle configs[1] / configs[2] measurement")
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
