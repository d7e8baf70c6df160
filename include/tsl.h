/*
 * tsl.h -- C ABI of libtsl.so, the B200 (sm_100a) thin-shell implicit-step engine.
 *
 * The reference (Genesis-Embodied-AI/ThinShellLab) has no FFI: its boundary is the Python object API
 * of code/engine/BaseScene.py + code/engine/analytic_grad_system.py as driven by
 * code/training/trajopt_*.py.  Each entry point below names the reference method(s) it replaces.
 * All functions return 0 on success and a negative tsl_status otherwise; tsl_last_error() gives text.
 * Pointers named *_dev are CUDA device pointers owned by the caller (torch tensors); *_host are host
 * pointers.  Work runs on a stream the library owns, ordered after everything already enqueued on the stream given to
 * tsl_set_stream (default: legacy stream 0) at entry, and that stream is made to wait for the library's work at exit.
 * One context per GPU / rank; a context is not re-entrant.
 */
#ifndef TSL_H
#define TSL_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tsl_ctx tsl_ctx;

enum tsl_status {
    TSL_OK = 0,
    TSL_ERR_INVALID = -1,       /* bad argument / call order */
    TSL_ERR_CUDA = -2,          /* a CUDA call failed */
    TSL_ERR_CAPACITY = -3,      /* max_n_constraints or grid capacity exceeded */
    TSL_ERR_UNSUPPORTED = -4,   /* configuration outside what this build implements (never silently wrong) */
    TSL_ERR_NUMERIC = -5        /* NaN / breakdown */
};

/* scene constants: BaseScene.__init__ / Scene_*.init_scene_parameters
 * (code/engine/BaseScene.py:31-53, code/task_scene/Scene_bouncing.py:37-53) */
typedef struct tsl_config {
    int struct_size;            /* = sizeof(tsl_config), ABI check */
    int n_verts;                /* tot_NV */
    double dt;                  /* self.dt == self.h */
    double k_contact, eps_contact, eps_v;
    double damping;             /* BaseScene.damping (velocity update) */
    double gravity[3];
    int max_n_constraints;      /* BaseScene.max_n_constraints */
    double grid_h;              /* geometry.grid_h   (code/engine/geometry.py:8), reference value 0.003 */
    int grid_n;                 /* geometry.grid_n   (:9), reference value 132 */
} tsl_config;

typedef struct tsl_step_stats {
    int newton_iters;           /* Newton iterations taken                      */
    int linear_iters;           /* total Krylov iterations                      */
    int linesearch_evals;       /* energy evaluations inside the line searches  */
    int n_contacts;             /* nc after contact_analysis                    */
    int converged;              /* 1 if delta < tol                             */
    int flags;                  /* bit0: negative curvature met in PCG; bit1: Krylov hit its cap; bit2: a step was accepted below the
                                   resolution of the energy sum (plain Newton step near convergence) */
    double delta;               /* last |p|_inf / h  (BaseScene.newton_step return value) */
    double energy;              /* E at the accepted point                      */
    double ms_contact, ms_assembly, ms_solve, ms_linesearch; /* host wall clock per phase (diagnostic) */
} tsl_step_stats;

typedef struct tsl_solve_stats {
    int iters;                  /* Krylov iterations; 0 = the adjoint system was solved directly (dense LU, see TSL_OPT_DIRECT_MAX_DOF) */
    int flags;                  /* bit0 breakdown / stall, bit1 iteration cap hit, bit3 the adjoint solve fell back from the multigrid to
                                   the block-Jacobi preconditioner (and converged unless bit0 / bit1 are set too) */
    double rel_residual;        /* forward PCG: |b - A x|_2 / |b|_2 of the recurrence; adjoint: of the TRUE residual */
} tsl_solve_stats;

/* ---- lifetime ---------------------------------------------------------------------------- */
int tsl_create(const tsl_config *cfg, tsl_ctx **out);
int tsl_destroy(tsl_ctx *ctx);
const char *tsl_last_error(tsl_ctx *ctx);
const char *tsl_version(void);
int tsl_set_stream(tsl_ctx *ctx, void *cuda_stream);

/* ---- scene description (cold path) --------------------------------------------------------- */
/* Cloth(N, dt, Len, tot_NV, rho, offset, is_square, M) + Cloth.init_mesh
 * (code/engine/model_fold_offset.py:11-106, 929-1018).  Returns the cloth id (>= 0).
 * ref_angle_dev: [2*N*M][3] f64 plastic rest angles, owned by the caller (Cloth.ref_angle).
 * Up to 4 cloths per context (Scene_card / Scene_sliding stack three: code/task_scene/Scene_card.py:60-63); the multigrid hierarchy of
 * the forward solve is built on the first cloth's grid, the others are smoothed on the fine level only. */
int tsl_add_cloth(tsl_ctx *ctx, int N, int M, int v_offset, double dx, double rho,
                  double Kl, double Ka, double Kb, double k_angle, double *ref_angle_dev);
/* Cloth.Kl/Ka/Kb/k_angle[None] = ...  (scripts set sys.cloths[0].Kb[None], trajopt_bouncing.py:46) */
int tsl_set_cloth_params(tsl_ctx *ctx, int cloth, double Kl, double Ka, double Kb, double k_angle);
/* Cloth.update_ref_angle / init_ref_angle at the bound positions (code/engine/model_fold_offset.py:176-185, 788-797): plastic flow of
 * the rest angles beyond k_angle.  time_step does this at the end of every step; scene construction calls it once on the folded strip. */
int tsl_cloth_update_ref_angle(tsl_ctx *ctx, int cloth);
/* topology read-back for tests / renderers: Cloth.f2v, counter_face, counter_point ([2NM][3] i32, host) */
int tsl_get_cloth_topology(tsl_ctx *ctx, int cloth, int *f2v_host, int *counter_face_host, int *counter_point_host);

/* Test hook for the one canonicalised decision of this library: Cloth.judge_angle / compute_angle
 * (code/engine/model_fold_offset.py:116,135,144) evaluate a side test on neighbour entries the reference's mesher mis-wires; there the
 * exact value is 0 and the reference's outcome is the sign of rounding noise.  The library treats those tests as "not negative";
 * ov_host [2NM][3] (1 = negative, else not negative; NULL = canonical rule) lets a parity test inject the outcomes of a reference run
 * so that every other term is compared on equal footing. */
int tsl_set_side_test_override(tsl_ctx *ctx, int cloth, const signed char *ov_host);

/* Elastic bodies: the neo-Hookean box (kind 0; code/engine/model_elastic_offset.py:11-93, init_pos :240-253) and the tactile pad / ball
 * (kind 1; code/engine/model_elastic_tactile.py:13-80, init_pos :215-229).  tets_host [n_cells][4] body-local vertex ids
 * (Elastic.F_vertices), B_host [n_cells][3][3] = F_B (inverse rest Ds), W_host [n_cells] = F_W (rest volume); mu, lam, alpha as the
 * reference's fields; gravity_host [3] = Elastic.gravity (NULL: the scene's; effector pads carry none, BaseScene.py:371-374).
 * Vertex masses come from the bound mass array.  Returns the body id (>= 0). */
int tsl_add_tets(tsl_ctx *ctx, int kind, int v_offset, int n_verts, int n_cells, const int *tets_host, const double *B_host,
                 const double *W_host, double mu, double lam, double alpha, const double *gravity_host);
/* Elastic.mu / lam [None] = ... */
int tsl_set_tet_params(tsl_ctx *ctx, int body, double mu, double lam);

/* BaseScene.faces + body_list (code/engine/BaseScene.py:81,91-99, init_faces :355-359):
 * faces_host [tot_nf][3] global vertex ids, bodies_host [n_bodies][4] = v_start, v_end, f_start, f_end. */
int tsl_set_surfaces(tsl_ctx *ctx, const int *faces_host, int tot_nf, const int *bodies_host, int n_bodies);
/* one contact_pair_analysis(b_idx, v_start, v_end, mu) line of a scene's contact_analysis()
 * (code/task_scene/Scene_bouncing.py:91-96, code/engine/BaseScene.py:778-816) */
int tsl_add_contact_pair(tsl_ctx *ctx, int surface_body, int v_start, int v_end, double mu);
int tsl_set_contact_mu(tsl_ctx *ctx, int pair, double mu);   /* sys.mu_cloth_elastic[None] = ... */

/* BaseScene.pos / prev_pos / vel ([n_verts][3] f64), mass [n_verts] f64, frozen [3 n_verts] i32,
 * border_flag [n_verts] i32 (may be NULL = all zero): borrowed device pointers */
int tsl_bind_state(tsl_ctx *ctx, double *pos_dev, double *prev_pos_dev, double *vel_dev,
                   const double *mass_dev, const int *frozen_dev, const int *border_flag_dev);
/* builds the block-sparse pattern and scratch; call once after the scene is described */
int tsl_finalize(tsl_ctx *ctx);
/* BaseScene.reset(): proj_flag.fill(0) (code/engine/BaseScene.py:268) -- sticky contact sides */
int tsl_reset_contact_state(tsl_ctx *ctx);

/* ---- hot path -------------------------------------------------------------------------------- */
/* BaseScene.calc_vn + geometry.projection_query + Scene.contact_analysis
 * (code/engine/BaseScene.py:837-850, code/engine/geometry.py:223-229, Scene_bouncing.py:91-96).
 * Uses the bound pos / prev_pos.  n_contacts_out may be NULL. */
int tsl_contact_detect(tsl_ctx *ctx, int *n_contacts_out);
/* BaseScene.compute_energy (code/engine/BaseScene.py:427-451) at the bound pos */
int tsl_energy(tsl_ctx *ctx, double *energy_out);
/* BaseScene.compute_residual_and_Hessian(spd) / compute_Hessian(spd) (code/engine/BaseScene.py:976-1052).
 * flags: bit0 residual into F, bit1 Hessian, bit2 spd projection (forward), bit3 symmetrise element
 * blocks, bit4 store Hessian in fp64 (adjoint) instead of fp32. */
#define TSL_ASM_RESIDUAL 1
#define TSL_ASM_HESSIAN 2
#define TSL_ASM_SPD 4
#define TSL_ASM_SYM 8
#define TSL_ASM_F64 16
/* bit5: assemble the engine's own forward Newton matrix (exact membrane Hessian + Gauss-Newton bending, DESIGN.md
 * section 4) instead of the reference's formulas; with it, TSL_ASM_SPD means "clamp the indefinite pieces". */
#define TSL_ASM_NEWTON 32
int tsl_assemble(tsl_ctx *ctx, int flags);
/* SparseMatrix.solve (code/engine/sparse_solver.py:85-105): x = H^-1 b with the last assembled Hessian.
 * fp32 Hessian (forward Newton matrix): PCG with fp64 vectors, preconditioned by one geometric-multigrid V-cycle over the cloth grid
 * (TSL_OPT_PRECOND = 0: block-Jacobi).  fp64 Hessian (the reference's un-projected, non-symmetric adjoint matrix): dense LU with
 * partial pivoting up to TSL_OPT_DIRECT_MAX_DOF unknowns, FGMRES(m) with the same V-cycle as flexible right preconditioner above;
 * returns TSL_ERR_NUMERIC when the requested tolerance is not reached (a direct solver never leaves that to the caller).
 * rhs_dev / x_dev: [3 n_verts] f64 device. */
int tsl_solve(tsl_ctx *ctx, const double *rhs_dev, double *x_dev, double rel_tol, int max_iters, tsl_solve_stats *stats);
/* BaseScene.time_step(f_contact, frame) with Scene_bouncing.timestep_finish
 * (code/engine/BaseScene.py:1327-1370, code/task_scene/Scene_bouncing.py:115-119) */
int tsl_step_forward(tsl_ctx *ctx, int max_newton, double tol, tsl_step_stats *stats);
/* same, end to end from HOST buffers: copies pos/vel host->device into the bound tensors, steps,
 * copies pos/vel back (the reference's Field.from_numpy / to_numpy around time_step) */
int tsl_step_forward_host(tsl_ctx *ctx, double *pos_host, double *vel_host, int max_newton, double tol, tsl_step_stats *stats);

/* analytic_grad_system.Grad.transfer_grad(step, sys, f_contact) (code/engine/analytic_grad_system.py:115-160)
 * for one step t.  Device pointers, f64:
 *   x_t, x_tm1            [n_verts][3]  pos_buffer[t], pos_buffer[t-1]
 *   ref_angle_tm1         [NF][3]       ref_angle_buffer[t-1]; with several cloths: [cloth][NF_c][3], one after the other, likewise
 *                                       angleref_grad_* (the layout of Grad.ref_angle_buffer[t] / angleref_grad[t])
 *   pos_grad_t/tm1/tm2    [n_verts][3]  pos_grad[t] (clamped + updated in place), pos_grad[t-1], pos_grad[t-2] (NULL if t < 2)
 *   angleref_grad_t/tm1   [NF][3]
 *   grad_kb_accum_dev     [1]           += sum_free z * d_kb   (Grad.get_parameters_grad :69-79; d_kb over every cloth, each with its
 *                                       own Kb: Scene_card.get_paramters_grad)
 *   z_out_dev             [3 n_verts]   adjoint solution (may be NULL)
 * clamp: +-1 (system-ID Grad, :104-108) or +-1000 (trajectory Grad). */
int tsl_step_backward(tsl_ctx *ctx, const double *x_t, const double *x_tm1, const double *ref_angle_tm1,
                      double *pos_grad_t, double *pos_grad_tm1, double *pos_grad_tm2,
                      const double *angleref_grad_t, double *angleref_grad_tm1,
                      double *grad_kb_accum_dev, double *z_out_dev, double clamp, double rel_tol, int max_iters,
                      tsl_solve_stats *stats);

/* Grad.transfer_grad of the trajectory optimiser (code/engine/analytic_grad_single.py:217-255): tsl_step_backward plus
 *   z_frozen_out_dev [3 n_verts]  BaseScene.tmp_z_frozen of the second, "counting" assembly (code/engine/BaseScene.py:399-405;
 *                                 analytic_grad_single.py:240-243); may be NULL
 * grad_kb_accum_dev may be NULL (no system-ID gradient); clamp_angleref > 0 also clamps angleref_grad_t in place
 * (clamp_grad, analytic_grad_single.py:177-185: +-1000 for both). */
int tsl_step_backward_ex(tsl_ctx *ctx, const double *x_t, const double *x_tm1, const double *ref_angle_tm1,
                         double *pos_grad_t, double *pos_grad_tm1, double *pos_grad_tm2,
                         double *angleref_grad_t, double *angleref_grad_tm1,
                         double *grad_kb_accum_dev, double *z_out_dev, double *z_frozen_out_dev, double clamp, double clamp_angleref,
                         double rel_tol, int max_iters, tsl_solve_stats *stats);

/* Material-parameter sensitivities of the elastic bodies at the bound positions: d_mu_dev, d_lam_dev [n_verts][3] f64 (overwritten) =
 * Elastic.compute_deri pushed up into BaseScene.d_mu / d_lam (code/engine/model_elastic_offset.py:423-438,
 * code/engine/model_elastic_tactile.py:329-347, code/engine/BaseScene.py:1523-1525); if z_dev [3 n_verts] is given, out2_host [2] =
 * sum over the free DOFs of z d_mu and z d_lam = one step's contribution to Grad.grad_mu / grad_lam
 * (code/engine/analytic_grad_system.py:69-75).  Call after tsl_step_backward* of the same step (positions x_t are still bound). */
int tsl_elastic_param_grad(tsl_ctx *ctx, const double *z_dev, double *d_mu_dev, double *d_lam_dev, double *out2_host);

/* Cloth.compute_deri (code/engine/model_fold_offset.py:1083-1129) pushed up into BaseScene.d_kl / d_ka / d_kb (BaseScene.get_paramters_grad,
 * code/engine/BaseScene.py:1513-1521): dF/dKl, dF/dKa, dF/dKb per vertex at the bound positions, [n_verts][3] f64 device (rows outside
 * the cloth are zeroed); any of the three pointers may be NULL. */
int tsl_cloth_param_deri(tsl_ctx *ctx, int cloth, double *d_kl_dev, double *d_ka_dev, double *d_kb_dev);
/* Friction-coefficient gradient of one adjoint step (Scene.contact_energy_backprop_friction, code/task_scene/Scene_sliding.py:140-177, called
 * from Grad.transfer_grad, code/engine/analytic_grad_system.py:150-151): sum over the constraints of the contact pairs [pair_begin, pair_end)
 * of the current set, and over their free DOFs, of z * w1 * (k f1(|u|) T^T u) / mu.  z_dev [3 n_verts] = the adjoint solution of the step. */
int tsl_friction_coef_grad(tsl_ctx *ctx, const double *z_dev, int pair_begin, int pair_end, double *out_host);
/* Elastic.get_force of one tetrahedral body at the bound positions (code/engine/model_elastic_tactile.py:145-164,
 * model_elastic_offset.py:187-210): F_f [body verts][3] f64 device = internal force + m g (what gather_force / check_early_stop read,
 * code/engine/BaseScene.py:1542-1585). */
int tsl_elastic_force(tsl_ctx *ctx, int body, double *Ff_dev);

/* Kinematic boundary of a pad (code/engine/gripper_single.py): gripper.get_vert_pos + update_bound + the scene's pushup
 * (:79-83, 157-161; Scene_folding.action, code/task_scene/Scene_folding.py:213-224) for the n_bound driven vertices of the body at
 * v_offset: pos[v_offset + bound_idx[i]] = p + R F_x[bound_idx[i]].  bound_idx_dev [n_bound] i32, Fx_dev [body verts][3] f64 (device),
 * pos3_host [3], rotmat9_host [9] fp32 row-major (the reference keeps rotmat in an f32 field). */
int tsl_gripper_apply(tsl_ctx *ctx, int v_offset, int n_bound, const int *bound_idx_dev, const double *Fx_dev,
                      const double *pos3_host, const float *rotmat9_host);
/* gripper.gather_grad (code/engine/gripper_single.py:133-150): out6_host = (d_pos, d_angle) = mean over the bound vertices of
 * tmp_z_frozen and of (R F_x) x tmp_z_frozen, clamped to +-clamp_pos / +-clamp_angle (10 / 100 in the reference). */
int tsl_gripper_gather(tsl_ctx *ctx, const double *z_frozen_dev, int v_offset, int n_bound, const int *bound_idx_dev,
                       const double *Fx_dev, const float *rotmat9_host, double clamp_pos, double clamp_angle, double *out6_host);

/* ---- strip partition over the GPUs of one node (forward step) ------------------------------------------------
 * One process and one context per GPU.  The context of rank r describes the grid rows [first owned - ghost_lo_rows, last owned +
 * ghost_hi_rows] of a longer sheet (cloth rows are contiguous in vertex numbering; 2 ghost rows on every inner side, 0 on the outer
 * sides of the first / last rank) in its own coordinate frame, plus whatever frozen bodies lie under that strip.  After tsl_dist_init,
 * tsl_step_forward / tsl_energy / tsl_solve (fp32 PCG) act on the global sheet: assembly is local (every element touching an owned
 * vertex is present), the preconditioner is the local multigrid cycle (block-Jacobi over the strips), and NCCL carries the ghost rows of
 * the PCG direction (send/recv with the two neighbours per iteration), the PCG scalars (p.Ap; (r.z, |r|^2) as one message) and the
 * Newton driver's energy / |p|_inf / F.p.  The reference has no multi-GPU path; this is north_star's partition.
 * tsl_dist_unique_id: 128-byte ncclUniqueId made by rank 0, to be broadcast by the caller (torch.distributed). */
int tsl_dist_unique_id(void *out128_host);
/* Geometry rules, checked (TSL_ERR_INVALID): ghost rows are 2 on an inner side and 0 on the outer side of the first / last strip; a strip
 * owns at least 2 rows; first_row_global -- the global grid row of the context's local row 0 (ghost rows included) -- is EVEN (the
 * mesher picks triangle diagonals from row parity, so an odd start would triangulate differently from the single-GPU sheet); all strips
 * have the same row length and tile the sheet without gaps (verified across ranks at init, so that the halo exchange cannot hang).
 * Bodies other than the cloth are rank-local (each rank describes the part of the table under its strip): a body replicated on several
 * ranks would be counted once per rank in the all-reduced energy. */
int tsl_dist_init(tsl_ctx *ctx, const void *id128_host, int rank, int world, int ghost_lo_rows, int ghost_hi_rows, int first_row_global);
int tsl_dist_stats(tsl_ctx *ctx, long long *halo_msgs_out, long long *allreduces_out);

/* ---- introspection used by the parity tests and the benchmark ---------------------------------- */
int tsl_get_residual(tsl_ctx *ctx, double *F_host);                     /* BaseScene.F [3 n_verts] */
int tsl_get_matrix_nnzb(tsl_ctx *ctx, int *nnzb_out);                   /* number of 3x3 blocks (unpadded) */
/* last assembled Hessian as block-CSR on the host: rowptr [n_verts+1], colidx [nnzb], val [nnzb][3][3] f64 */
int tsl_get_matrix(tsl_ctx *ctx, int *rowptr_host, int *colidx_host, double *val_host);
/* blocks of the last assembled Hessian that are not in the static pattern: the 12 off-diagonal 3x3 blocks of every constraint
 * against a triangle with free vertices (kept in a side buffer, applied after the sliced-ELL pass).  n_out = 12 * constraints
 * (0 when every contact surface is frozen); rows_host, cols_host [n_out] vertex ids, val_host [n_out][3][3] f64; may be NULL to
 * query n_out only. */
int tsl_get_contact_blocks(tsl_ctx *ctx, int *n_out, int *rows_host, int *cols_host, double *val_host);
/* contact candidates / constraints: proj_* rows of one surface body; const_* of the current set */
int tsl_get_projection(tsl_ctx *ctx, int surface_body, int *flag_host, int *dir_host, int *idx_host, double *w_host);
int tsl_get_constraints(tsl_ctx *ctx, int *n_out, int *idx_host, double *w_host, double *k_host, double *dx0_host,
                        double *T_host, double *n_host);
/* dimensions of the solver, for roofline arithmetic */
typedef struct tsl_sizes {
    int n_verts, n_tris, n_hinges, nnzb, nnzb_padded, n_contacts;
    long long bytes_matrix_f32, bytes_matrix_f64;
    int n_solve, nnzb_solve;    /* rows / blocks of the forward solve (trailing fully frozen bodies are skipped) */
} tsl_sizes;
int tsl_get_sizes(tsl_ctx *ctx, tsl_sizes *out);
/* benchmark hooks: run `iters` PCG iterations (no convergence test) on the last fp32 Hessian, and time
 * one kernel class with CUDA events on the context's stream; ms_out = average per launch.
 * what: 0 = PCG iteration (SpMV + vector kernels + preconditioner), 1 = SpMV only, 2 = energy, 3 = residual, 4 = Hessian (BOTH
 * forward Newton matrices, exact + clamped, as one Newton iteration assembles them), 5 = preconditioner application (one V-cycle),
 * 6 = multigrid setup (Galerkin products + eigenvalue estimates), 7 = one forward Newton matrix through the element-scatter kernels */
int tsl_bench_kernel(tsl_ctx *ctx, int what, int iters, float *ms_out);
/* solver options (the reference has none: its solve is a direct factorisation, code/engine/sparse_solver.py:85-105).
 * TSL_OPT_PRECOND: 0 = block-Jacobi, 1 = geometric multigrid V-cycle over the cloth grid (default);
 * TSL_OPT_MG_*: Chebyshev smoother degree (default 2), coarsest-grid sweep degree (8), eigenvalue interval ratio (8),
 * safety factor on the power-iteration estimate of lambda_max (1.2);
 * TSL_OPT_NEWTON_MODE: what the forward Newton iteration does when PCG meets negative curvature in the exact matrix --
 *   0 = redo the step with the clamped (projected) matrix and skip the exact attempt for a few iterations (the path
 *       closest to the reference's projected Newton), 1 = move along the direction of negative curvature and keep
 *       the multigrid hierarchy for several iterations (fewer iterations on buckling sheets; may settle in another
 *       local minimum than the reference's path), 2 = solve with the blend A_e + theta (A_c - A_e), theta the smallest of
 *       0, 1/16, ..., 1 that PCG accepts (default: fewest iterations measured; reproduces the reference's Scene_bouncing
 *       rollout like the others; on buckling steps it may settle in another local minimum than mode 0);
 * TSL_OPT_GRAPHS: 1 = replay the solver iterations as captured CUDA graphs (default), 0 = eager launches.
 * TSL_OPT_ADJOINT_SOLVER: 0 = automatic (dense LU when 3 * solved vertices <= TSL_OPT_DIRECT_MAX_DOF, default 12288, else FGMRES),
 *   1 = dense LU, 2 = FGMRES(TSL_OPT_GMRES_M, default 50), 3 = BiCGStab (the round-1 solver, kept for comparison).
 * TSL_OPT_FAST_ASSEMBLY: 1 (default for single-cloth scenes) = the cloth rows of the forward Newton matrices come from the
 *   owner-computes grid kernel (no atomics, deterministic, exact + clamped matrix in one pass); 0 = element scatter with atomics;
 *   2 = also the fp64 residual and energy by tiles (deterministic; slower than the element kernels, which are fp64-math bound). */
enum tsl_option { TSL_OPT_PRECOND = 0, TSL_OPT_MG_DEGREE = 1, TSL_OPT_MG_COARSE_DEGREE = 2, TSL_OPT_MG_RATIO = 3, TSL_OPT_MG_SAFETY = 4, TSL_OPT_GRAPHS = 5, TSL_OPT_NEWTON_MODE = 6,
                  TSL_OPT_ADJOINT_SOLVER = 7, TSL_OPT_DIRECT_MAX_DOF = 8, TSL_OPT_GMRES_M = 9, TSL_OPT_FAST_ASSEMBLY = 10 };
int tsl_set_option(tsl_ctx *ctx, int key, double value);
/* multigrid level read-back for tests: dims_host[3] = n0, n1, number of levels; lmax_host[1]; val_host [25][9][n0*n1] f32
 * (5x5 stencil of 3x3 blocks, slot-major; level 0 returns the stencil copy of the cloth block).  Any pointer may be NULL. */
int tsl_mg_get_level(tsl_ctx *ctx, int level, int *dims_host, float *lmax_host, float *val_host);
/* test hook of the dense LU behind the direct adjoint solve: x = A^-1 b for a host matrix (column-major [n][n]), factorised on the GPU */
int tsl_dense_solve_host(tsl_ctx *ctx, int n, const double *A_host, const double *b_host, double *x_host);
/* z = M^-1 b with the preconditioner built by the last tsl_assemble / step (b, z: [3 n_verts] f64 device) */
int tsl_precond_apply(tsl_ctx *ctx, const double *b_dev, double *z_dev);
/* number of kernels this library has launched since creation (bench.py's gpu_launches) */
long long tsl_launch_count(tsl_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* TSL_H */
