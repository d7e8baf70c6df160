// tsl_kernels.cuh -- host-side launchers shared between the translation units of libtsl.
#pragma once
#include "tsl_internal.cuh"

namespace tsl {

// tsl_physics.cu
void launch_face_normals(tsl_ctx *ctx, const ClothDev &c, const double *pos);
void launch_energy(tsl_ctx *ctx, const double *pos, double *out_dev);
void launch_residual(tsl_ctx *ctx, const double *pos);
void launch_cloth_param_deri(tsl_ctx *ctx, const ClothDev &c, const double *pos, double *d_kb, bool zero = true);
// into_clamped: the fp32 result goes to A.val32c (multigrid hierarchy / fallback operator) instead of A.val32
void launch_hessian(tsl_ctx *ctx, const double *pos, bool f64, int spd, int sym, int newton_model, bool into_clamped = false);
void launch_cloth_deri(tsl_ctx *ctx, const ClothDev &c, const double *pos, double *d_kl, double *d_ka, double *d_kb);
void launch_friction_coef_grad(tsl_ctx *ctx, const double *pos, const double *z, int c0, int c1, double *out_dev);
void launch_elastic_force(tsl_ctx *ctx, int body, const double *pos, double *Ff);
void launch_tets_param_grad(tsl_ctx *ctx, const double *pos, const double *z, double *d_mu, double *d_lam, double *out2_dev);
void launch_hessian_counting(tsl_ctx *ctx, const double *pos, const double *z, double *zf);
void launch_axpy_pos(tsl_ctx *ctx, const double *x1, const double *p, double alpha, double *pos);
void launch_axpy2_pos(tsl_ctx *ctx, const double *x1, const double *u, double a, const double *w, double b, double *pos);
void launch_update_vel(tsl_ctx *ctx);
void launch_update_ref_angle(tsl_ctx *ctx, const ClothDev &c);
void launch_absmax(tsl_ctx *ctx, const double *a, int n, double *out_dev);
void launch_dot(tsl_ctx *ctx, const double *a, const double *b, int n, double *out_dev);
void launch_refangle_a2ax(tsl_ctx *ctx, const ClothDev &c, const double *pos, const double *ag_step, double *ag_prev, double *pg_step);
void launch_refangle_x2a(tsl_ctx *ctx, const ClothDev &c, const double *pos, const double *z, double *ag_prev);
void launch_contact_backprop(tsl_ctx *ctx, const double *pos, const double *z, double *pg_prev);
void launch_clamp(tsl_ctx *ctx, double *a, int n, double lim);
void launch_adjoint_tail(tsl_ctx *ctx, const double *z, const double *d_kb, double *pg_tm1, double *pg_tm2, double *grad_kb_accum);

// forward Newton matrices A_e -> A.val32 and A_c -> A.val32c in one pass (owner-computes cloth rows + the other bodies' elements)
void launch_hessian_newton_pair(tsl_ctx *ctx, const double *pos);

// tsl_assembly.cu: owner-computes cloth rows on the structured grid
int assembly_init(tsl_ctx *ctx);
void launch_hessian_rows(tsl_ctx *ctx, const double *pos, float *val_e, float *val_c);
void launch_residual_rows(tsl_ctx *ctx, const double *pos);
void launch_energy_rows(tsl_ctx *ctx, const double *pos, double *out_dev);

// tsl_contact.cu
int contact_alloc(tsl_ctx *ctx);
int contact_detect(tsl_ctx *ctx, const double *pos, const double *prev_pos);

// tsl_linalg.cu
int linalg_alloc(tsl_ctx *ctx);
void launch_block_jacobi64(tsl_ctx *ctx);
void launch_blend(tsl_ctx *ctx, const float *a, const float *b, float t, float *out);   // out = a + t (b - a) over the matrix values
// forward solve: PCG (fp32 vectors, fp64 reductions) on `opval` (A.val32 or A.val32c), preconditioned by the hierarchy of the
// last mg_setup; st->flags bit0 = negative curvature met (x = last iterate before it, or the preconditioned gradient)
int solve_pcg32(tsl_ctx *ctx, const float *opval, const double *rhs, double *x, double rel_tol, int max_iters, tsl_solve_stats *st);
// adjoint solve: right-preconditioned BiCGStab in fp64 on A.val64, restarted on breakdown
int solve_bicgstab64(tsl_ctx *ctx, const double *rhs, double *x, double rel_tol, int max_iters, tsl_solve_stats *st);
// adjoint solve: FGMRES(m) in fp64 on A.val64 (+ contact side blocks), flexible right preconditioning by the fp32 V-cycle
int solve_fgmres64(tsl_ctx *ctx, const double *rhs, double *x, double rel_tol, int max_iters, tsl_solve_stats *st);
int adjoint_residual64(tsl_ctx *ctx, const double *rhs, const double *x, double *res, double *rr_out);
void adjoint_apply_minv_tail(tsl_ctx *ctx, int r0, int r1, const double *in, double *out);
void adjoint_axpy64(tsl_ctx *ctx, int n, const double *dx, double *x);
int bench_pcg_iterations(tsl_ctx *ctx, int iters, int what, float *ms_out);
int precond_apply_f64io(tsl_ctx *ctx, const double *in, double *out);
int probe_curvature(tsl_ctx *ctx, const float *opval, const double *dir, double *out);
void graphs_invalidate(tsl_ctx *ctx);
int mg_setup_replay(tsl_ctx *ctx);   // mg_setup through a captured graph after the first (cold) call

// tsl_dense.cu: dense fp64 LU (direct path of the adjoint solve)
int solve_dense64(tsl_ctx *ctx, const double *rhs, double *x, tsl_solve_stats *st);
int dense_solve_host(tsl_ctx *ctx, int n, const double *A_host, const double *b_host, double *x_host);
void dense_free(tsl_ctx *ctx);

// tsl_dist.cu: collectives of the strip-partitioned solve (no-ops returning TSL_OK when the context is not partitioned)
int dist_allreduce(tsl_ctx *ctx, double *dev, int n, bool max_op = false);   // in place, on ctx->stream
int dist_halo(tsl_ctx *ctx, double *vec3);                                   // ghost rows of a [3 n_rows] vector <- the neighbours' owned rows
void launch_zero_ghost(tsl_ctx *ctx, double *vec3);                          // zero the ghost rows of a [3 n_rows] vector
void dist_destroy(tsl_ctx *ctx);                                             // releases the NCCL communicator (tsl_destroy)

// tsl_mg.cu
int mg_alloc(tsl_ctx *ctx);
void mg_free(tsl_ctx *ctx);
int mg_setup(tsl_ctx *ctx);
// first_done: the caller already applied the first Chebyshev step of level 0 (d = c D^-1 b, x[0] = d) while producing b
int mg_apply(tsl_ctx *ctx, const float *b, float *z, double *acc_bz, bool first_done = false);
void mg_first_step_targets(tsl_ctx *ctx, const float **dinv, float **d, float **x0, const float **coef);
int mg_get_level(tsl_ctx *ctx, int level, int *dims, float *lmax, float *val_host);

}  // namespace tsl
