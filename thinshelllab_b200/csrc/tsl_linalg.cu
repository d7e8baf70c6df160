// tsl_linalg.cu -- block-sparse linear algebra of the implicit step (sm_100a).
//
// Replaces SparseMatrix.solve (code/engine/sparse_solver.py:85-105; cuSOLVER sparse QR through CuPy) with
//   * forward Newton:  block-Jacobi PCG on the fp32 sliced-ELL matrix (fp32 vectors, fp64 reductions),
//   * adjoint:         block-Jacobi BiCGStab in fp64 on the un-projected, non-symmetric reference Hessian.
// The whole iteration runs without the host: step lengths are formed on the device from reduction results
// kept in a KrylovScalars struct; the host only polls |r|^2 every `check_every` iterations.
// HBM-bound: the SpMV streams 40 B (fp32) / 76 B (fp64) per 3x3 block in full-line transactions
// (see SellMatrix), gathers the direction vector through L1/L2, and fuses the dot products it feeds.
#include "tsl_internal.cuh"
#include "tsl_kernels.cuh"

namespace tsl {

#define GRID(n, b) (unsigned)(((n) + (b) - 1) / (b))
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return TSL_ERR_CUDA; } } while (0)

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-level sum of up to two values, one atomicAdd per block and value
__device__ __forceinline__ void block_atomic_sum2(double a, double b, double *acc_a, double *acc_b)
{
    __shared__ double sa[32], sb[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    a = warp_sum_d(a); b = warp_sum_d(b);
    if (lane == 0) { sa[w] = a; sb[w] = b; }
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        a = lane < nw ? sa[lane] : 0.0; b = lane < nw ? sb[lane] : 0.0;
        a = warp_sum_d(a); b = warp_sum_d(b);
        if (lane == 0) { if (acc_a) atomicAdd(acc_a, a); if (acc_b) atomicAdd(acc_b, b); }
    }
}

// y = A x for one block row per thread; returns the 3 results in registers
template <typename T>
__device__ __forceinline__ void spmv_row(const int *__restrict__ slice_base, const int *__restrict__ colidx, const T *__restrict__ val,
                                         const T *__restrict__ x, int row, T &y0, T &y1, T &y2)
{
    int S = row >> 5, lane = row & 31;
    int b0 = slice_base[S], b1 = slice_base[S + 1];
    T a0 = 0, a1 = 0, a2 = 0;
#pragma unroll 2
    for (int b = b0; b < b1; b += 32) {
        int col = __ldg(colidx + b + lane);
        const T *v = val + (long long)b * 9 + lane;
        T x0 = x[3 * col], x1 = x[3 * col + 1], x2 = x[3 * col + 2];
        a0 += __ldg(v) * x0 + __ldg(v + 32) * x1 + __ldg(v + 64) * x2;
        a1 += __ldg(v + 96) * x0 + __ldg(v + 128) * x1 + __ldg(v + 160) * x2;
        a2 += __ldg(v + 192) * x0 + __ldg(v + 224) * x1 + __ldg(v + 256) * x2;
    }
    y0 = a0; y1 = a1; y2 = a2;
}

// y = A x;  acc_uy += u . y;  acc_yy += y . y   (u may alias x)
template <typename T>
__global__ void __launch_bounds__(256) k_spmv_dots(int n_rows, const int *__restrict__ slice_base, const int *__restrict__ colidx,
                                                   const T *__restrict__ val, const T *__restrict__ x, T *__restrict__ y,
                                                   const T *__restrict__ u, double *acc_uy, double *acc_yy, double *zero_a, double *zero_b)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row == 0) { if (zero_a) *zero_a = 0; if (zero_b) *zero_b = 0; }
    double uy = 0, yy = 0;
    if (row < n_rows) {
        T y0, y1, y2;
        spmv_row<T>(slice_base, colidx, val, x, row, y0, y1, y2);
        y[3 * row] = y0; y[3 * row + 1] = y1; y[3 * row + 2] = y2;
        if (u) uy = (double)u[3 * row] * y0 + (double)u[3 * row + 1] * y1 + (double)u[3 * row + 2] * y2;
        yy = (double)y0 * y0 + (double)y1 * y1 + (double)y2 * y2;
    }
    block_atomic_sum2(uy, yy, acc_uy, acc_yy);
}

// block-Jacobi: inverse of the diagonal 3x3 blocks
template <typename T>
__global__ void k_block_jacobi(int n_rows, const int *__restrict__ diag_pb, const T *__restrict__ val, T *minv)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    long long base = sell_addr(diag_pb[r], r & 31, 0);
    double a[9];
#pragma unroll
    for (int c = 0; c < 9; c++) a[c] = (double)val[base + c * 32];
    double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    double id = 1.0 / det;
    double inv[9] = { c00 * id, (a[2] * a[7] - a[1] * a[8]) * id, (a[1] * a[5] - a[2] * a[4]) * id,
                      c01 * id, (a[0] * a[8] - a[2] * a[6]) * id, (a[2] * a[3] - a[0] * a[5]) * id,
                      c02 * id, (a[1] * a[6] - a[0] * a[7]) * id, (a[0] * a[4] - a[1] * a[3]) * id };
#pragma unroll
    for (int c = 0; c < 9; c++) minv[9 * r + c] = (T)inv[c];
}
template <typename T>
__device__ __forceinline__ void apply_minv(const T *__restrict__ minv, int r, T r0, T r1, T r2, T &z0, T &z1, T &z2)
{
    const T *m = minv + 9 * r;
    z0 = m[0] * r0 + m[1] * r1 + m[2] * r2;
    z1 = m[3] * r0 + m[4] * r1 + m[5] * r2;
    z2 = m[6] * r0 + m[7] * r1 + m[8] * r2;
}

// ------------------------------------------------------------------------------------------------ PCG (fp32)
// Preconditioned CG with the preconditioner applied between kernels (mg_apply: V-cycle or block-Jacobi):
//   init      : x = 0, r = b, rr[0] = r.r                         z = M r, rz[0] = r.z (fused in the preconditioner)   p = z
//   iteration : q = A p, pq[c] = p.q | x += a p, r -= a q, rr[n] | z = M r, rz[n] = r.z                                | p = z + (rz[n]/rz[c]) p
// c = it & 1, n = c ^ 1.  Step lengths are formed on the device; the host reads the scalars once per iteration
// (an iteration is a whole V-cycle, so the read-back is noise) to test convergence / negative curvature.
__global__ void __launch_bounds__(256) k_pcg_init(int n_rows, int n_alloc, const double *__restrict__ b, float *x, float *r, KrylovScalars *ks)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    double rr = 0;
    if (row < n_alloc) {
        float r0 = 0, r1 = 0, r2 = 0;
        if (row < n_rows) { r0 = (float)b[3 * row]; r1 = (float)b[3 * row + 1]; r2 = (float)b[3 * row + 2]; }
        x[3 * row] = x[3 * row + 1] = x[3 * row + 2] = 0.f;
        r[3 * row] = r0; r[3 * row + 1] = r1; r[3 * row + 2] = r2;
        rr = (double)r0 * r0 + (double)r1 * r1 + (double)r2 * r2;
    }
    block_atomic_sum2(rr, 0.0, &ks->acc_rr[0], nullptr);
}
// alpha = rz[c] / pq[c];  x += alpha p;  r -= alpha q;  rr[n] += r.r
__global__ void __launch_bounds__(256) k_pcg_update(int n_rows, int it, const float *__restrict__ p, const float *__restrict__ q,
                                                    float *x, float *r, KrylovScalars *ks)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    int cur = it & 1, nxt = cur ^ 1;
    double pq = ks->acc_pq[cur], rzc = ks->acc_rz[cur];
    // negative curvature or breakdown freezes the iterate (truncated Newton).  This kernel only READS the flag
    // (k_pcg_direction publishes it), so every thread takes the same branch.
    const bool bad = !(pq > 0.0);
    const bool frozen = ((ks->flags & 1) != 0) || bad;
    if (bad && it == 0 && row < n_rows) {   // no progress yet: fall back to the preconditioned gradient direction
        x[3 * row] = p[3 * row]; x[3 * row + 1] = p[3 * row + 1]; x[3 * row + 2] = p[3 * row + 2];
    }
    if (row == 0) ks->acc_pq[nxt] = 0;      // consumed by iteration it-1, next written by iteration it+1
    if (frozen) {                           // keep the scalars of the frozen state so later iterations are no-ops
        if (row == 0) { ks->acc_rr[nxt] = ks->acc_rr[cur]; }
        return;
    }
    double rr = 0;
    if (row < n_rows) {
        float alpha = (float)(rzc / pq);
        float r0 = r[3 * row] - alpha * q[3 * row], r1 = r[3 * row + 1] - alpha * q[3 * row + 1], r2 = r[3 * row + 2] - alpha * q[3 * row + 2];
        x[3 * row] += alpha * p[3 * row]; x[3 * row + 1] += alpha * p[3 * row + 1]; x[3 * row + 2] += alpha * p[3 * row + 2];
        r[3 * row] = r0; r[3 * row + 1] = r1; r[3 * row + 2] = r2;
        rr = (double)r0 * r0 + (double)r1 * r1 + (double)r2 * r2;
    }
    block_atomic_sum2(rr, 0.0, &ks->acc_rr[nxt], nullptr);
}
// p = z + beta p with beta = rz[n] / rz[c]   (it < 0: p = z)
__global__ void __launch_bounds__(256) k_pcg_direction(int n, int it, const float *__restrict__ z, float *p, KrylovScalars *ks)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (it < 0) { if (i < n) p[i] = z[i]; return; }
    int cur = it & 1, nxt = cur ^ 1;
    const bool bad = !(ks->acc_pq[cur] > 0.0);     // same test as k_pcg_update (acc_pq[cur] is still intact)
    if (bad && i == 0) atomicOr(&ks->flags, 1);
    if ((ks->flags & 1) || bad) {
        if (i == 0) ks->acc_rz[nxt] = ks->acc_rz[cur];
        return;
    }
    float beta = (float)(ks->acc_rz[nxt] / ks->acc_rz[cur]);
    if (i < n) p[i] = z[i] + beta * p[i];
}
__global__ void k_f32_to_f64(int n, const float *__restrict__ a, double *b)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (double)a[i];
}
__global__ void k_f64_to_f32(int n, int n_alloc, const double *__restrict__ a, float *b)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_alloc) b[i] = i < n ? (float)a[i] : 0.f;
}

int linalg_alloc(tsl_ctx *ctx)
{
    int nr = ctx->A.n_slices * 32;
    size_t nb = sizeof(float) * 3 * (size_t)nr;
    CK(cudaMalloc(&ctx->cg_x, nb)); CK(cudaMalloc(&ctx->cg_r, nb)); CK(cudaMalloc(&ctx->cg_z, nb));
    CK(cudaMalloc(&ctx->cg_p, nb)); CK(cudaMalloc(&ctx->cg_q, nb)); CK(cudaMalloc(&ctx->cg_r64tmp, nb));
    CK(cudaMemset(ctx->cg_p, 0, nb)); CK(cudaMemset(ctx->cg_x, 0, nb)); CK(cudaMemset(ctx->cg_r, 0, nb));
    CK(cudaMemset(ctx->cg_z, 0, nb)); CK(cudaMemset(ctx->cg_q, 0, nb)); CK(cudaMemset(ctx->cg_r64tmp, 0, nb));
    CK(cudaMalloc(&ctx->minv32, sizeof(float) * 9 * (size_t)nr));
    CK(cudaMalloc(&ctx->ks, sizeof(KrylovScalars)));
    CK(cudaMallocHost(&ctx->ks_host, sizeof(KrylovScalars)));
    CK(cudaMalloc(&ctx->sol, sizeof(double) * 3 * (size_t)nr));
    return TSL_OK;
}

void launch_block_jacobi64(tsl_ctx *ctx)
{
    int n = ctx->cfg.n_verts;
    k_block_jacobi<double><<<GRID(n, 256), 256, 0, ctx->stream>>>(n, ctx->A.diag_pb, ctx->A.val64, ctx->minv64);
    ctx->launches++;
}

static int pcg_iteration(tsl_ctx *ctx, const float *opval, int it)
{
    int n = ctx->cfg.n_verts;
    const SellMatrix &A = ctx->A;
    KrylovScalars *ks = ctx->ks;
    int cur = it & 1, nxt = cur ^ 1;
    // q = A p, pq[cur] += p.q ; also clears rz/rr of the next parity (they were read by iteration it-1's direction update)
    k_spmv_dots<float><<<GRID(n, 256), 256, 0, ctx->stream>>>(n, A.slice_base, A.colidx, opval, ctx->cg_p, ctx->cg_q, ctx->cg_p,
                                                              &ks->acc_pq[cur], nullptr, &ks->acc_rz[nxt], &ks->acc_rr[nxt]);
    k_pcg_update<<<GRID(n, 256), 256, 0, ctx->stream>>>(n, it, ctx->cg_p, ctx->cg_q, ctx->cg_x, ctx->cg_r, ks);
    ctx->launches += 2;
    int rc = mg_apply(ctx, ctx->cg_r, ctx->cg_z, &ks->acc_rz[nxt]);
    if (rc != TSL_OK) return rc;
    k_pcg_direction<<<GRID(3 * n, 256), 256, 0, ctx->stream>>>(3 * n, it, ctx->cg_z, ctx->cg_p, ks);
    ctx->launches++;
    return TSL_OK;
}

static int pcg_start(tsl_ctx *ctx, const double *rhs)
{
    int n = ctx->cfg.n_verts, nr = ctx->A.n_slices * 32;
    cudaStream_t s = ctx->stream;
    CK(cudaMemsetAsync(ctx->ks, 0, sizeof(KrylovScalars), s));
    k_pcg_init<<<GRID(nr, 256), 256, 0, s>>>(n, nr, rhs, ctx->cg_x, ctx->cg_r, ctx->ks);
    ctx->launches++;
    int rc = mg_apply(ctx, ctx->cg_r, ctx->cg_z, &ctx->ks->acc_rz[0]);
    if (rc != TSL_OK) return rc;
    k_pcg_direction<<<GRID(3 * n, 256), 256, 0, s>>>(3 * n, -1, ctx->cg_z, ctx->cg_p, ctx->ks);
    ctx->launches++;
    return TSL_OK;
}

int solve_pcg32(tsl_ctx *ctx, const float *opval, const double *rhs, double *x, double rel_tol, int max_iters, tsl_solve_stats *st)
{
    int n = ctx->cfg.n_verts;
    cudaStream_t s = ctx->stream;
    int rc = pcg_start(ctx, rhs);
    if (rc != TSL_OK) return rc;
    CK(cudaMemcpyAsync(ctx->ks_host, ctx->ks, sizeof(KrylovScalars), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    double rr0 = ctx->ks_host->acc_rr[0];
    int it = 0, flags = 0;
    double rr = rr0;
    // the block-Jacobi iteration is three short kernels: poll less often there
    const int check_every = (ctx->precond == 0 || ctx->mg.n_levels == 0) ? 10 : 1;
    if (rr0 > 0) {
        while (it < max_iters) {
            int chunk = std::min(check_every, max_iters - it);
            for (int k = 0; k < chunk; k++, it++) { rc = pcg_iteration(ctx, opval, it); if (rc != TSL_OK) return rc; }
            CK(cudaMemcpyAsync(ctx->ks_host, ctx->ks, sizeof(KrylovScalars), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            rr = ctx->ks_host->acc_rr[it & 1];
            flags = ctx->ks_host->flags;
            if (!(rr == rr)) { ctx->err = "PCG produced NaN"; return TSL_ERR_NUMERIC; }
            if (flags & 1) break;
            if (rr <= rel_tol * rel_tol * rr0) break;
        }
        if (it >= max_iters && rr > rel_tol * rel_tol * rr0) flags |= 2;
    }
    k_f32_to_f64<<<GRID(3 * n, 256), 256, 0, s>>>(3 * n, ctx->cg_x, x);
    ctx->launches++;
    ctx->ks_host->rr0 = sqrt(rr0);      // |b|_2 for the caller's forcing term
    if (st) { st->iters = it; st->flags = flags; st->rel_residual = rr0 > 0 ? sqrt(rr / rr0) : 0.0; }
    CK(cudaGetLastError());
    return TSL_OK;
}

// what: 0 = full PCG iterations, 1 = SpMV only, 5 = V-cycle only
int bench_pcg_iterations(tsl_ctx *ctx, int iters, int what, float *ms_out)
{
    int n = ctx->cfg.n_verts;
    cudaStream_t s = ctx->stream;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // a well-defined state: r = F, p = z = M^-1 F
    int rc = pcg_start(ctx, ctx->F);
    if (rc != TSL_OK) return rc;
    CK(cudaEventRecord(e0, s));
    for (int it = 0; it < iters; it++) {
        if (what == 1) {
            k_spmv_dots<float><<<GRID(n, 256), 256, 0, s>>>(n, ctx->A.slice_base, ctx->A.colidx, ctx->A.val32, ctx->cg_p, ctx->cg_q, ctx->cg_p,
                                                            &ctx->ks->acc_pq[it & 1], nullptr, nullptr, nullptr);
            ctx->launches++;
        } else if (what == 5) {
            rc = mg_apply(ctx, ctx->cg_r, ctx->cg_z, nullptr);
            if (rc != TSL_OK) return rc;
        } else {
            rc = pcg_iteration(ctx, ctx->A.val32, it);
            if (rc != TSL_OK) return rc;
        }
    }
    CK(cudaEventRecord(e1, s));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return TSL_OK;
}

// ------------------------------------------------------------------------------------------------ BiCGStab (fp64)
// Right-preconditioned BiCGStab on the fp64 matrix (the reference's un-projected, non-symmetric Hessian):
//   p = r + beta (p - omega v);  y = M p;  v = A y;  alpha = rho / (rhat.v);  s = r - alpha v;  z = M s;  t = A z;
//   omega = (t.s)/(t.t);  dx += alpha y + omega z;  r = s - omega t
// M = the fp32 V-cycle of the clamped Newton matrix at the same state (or fp64 block-Jacobi, precond == 0).
// vectors: bi[0]=r, [1]=rhat, [2]=p, [3]=v, [4]=y, [5]=s, [6]=z, [7]=t ; the correction dx accumulates in ctx->sol,
// the outer loop adds it to x and restarts from the true residual (breakdown recovery + iterative refinement).
__global__ void __launch_bounds__(256) k_bi_init(int n, const double *__restrict__ b, double *x, double *r, double *rhat, double *p, double *v,
                                                 KrylovScalars *ks)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double rr = 0;
    if (i < n) {
        double bi = b[i];
        x[i] = 0; r[i] = bi; rhat[i] = bi; p[i] = 0; v[i] = 0;
        rr = bi * bi;
    }
    block_atomic_sum2(rr, rr, &ks->acc_rho[0], &ks->acc_rr[0]);
}
// K_a (iteration it): beta = (rho_new/rho_old)(alpha/omega); p = r + beta (p - omega v)
__global__ void __launch_bounds__(256) k_bi_a(int n, int it, const double *__restrict__ r, const double *__restrict__ v, double *p, KrylovScalars *ks)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cur = it & 1;
    if (ks->flags & 1) return;
    double beta = 0, omega = 0;
    if (it > 0) {
        double rho_new = ks->acc_rho[cur], rho_old = ks->acc_rho[cur ^ 1];
        omega = ks->acc_ts / ks->acc_tt;
        beta = (rho_new / rho_old) * (ks->alpha / omega);
    }
    if (i < n) p[i] = r[i] + beta * (p[i] - omega * v[i]);
}
// K_c: alpha = rho_new / (rhat.v); s = r - alpha v
__global__ void __launch_bounds__(256) k_bi_c(int n, int it, const double *__restrict__ r, const double *__restrict__ v, double *s, KrylovScalars *ks)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int cur = it & 1;
    if (ks->flags & 1) return;
    double rhv = ks->acc_rhv;
    if (!(rhv != 0.0) || !(rhv == rhv)) { if (i == 0) atomicOr(&ks->flags, 1); return; }
    double alpha = ks->acc_rho[cur] / rhv;
    if (i < n) s[i] = r[i] - alpha * v[i];
}
// K_e: omega = ts/tt; x += alpha y + omega z; r = s - omega t; rho[(it+1)&1] += rhat.r; rr[(it+1)&1] += r.r
__global__ void __launch_bounds__(256) k_bi_e(int n, int it, const double *__restrict__ y, const double *__restrict__ z,
                                              const double *__restrict__ s, const double *__restrict__ t, const double *__restrict__ rhat,
                                              double *x, double *r, KrylovScalars *ks)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int nxt = (it & 1) ^ 1;
    if (ks->flags & 1) {
        if (i == 0) { ks->acc_rr[nxt] = ks->acc_rr[nxt ^ 1]; ks->acc_rho[nxt] = ks->acc_rho[nxt ^ 1]; }
        return;
    }
    double tt = ks->acc_tt;
    double omega = tt > 0 ? ks->acc_ts / tt : 0.0;
    double alpha = ks->alpha;
    if (i == 0) ks->acc_rhv = 0;
    double rho = 0, rr = 0;
    if (i < n) {
        x[i] += alpha * y[i] + omega * z[i];
        double ri = s[i] - omega * t[i];
        r[i] = ri;
        rho = rhat[i] * ri; rr = ri * ri;
    }
    block_atomic_sum2(rho, rr, &ks->acc_rho[nxt], &ks->acc_rr[nxt]);
}
__global__ void k_bi_store_alpha(int it, KrylovScalars *ks)
{   // runs between K_c and the second SpMV: publishes alpha while rho/rhv are still intact
    if (ks->flags & 1) return;
    ks->alpha = ks->acc_rho[it & 1] / ks->acc_rhv;
}
// out = M in for fp64 vectors
template <typename T>
__global__ void __launch_bounds__(256) k_apply_minv64(int n_rows, const T *__restrict__ minv, const double *__restrict__ in, double *out)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    double z0, z1, z2;
    apply_minv<double>(minv, row, in[3 * row], in[3 * row + 1], in[3 * row + 2], z0, z1, z2);
    out[3 * row] = z0; out[3 * row + 1] = z1; out[3 * row + 2] = z2;
}
int precond_apply_f64io(tsl_ctx *ctx, const double *in, double *out);
static int precond64(tsl_ctx *ctx, const double *in, double *out)
{
    int n = ctx->cfg.n_verts;
    cudaStream_t s = ctx->stream;
    if (ctx->precond == 0 || ctx->mg.n_levels == 0) {
        k_apply_minv64<double><<<GRID(n, 256), 256, 0, s>>>(n, ctx->minv64, in, out);
        ctx->launches++;
        return TSL_OK;
    }
    return precond_apply_f64io(ctx, in, out);
}
// out = M in through the fp32 preconditioner of the last mg_setup (V-cycle, or fp32 block-Jacobi when precond == 0)
int precond_apply_f64io(tsl_ctx *ctx, const double *in, double *out)
{
    int n = ctx->cfg.n_verts, nr = ctx->A.n_slices * 32;
    cudaStream_t s = ctx->stream;
    k_f64_to_f32<<<GRID(3 * nr, 256), 256, 0, s>>>(3 * n, 3 * nr, in, ctx->cg_r64tmp);
    int rc = mg_apply(ctx, ctx->cg_r64tmp, ctx->cg_z, nullptr);
    if (rc != TSL_OK) return rc;
    k_f32_to_f64<<<GRID(3 * n, 256), 256, 0, s>>>(3 * n, ctx->cg_z, out);
    ctx->launches += 2;
    return TSL_OK;
}
// r = b - A x (fp64), rr = |r|^2 into acc
__global__ void __launch_bounds__(256) k_residual64(int n_rows, const int *__restrict__ slice_base, const int *__restrict__ colidx,
                                                    const double *__restrict__ val, const double *__restrict__ b, const double *__restrict__ x,
                                                    double *r, double *acc_rr)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    double rr = 0;
    if (row < n_rows) {
        double y0, y1, y2;
        spmv_row<double>(slice_base, colidx, val, x, row, y0, y1, y2);
        double r0 = b[3 * row] - y0, r1 = b[3 * row + 1] - y1, r2 = b[3 * row + 2] - y2;
        r[3 * row] = r0; r[3 * row + 1] = r1; r[3 * row + 2] = r2;
        rr = r0 * r0 + r1 * r1 + r2 * r2;
    }
    block_atomic_sum2(rr, 0.0, acc_rr, nullptr);
}
__global__ void k_axpy64(int n, const double *__restrict__ dx, double *x)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += dx[i];
}

int solve_bicgstab64(tsl_ctx *ctx, const double *rhs, double *x, double rel_tol, int max_iters, tsl_solve_stats *st)
{
    int n = ctx->cfg.n_verts, n3 = 3 * n;
    cudaStream_t s = ctx->stream;
    const SellMatrix &A = ctx->A;
    KrylovScalars *ks = ctx->ks;
    double *r = ctx->bi[0], *rhat = ctx->bi[1], *p = ctx->bi[2], *v = ctx->bi[3], *y = ctx->bi[4], *sv = ctx->bi[5], *z = ctx->bi[6], *t = ctx->bi[7];
    double *dx = ctx->sol;                 // correction of the current restart cycle
    double *res = ctx->adj_rhs;            // true residual b - A x
    CK(cudaMemsetAsync(x, 0, sizeof(double) * n3, s));
    CK(cudaMemcpyAsync(res, rhs, sizeof(double) * n3, cudaMemcpyDeviceToDevice, s));
    double rr00 = -1, rr = 0;
    int it = 0, flags = 0, restarts = 0;
    const int check_every = (ctx->precond == 0 || ctx->mg.n_levels == 0) ? 10 : 1;
    const int max_restarts = 8;
    while (true) {
        CK(cudaMemsetAsync(ks, 0, sizeof(KrylovScalars), s));
        k_bi_init<<<GRID(n3, 256), 256, 0, s>>>(n3, res, dx, r, rhat, p, v, ks);
        ctx->launches++;
        CK(cudaMemcpyAsync(ctx->ks_host, ks, sizeof(KrylovScalars), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        double rr0 = ctx->ks_host->acc_rr[0];
        if (rr00 < 0) rr00 = rr0;
        rr = rr0;
        if (!(rr0 == rr0)) { ctx->err = "BiCGStab: NaN residual"; return TSL_ERR_NUMERIC; }
        if (rr0 <= rel_tol * rel_tol * rr00 || rr00 == 0) break;
        int it_cycle = 0;
        bool broke = false;
        while (it < max_iters) {
            int chunk = std::min(check_every, max_iters - it);
            for (int k = 0; k < chunk; k++, it++, it_cycle++) {
                int cur = it_cycle & 1;
                k_bi_a<<<GRID(n3, 256), 256, 0, s>>>(n3, it_cycle, r, v, p, ks);
                int rc = precond64(ctx, p, y);
                if (rc != TSL_OK) return rc;
                // v = A y, rhv += rhat.v ; clears rho/rr of the next parity (K_a has read rho_old) and ts/tt (K_a has read omega)
                k_spmv_dots<double><<<GRID(n, 256), 256, 0, s>>>(n, A.slice_base, A.colidx, A.val64, y, v, rhat, &ks->acc_rhv, nullptr,
                                                                 &ks->acc_rho[cur ^ 1], &ks->acc_rr[cur ^ 1]);
                k_bi_c<<<GRID(n3, 256), 256, 0, s>>>(n3, it_cycle, r, v, sv, ks);
                k_bi_store_alpha<<<1, 1, 0, s>>>(it_cycle, ks);
                CK(cudaMemsetAsync(&ks->acc_ts, 0, 2 * sizeof(double), s));
                rc = precond64(ctx, sv, z);
                if (rc != TSL_OK) return rc;
                // t = A z, ts += s.t, tt += t.t
                k_spmv_dots<double><<<GRID(n, 256), 256, 0, s>>>(n, A.slice_base, A.colidx, A.val64, z, t, sv, &ks->acc_ts, &ks->acc_tt, nullptr, nullptr);
                k_bi_e<<<GRID(n3, 256), 256, 0, s>>>(n3, it_cycle, y, z, sv, t, rhat, dx, r, ks);
                ctx->launches += 6;
            }
            CK(cudaMemcpyAsync(ctx->ks_host, ks, sizeof(KrylovScalars), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            rr = ctx->ks_host->acc_rr[it_cycle & 1];
            if (!(rr == rr)) { broke = true; break; }          // dx is poisoned: drop this cycle's correction
            if (ctx->ks_host->flags & 1) break;                // breakdown: dx holds the last good iterate
            if (rr <= rel_tol * rel_tol * rr00) break;
        }
        // x += dx, true residual for the next cycle / the final report
        if (!broke) {
            k_axpy64<<<GRID(n3, 256), 256, 0, s>>>(n3, dx, x);
            ctx->launches++;
        }
        CK(cudaMemsetAsync(&ks->acc_rr[0], 0, sizeof(double), s));
        k_residual64<<<GRID(n, 256), 256, 0, s>>>(n, A.slice_base, A.colidx, A.val64, rhs, x, res, &ks->acc_rr[0]);
        ctx->launches++;
        CK(cudaMemcpyAsync(ctx->ks_host, ks, sizeof(KrylovScalars), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        rr = ctx->ks_host->acc_rr[0];
        if (!(rr == rr)) { ctx->err = "BiCGStab produced NaN"; return TSL_ERR_NUMERIC; }
        if (rr <= rel_tol * rel_tol * rr00) break;
        if (it >= max_iters) { flags |= 2; break; }
        if (++restarts > max_restarts) { flags |= 1; break; }
    }
    if (st) { st->iters = it; st->flags = flags; st->rel_residual = rr00 > 0 ? sqrt(rr / rr00) : 0.0; }
    CK(cudaGetLastError());
    return TSL_OK;
}

}  // namespace tsl
