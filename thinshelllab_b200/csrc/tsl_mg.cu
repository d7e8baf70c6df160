// tsl_mg.cu -- geometric multigrid preconditioner over the cloth's structured vertex grid (sm_100a).
//
// The reference solves every Newton / adjoint system with a sparse direct solver (cuSOLVER QR through CuPy,
// code/engine/sparse_solver.py:85-105).  The B200 path keeps the matrix block-sparse and iterates; what makes
// that competitive is this preconditioner: Cloth is always an (N+1) x (M+1) vertex grid
// (code/engine/model_fold_offset.py:929-1018), so the hierarchy is geometric --
//   * prolongation P: bilinear per coordinate (vertices with even (i, j) survive), applied per x/y/z component,
//   * coarse operators: Galerkin products P^T A P, recomputed on the device for every new matrix; a fine stencil of
//     radius 2 (triangle + hinge neighbours) gives coarse stencils of radius 2 on every level -> 5x5 blocks of 3x3,
//     stored per vertex as 225 contiguous floats and applied by one warp per vertex,
//   * smoother: Chebyshev polynomial in D^-1 A (D = 3x3 diagonal blocks) on [lmax/ratio, lmax]; lmax(D^-1 A) comes
//     from 10 power iterations per level with a safety factor (the iteration is sensitive to an under-estimate,
//     not to an over-estimate),
//   * coarsest grid (<= 6 vertices in one direction): a longer Chebyshev sweep, no direct solve.
// The V-cycle is symmetric (same polynomial before and after the coarse correction), so it is a valid PCG
// preconditioner.  Everything is fp32; all coefficients live in device memory, so setup and application never
// synchronise with the host.  Bound: every kernel streams its level's matrix once (HBM / L2), the small levels are
// launch-latency bound.
#include "tsl_internal.cuh"
#include <cuda_fp16.h>
#include "tsl_kernels.cuh"
#include "tsl_mg_kernels.cuh"

namespace tsl {

#define GRID(n, b) (unsigned)(((n) + (b) - 1) / (b))
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return TSL_ERR_CUDA; } } while (0)

// ------------------------------------------------------------------------------------------------ row operators
template <class VT>
struct SellOpT {
    const int *slice_base, *colidx;
    const VT *val;
    const float *sc;             // {scale, 1 / scale} of the stored values (NULL: stored = true value)
    __device__ __forceinline__ void mul(int row, const float *__restrict__ x, float &y0, float &y1, float &y2) const
    {
        int S = row >> 5, lane = row & 31;
        int b0 = slice_base[S], b1 = slice_base[S + 1];
        float a0 = 0, a1 = 0, a2 = 0;
#pragma unroll 2
        for (int b = b0; b < b1; b += 32) {
            int col = __ldg(colidx + b + lane);
            const VT *v = val + (long long)b * 9 + lane;
            float x0 = x[3 * col], x1 = x[3 * col + 1], x2 = x[3 * col + 2];
            a0 += mg_ld(v) * x0 + mg_ld(v + 32) * x1 + mg_ld(v + 64) * x2;
            a1 += mg_ld(v + 96) * x0 + mg_ld(v + 128) * x1 + mg_ld(v + 160) * x2;
            a2 += mg_ld(v + 192) * x0 + mg_ld(v + 224) * x1 + mg_ld(v + 256) * x2;
        }
        const float is = sc ? sc[1] : 1.f;
        y0 = a0 * is; y1 = a1 * is; y2 = a2 * is;
    }
    // the same row product by NT warps of a CTA (warp = part, lane = row of the CTA's slice): part p visits slice columns p, p + NT, ...
    // and the parts meet in shared memory; every thread of the CTA must call it (barrier inside); the result is valid for part 0
    template <int NT>
    __device__ __forceinline__ void mul_split(int row, int part, const float *__restrict__ x, float &y0, float &y1, float &y2) const
    {
        int S = row >> 5, lane = row & 31;
        int b0 = slice_base[S], b1 = slice_base[S + 1];
        float a0 = 0, a1 = 0, a2 = 0;
#pragma unroll 2
        for (int b = b0 + 32 * part; b < b1; b += 32 * NT) {
            int col = __ldg(colidx + b + lane);
            const VT *v = val + (long long)b * 9 + lane;
            float x0 = x[3 * col], x1 = x[3 * col + 1], x2 = x[3 * col + 2];
            a0 += mg_ld(v) * x0 + mg_ld(v + 32) * x1 + mg_ld(v + 64) * x2;
            a1 += mg_ld(v + 96) * x0 + mg_ld(v + 128) * x1 + mg_ld(v + 160) * x2;
            a2 += mg_ld(v + 192) * x0 + mg_ld(v + 224) * x1 + mg_ld(v + 256) * x2;
        }
        __shared__ float sp[NT][3][32];
        sp[part][0][lane] = a0; sp[part][1][lane] = a1; sp[part][2][lane] = a2;
        __syncthreads();
        if (part == 0) {
#pragma unroll
            for (int q = 1; q < NT; q++) { a0 += sp[q][0][lane]; a1 += sp[q][1][lane]; a2 += sp[q][2][lane]; }
        }
        const float is = sc ? sc[1] : 1.f;
        y0 = a0 * is; y1 = a1 * is; y2 = a2 * is;
    }
};
typedef SellOpT<float> SellOp;
// Stencil levels: one WARP per vertex.  The 225 values of a vertex row (25 slots x 3x3) are contiguous, so the warp
// streams them with 8 fully coalesced loads; lane e handles element e = slot*9 + comp of each 32-chunk, multiplies by
// the matching component of the neighbour's x and the three row sums are formed with a shuffle reduction.  (A
// thread-per-vertex walk over 25 slots is a 35-45 us latency chain even on a 36-vertex grid: measured, profiles/.)
// Slots that point outside the grid hold zeros (k_galerkin writes them), so neighbour indices are clamped, not branched.
// Two layouts, chosen per level by size (MgLevel::sv / se: element (v, e) lives at val[v*sv + e*se]):
//   small levels  (sv, se) = (225, 1): row-contiguous, one warp per vertex -- latency bound, wants few loads per thread;
//   large levels  (sv, se) = (1, nvp): element-major, one THREAD per vertex with the 25 slots fully unrolled and
//                 clamped neighbour indices -- 225 independent coalesced loads per thread, bandwidth bound.
template <class VT>
struct StencilOpT {
    const VT *val;
    int n0, n1;
    long long sv, se;
    const float *sc;             // {scale, 1 / scale} of the stored values (NULL: stored = true value)
};
typedef StencilOpT<float> StencilOp;
__device__ __forceinline__ void stencil_row_warp(const StencilOp &A, int v, int lane, const float *__restrict__ x, float &y0, float &y1, float &y2)
{
    int I = v / A.n1, J = v - I * A.n1;
    const float *row = A.val + (size_t)v * 225;      // warp kernels run on row-contiguous levels only
    float a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int e = k * 32 + lane;
        if (e < 225) {
            int slot = e / 9, c = e - slot * 9;
            int q = slot / 5;
            int ii = min(max(I + q - 2, 0), A.n0 - 1), jj = min(max(J + slot - q * 5 - 2, 0), A.n1 - 1);
            int r = c / 3;
            float t = __ldg(row + e) * x[3 * (ii * A.n1 + jj) + (c - r * 3)];
            a0 += (r == 0) ? t : 0.f; a1 += (r == 1) ? t : 0.f; a2 += (r == 2) ? t : 0.f;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    y0 = a0; y1 = a1; y2 = a2;
}

__device__ __forceinline__ void block_atomic_sum(double a, double *acc)
{
    __shared__ double sa[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) sa[w] = a;
    __syncthreads();
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        a = lane < nw ? sa[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) atomicAdd(acc, a);
    }
}

// ------------------------------------------------------------------------------------------------ smoother kernels
// d = c D^-1 b ; x_out = d                     (first Chebyshev step from a zero guess: no matrix pass)
// acc (optional) += b . x_out
__global__ void __launch_bounds__(256) k_cheb_first(int nrows, const float *__restrict__ dinv, const float *__restrict__ b, float *d,
                                                    float *x_out, const float *__restrict__ coef, double *acc)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0;
    if (row < nrows) {
        float c = coef[1];
        float r0 = b[3 * row], r1 = b[3 * row + 1], r2 = b[3 * row + 2];
        const float *m = dinv + 9 * (size_t)row;
        float d0 = c * (m[0] * r0 + m[1] * r1 + m[2] * r2), d1 = c * (m[3] * r0 + m[4] * r1 + m[5] * r2), d2 = c * (m[6] * r0 + m[7] * r1 + m[8] * r2);
        if (d) { d[3 * row] = d0; d[3 * row + 1] = d1; d[3 * row + 2] = d2; }
        x_out[3 * row] = d0; x_out[3 * row + 1] = d1; x_out[3 * row + 2] = d2;
        s = (double)r0 * d0 + (double)r1 * d1 + (double)r2 * d2;
    }
    if (acc) block_atomic_sum(s, acc);
}
// d = a d + c D^-1 (b - A x_in) ; x_out = x_in + d     (b == nullptr: b = 0; x_out == nullptr: not stored)
// acc_mode 1: acc += b . x_out      acc_mode 2: acc += d . d
// thread-per-vertex product for element-major levels
template <class VT>
__device__ __forceinline__ void stencil_row_thread(const StencilOpT<VT> &A, int v, const float *__restrict__ x, float &y0, float &y1, float &y2)
{
    int I = v / A.n1, J = v - I * A.n1;
    const VT *a = A.val + (size_t)v * A.sv;
    const long long se = A.se;
    float a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int slot = 0; slot < 25; slot++) {
        const int dI = slot / 5 - 2, dJ = slot % 5 - 2;
        int ii = min(max(I + dI, 0), A.n0 - 1), jj = min(max(J + dJ, 0), A.n1 - 1);     // out-of-grid slots hold zeros
        int u = ii * A.n1 + jj;
        float x0 = x[3 * u], x1 = x[3 * u + 1], x2 = x[3 * u + 2];
        const VT *q = a + (size_t)(slot * 9) * se;
        a0 += mg_ld(q) * x0 + mg_ld(q + se) * x1 + mg_ld(q + 2 * se) * x2;
        a1 += mg_ld(q + 3 * se) * x0 + mg_ld(q + 4 * se) * x1 + mg_ld(q + 5 * se) * x2;
        a2 += mg_ld(q + 6 * se) * x0 + mg_ld(q + 7 * se) * x1 + mg_ld(q + 8 * se) * x2;
    }
    const float is = A.sc ? A.sc[1] : 1.f;
    y0 = a0 * is; y1 = a1 * is; y2 = a2 * is;
}
// NT threads per vertex (slots part, part + NT, ...): a 354 x 354 level has only 125 k vertices = 41 % of the resident thread slots of
// 148 SMs and a 177 x 177 level 10 % -- these levels are bound by the LATENCY of the 25 dependent slot visits of a thread, not by bytes
// (ncu, 177 x 177 with two threads per vertex: 13 resident warps per SM, 27 of 31 cycles per instruction on the long scoreboard), so the
// work per thread is cut until the loads in flight cover the memory latency.  CTA = NT warps x 32 rows.
template <class VT, int NT>
__device__ __forceinline__ void stencil_row_split(const StencilOpT<VT> &A, int v, int part, const float *__restrict__ x, float &y0, float &y1, float &y2)
{
    int I = v / A.n1, J = v - I * A.n1;
    const VT *a = A.val + (size_t)v * A.sv;
    const long long se = A.se;
    float a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int q = 0; q < (25 + NT - 1) / NT; q++) {
        int slot = NT * q + part;
        if (slot < 25) {
            int dI = slot / 5 - 2, dJ = slot - (slot / 5) * 5 - 2;
            int ii = min(max(I + dI, 0), A.n0 - 1), jj = min(max(J + dJ, 0), A.n1 - 1);
            int u = ii * A.n1 + jj;
            float x0 = x[3 * u], x1 = x[3 * u + 1], x2 = x[3 * u + 2];
            const VT *p = a + (size_t)(slot * 9) * se;
            a0 += mg_ld(p) * x0 + mg_ld(p + se) * x1 + mg_ld(p + 2 * se) * x2;
            a1 += mg_ld(p + 3 * se) * x0 + mg_ld(p + 4 * se) * x1 + mg_ld(p + 5 * se) * x2;
            a2 += mg_ld(p + 6 * se) * x0 + mg_ld(p + 7 * se) * x1 + mg_ld(p + 8 * se) * x2;
        }
    }
    // warp `part` of the CTA holds the partial sums of the CTA's 32 rows (lane = row): consecutive lanes read consecutive elements, so
    // every load is a full line whatever NT is; the parts meet in shared memory
    __shared__ float sp[NT][3][32];
    const int lane = threadIdx.x & 31;
    sp[part][0][lane] = a0; sp[part][1][lane] = a1; sp[part][2][lane] = a2;
    __syncthreads();
    if (part == 0) {
#pragma unroll
        for (int q = 1; q < NT; q++) { a0 += sp[q][0][lane]; a1 += sp[q][1][lane]; a2 += sp[q][2][lane]; }
    }
    const float is = A.sc ? A.sc[1] : 1.f;
    y0 = a0 * is; y1 = a1 * is; y2 = a2 * is;
}
template <class VT, int NT>
__global__ void __launch_bounds__(32 * NT) k_cheb_step_stencil_t2(StencilOpT<VT> A, int nv, const float *__restrict__ dinv, const float *__restrict__ b,
                                                              const float *__restrict__ x_in, float *d, float *x_out,
                                                              const float *__restrict__ coef, double *acc, int acc_mode)
{
    int row = blockIdx.x * 32 + (threadIdx.x & 31), half = threadIdx.x >> 5;
    double s = 0;
    float y0 = 0, y1 = 0, y2 = 0;
    bool live = row < nv;
    stencil_row_split<VT, NT>(A, live ? row : nv - 1, half, x_in, y0, y1, y2);      // every thread reaches the barrier inside
    if (live && half == 0) {
        float a = coef[0], c = coef[1];
        float b0 = 0, b1 = 0, b2 = 0;
        if (b) { b0 = b[3 * row]; b1 = b[3 * row + 1]; b2 = b[3 * row + 2]; }
        float r0 = b0 - y0, r1 = b1 - y1, r2 = b2 - y2;
        const float *m = dinv + 9 * (size_t)row;
        float d0 = c * (m[0] * r0 + m[1] * r1 + m[2] * r2), d1 = c * (m[3] * r0 + m[4] * r1 + m[5] * r2), d2 = c * (m[6] * r0 + m[7] * r1 + m[8] * r2);
        if (a != 0.f) { d0 += a * d[3 * row]; d1 += a * d[3 * row + 1]; d2 += a * d[3 * row + 2]; }
        d[3 * row] = d0; d[3 * row + 1] = d1; d[3 * row + 2] = d2;
        if (x_out) {
            float o0 = x_in[3 * row] + d0, o1 = x_in[3 * row + 1] + d1, o2 = x_in[3 * row + 2] + d2;
            x_out[3 * row] = o0; x_out[3 * row + 1] = o1; x_out[3 * row + 2] = o2;
            if (acc_mode == 1) s = (double)b0 * o0 + (double)b1 * o1 + (double)b2 * o2;
        }
        if (acc_mode == 2) s = (double)d0 * d0 + (double)d1 * d1 + (double)d2 * d2;
    }
    if (acc) block_atomic_sum(s, acc);
}
template <class VT, int NT>
__global__ void __launch_bounds__(32 * NT) k_mg_residual_stencil_t2(StencilOpT<VT> A, int nv, const float *__restrict__ b, const float *__restrict__ x, float *r)
{
    int row = blockIdx.x * 32 + (threadIdx.x & 31), half = threadIdx.x >> 5;
    float y0, y1, y2;
    bool live = row < nv;
    stencil_row_split<VT, NT>(A, live ? row : nv - 1, half, x, y0, y1, y2);
    if (live && half == 0) { r[3 * row] = b[3 * row] - y0; r[3 * row + 1] = b[3 * row + 1] - y1; r[3 * row + 2] = b[3 * row + 2] - y2; }
}
template <class VT>
__global__ void __launch_bounds__(128) k_cheb_step_stencil_t(StencilOpT<VT> A, int nv, const float *__restrict__ dinv, const float *__restrict__ b,
                                                             const float *__restrict__ x_in, float *d, float *x_out,
                                                             const float *__restrict__ coef, double *acc, int acc_mode)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0;
    if (row < nv) {
        float a = coef[0], c = coef[1];
        float y0, y1, y2;
        stencil_row_thread(A, row, x_in, y0, y1, y2);
        float b0 = 0, b1 = 0, b2 = 0;
        if (b) { b0 = b[3 * row]; b1 = b[3 * row + 1]; b2 = b[3 * row + 2]; }
        float r0 = b0 - y0, r1 = b1 - y1, r2 = b2 - y2;
        const float *m = dinv + 9 * (size_t)row;
        float d0 = c * (m[0] * r0 + m[1] * r1 + m[2] * r2), d1 = c * (m[3] * r0 + m[4] * r1 + m[5] * r2), d2 = c * (m[6] * r0 + m[7] * r1 + m[8] * r2);
        if (a != 0.f) { d0 += a * d[3 * row]; d1 += a * d[3 * row + 1]; d2 += a * d[3 * row + 2]; }
        d[3 * row] = d0; d[3 * row + 1] = d1; d[3 * row + 2] = d2;
        if (x_out) {
            float o0 = x_in[3 * row] + d0, o1 = x_in[3 * row + 1] + d1, o2 = x_in[3 * row + 2] + d2;
            x_out[3 * row] = o0; x_out[3 * row + 1] = o1; x_out[3 * row + 2] = o2;
            if (acc_mode == 1) s = (double)b0 * o0 + (double)b1 * o1 + (double)b2 * o2;
        }
        if (acc_mode == 2) s = (double)d0 * d0 + (double)d1 * d1 + (double)d2 * d2;
    }
    if (acc) block_atomic_sum(s, acc);
}
template <class VT>
__global__ void __launch_bounds__(128) k_mg_residual_stencil_t(StencilOpT<VT> A, int nv, const float *__restrict__ b, const float *__restrict__ x, float *r)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nv) return;
    float y0, y1, y2;
    stencil_row_thread(A, row, x, y0, y1, y2);
    r[3 * row] = b[3 * row] - y0; r[3 * row + 1] = b[3 * row + 1] - y1; r[3 * row + 2] = b[3 * row + 2] - y2;
}
template <class VT, int NT>
__global__ void __launch_bounds__(NT == 1 ? 256 : 32 * NT) k_cheb_step_sell(SellOpT<VT> A, int nrows, const float *__restrict__ dinv, const float *__restrict__ b,
                                                                            const float *__restrict__ x_in, float *d, float *x_out,
                                                                            const float *__restrict__ coef, double *acc, int acc_mode)
{
    // NT == 1: one thread per row; else CTA = one slice (32 rows) x NT warps splitting its columns (rows are allocated in whole slices)
    int row = NT == 1 ? blockIdx.x * blockDim.x + threadIdx.x : blockIdx.x * 32 + (threadIdx.x & 31);
    const int part = NT == 1 ? 0 : (threadIdx.x >> 5);
    double s = 0;
    float y0 = 0, y1 = 0, y2 = 0;
    if (NT > 1) A.template mul_split<NT>(row, part, x_in, y0, y1, y2);
    if (row < nrows && part == 0) {
        float a = coef[0], c = coef[1];
        if (NT == 1) A.mul(row, x_in, y0, y1, y2);
        float b0 = 0, b1 = 0, b2 = 0;
        if (b) { b0 = b[3 * row]; b1 = b[3 * row + 1]; b2 = b[3 * row + 2]; }
        float r0 = b0 - y0, r1 = b1 - y1, r2 = b2 - y2;
        const float *m = dinv + 9 * (size_t)row;
        float d0 = c * (m[0] * r0 + m[1] * r1 + m[2] * r2), d1 = c * (m[3] * r0 + m[4] * r1 + m[5] * r2), d2 = c * (m[6] * r0 + m[7] * r1 + m[8] * r2);
        if (a != 0.f) { d0 += a * d[3 * row]; d1 += a * d[3 * row + 1]; d2 += a * d[3 * row + 2]; }
        d[3 * row] = d0; d[3 * row + 1] = d1; d[3 * row + 2] = d2;
        if (x_out) {
            float o0 = x_in[3 * row] + d0, o1 = x_in[3 * row + 1] + d1, o2 = x_in[3 * row + 2] + d2;
            x_out[3 * row] = o0; x_out[3 * row + 1] = o1; x_out[3 * row + 2] = o2;
            if (acc_mode == 1) s = (double)b0 * o0 + (double)b1 * o1 + (double)b2 * o2;
        }
        if (acc_mode == 2) s = (double)d0 * d0 + (double)d1 * d1 + (double)d2 * d2;
    }
    if (acc) block_atomic_sum(s, acc);
}
// the same update on a stencil level, one warp per vertex (lanes 0..2 finish the three components)
__global__ void __launch_bounds__(256) k_cheb_step_stencil(StencilOp A, int nv, const float *__restrict__ dinv, const float *__restrict__ b,
                                                           const float *__restrict__ x_in, float *d, float *x_out,
                                                           const float *__restrict__ coef, double *acc, int acc_mode)
{
    int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    double s = 0;
    if (v < nv) {
        float y0, y1, y2;
        stencil_row_warp(A, v, lane, x_in, y0, y1, y2);
        if (lane < 3) {
            float a = coef[0], c = coef[1];
            float r0 = -y0, r1 = -y1, r2 = -y2;
            float bq = 0.f;
            if (b) { r0 += b[3 * v]; r1 += b[3 * v + 1]; r2 += b[3 * v + 2]; bq = b[3 * v + lane]; }
            const float *m = dinv + 9 * (size_t)v + 3 * lane;
            float dq = c * (m[0] * r0 + m[1] * r1 + m[2] * r2);
            if (a != 0.f) dq += a * d[3 * v + lane];
            d[3 * v + lane] = dq;
            if (x_out) {
                float o = x_in[3 * v + lane] + dq;
                x_out[3 * v + lane] = o;
                if (acc_mode == 1) s = (double)bq * o;
            }
            if (acc_mode == 2) s = (double)dq * dq;
        }
    }
    if (acc) block_atomic_sum(s, acc);
}
template <class VT, int NT>
__global__ void __launch_bounds__(NT == 1 ? 256 : 32 * NT) k_mg_residual_sell(SellOpT<VT> A, int nrows, const float *__restrict__ b, const float *__restrict__ x, float *r)
{
    int row = NT == 1 ? blockIdx.x * blockDim.x + threadIdx.x : blockIdx.x * 32 + (threadIdx.x & 31);
    const int part = NT == 1 ? 0 : (threadIdx.x >> 5);
    float y0 = 0, y1 = 0, y2 = 0;
    if (NT > 1) A.template mul_split<NT>(row, part, x, y0, y1, y2);
    if (row >= nrows || part != 0) return;
    if (NT == 1) A.mul(row, x, y0, y1, y2);
    r[3 * row] = b[3 * row] - y0; r[3 * row + 1] = b[3 * row + 1] - y1; r[3 * row + 2] = b[3 * row + 2] - y2;
}
__global__ void __launch_bounds__(256) k_mg_residual_stencil(StencilOp A, int nv, const float *__restrict__ b, const float *__restrict__ x, float *r)
{
    int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (v >= nv) return;
    float y0, y1, y2;
    stencil_row_warp(A, v, lane, x, y0, y1, y2);
    if (lane < 3) r[3 * v + lane] = b[3 * v + lane] - (lane == 0 ? y0 : (lane == 1 ? y1 : y2));
}
// z = D^-1 b, acc += b . z  (block-Jacobi preconditioner, precond == 0)
__global__ void __launch_bounds__(256) k_apply_dinv(int nrows, const float *__restrict__ dinv, const float *__restrict__ b, float *z, double *acc)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0;
    if (row < nrows) {
        float r0 = b[3 * row], r1 = b[3 * row + 1], r2 = b[3 * row + 2];
        const float *m = dinv + 9 * (size_t)row;
        float z0 = m[0] * r0 + m[1] * r1 + m[2] * r2, z1 = m[3] * r0 + m[4] * r1 + m[5] * r2, z2 = m[6] * r0 + m[7] * r1 + m[8] * r2;
        z[3 * row] = z0; z[3 * row + 1] = z1; z[3 * row + 2] = z2;
        s = (double)r0 * z0 + (double)r1 * z1 + (double)r2 * z2;
    }
    if (acc) block_atomic_sum(s, acc);
}

// ------------------------------------------------------------------------------------------------ fused coarse tail
// The levels with at most 1024 vertices (<= 32 x 32) are ~7 kernels of 3-5 us each in a graph -- pure launch latency.
// k_vcycle_tail runs the whole sub-cycle of those levels in ONE thread block (32 warps, one warp per vertex,
// __syncthreads() between phases; the vectors live in L1/L2).  No __restrict__ / __ldg on vectors here: they are
// rewritten between phases of the same kernel.
#define TSL_MG_TAIL_MAX 6
struct TailLevel { const float *val, *dinv; float *x0, *x1, *b, *r, *d; int n0, n1, nv; };
struct TailArgs { TailLevel lev[TSL_MG_TAIL_MAX]; int n; int degree, coarse_degree; const float *coef; int first_level; };

__device__ __forceinline__ void tail_row(const TailLevel &L, int v, int lane, const float *x, float &y0, float &y1, float &y2)
{
    int I = v / L.n1, J = v - I * L.n1;
    const float *row = L.val + (size_t)v * 225;
    float a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int e = k * 32 + lane;
        if (e < 225) {
            int slot = e / 9, c = e - slot * 9;
            int q = slot / 5;
            int ii = min(max(I + q - 2, 0), L.n0 - 1), jj = min(max(J + slot - q * 5 - 2, 0), L.n1 - 1);
            int r = c / 3;
            float t = __ldg(row + e) * x[3 * (ii * L.n1 + jj) + (c - r * 3)];
            a0 += (r == 0) ? t : 0.f; a1 += (r == 1) ? t : 0.f; a2 += (r == 2) ? t : 0.f;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    y0 = a0; y1 = a1; y2 = a2;
}
// The same sub-cycle runs either in ONE thread block (CL = false: barrier = __syncthreads) or in one thread-block CLUSTER (CL = true:
// the work of a phase is dealt over every warp of the cluster, barrier = barrier.cluster with release / acquire, which also orders the
// global-memory vectors the CTAs exchange).  A cluster of 16 SMs walks a 45 x 45 level in ~2.5 us per phase where a graph node costs
// ~3.5 us before it does anything; one SM alone was measured slower than the nodes it replaces.
struct TailGeo { int tid, nt, warp, nw, lane; };
template <bool CL>
__device__ __forceinline__ TailGeo tail_geo()
{
    TailGeo g;
    unsigned rank = 0, nranks = 1;
    if (CL) {
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
        asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(nranks));
    }
    g.tid = (int)(rank * blockDim.x + threadIdx.x); g.nt = (int)(nranks * blockDim.x);
    g.warp = g.tid >> 5; g.nw = g.nt >> 5; g.lane = threadIdx.x & 31;
    return g;
}
template <bool CL>
__device__ __forceinline__ void tail_sync()
{
    if (CL) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();
}
// d = a d + c D^-1 (b - A x_in), x_out = x_in + d   (first: x_in = 0, no matrix pass)
template <bool CL>
__device__ __forceinline__ void tail_step(const TailGeo &g, const TailLevel &L, const float *b, const float *x_in, float *x_out, float a, float c, bool first)
{
    const int warp = g.warp, lane = g.lane, nw = g.nw;
    for (int v = warp; v < L.nv; v += nw) {
        float y0 = 0, y1 = 0, y2 = 0;
        if (!first) tail_row(L, v, lane, x_in, y0, y1, y2);
        if (lane < 3) {
            float r0 = b[3 * v] - y0, r1 = b[3 * v + 1] - y1, r2 = b[3 * v + 2] - y2;
            const float *m = L.dinv + 9 * (size_t)v + 3 * lane;
            float dq = c * (m[0] * r0 + m[1] * r1 + m[2] * r2);
            if (a != 0.f) dq += a * L.d[3 * v + lane];
            L.d[3 * v + lane] = dq;
            x_out[3 * v + lane] = (first ? 0.f : x_in[3 * v + lane]) + dq;
        }
    }
    tail_sync<CL>();
}
template <bool CL>
__global__ void __launch_bounds__(1024) k_vcycle_tail(TailArgs A)
{
    const int last = A.n - 1;
    const TailGeo g = tail_geo<CL>();
    const int lane = g.lane, warp = g.warp, nw = g.nw;
    // ---- downward: pre-smooth, residual, restrict (the coarsest level is solved by the long sweep)
    for (int l = 0; l <= last; l++) {
        const TailLevel &L = A.lev[l];
        const float *coef = A.coef + (size_t)(A.first_level + l) * TSL_MG_MAX_DEGREE * 2;
        const int deg = (l == last) ? A.coarse_degree : A.degree;
        for (int k = 0; k < deg; k++) {
            float *out = (k & 1) ? L.x1 : L.x0;
            const float *in = (k & 1) ? L.x0 : L.x1;
            tail_step<CL>(g, L, L.b, in, out, coef[2 * k], coef[2 * k + 1], k == 0);
        }
        if (l == last) break;
        const float *cur = ((deg - 1) & 1) ? L.x1 : L.x0;
        for (int v = warp; v < L.nv; v += nw) {
            float y0, y1, y2;
            tail_row(L, v, lane, cur, y0, y1, y2);
            if (lane < 3) L.r[3 * v + lane] = L.b[3 * v + lane] - (lane == 0 ? y0 : (lane == 1 ? y1 : y2));
        }
        tail_sync<CL>();
        const TailLevel &C = A.lev[l + 1];
        for (int cv = g.tid; cv < C.nv; cv += g.nt) {
            int I = cv / C.n1, J = cv - I * C.n1;
            float s0 = 0, s1 = 0, s2 = 0;
            for (int a = -1; a <= 1; a++) {
                int i = 2 * I + a;
                if ((unsigned)i >= (unsigned)L.n0) continue;
                float wi = a == 0 ? 1.f : (a < 0 ? 0.5f : (I + 1 < C.n0 ? 0.5f : 1.f));
                for (int b = -1; b <= 1; b++) {
                    int j = 2 * J + b;
                    if ((unsigned)j >= (unsigned)L.n1) continue;
                    float w = wi * (b == 0 ? 1.f : (b < 0 ? 0.5f : (J + 1 < C.n1 ? 0.5f : 1.f)));
                    int fr = 3 * (i * L.n1 + j);
                    s0 += w * L.r[fr]; s1 += w * L.r[fr + 1]; s2 += w * L.r[fr + 2];
                }
            }
            C.b[3 * cv] = s0; C.b[3 * cv + 1] = s1; C.b[3 * cv + 2] = s2;
        }
        tail_sync<CL>();
    }
    // ---- upward: prolong, post-smooth
    for (int l = last - 1; l >= 0; l--) {
        const TailLevel &L = A.lev[l], &C = A.lev[l + 1];
        const float *coef = A.coef + (size_t)(A.first_level + l) * TSL_MG_MAX_DEGREE * 2;
        const int cdeg = (l + 1 == last) ? A.coarse_degree : A.degree;
        // result of the coarser level: after `cdeg` pre steps (+ `cdeg` post steps unless it is the coarsest)
        int steps_c = (l + 1 == last) ? cdeg : 2 * cdeg;
        const float *xc = ((steps_c - 1) & 1) ? C.x1 : C.x0;
        float *cur = ((A.degree - 1) & 1) ? L.x1 : L.x0;
        for (int fv = g.tid; fv < L.nv; fv += g.nt) {
            int i = fv / L.n1, j = fv - i * L.n1;
            int I0 = i >> 1, J0 = j >> 1, nI = 1, nJ = 1;
            float wI = 1.f, wJ = 1.f;
            if ((i & 1) && I0 + 1 < C.n0) { nI = 2; wI = 0.5f; }
            if ((j & 1) && J0 + 1 < C.n1) { nJ = 2; wJ = 0.5f; }
            float s0 = 0, s1 = 0, s2 = 0;
            for (int a = 0; a < nI; a++)
                for (int b = 0; b < nJ; b++) {
                    int cv = 3 * ((I0 + a) * C.n1 + (J0 + b));
                    s0 += xc[cv]; s1 += xc[cv + 1]; s2 += xc[cv + 2];
                }
            float w = wI * wJ;
            cur[3 * fv] += w * s0; cur[3 * fv + 1] += w * s1; cur[3 * fv + 2] += w * s2;
        }
        tail_sync<CL>();
        for (int k = 0; k < A.degree; k++) {
            int kk = A.degree + k;                 // continues the ping-pong of the pre-smoothing steps
            float *out = (kk & 1) ? L.x1 : L.x0;
            const float *in = (kk & 1) ? L.x0 : L.x1;
            tail_step<CL>(g, L, L.b, in, out, coef[2 * k], coef[2 * k + 1], false);
        }
    }
}

// ------------------------------------------------------------------------------------------------ transfer operators
// b_c = P^T r_f.  off: first row of the grid inside the fine vectors (level 0: cloth vertex offset); mask: frozen
// flags [3 * rows] of the fine vectors or nullptr (frozen DOFs do not take part in the coarse correction)
// Optionally fused with the first Chebyshev step of the coarse level (zero guess: d = c D^-1 b, x = d), which only needs the
// vertex's own right-hand side: one graph node less per level.
__global__ void k_restrict(int n0f, int n1f, int off, const float *__restrict__ r_f, const int *__restrict__ mask,
                           int n0c, int n1c, float *b_c, const float *__restrict__ dinv_c, float *d_c, float *x0_c, const float *__restrict__ coef_c)
{
    int cv = blockIdx.x * blockDim.x + threadIdx.x;
    if (cv >= n0c * n1c) return;
    int I = cv / n1c, J = cv - I * n1c;
    float s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int a = -1; a <= 1; a++) {
        int i = 2 * I + a;
        if ((unsigned)i >= (unsigned)n0f) continue;
        float wi = pw1(a, I, n0c);
#pragma unroll
        for (int b = -1; b <= 1; b++) {
            int j = 2 * J + b;
            if ((unsigned)j >= (unsigned)n1f) continue;
            float w = wi * pw1(b, J, n1c);
            size_t fr = 3 * (size_t)(off + i * n1f + j);
            float m0 = 1, m1 = 1, m2 = 1;
            if (mask) { m0 = mask[fr] ? 0.f : 1.f; m1 = mask[fr + 1] ? 0.f : 1.f; m2 = mask[fr + 2] ? 0.f : 1.f; }
            s0 += w * m0 * r_f[fr]; s1 += w * m1 * r_f[fr + 1]; s2 += w * m2 * r_f[fr + 2];
        }
    }
    b_c[3 * cv] = s0; b_c[3 * cv + 1] = s1; b_c[3 * cv + 2] = s2;
    if (dinv_c) {
        float c = coef_c[1];
        const float *m = dinv_c + 9 * (size_t)cv;
        float d0 = c * (m[0] * s0 + m[1] * s1 + m[2] * s2), d1 = c * (m[3] * s0 + m[4] * s1 + m[5] * s2), d2 = c * (m[6] * s0 + m[7] * s1 + m[8] * s2);
        d_c[3 * cv] = d0; d_c[3 * cv + 1] = d1; d_c[3 * cv + 2] = d2;
        x0_c[3 * cv] = d0; x0_c[3 * cv + 1] = d1; x0_c[3 * cv + 2] = d2;
    }
}
// x_f += P x_c
__global__ void k_prolong_add(int n0f, int n1f, int off, float *x_f, const int *__restrict__ mask, int n0c, int n1c,
                              const float *__restrict__ x_c)
{
    int fv = blockIdx.x * blockDim.x + threadIdx.x;
    if (fv >= n0f * n1f) return;
    int i = fv / n1f, j = fv - i * n1f;
    int I0 = i >> 1, J0 = j >> 1;
    int nI = 1, nJ = 1;
    float wI[2] = { 1.f, 0.f }, wJ[2] = { 1.f, 0.f };
    if (i & 1) { if (I0 + 1 < n0c) { nI = 2; wI[0] = wI[1] = 0.5f; } }
    if (j & 1) { if (J0 + 1 < n1c) { nJ = 2; wJ[0] = wJ[1] = 0.5f; } }
    float s0 = 0, s1 = 0, s2 = 0;
    for (int a = 0; a < nI; a++)
        for (int b = 0; b < nJ; b++) {
            int cv = (I0 + a) * n1c + (J0 + b);
            float w = wI[a] * wJ[b];
            s0 += w * x_c[3 * cv]; s1 += w * x_c[3 * cv + 1]; s2 += w * x_c[3 * cv + 2];
        }
    size_t fr = 3 * (size_t)(off + fv);
    if (mask) { if (mask[fr]) s0 = 0; if (mask[fr + 1]) s1 = 0; if (mask[fr + 2]) s2 = 0; }
    x_f[fr] += s0; x_f[fr + 1] += s1; x_f[fr + 2] += s2;
}

// ------------------------------------------------------------------------------------------------ setup kernels
__device__ __forceinline__ void inv3_guarded(const float *a, float *inv)
{
    double c00 = (double)a[4] * a[8] - (double)a[5] * a[7], c01 = (double)a[5] * a[6] - (double)a[3] * a[8], c02 = (double)a[3] * a[7] - (double)a[4] * a[6];
    double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    double sc = fabs((double)a[0] * a[4] * a[8]);
    if (!(fabs(det) > 1e-12 * sc) || !(det == det)) {
        // singular block (all DOFs of the vertex masked out): the vertex takes no part in this level
#pragma unroll
        for (int c = 0; c < 9; c++) inv[c] = 0.f;
        return;
    }
    double id = 1.0 / det;
    inv[0] = (float)(c00 * id); inv[1] = (float)(((double)a[2] * a[7] - (double)a[1] * a[8]) * id); inv[2] = (float)(((double)a[1] * a[5] - (double)a[2] * a[4]) * id);
    inv[3] = (float)(c01 * id); inv[4] = (float)(((double)a[0] * a[8] - (double)a[2] * a[6]) * id); inv[5] = (float)(((double)a[2] * a[3] - (double)a[0] * a[5]) * id);
    inv[6] = (float)(c02 * id); inv[7] = (float)(((double)a[1] * a[6] - (double)a[0] * a[7]) * id); inv[8] = (float)(((double)a[0] * a[4] - (double)a[1] * a[3]) * id);
}
template <class VT>
__global__ void k_dinv_stencil(int nv, long long sv, long long se, const VT *__restrict__ val, const float *__restrict__ sc, float *dinv)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    float a[9], inv[9];
    const float is = sc ? sc[1] : 1.f;
#pragma unroll
    for (int c = 0; c < 9; c++) a[c] = mg_ld(val + (size_t)v * sv + (size_t)(12 * 9 + c) * se) * is;
    inv3_guarded(a, inv);
#pragma unroll
    for (int c = 0; c < 9; c++) dinv[9 * (size_t)v + c] = inv[c];
}
__global__ void k_dinv_sell(int n_rows, int n_alloc, const int *__restrict__ diag_pb, const float *__restrict__ val, float *dinv, unsigned int *maxdiag)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_alloc) return;
    float a[9], inv[9];
    if (r < n_rows) {
        long long base = sell_addr(diag_pb[r], r & 31, 0);
#pragma unroll
        for (int c = 0; c < 9; c++) a[c] = val[base + c * 32];
        inv3_guarded(a, inv);
        if (maxdiag) {
            // positive floats order like their bit patterns; one atomic per warp
            float m = fmaxf(fmaxf(a[0], a[4]), a[8]);
            m = (m > 0.f && m < 3e38f) ? m : 0.f;
            const unsigned act = __activemask();
            const unsigned mm = __reduce_max_sync(act, __float_as_uint(m));
            if ((int)(threadIdx.x & 31) == __ffs(act) - 1) atomicMax(maxdiag, mm);
        }
    } else {
#pragma unroll
        for (int c = 0; c < 9; c++) inv[c] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < 9; c++) dinv[9 * (size_t)r + c] = inv[c];
}
// scale of every fp16 level: the level-0 operator is positive semi-definite, so no entry exceeds its largest diagonal entry m; a
// Galerkin product with bilinear P grows entries by at most (sum of a parent's weights)^2 = 16 per level.  s_l = 2^e / 16^l with
// 2^e m <= 8192 keeps every stored value below 65504 with a factor 8 to spare; fp32 levels are stored unscaled.
__global__ void k_mg_scales(int n_levels, unsigned half_mask, const unsigned int *__restrict__ maxdiag, float *scale)
{
    int l = threadIdx.x;
    if (l >= n_levels) return;
    float s = 1.f;
    if (half_mask & (1u << l)) {
        float m = __uint_as_float(*maxdiag);
        if (!(m > 0.f) || !(m < 3e38f)) m = 1.f;
        int e = (int)floorf(log2f(8192.f / m)) - 4 * l;
        s = exp2f((float)max(min(e, 60), -60));
    }
    scale[2 * l] = s; scale[2 * l + 1] = 1.f / s;
}
__global__ void k_sell_to_half(long long n, const float *__restrict__ src, const float *__restrict__ sc, __half *dst)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float s = sc[0];
    // four values per thread: 16-byte loads, 8-byte stores
    if (4 * i + 3 < n) {
        float4 v = *reinterpret_cast<const float4 *>(src + 4 * i);
        __half2 lo = __floats2half2_rn(v.x * s, v.y * s), hi = __floats2half2_rn(v.z * s, v.w * s);
        uint2 o; o.x = *reinterpret_cast<unsigned *>(&lo); o.y = *reinterpret_cast<unsigned *>(&hi);
        *reinterpret_cast<uint2 *>(dst + 4 * i) = o;
    } else {
        for (long long k = 4 * i; k < n; k++) dst[k] = __float2half_rn(src[k] * s);
    }
}
// ---- exact solve on the coarsest level.  One thread block inverts the level's dense operator (n = 3 x vertices <= 120) by Gauss-Jordan
// elimination on [A | I] in shared memory.  No pivoting: the operator is positive semi-definite; a DOF whose pivot has collapsed
// (masked / frozen vertices leave zero rows and columns in the Galerkin product) is taken out, as inv3_guarded does for D^-1.
#define TSL_MG_DIRECT_MAX 120
__global__ void __launch_bounds__(1024) k_coarse_inverse(const float *__restrict__ val, int n0, int n1, float *inv)
{
    extern __shared__ float sm[];
    const int nv = n0 * n1, n = 3 * nv, w = 2 * n;
    float *M = sm;                       // [n][2 n]
    float *d0 = sm + (size_t)n * w;      // [n] original diagonal
    for (int t = threadIdx.x; t < n * w; t += blockDim.x) M[t] = 0.f;
    __syncthreads();
    for (int t = threadIdx.x; t < nv * 225; t += blockDim.x) {
        int v = t / 225, e = t - v * 225, slot = e / 9, c = e - slot * 9;
        int I = v / n1 + slot / 5 - 2, J = v % n1 + slot % 5 - 2;
        if ((unsigned)I < (unsigned)n0 && (unsigned)J < (unsigned)n1) M[(size_t)(3 * v + c / 3) * w + 3 * (I * n1 + J) + c % 3] = val[(size_t)v * 225 + e];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) { M[(size_t)t * w + n + t] = 1.f; d0[t] = M[(size_t)t * w + t]; }
    __syncthreads();
    // block Gauss-Jordan, one vertex (3 x 3 pivot block) per step: nv sequential steps instead of 3 nv.  At step K the A part is already
    // the identity left of column 3 K and the inverse part is still zero right of column n + 3 K + 2: only columns [3 K, n + 3 K + 3) move.
    __shared__ float s_pi[9];
    for (int K = 0; K < nv; K++) {
        const int k0 = 3 * K, c_lo = k0, c_hi = n + k0 + 3;
        if (threadIdx.x == 0) {
            float a[9], iv[9];
#pragma unroll
            for (int q = 0; q < 9; q++) a[q] = M[(size_t)(k0 + q / 3) * w + k0 + q % 3];
            // DOFs whose pivot has collapsed leave the system: the rest of the block is inverted on its own
            bool dead[3];
            for (int q = 0; q < 3; q++) dead[q] = !(a[4 * q] > 1e-6f * d0[k0 + q]) || !(d0[k0 + q] > 0.f);
            for (int q = 0; q < 3; q++)
                if (dead[q]) { for (int r = 0; r < 3; r++) { a[3 * q + r] = 0.f; a[3 * r + q] = 0.f; } a[4 * q] = 1.f; }
            inv3_guarded(a, iv);
            for (int q = 0; q < 3; q++)
                if (dead[q]) { for (int r = 0; r < 3; r++) { iv[3 * q + r] = 0.f; iv[3 * r + q] = 0.f; } }
#pragma unroll
            for (int q = 0; q < 9; q++) s_pi[q] = iv[q];
        }
        __syncthreads();
        // row block K <- P^-1 row block K (dead DOFs: zero rows); columns of the dead DOFs are zeroed in every row below
        for (int c = c_lo + (int)threadIdx.x; c < c_hi; c += blockDim.x) {
            float r0 = M[(size_t)k0 * w + c], r1 = M[(size_t)(k0 + 1) * w + c], r2 = M[(size_t)(k0 + 2) * w + c];
            M[(size_t)k0 * w + c] = s_pi[0] * r0 + s_pi[1] * r1 + s_pi[2] * r2;
            M[(size_t)(k0 + 1) * w + c] = s_pi[3] * r0 + s_pi[4] * r1 + s_pi[5] * r2;
            M[(size_t)(k0 + 2) * w + c] = s_pi[6] * r0 + s_pi[7] * r1 + s_pi[8] * r2;
        }
        __syncthreads();
        // every other row i: row_i -= sum_q A[i][k0 + q] row_{k0 + q}; a warp per row, lanes stride the active columns
        for (int t = threadIdx.x; t < n * 32; t += blockDim.x) {
            int i = t >> 5, lane = t & 31;
            if (i >= k0 && i < k0 + 3) continue;
            float f0 = M[(size_t)i * w + k0], f1 = M[(size_t)i * w + k0 + 1], f2 = M[(size_t)i * w + k0 + 2];
            __syncwarp();
            if (f0 != 0.f || f1 != 0.f || f2 != 0.f)
                for (int c = c_lo + 3 + lane; c < c_hi; c += 32)
                    M[(size_t)i * w + c] -= f0 * M[(size_t)k0 * w + c] + f1 * M[(size_t)(k0 + 1) * w + c] + f2 * M[(size_t)(k0 + 2) * w + c];
            __syncwarp();
            if (lane < 3) M[(size_t)i * w + k0 + lane] = 0.f;
        }
        __syncthreads();
    }
    // symmetrised inverse (rounding leaves it slightly unsymmetric; the preconditioner must be symmetric for PCG)
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
        int i = t / n, j = t - i * n;
        inv[t] = 0.5f * (M[(size_t)i * w + n + j] + M[(size_t)j * w + n + i]);
    }
}
// x = inv b on the coarsest level (one warp per row)
__global__ void __launch_bounds__(256) k_coarse_apply(int n, const float *__restrict__ inv, const float *__restrict__ b, float *x)
{
    int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    float a = 0.f;
    for (int c = lane; c < n; c += 32) a += __ldg(inv + (size_t)row * n + c) * b[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) x[row] = a;
}
__global__ void k_fill_hash(int n, float *v, unsigned seed)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned h = (unsigned)i * 2654435761u ^ seed;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    v[i] = (float)(h & 0xffffff) / 8388608.f - 1.f;
}
// Chebyshev coefficients of every level from the power-iteration norms; also the scale of the next warm start
__global__ void k_mg_coeffs(int n_levels, const double *__restrict__ pow_acc, int k_last, float safety, float ratio, float coarse_ratio,
                            int degree, int coarse_degree, float *coef, float *powc, float *lmax_out, int extra_slot)
{
    int l = threadIdx.x;
    if (l >= n_levels) return;
    double n1 = pow_acc[l * 16 + k_last], n0 = pow_acc[l * 16 + k_last - 1];
    double lam = sqrt(n1 / n0);
    if (!(lam > 1e-6) || !(lam < 1e6)) lam = 4.0;             // degenerate level (e.g. everything masked): any finite value
    if (l == 0 && extra_slot >= 0) {
        // second estimate from a start vector confined to the non-cloth rows (tetrahedral bodies): their stiffest cells carry the
        // largest eigenvalues of D^-1 A, which ten iterations from a vector spread over the whole system under-estimate
        double m1 = pow_acc[extra_slot * 16 + k_last], m0 = pow_acc[extra_slot * 16 + k_last - 1];
        double lam2 = sqrt(m1 / m0);
        if (lam2 > lam && lam2 < 1e6) lam = lam2;
    }
    double lmax = safety * lam;
    bool coarsest = (l == n_levels - 1);
    double lmin = lmax / (coarsest ? coarse_ratio : ratio);
    int deg = coarsest ? coarse_degree : degree;
    double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta, rho = 1.0 / sigma;
    float *c = coef + (size_t)l * TSL_MG_MAX_DEGREE * 2;
    c[0] = 0.f; c[1] = (float)(1.0 / theta);
    for (int k = 1; k < deg; k++) {
        double rho_new = 1.0 / (2.0 * sigma - rho);
        c[2 * k] = (float)(rho_new * rho); c[2 * k + 1] = (float)(2.0 * rho_new / delta);
        rho = rho_new;
    }
    powc[l * 4 + 0] = 0.f; powc[l * 4 + 1] = (n1 > 0 && n1 < 1e300) ? (float)(-1.0 / sqrt(n1)) : -1.f;
    powc[l * 4 + 2] = 0.f; powc[l * 4 + 3] = -1.f;
    lmax_out[l] = (float)lmax;
}

// ------------------------------------------------------------------------------------------------ host side
static int pad32(int n) { return (n + 31) / 32 * 32; }

int mg_alloc(tsl_ctx *ctx)
{
    MgDev &mg = ctx->mg;
    mg.n_levels = 0;
    if (ctx->cloths.empty()) return TSL_OK;
    const ClothDev &c = ctx->cloths[0];
    mg.cloth_offset = c.offset;
    int n0 = c.N + 1, n1 = c.M + 1;
    int nrows0 = ctx->A.n_slices * 32;
    { const char *e = getenv("TSL_MG_HALF"); mg.use_half = e ? atoi(e) : 1; }
    { const char *e = getenv("TSL_MG_DIRECT"); mg.coarse_direct = e ? atoi(e) : 1; }
    { const char *e = getenv("TSL_MG_SELL_SPLIT"); mg.sell_split = e ? atoi(e) : 4; if (mg.sell_split != 2 && mg.sell_split != 4) mg.sell_split = 1; }
    for (int l = 0; l < TSL_MG_MAX_LEVELS; l++) {
        MgLevel &L = mg.lev[l];
        L.n0 = n0; L.n1 = n1; L.nv = n0 * n1; L.nvp = pad32(L.nv);
        if (L.nv >= 8192) { L.sv = 1; L.se = L.nvp; } else { L.sv = 225; L.se = 1; }
        L.nrows = (l == 0) ? ctx->n_solve : L.nv;       // level 0 works on the rows of the forward solve
        size_t vb = sizeof(float) * 3 * (size_t)std::max(l == 0 ? nrows0 : L.nrows, 32);
        CK(cudaMalloc(&L.val, sizeof(float) * 225 * (size_t)L.nvp));
        CK(cudaMemset(L.val, 0, sizeof(float) * 225 * (size_t)L.nvp));
        L.half = (mg.use_half && L.sv == 1) ? 1 : 0;                 // the bandwidth-bound (element-major) levels; small ones are latency-bound
        if (L.half && l > 0) {
            CK(cudaMalloc(&L.val16, sizeof(__half) * 225 * (size_t)L.nvp));
            CK(cudaMemset(L.val16, 0, sizeof(__half) * 225 * (size_t)L.nvp));
        }
        CK(cudaMalloc(&L.dinv, sizeof(float) * 9 * (size_t)std::max(l == 0 ? nrows0 : L.nrows, 32)));
        for (int q = 0; q < 2; q++) {
            CK(cudaMalloc(&L.x[q], vb)); CK(cudaMemset(L.x[q], 0, vb));
            CK(cudaMalloc(&L.pv[q], vb)); CK(cudaMemset(L.pv[q], 0, vb));
        }
        CK(cudaMalloc(&L.b, vb)); CK(cudaMalloc(&L.r, vb)); CK(cudaMalloc(&L.d, vb));
        CK(cudaMemset(L.b, 0, vb)); CK(cudaMemset(L.r, 0, vb)); CK(cudaMemset(L.d, 0, vb));
        mg.n_levels = l + 1;
        if (std::min(n0, n1) <= 6) break;
        n0 = (n0 - 1) / 2 + 1; n1 = (n1 - 1) / 2 + 1;
    }
    {
        const MgLevel &LL = mg.lev[mg.n_levels - 1];
        const int n = 3 * LL.nv;
        if (mg.coarse_direct && mg.n_levels > 1 && LL.sv != 1 && n <= TSL_MG_DIRECT_MAX) {
            CK(cudaMalloc(&mg.coarse_inv, sizeof(float) * (size_t)n * n));
            CK(cudaMemset(mg.coarse_inv, 0, sizeof(float) * (size_t)n * n));
            CK(cudaFuncSetAttribute(k_coarse_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * ((size_t)n * 2 * n + n))));
        } else mg.coarse_direct = 0;
    }
    {
        std::vector<float> ones(2 * TSL_MG_MAX_LEVELS, 1.f);
        CK(cudaMalloc(&mg.scale, sizeof(float) * ones.size()));
        CK(cudaMemcpy(mg.scale, ones.data(), sizeof(float) * ones.size(), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&mg.maxdiag, sizeof(unsigned int)));
        CK(cudaMemset(mg.maxdiag, 0, sizeof(unsigned int)));
        if (mg.lev[0].half) {
            CK(cudaMalloc(&ctx->A.val16m, sizeof(__half) * 9 * (size_t)ctx->A.nnzb_pad));
            CK(cudaMemset(ctx->A.val16m, 0, sizeof(__half) * 9 * (size_t)ctx->A.nnzb_pad));
        }
    }
    // levels from tail_level on (<= 1024 vertices each, row-major, at most TSL_MG_TAIL_MAX of them) run in one fused kernel
    { const char *e = getenv("TSL_MG_PAIR"); mg.pair_threads = e ? atoi(e) : 1; }       // 0: one thread per vertex, 1: automatic split, 2 / 4 / 8: fixed
    if (mg.pair_threads != 0 && mg.pair_threads != 1 && mg.pair_threads != 2 && mg.pair_threads != 4 && mg.pair_threads != 8 && mg.pair_threads != 16) mg.pair_threads = 1;
    // TSL_MG_TAIL = 0: off; 1: one thread block (levels <= 1024 vertices); N >= 2: one cluster of N thread blocks (N = 8 is portable, 16 needs
    // the non-portable attribute), levels <= TSL_MG_TAIL_NV vertices (default 2100: from the 45 x 45 level down)
    // Measured at 1 M / 50 k triangles (PCG iteration, profiles/README.md): separate graph nodes 524 / 128 us, one block 784 / -- us,
    // cluster of 8: 614 / 184 us, of 16: 568 / 158 us -- a phase that ends in a cluster barrier and starts with an L2 round trip costs
    // MORE than a ~3.5 us graph node, so the fused tail stays opt-in.
    { const char *e = getenv("TSL_MG_TAIL"); mg.tail_cluster = e ? atoi(e) : 0; }
    int tail_nv = mg.tail_cluster > 1 ? 2100 : 1024;
    { const char *e = getenv("TSL_MG_TAIL_NV"); if (e) tail_nv = atoi(e); }
    mg.tail_level = -1;
    for (int l = 1; l < mg.n_levels; l++)
        if (mg.lev[l].sv != 1 && mg.lev[l].nv <= tail_nv && mg.n_levels - l <= TSL_MG_TAIL_MAX) { mg.tail_level = l; break; }
    if (mg.tail_cluster > 8) {
        if (cudaFuncSetAttribute(k_vcycle_tail<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); mg.tail_cluster = 8; }
    }
    if (mg.tail_cluster == 0) mg.tail_level = -1;
    CK(cudaMalloc(&mg.coef, sizeof(float) * TSL_MG_MAX_LEVELS * TSL_MG_MAX_DEGREE * 2));
    CK(cudaMalloc(&mg.powc, sizeof(float) * TSL_MG_MAX_LEVELS * 4));
    CK(cudaMalloc(&mg.pow_acc, sizeof(double) * TSL_MG_MAX_LEVELS * 16));
    CK(cudaMalloc(&mg.lmax, sizeof(float) * TSL_MG_MAX_LEVELS));
    std::vector<float> pc(TSL_MG_MAX_LEVELS * 4);
    for (int l = 0; l < TSL_MG_MAX_LEVELS; l++) { pc[4 * l] = 0; pc[4 * l + 1] = -1; pc[4 * l + 2] = 0; pc[4 * l + 3] = -1; }
    CK(cudaMemcpy(mg.powc, pc.data(), sizeof(float) * pc.size(), cudaMemcpyHostToDevice));
    mg.setups = 0;
    { const char *e = getenv("TSL_MG_TILED"); mg.tiled_galerkin = e ? atoi(e) : 1; }
    { const char *e = getenv("TSL_MG_FORK"); mg.fork = e ? atoi(e) : 1; }
    if (mg.fork)
        for (int l = 0; l < mg.n_levels; l++) {
            CK(cudaStreamCreateWithFlags(&mg.side[l], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&mg.ev_ready[l], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&mg.ev_done[l], cudaEventDisableTiming));
        }
    CK(cudaFuncSetAttribute(k_galerkin_tiled<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TSL_GAL_SMEM));
    CK(cudaFuncSetAttribute(k_galerkin_sell_tiled<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TSL_GAL_SMEM));
    return TSL_OK;
}

void mg_free(tsl_ctx *ctx)
{
    MgDev &mg = ctx->mg;
    for (int l = 0; l < mg.n_levels; l++) {
        MgLevel &L = mg.lev[l];
        cudaFree(L.val); cudaFree(L.val16); cudaFree(L.dinv); cudaFree(L.b); cudaFree(L.r); cudaFree(L.d);
        L.val16 = nullptr;
        for (int q = 0; q < 2; q++) { cudaFree(L.x[q]); cudaFree(L.pv[q]); }
    }
    cudaFree(mg.coef); cudaFree(mg.powc); cudaFree(mg.pow_acc); cudaFree(mg.lmax); cudaFree(mg.scale); cudaFree(mg.maxdiag);
    cudaFree(ctx->A.val16m); ctx->A.val16m = nullptr;
    cudaFree(mg.coarse_inv); mg.coarse_inv = nullptr;
    for (int l = 0; l < TSL_MG_MAX_LEVELS; l++) {
        if (mg.side[l]) cudaStreamDestroy(mg.side[l]);
        if (mg.ev_ready[l]) cudaEventDestroy(mg.ev_ready[l]);
        if (mg.ev_done[l]) cudaEventDestroy(mg.ev_done[l]);
        mg.side[l] = nullptr; mg.ev_ready[l] = mg.ev_done[l] = nullptr;
    }
    mg.n_levels = 0;
}

static SellOp sell_op(tsl_ctx *ctx, const float *val) { SellOp o; o.slice_base = ctx->A.slice_base; o.colidx = ctx->A.colidx; o.val = val; o.sc = nullptr; return o; }
static SellOpT<__half> sell_op16(tsl_ctx *ctx)
{
    SellOpT<__half> o; o.slice_base = ctx->A.slice_base; o.colidx = ctx->A.colidx; o.val = (const __half *)ctx->A.val16m; o.sc = ctx->mg.scale; return o;
}
// threads per vertex on an element-major level (TSL_MG_PAIR = 2 / 4 / 8 / 16 overrides)
static int split_threads(tsl_ctx *ctx, const MgLevel &L)
{
    if (ctx->mg.pair_threads > 1) return ctx->mg.pair_threads;
    return 16;                          // measured at 1 M triangles (levels 354^2, 177^2): PCG iteration 602 / 571 / 558 / 525 us for 2 / 4 / 8 / 16
                                        // threads per vertex (the last two with the level-0 slices split over 4 warps: 543 / 525)
}
static StencilOp stencil_op(const MgLevel &L) { StencilOp o; o.val = L.val; o.n0 = L.n0; o.n1 = L.n1; o.sv = L.sv; o.se = L.se; o.sc = nullptr; return o; }
static StencilOpT<__half> stencil_op16(tsl_ctx *ctx, int l)
{
    const MgLevel &L = ctx->mg.lev[l];
    StencilOpT<__half> o; o.val = (const __half *)L.val16; o.n0 = L.n0; o.n1 = L.n1; o.sv = L.sv; o.se = L.se; o.sc = ctx->mg.scale + 2 * l; return o;
}

// d = a d + c D^-1 (b - A x_in), x_out = x_in + d on level l (level 0 smooths with the clamped matrix)
static void launch_step(tsl_ctx *ctx, int l, const float *b, const float *x_in, float *d, float *x_out, const float *coef, double *acc, int mode)
{
    MgLevel &L = ctx->mg.lev[l];
    const int sn = ctx->mg.sell_split;
    if (l == 0 && L.half && sn == 4)
        k_cheb_step_sell<__half, 4><<<GRID(L.nrows, 32), 128, 0, ctx->stream>>>(sell_op16(ctx), L.nrows, L.dinv, b, x_in, d, x_out, coef, acc, mode);
    else if (l == 0 && L.half && sn == 2)
        k_cheb_step_sell<__half, 2><<<GRID(L.nrows, 32), 64, 0, ctx->stream>>>(sell_op16(ctx), L.nrows, L.dinv, b, x_in, d, x_out, coef, acc, mode);
    else if (l == 0 && L.half)
        k_cheb_step_sell<__half, 1><<<GRID(L.nrows, 256), 256, 0, ctx->stream>>>(sell_op16(ctx), L.nrows, L.dinv, b, x_in, d, x_out, coef, acc, mode);
    else if (l == 0 && sn == 4)
        k_cheb_step_sell<float, 4><<<GRID(L.nrows, 32), 128, 0, ctx->stream>>>(sell_op(ctx, ctx->A.val32m), L.nrows, L.dinv, b, x_in, d, x_out, coef, acc, mode);
    else if (l == 0 && sn == 2)
        k_cheb_step_sell<float, 2><<<GRID(L.nrows, 32), 64, 0, ctx->stream>>>(sell_op(ctx, ctx->A.val32m), L.nrows, L.dinv, b, x_in, d, x_out, coef, acc, mode);
    else if (l == 0)
        k_cheb_step_sell<float, 1><<<GRID(L.nrows, 256), 256, 0, ctx->stream>>>(sell_op(ctx, ctx->A.val32m), L.nrows, L.dinv, b, x_in, d, x_out, coef, acc, mode);
    else if (L.sv == 1 && L.half && ctx->mg.pair_threads) {
        const int nt = split_threads(ctx, L);
        if (nt == 16) k_cheb_step_stencil_t2<__half, 16><<<GRID(L.nv, 32), 512, 0, ctx->stream>>>(stencil_op16(ctx, l), L.nv, L.dinv, b, x_in, d, x_out, coef, acc, mode);
        else if (nt == 8) k_cheb_step_stencil_t2<__half, 8><<<GRID(L.nv, 32), 256, 0, ctx->stream>>>(stencil_op16(ctx, l), L.nv, L.dinv, b, x_in, d, x_out, coef, acc, mode);
        else if (nt == 4) k_cheb_step_stencil_t2<__half, 4><<<GRID(L.nv, 32), 128, 0, ctx->stream>>>(stencil_op16(ctx, l), L.nv, L.dinv, b, x_in, d, x_out, coef, acc, mode);
        else k_cheb_step_stencil_t2<__half, 2><<<GRID(L.nv, 32), 64, 0, ctx->stream>>>(stencil_op16(ctx, l), L.nv, L.dinv, b, x_in, d, x_out, coef, acc, mode);
    }
    else if (L.sv == 1 && L.half)
        k_cheb_step_stencil_t<__half><<<GRID(L.nv, 128), 128, 0, ctx->stream>>>(stencil_op16(ctx, l), L.nv, L.dinv, b, x_in, d, x_out, coef, acc, mode);
    else if (L.sv == 1 && ctx->mg.pair_threads) {
        const int nt = split_threads(ctx, L);
        if (nt == 16) k_cheb_step_stencil_t2<float, 16><<<GRID(L.nv, 32), 512, 0, ctx->stream>>>(stencil_op(L), L.nv, L.dinv, b, x_in, d, x_out, coef, acc, mode);
        else if (nt == 8) k_cheb_step_stencil_t2<float, 8><<<GRID(L.nv, 32), 256, 0, ctx->stream>>>(stencil_op(L), L.nv, L.dinv, b, x_in, d, x_out, coef, acc, mode);
        else if (nt == 4) k_cheb_step_stencil_t2<float, 4><<<GRID(L.nv, 32), 128, 0, ctx->stream>>>(stencil_op(L), L.nv, L.dinv, b, x_in, d, x_out, coef, acc, mode);
        else k_cheb_step_stencil_t2<float, 2><<<GRID(L.nv, 32), 64, 0, ctx->stream>>>(stencil_op(L), L.nv, L.dinv, b, x_in, d, x_out, coef, acc, mode);
    }
    else if (L.sv == 1)
        k_cheb_step_stencil_t<float><<<GRID(L.nv, 128), 128, 0, ctx->stream>>>(stencil_op(L), L.nv, L.dinv, b, x_in, d, x_out, coef, acc, mode);
    else
        k_cheb_step_stencil<<<GRID(32LL * L.nv, 256), 256, 0, ctx->stream>>>(stencil_op(L), L.nv, L.dinv, b, x_in, d, x_out, coef, acc, mode);
    ctx->launches++;
}

// Builds the hierarchy for the matrix currently in A.val32c (snapshotted into A.val32m).  No host synchronisation.
int mg_setup(tsl_ctx *ctx)
{
    MgDev &mg = ctx->mg;
    cudaStream_t s = ctx->stream;
    const SellMatrix &A = ctx->A;
    int nrows0 = A.n_slices * 32;
    // the hierarchy is built from (and its level-0 smoother keeps using) a snapshot, so that the caller may refresh
    // A.val32c every Newton iteration while the preconditioner stays self-consistent until the next build.  The snapshot is fp16
    // (A.val16m, scaled) when the level-0 operator is stored in half precision, else a plain fp32 copy (A.val32m).
    const bool half0 = mg.n_levels > 0 && mg.lev[0].half;
    CK(cudaMemcpyAsync(A.val32m, A.val32c, sizeof(float) * 9 * (size_t)A.nnzb_pad, cudaMemcpyDeviceToDevice, s));
    if (mg.n_levels == 0) {          // no cloth: block-Jacobi only
        k_dinv_sell<<<GRID(nrows0, 256), 256, 0, s>>>(A.n_rows, nrows0, A.diag_pb, A.val32m, ctx->minv32, nullptr);
        ctx->launches++;
        return TSL_OK;
    }
    const ClothDev &c = ctx->cloths[0];
    MgLevel &L0 = mg.lev[0];
    unsigned half_mask = 0;
    for (int l = 0; l < mg.n_levels; l++) if (mg.lev[l].half) half_mask |= 1u << l;
    if (half_mask) CK(cudaMemsetAsync(mg.maxdiag, 0, sizeof(unsigned int), s));
    k_dinv_sell<<<GRID(nrows0, 256), 256, 0, s>>>(A.n_rows, nrows0, A.diag_pb, A.val32m, L0.dinv, half_mask ? mg.maxdiag : nullptr);
    ctx->launches++;
    if (half_mask) {
        k_mg_scales<<<1, 32, 0, s>>>(mg.n_levels, half_mask, mg.maxdiag, mg.scale);
        ctx->launches++;
    }
    if (half0) {
        long long n = 9LL * A.nnzb_pad;
        k_sell_to_half<<<GRID((n + 3) / 4, 256), 256, 0, s>>>(n, A.val32m, mg.scale, (__half *)A.val16m);
        ctx->launches++;
    }
    const bool fork = mg.fork && mg.side[0];
    CK(cudaMemsetAsync(mg.pow_acc, 0, sizeof(double) * TSL_MG_MAX_LEVELS * 16, s));
    if (fork) CK(cudaEventRecord(mg.ev_ready[0], s));
    // Galerkin products, tiled (tsl_mg_kernels.cuh), all in fp32: level 0 -> 1 straight from the sliced-ELL snapshot (no stencil copy of
    // the fine level), the others from the level's own stencil layout; a level whose smoother reads fp16 gets that copy written in the
    // same pass.  TSL_MG_TILED=0 (fp32 storage only) falls back to the entrywise kernels
    for (int l = 0; l + 1 < mg.n_levels; l++) {
        MgLevel &F = mg.lev[l], &C = mg.lev[l + 1];
        const float *scC = mg.scale + 2 * (l + 1);
        if (mg.tiled_galerkin || half_mask) {
            dim3 grid(GRID(C.n1, TSL_TCJ), GRID(C.n0, TSL_TCI));
            const int *fz = ctx->frozen + 3 * (size_t)c.offset;
            __half *h = C.half ? (__half *)C.val16 : nullptr;
            if (l == 0)
                k_galerkin_sell_tiled<__half><<<grid, 256, TSL_GAL_SMEM, s>>>(c.offset, F.n0, F.n1, A.slice_base, A.colidx, A.val32m, A.diag_pb, fz,
                                                                             C.val, h, C.n0, C.n1, C.sv, C.se, scC);
            else
                k_galerkin_tiled<__half><<<grid, 256, TSL_GAL_SMEM, s>>>(F.val, F.n0, F.n1, F.sv, F.se, C.val, h, C.n0, C.n1, C.sv, C.se, scC);
            ctx->launches++;
        } else {
            if (l == 0) {
                k_sell_to_stencil<<<GRID(L0.nv, 128), 128, 0, s>>>(c.offset, L0.nv, L0.n1, A.slice_base, A.colidx, A.val32m, A.diag_pb, L0.val, L0.sv, L0.se);
                ctx->launches++;
            }
            long long nt = 25LL * C.nv;
            if (l == 0)
                k_galerkin<true><<<GRID(nt, 128), 128, 0, s>>>(F.val, F.n0, F.n1, F.sv, F.se, ctx->frozen + 3 * (size_t)c.offset, C.val, C.n0, C.n1, C.sv, C.se);
            else
                k_galerkin<false><<<GRID(nt, 128), 128, 0, s>>>(F.val, F.n0, F.n1, F.sv, F.se, nullptr, C.val, C.n0, C.n1, C.sv, C.se);
            ctx->launches++;
        }
        k_dinv_stencil<float><<<GRID(C.nv, 256), 256, 0, s>>>(C.nv, C.sv, C.se, C.val, nullptr, C.dinv);
        ctx->launches++;
        if (fork) CK(cudaEventRecord(mg.ev_ready[l + 1], s));
    }
    if (mg.coarse_direct) {
        const MgLevel &LL = mg.lev[mg.n_levels - 1];
        const int n = 3 * LL.nv;
        k_coarse_inverse<<<1, 1024, sizeof(float) * ((size_t)n * 2 * n + n), s>>>(LL.val, LL.n0, LL.n1, mg.coarse_inv);
        ctx->launches++;
    }
    // lambda_max(D^-1 A) per level: 10 power iterations from a fixed pseudo-random vector.  (Warm-starting from the
    // previous setup's vector was measured to UNDER-estimate after the contact set changes -- the old dominant mode
    // has almost no overlap with the new one -- and an under-estimate is what the Chebyshev smoother cannot tolerate.)
    // The iteration of level l runs on its own side stream as soon as that level's operator exists (works the same inside a stream
    // capture: the event edges become graph dependencies), so the latency-bound small levels and the bandwidth-bound fine level hide
    // behind the Galerkin chain.
    const int its = 10;
    int extra_slot = -1;
    for (int l = 0; l < mg.n_levels; l++) {
        if (mg.coarse_direct && mg.tail_level < 0 && l == mg.n_levels - 1) continue;       // solved exactly: no smoother, no eigenvalue estimate
                                                                                             // (the opt-in fused tail still sweeps it)
        MgLevel &L = mg.lev[l];
        cudaStream_t q = fork ? mg.side[l] : s;
        if (fork) CK(cudaStreamWaitEvent(q, mg.ev_ready[l], 0));
        ctx->stream = q;                                   // launch_step launches on the context's stream
        k_fill_hash<<<GRID(3 * L.nrows, 256), 256, 0, q>>>(3 * L.nrows, L.pv[0], 0x9e3779b9u * (l + 1));
        ctx->launches++;
        for (int k = 0; k < its; k++)
            launch_step(ctx, l, nullptr, L.pv[k & 1], L.pv[(k & 1) ^ 1], nullptr, mg.powc + 4 * l + 2, mg.pow_acc + 16 * l + k, 2);
        if (l == 0 && !ctx->tets.empty() && mg.n_levels < TSL_MG_MAX_LEVELS && ctx->n_solve > c.offset + c.NV) {
            // the sliced-ELL matrix has no cloth <-> solid blocks (contacts live in the side buffer), so the iteration stays on the solids
            extra_slot = TSL_MG_MAX_LEVELS - 1;
            k_fill_hash<<<GRID(3 * L.nrows, 256), 256, 0, q>>>(3 * L.nrows, L.pv[0], 0x51ed270bu);
            cudaMemsetAsync(L.pv[0] + 3 * (size_t)c.offset, 0, sizeof(float) * 3 * (size_t)c.NV, q);
            ctx->launches++;
            for (int k = 0; k < its; k++)
                launch_step(ctx, 0, nullptr, L.pv[k & 1], L.pv[(k & 1) ^ 1], nullptr, mg.powc + 2, mg.pow_acc + 16 * extra_slot + k, 2);
        }
        ctx->stream = s;
        if (fork) { CK(cudaEventRecord(mg.ev_done[l], q)); CK(cudaStreamWaitEvent(s, mg.ev_done[l], 0)); }
    }
    k_mg_coeffs<<<1, 32, 0, s>>>(mg.n_levels, mg.pow_acc, its - 1, mg.safety, mg.ratio, mg.coarse_ratio, mg.degree, mg.coarse_degree,
                                 mg.coef, mg.powc, mg.lmax, extra_slot);
    ctx->launches++;
    mg.setups++;
    CK(cudaGetLastError());
    return TSL_OK;
}

// z = M^-1 b : one V-cycle (or block-Jacobi when precond == 0); acc_bz (optional, device) += b . z
static float *vcycle_level(tsl_ctx *ctx, int l, const float *b, float *z_out, double *acc, bool first_done = false)
{
    MgDev &mg = ctx->mg;
    MgLevel &L = mg.lev[l];
    cudaStream_t s = ctx->stream;
    if (l > 0 && l == mg.tail_level && b == L.b) {
        // the remaining levels fit one thread block: one kernel for the whole sub-cycle
        TailArgs T;
        T.n = mg.n_levels - l; T.degree = mg.degree; T.coarse_degree = mg.coarse_degree; T.coef = mg.coef; T.first_level = l;
        for (int q = 0; q < T.n; q++) {
            MgLevel &Q = mg.lev[l + q];
            T.lev[q].val = Q.val; T.lev[q].dinv = Q.dinv; T.lev[q].x0 = Q.x[0]; T.lev[q].x1 = Q.x[1];
            T.lev[q].b = Q.b; T.lev[q].r = Q.r; T.lev[q].d = Q.d; T.lev[q].n0 = Q.n0; T.lev[q].n1 = Q.n1; T.lev[q].nv = Q.nv;
        }
        if (mg.tail_cluster > 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)mg.tail_cluster); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)mg.tail_cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, k_vcycle_tail<true>, T);
        } else
            k_vcycle_tail<false><<<1, 1024, 0, s>>>(T);
        ctx->launches++;
        int steps = (T.n == 1) ? mg.coarse_degree : 2 * mg.degree;
        return ((steps - 1) & 1) ? L.x[1] : L.x[0];
    }
    const float *coef = mg.coef + (size_t)l * TSL_MG_MAX_DEGREE * 2;
    const bool last = (l == mg.n_levels - 1);
    if (last && l > 0 && mg.coarse_direct) {
        const int n = 3 * L.nv;
        k_coarse_apply<<<GRID(32 * n, 256), 256, 0, s>>>(n, mg.coarse_inv, b, L.x[0]);
        ctx->launches++;
        return L.x[0];
    }
    const int deg = last ? mg.coarse_degree : mg.degree;
    // pre-smoothing (or the coarsest-grid sweep) from a zero guess
    float *cur = nullptr;
    for (int k = 0; k < deg; k++) {
        bool final_k = last && (k == deg - 1);
        float *out = (final_k && z_out) ? z_out : L.x[k & 1];
        double *a = final_k ? acc : nullptr;
        if (k == 0) {
            if (!first_done) {         // otherwise the producer of b (k_restrict / k_pcg_update) already wrote d and x[0]
                k_cheb_first<<<GRID(L.nrows, 256), 256, 0, s>>>(L.nrows, L.dinv, b, L.d, out, coef, a);
                ctx->launches++;
            }
        } else launch_step(ctx, l, b, cur, L.d, out, coef + 2 * k, a, a ? 1 : 0);
        cur = out;
    }
    if (last) return cur;
    MgLevel &C = mg.lev[l + 1];
    int off = (l == 0) ? mg.cloth_offset : 0;
    const int *mask = (l == 0) ? ctx->frozen : nullptr;
    // the residual that goes down to the coarse grid is taken with the fp32 operator (see tsl_mg_kernels.cuh)
    if (l == 0 && mg.sell_split == 4) k_mg_residual_sell<float, 4><<<GRID(L.nrows, 32), 128, 0, s>>>(sell_op(ctx, ctx->A.val32m), L.nrows, b, cur, L.r);
    else if (l == 0 && mg.sell_split == 2) k_mg_residual_sell<float, 2><<<GRID(L.nrows, 32), 64, 0, s>>>(sell_op(ctx, ctx->A.val32m), L.nrows, b, cur, L.r);
    else if (l == 0) k_mg_residual_sell<float, 1><<<GRID(L.nrows, 256), 256, 0, s>>>(sell_op(ctx, ctx->A.val32m), L.nrows, b, cur, L.r);
    else if (L.sv == 1 && mg.pair_threads) {
        const int nt = split_threads(ctx, L);
        if (nt == 16) k_mg_residual_stencil_t2<float, 16><<<GRID(L.nv, 32), 512, 0, s>>>(stencil_op(L), L.nv, b, cur, L.r);
        else if (nt == 8) k_mg_residual_stencil_t2<float, 8><<<GRID(L.nv, 32), 256, 0, s>>>(stencil_op(L), L.nv, b, cur, L.r);
        else if (nt == 4) k_mg_residual_stencil_t2<float, 4><<<GRID(L.nv, 32), 128, 0, s>>>(stencil_op(L), L.nv, b, cur, L.r);
        else k_mg_residual_stencil_t2<float, 2><<<GRID(L.nv, 32), 64, 0, s>>>(stencil_op(L), L.nv, b, cur, L.r);
    }
    else if (L.sv == 1) k_mg_residual_stencil_t<float><<<GRID(L.nv, 128), 128, 0, s>>>(stencil_op(L), L.nv, b, cur, L.r);
    else k_mg_residual_stencil<<<GRID(32LL * L.nv, 256), 256, 0, s>>>(stencil_op(L), L.nv, b, cur, L.r);
    const bool c_last = (l + 1 == mg.n_levels - 1);
    const bool fuse_first = (l + 1 != mg.tail_level) && ((c_last ? mg.coarse_degree : mg.degree) >= 2);
    const float *coef_c = mg.coef + (size_t)(l + 1) * TSL_MG_MAX_DEGREE * 2;
    k_restrict<<<GRID(C.nv, 128), 128, 0, s>>>(L.n0, L.n1, off, L.r, mask, C.n0, C.n1, C.b, fuse_first ? C.dinv : nullptr, C.d, C.x[0], coef_c);
    ctx->launches += 2;
    float *xc = vcycle_level(ctx, l + 1, C.b, nullptr, nullptr, fuse_first);
    k_prolong_add<<<GRID(L.nv, 256), 256, 0, s>>>(L.n0, L.n1, off, cur, mask, C.n0, C.n1, xc);
    ctx->launches++;
    // post-smoothing with the same polynomial (symmetric cycle)
    for (int k = 0; k < deg; k++) {
        bool final_k = (k == deg - 1);
        float *out = (final_k && z_out) ? z_out : (cur == L.x[0] ? L.x[1] : L.x[0]);
        double *a = (final_k && l == 0) ? acc : nullptr;
        launch_step(ctx, l, b, cur, L.d, out, coef + 2 * k, a, a ? 1 : 0);
        cur = out;
    }
    return cur;
}

// level-0 buffers a producer of the right-hand side may fill with the first Chebyshev step itself (see k_pcg_update)
void mg_first_step_targets(tsl_ctx *ctx, const float **dinv, float **d, float **x0, const float **coef)
{
    MgDev &mg = ctx->mg;
    bool ok = mg.n_levels > 0 && ctx->precond != 0 && ((mg.n_levels == 1 ? mg.coarse_degree : mg.degree) >= 2);
    *dinv = ok ? mg.lev[0].dinv : nullptr; *d = ok ? mg.lev[0].d : nullptr; *x0 = ok ? mg.lev[0].x[0] : nullptr; *coef = ok ? mg.coef : nullptr;
}

int mg_apply(tsl_ctx *ctx, const float *b, float *z, double *acc_bz, bool first_done)
{
    MgDev &mg = ctx->mg;
    if (mg.n_levels == 0 || ctx->precond == 0) {
        int nrows0 = ctx->A.n_slices * 32;
        const float *dinv = (mg.n_levels == 0) ? ctx->minv32 : mg.lev[0].dinv;
        k_apply_dinv<<<GRID(nrows0, 256), 256, 0, ctx->stream>>>(nrows0, dinv, b, z, acc_bz);   // block-Jacobi: every row
        ctx->launches++;
        return TSL_OK;
    }
    vcycle_level(ctx, 0, b, z, acc_bz, first_done);
    CK(cudaGetLastError());
    return TSL_OK;
}

// test / diagnostic read-back of one level: grid, lambda_max estimate, stencil operator [25][9][nv] (slot-major)
int mg_get_level(tsl_ctx *ctx, int level, int *dims, float *lmax, float *val_host)
{
    MgDev &mg = ctx->mg;
    if (level < 0 || level >= mg.n_levels) { ctx->err = "mg_get_level: no such level"; return TSL_ERR_INVALID; }
    MgLevel &L = mg.lev[level];
    if (level == 0 && val_host && mg.n_levels > 1) {
        // the stencil copy of the fine level is no longer part of the setup (the tiled Galerkin product reads the sliced-ELL matrix): made on demand
        const ClothDev &c = ctx->cloths[0];
        k_sell_to_stencil<<<GRID(L.nv, 128), 128, 0, ctx->stream>>>(c.offset, L.nv, L.n1, ctx->A.slice_base, ctx->A.colidx, ctx->A.val32m, ctx->A.diag_pb,
                                                                   L.val, L.sv, L.se);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (dims) { dims[0] = L.n0; dims[1] = L.n1; dims[2] = mg.n_levels; }
    if (lmax) CK(cudaMemcpy(lmax, mg.lmax + level, sizeof(float), cudaMemcpyDeviceToHost));
    if (val_host) {
        std::vector<float> tmp((size_t)225 * L.nvp);
        CK(cudaMemcpy(tmp.data(), L.val, sizeof(float) * tmp.size(), cudaMemcpyDeviceToHost));
        for (int sc = 0; sc < 225; sc++)
            for (int v = 0; v < L.nv; v++) val_host[(size_t)sc * L.nv + v] = tmp[(size_t)v * L.sv + (size_t)sc * L.se];
    }
    return TSL_OK;
}

}  // namespace tsl
