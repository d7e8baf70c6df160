// tsl_dense_kernels.cuh -- kernels of the dense fp64 LU (see tsl_dense.cu for the design).  Only CUDA built-ins are used, so the CPU
// test suite can run these through tests/csrc/cuda_emu.h (one std::thread per CUDA thread) without a GPU.
#pragma once
#ifndef TSL_CUDA_EMU
#include <cuda_runtime.h>
#endif

namespace tsl {

constexpr int LU_NB = 32;

__global__ void __launch_bounds__(1024) k_lu_panel(double *A, int lda, int n, int k0, int nb, int *ipiv, int *info)
{
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_p;
    __shared__ double s_row[LU_NB];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
    for (int j = 0; j < nb; j++) {
        const int col = k0 + j;
        double *c = A + (size_t)col * lda;
        // pivot: largest |c[i]|, i in [col, n); the smaller row index wins ties (deterministic)
        double best = -1.0;
        int bi = 0x7fffffff;
        for (int i = col + tid; i < n; i += blockDim.x) {
            double v = fabs(c[i]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_down_sync(0xffffffffu, best, o);
            int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[w] = best; s_idx[w] = bi; }
        __syncthreads();
        if (w == 0) {
            best = lane < nw ? s_val[lane] : -1.0;
            bi = lane < nw ? s_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_down_sync(0xffffffffu, best, o);
                int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) {
                if (!(best > 0.0) || bi >= n) { bi = col; atomicOr(info, 1); }   // singular (or NaN) column
                s_p = bi;
                ipiv[col] = bi;
            }
        }
        __syncthreads();
        const int p = s_p;
        // swap rows col <-> p inside the panel; s_row = the new pivot row
        if (tid < nb) {
            double *q = A + (size_t)(k0 + tid) * lda;
            double a = q[col], b = q[p];
            if (p != col) { q[col] = b; q[p] = a; }
            s_row[tid] = (p != col) ? b : a;
        }
        __syncthreads();
        const double piv = s_row[j];
        const double inv = (piv != 0.0) ? 1.0 / piv : 0.0;
        for (int i = col + 1 + tid; i < n; i += blockDim.x) {
            double l = c[i] * inv;
            c[i] = l;
            for (int jj = j + 1; jj < nb; jj++) A[(size_t)(k0 + jj) * lda + i] -= l * s_row[jj];
        }
        __syncthreads();
    }
}

__global__ void k_lu_swap(double *A, int lda, int n, int k0, int nb, const int *__restrict__ ipiv)
{
    int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n || (col >= k0 && col < k0 + nb)) return;
    double *c = A + (size_t)col * lda;
    for (int j = 0; j < nb; j++) {
        int r = k0 + j, p = ipiv[r];
        if (p != r) { double t = c[r]; c[r] = c[p]; c[p] = t; }
    }
}

__global__ void __launch_bounds__(128) k_lu_trsm(double *A, int lda, int n, int k0, int nb)
{
    __shared__ double L[LU_NB][LU_NB + 1];
    for (int t = threadIdx.x; t < nb * nb; t += blockDim.x) {
        int i = t % nb, j = t / nb;
        L[i][j] = A[(size_t)(k0 + j) * lda + k0 + i];
    }
    __syncthreads();
    int col = k0 + nb + blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= n) return;
    double *c = A + (size_t)col * lda + k0;
    double u[LU_NB];
#pragma unroll
    for (int i = 0; i < LU_NB; i++) u[i] = i < nb ? c[i] : 0.0;
#pragma unroll
    for (int i = 1; i < LU_NB; i++) {
        double s = u[i];
#pragma unroll
        for (int jj = 0; jj < i; jj++) s -= L[i][jj] * u[jj];     // rows >= nb of L are never loaded: guarded by the store below
        u[i] = i < nb ? s : 0.0;
    }
#pragma unroll
    for (int i = 0; i < LU_NB; i++) if (i < nb) c[i] = u[i];
}

__global__ void __launch_bounds__(256) k_lu_gemm(double *A, int lda, int n, int k0, int nb)
{
    __shared__ double sL[LU_NB][64 + 1];      // [k][i]
    __shared__ double sU[LU_NB][64 + 1];      // [k][j]
    const int tid = threadIdx.x;
    const int r0 = k0 + nb + blockIdx.x * 64, c0 = k0 + nb + blockIdx.y * 64;
    for (int t = tid; t < 64 * LU_NB; t += 256) {
        int i = t & 63, k = t >> 6;
        sL[k][i] = (k < nb && r0 + i < n) ? A[(size_t)(k0 + k) * lda + r0 + i] : 0.0;
    }
    for (int t = tid; t < 64 * LU_NB; t += 256) {
        int k = t & (LU_NB - 1), j = t / LU_NB;
        sU[k][j] = (k < nb && c0 + j < n) ? A[(size_t)(c0 + j) * lda + k0 + k] : 0.0;
    }
    __syncthreads();
    const int ti = tid & 15, tj = tid >> 4;
    double acc[4][4];
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int r = 0; r < 4; r++) acc[q][r] = 0.0;
#pragma unroll 8
    for (int k = 0; k < LU_NB; k++) {
        double a[4], b[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { a[q] = sL[k][ti + 16 * q]; b[q] = sU[k][tj + 16 * q]; }
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int r = 0; r < 4; r++) acc[q][r] += a[q] * b[r];
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
        int j = c0 + tj + 16 * r;
        if (j >= n) continue;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int i = r0 + ti + 16 * q;
            if (i < n) A[(size_t)j * lda + i] -= acc[q][r];
        }
    }
}

// x <- A^-1 x with the factors of k_lu_*: row swaps, L y = P x (unit lower), U x = y.  One CTA; the vector stays in global memory
// (the CTA's own writes are visible to it after __syncthreads()).
__global__ void __launch_bounds__(1024) k_lu_solve(const double *__restrict__ A, int lda, int n, const int *__restrict__ ipiv, double *x)
{
    __shared__ double s_y[LU_NB];
    __shared__ double s_D[LU_NB][LU_NB + 1];
    const int tid = threadIdx.x;
    if (tid == 0)
        for (int r = 0; r < n; r++) { int p = ipiv[r]; if (p != r) { double t = x[r]; x[r] = x[p]; x[p] = t; } }
    __syncthreads();
    // forward substitution, 32 columns at a time: the diagonal block is solved serially from shared memory, the rows below in parallel
    for (int k0 = 0; k0 < n; k0 += LU_NB) {
        int nb = min(LU_NB, n - k0);
        for (int t = tid; t < nb * nb; t += blockDim.x) { int i = t % nb, j = t / nb; s_D[i][j] = A[(size_t)(k0 + j) * lda + k0 + i]; }
        __syncthreads();
        if (tid == 0) {
            for (int i = 0; i < nb; i++) {
                double s = x[k0 + i];
                for (int j = 0; j < i; j++) s -= s_D[i][j] * s_y[j];
                s_y[i] = s;
                x[k0 + i] = s;
            }
        }
        __syncthreads();
        for (int i = k0 + nb + tid; i < n; i += blockDim.x) {
            double s = x[i];
            for (int j = 0; j < nb; j++) s -= A[(size_t)(k0 + j) * lda + i] * s_y[j];
            x[i] = s;
        }
        __syncthreads();
    }
    // backward substitution
    for (int k1 = n; k1 > 0; k1 -= LU_NB) {
        int k0 = max(k1 - LU_NB, 0), nb = k1 - k0;
        for (int t = tid; t < nb * nb; t += blockDim.x) { int i = t % nb, j = t / nb; s_D[i][j] = A[(size_t)(k0 + j) * lda + k0 + i]; }
        __syncthreads();
        if (tid == 0) {
            for (int i = nb - 1; i >= 0; i--) {
                double s = x[k0 + i];
                for (int j = i + 1; j < nb; j++) s -= s_D[i][j] * s_y[j];
                s /= s_D[i][i];
                s_y[i] = s;
                x[k0 + i] = s;
            }
        }
        __syncthreads();
        for (int i = tid; i < k0; i += blockDim.x) {
            double s = x[i];
            for (int j = 0; j < nb; j++) s -= A[(size_t)(k0 + j) * lda + i] * s_y[j];
            x[i] = s;
        }
        __syncthreads();
    }
}

}  // namespace tsl
