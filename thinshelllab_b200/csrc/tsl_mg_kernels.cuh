// tsl_mg_kernels.cuh -- setup kernels of the geometric multigrid hierarchy (tsl_mg.cu): Galerkin products P^T A P on the cloth grid.
// Only CUDA built-ins are used, so the CPU test suite runs them through tests/csrc/cuda_emu.h.
#pragma once
#ifndef TSL_CUDA_EMU
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#endif

namespace tsl {

// The SMOOTHER's matrix passes on the bandwidth-bound levels (level-0 snapshot, element-major coarse levels) may read an fp16 copy of the
// operator, scaled by a power of two per level so that the largest entry stays far from 65504 (tsl_mg.cu: k_mg_scales): damping the
// high-frequency error does not need more than 11 bits of the operator.  Everything that feeds a coarse-grid SOLVE -- the Galerkin
// chain and the residual that is restricted -- stays in fp32: with all-fp16 operators the 1 M-triangle sheet at rest (lambda_min / a_ii
// ~ 1e-3) needed 13 200 PCG iterations for a step that takes 60 (measured, profiles/README.md).  mg_ld / mg_st hide the storage type.
__device__ __forceinline__ float mg_ld(const float *p) { return *p; }
__device__ __forceinline__ void mg_st(float *p, float v) { *p = v; }
#ifndef TSL_CUDA_EMU
__device__ __forceinline__ float mg_ld(const __half *p) { return __half2float(*p); }
__device__ __forceinline__ void mg_st(__half *p, float v) { *p = __float2half_rn(v); }
#endif

// 1-D bilinear weight of fine index 2I + a (a in -1..1) towards coarse parent I of nc coarse points
__device__ __forceinline__ float pw1(int a, int I, int nc) { return a == 0 ? 1.f : (a < 0 ? 0.5f : (I + 1 < nc ? 0.5f : 1.f)); }

// stencil copy of the cloth block of the sliced-ELL matrix (input of the first Galerkin product)
__global__ void k_sell_to_stencil(int off, int nvc, int n1, const int *__restrict__ slice_base, const int *__restrict__ colidx,
                                  const float *__restrict__ val, const int *__restrict__ diag_pb, float *out, long long sv, long long se)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nvc) return;
    int row = off + v;
    int S = row >> 5, lane = row & 31;
    int i = v / n1, j = v - i * n1;
    int b0 = slice_base[S], b1 = slice_base[S + 1];
    int dpb = diag_pb[row];
    for (int b = b0; b < b1; b += 32) {
        int pb = b + lane;
        int col = colidx[pb];
        int cv = col - off;
        if (cv < 0 || cv >= nvc) continue;
        if (col == row && pb != dpb) continue;          // ELL padding (zero block pointing at the diagonal)
        int ip = cv / n1, jp = cv - ip * n1;
        int di = ip - i, dj = jp - j;
        if (di < -2 || di > 2 || dj < -2 || dj > 2) continue;
        int slot = (di + 2) * 5 + (dj + 2);
        const float *src = val + (long long)b * 9 + lane;
        float *dst = out + (size_t)v * sv + (size_t)(slot * 9) * se;
#pragma unroll
        for (int c = 0; c < 9; c++) dst[(size_t)c * se] = src[c * 32];
    }
}
// A_c = P^T A_f P, one thread per (coarse vertex, coarse stencil slot).  mask: frozen flags of the fine grid's
// DOFs ([3 * nvf], level 0 only) -- frozen DOFs are left out of the coarse spaces.
template <bool MASK>
__global__ void __launch_bounds__(128) k_galerkin(const float *__restrict__ val_f, int n0f, int n1f, long long svf, long long sef, const int *__restrict__ mask,
                                                  float *val_c, int n0c, int n1c, long long svc, long long sec)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int nvc = n0c * n1c;
    if (t >= nvc * 25) return;
    // element-major coarse level: consecutive threads = consecutive vertices of one slot; row-major: consecutive slots
    int cv, slot;
    if (svc == 1) { slot = t / nvc; cv = t - slot * nvc; } else { cv = t / 25; slot = t - cv * 25; }
    int I = cv / n1c, J = cv - I * n1c;
    int Ip = I + slot / 5 - 2, Jp = J + slot % 5 - 2;
    float acc[9];
#pragma unroll
    for (int c = 0; c < 9; c++) acc[c] = 0.f;
    if ((unsigned)Ip < (unsigned)n0c && (unsigned)Jp < (unsigned)n1c) {
        for (int a = -1; a <= 1; a++) {
            int i = 2 * I + a;
            if ((unsigned)i >= (unsigned)n0f) continue;
            float wi = pw1(a, I, n0c);
            for (int b = -1; b <= 1; b++) {
                int j = 2 * J + b;
                if ((unsigned)j >= (unsigned)n1f) continue;
                float wr = wi * pw1(b, J, n1c);
                int fv = i * n1f + j;
                for (int ap = -1; ap <= 1; ap++) {
                    int ip = 2 * Ip + ap;
                    int di = ip - i;
                    if ((unsigned)ip >= (unsigned)n0f || di < -2 || di > 2) continue;
                    float wip = wr * pw1(ap, Ip, n0c);
                    for (int bp = -1; bp <= 1; bp++) {
                        int jp = 2 * Jp + bp;
                        int dj = jp - j;
                        if ((unsigned)jp >= (unsigned)n1f || dj < -2 || dj > 2) continue;
                        float w = wip * pw1(bp, Jp, n1c);
                        const float *src = val_f + (size_t)fv * svf + (size_t)(((di + 2) * 5 + (dj + 2)) * 9) * sef;
                        if (MASK) {
                            int fc = ip * n1f + jp;
                            float mr[3], mc[3];
#pragma unroll
                            for (int q = 0; q < 3; q++) { mr[q] = mask[3 * fv + q] ? 0.f : w; mc[q] = mask[3 * fc + q] ? 0.f : 1.f; }
#pragma unroll
                            for (int c = 0; c < 9; c++) acc[c] += mr[c / 3] * mc[c % 3] * __ldg(src + (size_t)c * sef);
                        } else {
#pragma unroll
                            for (int c = 0; c < 9; c++) acc[c] += w * __ldg(src + (size_t)c * sef);
                        }
                    }
                }
            }
        }
    }
    float *dst = val_c + (size_t)cv * svc + (size_t)(slot * 9) * sec;
#pragma unroll
    for (int c = 0; c < 9; c++) dst[(size_t)c * sec] = acc[c];
}

// ------------------------------------------------------------------------------------------------ tiled Galerkin products
// k_galerkin above re-reads every fine stencil row ~25 x 9 times through L1 / L2 (one thread per (coarse vertex, slot)): at 1 M
// triangles the level 0 -> 1 product moved 5.6 GB through the L2 and took 793 us (profiles/r2_launches_kernels707.md).  The tiled
// kernels stage the fine rows a tile of TCI x TCJ coarse vertices needs ((2 TCI + 1) x (2 TCJ + 1) fine rows x 225 floats) in shared
// memory once and read them from there; TCJ = 8 makes the element-major output runs full 32-byte sectors.
#define TSL_TCI 2
#define TSL_TCJ 8
#define TSL_TFI (2 * TSL_TCI + 1)
#define TSL_TFJ (2 * TSL_TCJ + 1)
#define TSL_GAL_SMEM (sizeof(float) * TSL_TFI * TSL_TFJ * 225)
#ifndef TSL_CUDA_EMU
#define TSL_DYN_SMEM_F(name) extern __shared__ __align__(16) unsigned char name##_bytes[]; float *name = (float *)name##_bytes
#else
#define TSL_DYN_SMEM_F(name) static float name[1 << 16]
#endif

// the Galerkin sums of one CTA from the staged fine rows sA[(fi * TFJ + fj) * 225 + slot * 9 + c]; fine window starts at (fi0, fj0)
// val_c: the coarse operator in fp32 (the Galerkin chain and the residuals stay in fp32: the smooth modes a coarse level exists for live
// on cancellations between entries that 11 bits do not preserve); val_h (optional): a scaled fp16 copy for the smoother's matrix passes
template <bool MASK, class HT>
__device__ __forceinline__ void galerkin_tile_compute(const float *sA, int fi0, int fj0, int I0, int J0, int n0f, int n1f, const int *__restrict__ mask,
                                                      float *val_c, HT *val_h, int n0c, int n1c, long long svc, long long sec, float h_mul)
{
    for (int item = threadIdx.x; item < TSL_TCI * TSL_TCJ * 25; item += blockDim.x) {
        // consecutive threads = consecutive J of one (I, slot): coalesced stores on element-major levels
        int tJ = item % TSL_TCJ, rest = item / TSL_TCJ, slot = rest % 25, tI = rest / 25;
        int I = I0 + tI, J = J0 + tJ;
        if (I >= n0c || J >= n1c) continue;
        int Ip = I + slot / 5 - 2, Jp = J + slot % 5 - 2;
        float acc[9];
#pragma unroll
        for (int c = 0; c < 9; c++) acc[c] = 0.f;
        if ((unsigned)Ip < (unsigned)n0c && (unsigned)Jp < (unsigned)n1c) {
            for (int a = -1; a <= 1; a++) {
                int i = 2 * I + a;
                if ((unsigned)i >= (unsigned)n0f) continue;
                float wi = pw1(a, I, n0c);
                for (int b = -1; b <= 1; b++) {
                    int j = 2 * J + b;
                    if ((unsigned)j >= (unsigned)n1f) continue;
                    float wr = wi * pw1(b, J, n1c);
                    const float *row = sA + (size_t)((i - fi0) * TSL_TFJ + (j - fj0)) * 225;
                    int fv = i * n1f + j;
                    for (int ap = -1; ap <= 1; ap++) {
                        int ip = 2 * Ip + ap, di = ip - i;
                        if ((unsigned)ip >= (unsigned)n0f || di < -2 || di > 2) continue;
                        float wip = wr * pw1(ap, Ip, n0c);
                        for (int bp = -1; bp <= 1; bp++) {
                            int jp = 2 * Jp + bp, dj = jp - j;
                            if ((unsigned)jp >= (unsigned)n1f || dj < -2 || dj > 2) continue;
                            float w = wip * pw1(bp, Jp, n1c);
                            const float *src = row + ((di + 2) * 5 + (dj + 2)) * 9;
                            if (MASK) {
                                int fc = ip * n1f + jp;
                                float mr[3], mc[3];
#pragma unroll
                                for (int q = 0; q < 3; q++) { mr[q] = mask[3 * fv + q] ? 0.f : w; mc[q] = mask[3 * fc + q] ? 0.f : 1.f; }
#pragma unroll
                                for (int c = 0; c < 9; c++) acc[c] += mr[c / 3] * mc[c % 3] * src[c];
                            } else {
#pragma unroll
                                for (int c = 0; c < 9; c++) acc[c] += w * src[c];
                            }
                        }
                    }
                }
            }
        }
        const size_t o = (size_t)(I * n1c + J) * svc + (size_t)(slot * 9) * sec;
#pragma unroll
        for (int c = 0; c < 9; c++) val_c[o + (size_t)c * sec] = acc[c];
        if (val_h) {
#pragma unroll
            for (int c = 0; c < 9; c++) mg_st(val_h + o + (size_t)c * sec, acc[c] * h_mul);
        }
    }
}

// level l >= 1 -> l + 1: fine operator in the level's stencil layout (element (v, e) at val_f[v * svf + e * sef])
// val_h / sc_c (optional): fp16 copy of the coarse operator, stored x sc_c[0]
template <class HT>
__global__ void __launch_bounds__(256) k_galerkin_tiled(const float *__restrict__ val_f, int n0f, int n1f, long long svf, long long sef,
                                                        float *val_c, HT *val_h, int n0c, int n1c, long long svc, long long sec,
                                                        const float *__restrict__ sc_c)
{
    TSL_DYN_SMEM_F(sA);
    const int I0 = blockIdx.y * TSL_TCI, J0 = blockIdx.x * TSL_TCJ;
    const int fi0 = 2 * I0 - 1, fj0 = 2 * J0 - 1;
    // staging: consecutive threads = consecutive fine columns of one (row, element) -- the contiguous direction of element-major levels
    for (int t = threadIdx.x; t < TSL_TFI * 225 * TSL_TFJ; t += blockDim.x) {
        int fj = t % TSL_TFJ, rest = t / TSL_TFJ, e = rest % 225, fi = rest / 225;
        int i = fi0 + fi, j = fj0 + fj;
        float v = 0.f;
        if ((unsigned)i < (unsigned)n0f && (unsigned)j < (unsigned)n1f) v = val_f[(size_t)(i * n1f + j) * svf + (size_t)e * sef];
        sA[(size_t)(fi * TSL_TFJ + fj) * 225 + e] = v;
    }
    __syncthreads();
    galerkin_tile_compute<false, HT>(sA, fi0, fj0, I0, J0, n0f, n1f, nullptr, val_c, val_h, n0c, n1c, svc, sec, sc_c ? sc_c[0] : 1.f);
}

// level 0 -> 1 straight from the sliced-ELL matrix (no stencil copy of the fine level): rows [off, off + n0f * n1f) of the matrix are
// the cloth grid; mask = frozen flags of those rows' DOFs (relative to `off`): frozen DOFs are left out of the coarse spaces
template <class HT>
__global__ void __launch_bounds__(256) k_galerkin_sell_tiled(int off, int n0f, int n1f, const int *__restrict__ slice_base, const int *__restrict__ colidx,
                                                             const float *__restrict__ val, const int *__restrict__ diag_pb, const int *__restrict__ mask,
                                                             float *val_c, HT *val_h, int n0c, int n1c, long long svc, long long sec, const float *__restrict__ sc_c)
{
    TSL_DYN_SMEM_F(sA);
    const int I0 = blockIdx.y * TSL_TCI, J0 = blockIdx.x * TSL_TCJ;
    const int fi0 = 2 * I0 - 1, fj0 = 2 * J0 - 1;
    for (int t = threadIdx.x; t < TSL_TFI * TSL_TFJ * 225; t += blockDim.x) sA[t] = 0.f;
    __syncthreads();
    const int nvc = n0f * n1f;
    // one thread per (fine row of the window, sliced-ELL slot k <= 13): consecutive threads = consecutive rows = consecutive lanes of a slice
    for (int t = threadIdx.x; t < TSL_TFI * 16 * TSL_TFJ; t += blockDim.x) {
        int fj = t % TSL_TFJ, rest = t / TSL_TFJ, k = rest % 16, fi = rest / 16;
        int i = fi0 + fi, j = fj0 + fj;
        if ((unsigned)i >= (unsigned)n0f || (unsigned)j >= (unsigned)n1f) continue;
        int row = off + i * n1f + j;
        int S = row >> 5, lane = row & 31;
        int b0 = slice_base[S], b1 = slice_base[S + 1];
        int pb = b0 + 32 * k + lane;
        if (pb >= b1) continue;
        int col = colidx[pb], cv = col - off;
        if (cv < 0 || cv >= nvc) continue;
        if (col == row && pb != diag_pb[row]) continue;          // ELL padding
        int ip = cv / n1f, jp = cv - ip * n1f;
        int di = ip - i, dj = jp - j;
        if (di < -2 || di > 2 || dj < -2 || dj > 2) continue;
        const float *src = val + (long long)(pb - lane) * 9 + lane;
        float *dst = sA + (size_t)(fi * TSL_TFJ + fj) * 225 + ((di + 2) * 5 + (dj + 2)) * 9;
#pragma unroll
        for (int c = 0; c < 9; c++) dst[c] = src[c * 32];
    }
    __syncthreads();
    galerkin_tile_compute<true, HT>(sA, fi0, fj0, I0, J0, n0f, n1f, mask, val_c, val_h, n0c, n1c, svc, sec, sc_c ? sc_c[0] : 1.f);
}

}  // namespace tsl
