// tsl_assembly_kernels.cuh -- owner-computes assembly of the cloth's forward Newton matrices on the structured grid (sm_100a).
//
// Replaces, for the forward step, the scatter of BaseScene.compute_Hessian (code/engine/BaseScene.py:1042-1052 ->
// Cloth.compute_Hessian_me / _ma / _bending, code/engine/model_fold_offset.py:467-637 -> H.add atomics): every 3x3 block of a cloth row
// is produced by ONE thread that gathers the elements containing both vertices, so there are no atomics, the result is bit-for-bit
// reproducible, every value is written exactly once with full 128-byte lines, and the exact (A_e) and clamped (A_c) Newton matrices
//  leave the same pass.
//
// One CTA owns a TI x TJ tile of grid vertices (TJ = 32 consecutive vertex ids = one warp-wide store per (row, slot, component)):
//   phase 0  fp64 positions of the tile and a halo of 2 are staged in shared memory;
//   phase 1  the gradient of the dihedral angle of every hinge that touches the tile is computed once, in fp32 from fp64 edge
//            differences, into shared memory (12 floats per hinge; hinges that do not exist hold zeros);
//   phase 2  one thread per (tile vertex, sliced-ELL slot): the column vertex comes from the matrix' own colidx, the list of
//            contributing triangles / hinges from the constant-memory tables of tsl_grid.h (read off the reference's mesher),
//            triangle terms are evaluated on the fly, hinge terms are rank-1 products of the staged gradients.
// Bound: HBM writes (72 B per block for the two matrices); arithmetic is fp32 except the position differences.
// Only CUDA built-ins are used: tests/csrc/cuda_emu.h runs these kernels on the CPU.
#pragma once
#ifndef TSL_CUDA_EMU
#include <cuda_runtime.h>
#define TSL_CONSTANT __constant__
#else
#define TSL_CONSTANT
#endif
#include "tsl_grid.h"

namespace tsl {

static TSL_CONSTANT GridTables c_gt;

struct ClothGrid {
    int N, M, NV, offset;            // (N+1) x (M+1) vertices; first matrix row of the cloth
    float Kl, Ka, Kb, dx, mass_dt2;  // mass / dt^2 of a cloth vertex
};

#define TSL_TI 4
#define TSL_TJ 32
#define TSL_PW (TSL_TJ + 4)          // position window: rows i0-2 .. i0+TI+1, columns j0-2 .. j0+TJ+1
#define TSL_PH (TSL_TI + 4)
#define TSL_HW (TSL_TJ + 2)          // hinge window: anchors rows i0-1 .. i0+TI, columns j0-1 .. j0+TJ
#define TSL_HH (TSL_TI + 2)

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross3(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float comp3(f3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// fp32 edge vector b - a from the staged fp64 positions (the subtraction is exact to fp64, the strain keeps its digits)
__device__ __forceinline__ f3 edge32(const double *a, const double *b) { return mk3((float)(b[0] - a[0]), (float)(b[1] - a[1]), (float)(b[2] - a[2])); }

// gradient of the dihedral angle w.r.t. the hinge vertices p0 (opposite, owner face), p1, p2 (shared edge), p3 (opposite, neighbour):
// Cloth.compute_bending_grad (model_fold_offset.py:379-402).  Faces are consistently oriented: (p0,p1,p2) and (p3,p2,p1).
__device__ __forceinline__ void hinge_grad32(const double *p0, const double *p1, const double *p2, const double *p3, float *g)
{
    f3 e01 = edge32(p0, p1), e02 = edge32(p0, p2), e12 = edge32(p1, p2), e31 = edge32(p3, p1), e32 = edge32(p3, p2);
    f3 n1 = cross3(e01, e12);                       // (p1 - p0) x (p2 - p1)
    f3 n2 = cross3(e32, mk3(-e12.x, -e12.y, -e12.z));                  // (p2 - p3) x (p1 - p2)
    float A1 = sqrtf(dot3(n1, n1)), A2 = sqrtf(dot3(n2, n2));         // twice the face areas
    n1 = (1.f / A1) * n1; n2 = (1.f / A2) * n2;
    float l12 = sqrtf(dot3(e12, e12)), l02 = sqrtf(dot3(e02, e02)), l01 = sqrtf(dot3(e01, e01)), l32 = sqrtf(dot3(e32, e32)), l31 = sqrtf(dot3(e31, e31));
    float h1_p0 = A1 / l12, h1_p1 = A1 / l02, h1_p2 = A1 / l01;
    float h2_p3 = A2 / l12, h2_p1 = A2 / l32, h2_p2 = A2 / l31;
    // cosines of the interior angles at p1 and p2 in both faces
    float c1_p1 = dot3(mk3(-e01.x, -e01.y, -e01.z), e12) / (l01 * l12);            // (p0 - p1) . (p2 - p1)
    float c1_p2 = dot3(mk3(-e02.x, -e02.y, -e02.z), mk3(-e12.x, -e12.y, -e12.z)) / (l02 * l12);   // (p0 - p2) . (p1 - p2)
    float c2_p1 = dot3(mk3(-e31.x, -e31.y, -e31.z), e12) / (l31 * l12);
    float c2_p2 = dot3(mk3(-e32.x, -e32.y, -e32.z), mk3(-e12.x, -e12.y, -e12.z)) / (l32 * l12);
    f3 ga = (-1.f / h1_p0) * n1;
    f3 gd = (-1.f / h2_p3) * n2;
    f3 gb = (c1_p2 / h1_p1) * n1 + (c2_p2 / h2_p1) * n2;
    f3 gc = (c1_p1 / h1_p2) * n1 + (c2_p1 / h2_p2) * n2;
    g[0] = ga.x; g[1] = ga.y; g[2] = ga.z; g[3] = gb.x; g[4] = gb.y; g[5] = gb.z;
    g[6] = gc.x; g[7] = gc.y; g[8] = gc.z; g[9] = gd.x; g[10] = gd.y; g[11] = gd.z;
}

// 3x3 block of one edge spring (delta = x_l - x_{l+1}) of the Newton model, exact (He) and clamped (Hc):
//   dE/dl / l (I - d d^T) + d2E/dl2 d d^T, the first term dropped when the edge is compressed
__device__ __forceinline__ void edge_block32(const ClothGrid &G, f3 dv, int l, float sign, float *He, float *Hc)
{
    float lt = sqrtf(dot3(dv, dv));
    float base = (l == 2) ? G.dx * 1.41421356237309515f : G.dx;
    float dl = -G.Kl * 2.f * (1.f - lt / base), dl2 = G.Kl * 2.f / base;
    float ge = dl / lt, gc = dl > 0.f ? ge : 0.f;
    float d[3] = { dv.x / lt, dv.y / lt, dv.z / lt };
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float dd = d[j] * d[k], id = (j == k) ? 1.f : 0.f;
            He[j * 3 + k] += sign * (ge * (id - dd) + dl2 * dd);
            Hc[j * 3 + k] += sign * (gc * (id - dd) + dl2 * dd);
        }
}

__global__ void __launch_bounds__(256) k_hessian_rows(ClothGrid G, const double *__restrict__ pos, const int *__restrict__ frozen,
                                                      const int *__restrict__ slice_base, const int *__restrict__ colidx,
                                                      const int *__restrict__ diag_pb, float *__restrict__ val_e, float *__restrict__ val_c)
{
    __shared__ double spos[TSL_PH][TSL_PW][3];
    __shared__ float shin[3][TSL_HH][TSL_HW][12];
    const int W = G.M + 1;
    const int i0 = blockIdx.y * TSL_TI, j0 = blockIdx.x * TSL_TJ;
    const int tid = threadIdx.x;
    // ---- phase 0: positions (outside the grid: zeros, never used by an existing element)
    for (int t = tid; t < TSL_PH * TSL_PW; t += blockDim.x) {
        int r = t / TSL_PW, c = t - r * TSL_PW;
        int i = i0 - 2 + r, j = j0 - 2 + c;
        double x = 0, y = 0, z = 0;
        if (i >= 0 && i <= G.N && j >= 0 && j <= G.M) {
            const double *p = pos + 3 * (size_t)(G.offset + i * W + j);
            x = p[0]; y = p[1]; z = p[2];
        }
        spos[r][c][0] = x; spos[r][c][1] = y; spos[r][c][2] = z;
    }
    __syncthreads();
    // ---- phase 1: hinge gradients
    for (int t = tid; t < 3 * TSL_HH * TSL_HW; t += blockDim.x) {
        int type = t / (TSL_HH * TSL_HW), rem = t - type * (TSL_HH * TSL_HW);
        int r = rem / TSL_HW, c = rem - r * TSL_HW;
        int ai = i0 - 1 + r, aj = j0 - 1 + c;
        float g[12];
#pragma unroll
        for (int q = 0; q < 12; q++) g[q] = 0.f;
        int par = (ai + aj) & 1;
        bool ex = true;
        const double *pp[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int vi = ai + c_gt.hin_v[type][par][q][0], vj = aj + c_gt.hin_v[type][par][q][1];
            ex = ex && vi >= 0 && vi <= G.N && vj >= 0 && vj <= G.M;
            int pr = vi - (i0 - 2), pc = vj - (j0 - 2);
            ex = ex && pr >= 0 && pr < TSL_PH && pc >= 0 && pc < TSL_PW;       // (always true for a hinge the tile needs)
            pp[q] = ex ? &spos[pr][pc][0] : &spos[0][0][0];
        }
        if (ex) hinge_grad32(pp[0], pp[1], pp[2], pp[3], g);
#pragma unroll
        for (int q = 0; q < 12; q++) shin[type][r][c][q] = g[q];
    }
    __syncthreads();
    // ---- phase 2: one thread per (tile vertex, matrix slot)
    const float d2 = 2.f * G.Kb * G.dx * G.dx * (1.f / 3.f);
    const float V = 0.5f * G.dx * G.dx, da2 = G.Ka * 2.f / V;
    for (int item = tid; item < TSL_TI * 13 * TSL_TJ; item += blockDim.x) {
        int tj = item % TSL_TJ, rest = item / TSL_TJ, k = rest % 13, ti = rest / 13;
        int i = i0 + ti, j = j0 + tj;
        if (i > G.N || j > G.M) continue;
        int v = i * W + j, row = G.offset + v;
        int S = row >> 5, lane = row & 31;
        int b0 = slice_base[S], b1 = slice_base[S + 1];
        int pb = b0 + 32 * k + lane;
        if (pb >= b1) continue;
        int col = colidx[pb];
        if (col == row && pb != diag_pb[row]) continue;                      // ELL padding: stays zero
        int u = col - G.offset;
        if (u < 0 || u >= G.NV) continue;
        int ui = u / W, uj = u - ui * W;
        int di = ui - i, dj = uj - j;
        if (di < -2 || di > 2 || dj < -2 || dj > 2) continue;
        int slot = (di + 2) * 5 + (dj + 2), p = (i + j) & 1;
        float Be[9], Bc[9];
#pragma unroll
        for (int q = 0; q < 9; q++) { Be[q] = 0.f; Bc[q] = 0.f; }
        // triangles containing both vertices
        int nt = c_gt.n_tri[p][slot];
        for (int n = 0; n < nt; n++) {
            GridTables::TriE e = c_gt.tri[p][slot][n];
            int qi = i + e.qi, qj = j + e.qj;
            if (qi < 0 || qi >= G.N || qj < 0 || qj >= G.M) continue;
            int qp = (qi + qj) & 1;
            const double *x[3];
#pragma unroll
            for (int l = 0; l < 3; l++)
                x[l] = &spos[qi + c_gt.tri_v[qp][e.t][l][0] - (i0 - 2)][qj + c_gt.tri_v[qp][e.t][l][1] - (j0 - 2)][0];
            f3 e1 = edge32(x[0], x[1]), e2 = edge32(x[0], x[2]);
            f3 nn = cross3(e1, e2);
            float nl = sqrtf(dot3(nn, nn));
            float da = -G.Ka * 2.f * (1.f - 0.5f * nl / V);
            float sae = da / (2.f * nl), sac = da > 0.f ? sae : 0.f;
            f3 nh = (1.f / nl) * nn;
            // w_a: the edge opposite to local vertex a (J_a[j] = e_j x w_a); g_a = 0.5 w_a x nh
            f3 w[3] = { e1 - e2, e2, mk3(-e1.x, -e1.y, -e1.z) };
            const int a = e.a, b = e.b;
            f3 wa = w[a], wb = w[b];
            f3 ga = 0.5f * cross3(wa, nh), gb = 0.5f * cross3(wb, nh);
            float wab = dot3(wa, wb);
#pragma unroll
            for (int jj = 0; jj < 3; jj++)
#pragma unroll
                for (int kk = 0; kk < 3; kk++) {
                    float gg = comp3(ga, jj) * comp3(gb, kk);
                    float jtj = ((jj == kk) ? wab : 0.f) - comp3(wb, jj) * comp3(wa, kk);
                    Be[jj * 3 + kk] += da2 * gg + sae * (jtj - 4.f * gg);
                    Bc[jj * 3 + kk] += da2 * gg + sac * (jtj - 4.f * gg);
                }
            if (a == b) {
                const int lp = (a + 2) % 3;
                edge_block32(G, edge32(x[(a + 1) % 3], x[a]), a, 1.f, Be, Bc);          // edge a joins (a, a+1): delta = x_a - x_{a+1}
                edge_block32(G, edge32(x[a], x[lp]), lp, 1.f, Be, Bc);                  // edge lp joins (lp, a): delta = x_lp - x_a
            } else {
                const int l = ((a + 1) % 3 == b) ? a : b;
                edge_block32(G, edge32(x[(l + 1) % 3], x[l]), l, -1.f, Be, Bc);
            }
        }
        // hinges containing both vertices: d2E/dtheta2 grad(theta)_j grad(theta)_k^T (the same in both matrices)
        int nh_ = c_gt.n_hin[p][slot];
        for (int n = 0; n < nh_; n++) {
            GridTables::HinE e = c_gt.hin[p][slot][n];
            int r = i + e.ei - (i0 - 1), c = j + e.ej - (j0 - 1);
            const float *g = &shin[e.type][r][c][0];
            const float *gj = g + 3 * e.j, *gk = g + 3 * e.k;
#pragma unroll
            for (int jj = 0; jj < 3; jj++)
#pragma unroll
                for (int kk = 0; kk < 3; kk++) {
                    float h = d2 * gj[jj] * gk[kk];
                    Be[jj * 3 + kk] += h; Bc[jj * 3 + kk] += h;
                }
        }
        // frozen mask (BaseScene.add_H :399-402), then the mass diagonal on every DOF (H.add, quirk Q6)
        int fr[3] = { frozen[3 * row], frozen[3 * row + 1], frozen[3 * row + 2] };
        int fc[3] = { frozen[3 * col], frozen[3 * col + 1], frozen[3 * col + 2] };
#pragma unroll
        for (int jj = 0; jj < 3; jj++)
#pragma unroll
            for (int kk = 0; kk < 3; kk++)
                if (fr[jj] || fc[kk]) { Be[jj * 3 + kk] = 0.f; Bc[jj * 3 + kk] = 0.f; }
        if (col == row) {
            Be[0] += G.mass_dt2; Be[4] += G.mass_dt2; Be[8] += G.mass_dt2;
            Bc[0] += G.mass_dt2; Bc[4] += G.mass_dt2; Bc[8] += G.mass_dt2;
        }
        long long base = (long long)(pb - lane) * 9 + lane;
#pragma unroll
        for (int q = 0; q < 9; q++) { val_e[base + q * 32] = Be[q]; val_c[base + q * 32] = Bc[q]; }
    }
}

}  // namespace tsl
