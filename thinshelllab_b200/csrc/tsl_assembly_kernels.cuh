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
#define TSL_DYN_SMEM(T, name) extern __shared__ __align__(16) unsigned char name##_bytes[]; T *name = (T *)name##_bytes
#else
#define TSL_CONSTANT
#define TSL_DYN_SMEM(T, name) static T name[1 << 16]
#endif
#include "tsl_elements.cuh"
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

#define TSL_HREC 13                  // floats per staged hinge record (12 used; the odd stride keeps shared-memory loads conflict-free)
#define TSL_TREC 21                  // floats per staged triangle record (20 used)
#define TSL_TW (TSL_TJ + 1)          // triangle window: quads rows i0-1 .. i0+TI-1, columns j0-1 .. j0+TJ-1
#define TSL_TH (TSL_TI + 1)
#define TSL_HESS_SMEM (sizeof(double) * TSL_PH * TSL_PW * 3 + sizeof(float) * (3 * TSL_HH * TSL_HW * TSL_HREC + TSL_TH * TSL_TW * 2 * TSL_TREC))

// Staged triangle record (Newton model), vertices x0, x1, x2 in f2v order:
//   [0..2] w1 = x2 - x0, [3..5] w2 = x0 - x1  (w_a = the edge opposite to local vertex a; w0 = -(w1 + w2); d n / d x_a,j = e_j x w_a)
//   [6..8] unit normal, [9] sa exact = dE/dA / (2 |n|), [10] sa clamped (dropped when the triangle is compressed)
//   [11 + 3l ..] edge l (joins l, l+1; direction w_{(l+2)%3}): 1 / length, dE/dl / l exact, the same clamped (dropped when compressed)
__device__ __forceinline__ void tri_record32(const ClothGrid &G, const double *x0, const double *x1, const double *x2, float *T)
{
    f3 e1 = edge32(x0, x1), e2 = edge32(x0, x2);
    f3 nn = cross3(e1, e2);
    float nl = sqrtf(dot3(nn, nn)), inl = 1.f / nl;
    const float V = 0.5f * G.dx * G.dx;
    float da = -G.Ka * 2.f * (1.f - 0.5f * nl / V);
    float sae = 0.5f * da * inl;
    T[0] = e2.x; T[1] = e2.y; T[2] = e2.z; T[3] = -e1.x; T[4] = -e1.y; T[5] = -e1.z;
    T[6] = nn.x * inl; T[7] = nn.y * inl; T[8] = nn.z * inl;
    T[9] = sae; T[10] = da > 0.f ? sae : 0.f;
    f3 w0 = e1 - e2;
    f3 dv[3] = { mk3(-e1.x, -e1.y, -e1.z), w0, e2 };          // edge 0: x0 - x1, edge 1: x1 - x2, edge 2: x2 - x0
#pragma unroll
    for (int l = 0; l < 3; l++) {
        float lt = sqrtf(dot3(dv[l], dv[l])), il = 1.f / lt;
        float base = (l == 2) ? G.dx * 1.41421356237309515f : G.dx;
        float dl = -G.Kl * 2.f * (1.f - lt / base);
        float ge = dl * il;
        T[11 + 3 * l] = il; T[12 + 3 * l] = ge; T[13 + 3 * l] = dl > 0.f ? ge : 0.f;
    }
}

// += sign * [ g (I - d d^T) + dl2 d d^T ] of staged edge l, exact (ge) and clamped (gc)
__device__ __forceinline__ void edge_add32(const float *T, f3 w0, f3 w1, f3 w2, int l, float dl2, float sign, float *Be, float *Bc)
{
    f3 dv = (l == 0) ? w2 : (l == 1 ? w0 : w1);
    float il = T[11 + 3 * l], ge = T[12 + 3 * l], gc = T[13 + 3 * l];
    float d[3] = { dv.x * il, dv.y * il, dv.z * il };
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float dd = d[j] * d[k], id = (j == k) ? 1.f : 0.f;
            Be[j * 3 + k] += sign * (ge * (id - dd) + dl2 * dd);
            Bc[j * 3 + k] += sign * (gc * (id - dd) + dl2 * dd);
        }
}

__global__ void __launch_bounds__(256) k_hessian_rows(ClothGrid G, const double *__restrict__ pos, const int *__restrict__ frozen,
                                                      const int *__restrict__ slice_base, const int *__restrict__ colidx,
                                                      const int *__restrict__ diag_pb, float *__restrict__ val_e, float *__restrict__ val_c)
{
    TSL_DYN_SMEM(double, smem_raw);
    double (*spos)[TSL_PW][3] = (double (*)[TSL_PW][3])smem_raw;
    float (*shin)[TSL_HH][TSL_HW][TSL_HREC] = (float (*)[TSL_HH][TSL_HW][TSL_HREC])(smem_raw + TSL_PH * TSL_PW * 3);
    float (*stri)[TSL_TW][2][TSL_TREC] = (float (*)[TSL_TW][2][TSL_TREC])((float *)(smem_raw + TSL_PH * TSL_PW * 3) + 3 * TSL_HH * TSL_HW * TSL_HREC);
    __shared__ int s_frozen;
    const int W = G.M + 1;
    const int i0 = blockIdx.y * TSL_TI, j0 = blockIdx.x * TSL_TJ;
    const int tid = threadIdx.x;
    if (tid == 0) s_frozen = 0;
    __syncthreads();
    // ---- phase 0: positions (outside the grid: zeros, never used by an existing element); does the tile see a frozen DOF at all?
    for (int t = tid; t < TSL_PH * TSL_PW; t += blockDim.x) {
        int r = t / TSL_PW, c = t - r * TSL_PW;
        int i = i0 - 2 + r, j = j0 - 2 + c;
        double x = 0, y = 0, z = 0;
        if (i >= 0 && i <= G.N && j >= 0 && j <= G.M) {
            size_t row = (size_t)(G.offset + i * W + j);
            const double *p = pos + 3 * row;
            x = p[0]; y = p[1]; z = p[2];
            if (frozen[3 * row] | frozen[3 * row + 1] | frozen[3 * row + 2]) atomicOr(&s_frozen, 1);
        }
        spos[r][c][0] = x; spos[r][c][1] = y; spos[r][c][2] = z;
    }
    __syncthreads();
    // ---- phase 1: hinge gradients and triangle records, each once per tile
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
    for (int t = tid; t < TSL_TH * TSL_TW * 2; t += blockDim.x) {
        int tt = t & 1, rem = t >> 1;
        int r = rem / TSL_TW, c = rem - r * TSL_TW;
        int qi = i0 - 1 + r, qj = j0 - 1 + c;
        float T[20];
#pragma unroll
        for (int q = 0; q < 20; q++) T[q] = 0.f;
        if (qi >= 0 && qi < G.N && qj >= 0 && qj < G.M) {
            int qp = (qi + qj) & 1;
            const double *x[3];
#pragma unroll
            for (int l = 0; l < 3; l++) x[l] = &spos[r + 1 + c_gt.tri_v[qp][tt][l][0]][c + 1 + c_gt.tri_v[qp][tt][l][1]][0];
            tri_record32(G, x[0], x[1], x[2], T);
        }
#pragma unroll
        for (int q = 0; q < 20; q++) stri[r][c][tt][q] = T[q];
    }
    __syncthreads();
    // ---- phase 2: one thread per (tile vertex, matrix slot).  A warp = one slot k, one vertex parity, two tile rows: lanes 0-15 the
    // vertices of that parity in row 2 * pair, lanes 16-31 in row 2 * pair + 1 -- every lane walks the same contribution lists
    const float d2 = 2.f * G.Kb * G.dx * G.dx * (1.f / 3.f);
    const float da2 = G.Ka * 2.f / (0.5f * G.dx * G.dx);
    const float dl2_axis = G.Kl * 2.f / G.dx, dl2_diag = G.Kl * 2.f / (G.dx * 1.41421356237309515f);
    const int lane_ = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const bool any_frozen = s_frozen != 0;
    for (int witem = warp; witem < (TSL_TI / 2) * 13 * 2; witem += nwarps) {
        const int pair = witem / 26, rem = witem - pair * 26, k = rem >> 1, p = rem & 1;
        const int ti = 2 * pair + (lane_ >> 4);
        const int i = i0 + ti;
        const int tj = 2 * (lane_ & 15) + ((p + i) & 1);
        const int j = j0 + tj;
        if (i > G.N || j > G.M) continue;
        const int row = G.offset + i * W + j;
        const int S = row >> 5, lane = row & 31;
        const int b0 = slice_base[S], b1 = slice_base[S + 1];
        const int pb = b0 + 32 * k + lane;
        if (pb >= b1) continue;
        const int col = colidx[pb];
        if (col == row && pb != diag_pb[row]) continue;                      // ELL padding: stays zero
        const int u = col - G.offset;
        if (u < 0 || u >= G.NV) continue;
        const int ui = u / W, uj = u - ui * W;                                // (exact for every grid width, also W < 5)
        const int di = ui - i, dj = uj - j;
        if (di < -2 || di > 2 || dj < -2 || dj > 2) continue;
        const int slot = (di + 2) * 5 + (dj + 2);
        float Be[9], Bc[9];
#pragma unroll
        for (int q = 0; q < 9; q++) { Be[q] = 0.f; Bc[q] = 0.f; }
        // triangles containing both vertices (records of triangles outside the grid are zero)
        const int nt = c_gt.n_tri[p][slot];
        for (int n = 0; n < nt; n++) {
            GridTables::TriE e = c_gt.tri[p][slot][n];
            const float *T = &stri[ti + e.qi + 1][tj + e.qj + 1][e.t][0];
            f3 w1 = mk3(T[0], T[1], T[2]), w2 = mk3(T[3], T[4], T[5]), nh = mk3(T[6], T[7], T[8]);
            f3 w0 = mk3(-w1.x - w2.x, -w1.y - w2.y, -w1.z - w2.z);
            const int a = e.a, b = e.b;
            f3 wa = (a == 0) ? w0 : (a == 1 ? w1 : w2), wb = (b == 0) ? w0 : (b == 1 ? w1 : w2);
            f3 ga = 0.5f * cross3(wa, nh), gb = 0.5f * cross3(wb, nh);
            float wab = dot3(wa, wb), sae = T[9], sac = T[10];
            float ce = da2 - 4.f * sae, cc = da2 - 4.f * sac;
#pragma unroll
            for (int jj = 0; jj < 3; jj++)
#pragma unroll
                for (int kk = 0; kk < 3; kk++) {
                    float gg = comp3(ga, jj) * comp3(gb, kk);
                    float jtj = ((jj == kk) ? wab : 0.f) - comp3(wb, jj) * comp3(wa, kk);
                    Be[jj * 3 + kk] += ce * gg + sae * jtj;
                    Bc[jj * 3 + kk] += cc * gg + sac * jtj;
                }
            if (a == b) {
                const int lp = (a + 2) % 3;
                edge_add32(T, w0, w1, w2, a, a == 2 ? dl2_diag : dl2_axis, 1.f, Be, Bc);
                edge_add32(T, w0, w1, w2, lp, lp == 2 ? dl2_diag : dl2_axis, 1.f, Be, Bc);
            } else {
                const int l = ((a + 1) % 3 == b) ? a : b;
                edge_add32(T, w0, w1, w2, l, l == 2 ? dl2_diag : dl2_axis, -1.f, Be, Bc);
            }
        }
        // hinges containing both vertices: d2E/dtheta2 grad(theta)_j grad(theta)_k^T (the same in both matrices; zero records outside)
        const int nh_ = c_gt.n_hin[p][slot];
        for (int n = 0; n < nh_; n++) {
            GridTables::HinE e = c_gt.hin[p][slot][n];
            const float *g = &shin[e.type][ti + e.ei + 1][tj + e.ej + 1][0];
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
        if (any_frozen) {
            int fr[3] = { frozen[3 * row], frozen[3 * row + 1], frozen[3 * row + 2] };
            int fc[3] = { frozen[3 * col], frozen[3 * col + 1], frozen[3 * col + 2] };
#pragma unroll
            for (int jj = 0; jj < 3; jj++)
#pragma unroll
                for (int kk = 0; kk < 3; kk++)
                    if (fr[jj] || fc[kk]) { Be[jj * 3 + kk] = 0.f; Bc[jj * 3 + kk] = 0.f; }
        }
        if (col == row) {
            Be[0] += G.mass_dt2; Be[4] += G.mass_dt2; Be[8] += G.mass_dt2;
            Bc[0] += G.mass_dt2; Bc[4] += G.mass_dt2; Bc[8] += G.mass_dt2;
        }
        const long long base = (long long)(pb - lane) * 9 + lane;
#pragma unroll
        for (int q = 0; q < 9; q++) { val_e[base + q * 32] = Be[q]; val_c[base + q * 32] = Bc[q]; }
    }
}

// ================================================================================================ fp64 residual and energy
// The residual is the reference's exact gradient (Cloth.compute_residual, model_fold_offset.py:640-687) and decides the fixed point
// of the step, the energy (Cloth.compute_energy :190-218) decides every line search: both stay fp64.  Same tiles as above;
// the dihedral angle (acos, two face normals), its side test (quirk Q3) and its gradient are evaluated once per hinge and tile.
struct ClothGrid64 {
    int N, M, NV, offset;
    double Kl, Ka, Kb, dx, dt, mass;
    double g[3];                     // the scene's gravity (used when no per-vertex gravity array is bound)
};

struct Hinge64 { d3 pt[4]; d3 n1, n2; double th; };
// signed dihedral angle of the hinge anchored at (ai, aj) minus nothing: returns false when the hinge does not exist.
// *face_l = 3 * owner face + owner slot (index into ref_angle)
__device__ __forceinline__ bool load_hinge64(const ClothGrid64 &G, const double (*spos)[TSL_PW][3], int i0, int j0, int type, int ai, int aj,
                                             Hinge64 &h, int *face_l)
{
    if (ai < 0 || aj < 0 || ai > G.N || aj > G.M) return false;
    const int par = (ai + aj) & 1;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        int vi = ai + c_gt.hin_v[type][par][q][0], vj = aj + c_gt.hin_v[type][par][q][1];
        if (vi < 0 || vi > G.N || vj < 0 || vj > G.M) return false;
        const double *p = &spos[vi - (i0 - 2)][vj - (j0 - 2)][0];
        h.pt[q] = mk(p[0], p[1], p[2]);
    }
    h.n1 = face_normal(h.pt[0], h.pt[1], h.pt[2]);              // owner face, a rotation of its f2v order
    h.n2 = face_normal(h.pt[3], h.pt[2], h.pt[1]);              // neighbour face (consistent orientation)
    double th = hinge_theta_abs(h.n1, h.n2);
    const int l = c_gt.hin_owner[type][par][3];
    // Cloth.compute_angle's side test: n2 . (f.p[(l+1) % 2] - f.p[l]) with the reference's `% 2` (Q3); f.p[l] = pt0,
    // f.p[(l+1) % 2] = pt1 for l = 0 and pt2 for l = 1, 2
    d3 e = (l == 0 ? h.pt[1] : h.pt[2]) - h.pt[0];
    if (dot(h.n2, e) < 0) th = -th;
    h.th = th;
    int face = ((ai + c_gt.hin_owner[type][par][0]) * G.M + aj + c_gt.hin_owner[type][par][1]) * 2 + c_gt.hin_owner[type][par][2];
    *face_l = 3 * face + l;
    return true;
}

// F[cloth rows] = vertex term + membrane + bending gradient; plain stores (one thread per vertex gathers its elements)
__global__ void __launch_bounds__(128) k_residual_rows(ClothGrid64 G, const double *__restrict__ pos, const double *__restrict__ prev_pos,
                                                       const double *__restrict__ vel, const double *__restrict__ vgrav,
                                                       const double *__restrict__ ref_angle, double *__restrict__ F)
{
    __shared__ double spos[TSL_PH][TSL_PW][3];
    TSL_DYN_SMEM(double, shin_raw);                              // [3][TSL_HH][TSL_HW][12]: dE/dtheta * grad(theta) of pt0..pt3
    double (*shin)[TSL_HH][TSL_HW][12] = (double (*)[TSL_HH][TSL_HW][12])shin_raw;
    const int W = G.M + 1;
    const int i0 = blockIdx.y * TSL_TI, j0 = blockIdx.x * TSL_TJ;
    const int tid = threadIdx.x;
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
    ClothParams P;
    P.dx = G.dx; P.dt = G.dt; P.mass = G.mass; P.Kl = G.Kl; P.Ka = G.Ka; P.Kb = G.Kb; P.k_angle = 0;
    for (int t = tid; t < 3 * TSL_HH * TSL_HW; t += blockDim.x) {
        int type = t / (TSL_HH * TSL_HW), rem = t - type * (TSL_HH * TSL_HW);
        int r = rem / TSL_HW, c = rem - r * TSL_HW;
        Hinge64 h;
        int fl = 0;
        double *out = &shin[type][r][c][0];
        if (load_hinge64(G, spos, i0, j0, type, i0 - 1 + r, j0 - 1 + c, h, &fl)) {
            d3 g[4];
            hinge_grad(h.pt[0], h.pt[1], h.pt[2], h.pt[3], h.n1, h.n2, g[0], g[1], g[2], g[3]);
            double dth = 2.0 * G.Kb * (h.th - ref_angle[fl]) * G.dx * G.dx * 1.0 / 3.0;
#pragma unroll
            for (int q = 0; q < 4; q++) { out[3 * q] = dth * g[q].x; out[3 * q + 1] = dth * g[q].y; out[3 * q + 2] = dth * g[q].z; }
        } else {
#pragma unroll
            for (int q = 0; q < 12; q++) out[q] = 0.0;
        }
    }
    __syncthreads();
    for (int item = tid; item < TSL_TI * TSL_TJ; item += blockDim.x) {
        int ti = item / TSL_TJ, tj = item - ti * TSL_TJ;
        int i = i0 + ti, j = j0 + tj;
        if (i > G.N || j > G.M) continue;
        const int row = G.offset + i * W + j, p = (i + j) & 1;
        const double *xs = &spos[ti + 2][tj + 2][0];
        d3 x = mk(xs[0], xs[1], xs[2]);
        d3 xp = ld3(prev_pos, row), v = ld3(vel, row);
        d3 gv = vgrav ? ld3(vgrav, row) : mk(G.g[0], G.g[1], G.g[2]);
        d3 f = (G.mass / (G.dt * G.dt)) * (x - xp - G.dt * v) - G.mass * gv;
        const int nt = c_gt.n_tri[p][12];
        for (int n = 0; n < nt; n++) {
            GridTables::TriE e = c_gt.tri[p][12][n];
            int qi = i + e.qi, qj = j + e.qj;
            if (qi < 0 || qi >= G.N || qj < 0 || qj >= G.M) continue;
            int qp = (qi + qj) & 1;
            Tri t;
            d3 xv[3];
#pragma unroll
            for (int l = 0; l < 3; l++) {
                const double *q = &spos[qi + c_gt.tri_v[qp][e.t][l][0] - (i0 - 2)][qj + c_gt.tri_v[qp][e.t][l][1] - (j0 - 2)][0];
                t.p[l][0] = q[0]; t.p[l][1] = q[1]; t.p[l][2] = q[2];
                xv[l] = mk(q[0], q[1], q[2]);
            }
            const int a = e.a, an = (a + 1) % 3, ap = (a + 2) % 3;
            // edge a joins (a, a+1): + grad; edge a-1 joins (a-1, a): - grad (compute_residual :658-665)
            f = f + edge_grad(P, xv[a] - xv[an], rest_len(P, a)) - edge_grad(P, xv[ap] - xv[a], rest_len(P, ap));
            double area = tri_area(t), V = rest_area(P);
            double da = -G.Ka * 2.0 * (1.0 - area / V);
            f = f + da * mk(area_dx(2 * area, t.p[a], t.p[an], t.p[ap], 0), area_dx(2 * area, t.p[a], t.p[an], t.p[ap], 1),
                            area_dx(2 * area, t.p[a], t.p[an], t.p[ap], 2));
        }
        const int nh = c_gt.n_hin[p][12];
        for (int n = 0; n < nh; n++) {
            GridTables::HinE e = c_gt.hin[p][12][n];
            const double *g = &shin[e.type][i + e.ei - (i0 - 1)][j + e.ej - (j0 - 1)][3 * e.j];
            f = f + mk(g[0], g[1], g[2]);
        }
        F[3 * (size_t)row] = f.x; F[3 * (size_t)row + 1] = f.y; F[3 * (size_t)row + 2] = f.z;
    }
}

// energy of the cloth: vertex terms of the tile's vertices, the two triangles of every quad anchored in the tile, every hinge
// anchored in the tile (each element has exactly one anchor).  Deterministic: per-CTA partial, the last CTA adds them in order.
__global__ void __launch_bounds__(128) k_energy_rows(ClothGrid64 G, const double *__restrict__ pos, const double *__restrict__ prev_pos,
                                                     const double *__restrict__ vel, const double *__restrict__ vgrav,
                                                     const double *__restrict__ ref_angle, double *partial, unsigned int *ticket, double *out)
{
    __shared__ double spos[TSL_PH][TSL_PW][3];
    __shared__ double sred[4];
    __shared__ bool last;
    const int W = G.M + 1;
    const int i0 = blockIdx.y * TSL_TI, j0 = blockIdx.x * TSL_TJ;
    const int tid = threadIdx.x;
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
    ClothParams P;
    P.dx = G.dx; P.dt = G.dt; P.mass = G.mass; P.Kl = G.Kl; P.Ka = G.Ka; P.Kb = G.Kb; P.k_angle = 0;
    double E = 0;
    for (int item = tid; item < TSL_TI * TSL_TJ; item += blockDim.x) {
        int ti = item / TSL_TJ, tj = item - ti * TSL_TJ;
        int i = i0 + ti, j = j0 + tj;
        if (i > G.N || j > G.M) continue;
        const int row = G.offset + i * W + j;
        const double *xs = &spos[ti + 2][tj + 2][0];
        d3 x = mk(xs[0], xs[1], xs[2]);
        d3 X = x - ld3(prev_pos, row) - G.dt * ld3(vel, row);
        d3 gv = vgrav ? ld3(vgrav, row) : mk(G.g[0], G.g[1], G.g[2]);
        E += -G.mass * dot(x, gv) + 0.5 * G.mass * dot(X, X) / (G.dt * G.dt);
        if (i < G.N && j < G.M) {
            const int qp = (i + j) & 1;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                d3 xv[3];
#pragma unroll
                for (int l = 0; l < 3; l++) {
                    const double *q = &spos[ti + 2 + c_gt.tri_v[qp][t][l][0]][tj + 2 + c_gt.tri_v[qp][t][l][1]][0];
                    xv[l] = mk(q[0], q[1], q[2]);
                }
                double area = 0.5 * norm(cross(xv[1] - xv[0], xv[2] - xv[0])), V = rest_area(P);
                E += G.Ka * (1 - area / V) * (1 - area / V) * V;
#pragma unroll
                for (int l = 0; l < 3; l++) E += edge_energy(P, xv[(l + 1) % 3] - xv[l], rest_len(P, l));
            }
        }
        for (int type = 0; type < 3; type++) {
            Hinge64 h;
            int fl = 0;
            if (load_hinge64(G, spos, i0, j0, type, i, j, h, &fl)) {
                double th = h.th - ref_angle[fl];
                E += G.Kb * th * th * G.dx * G.dx * 1.0 / 3.0;
            }
        }
    }
    // block sum (4 warps) -> deterministic grid reduction
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) E += __shfl_xor_sync(0xffffffffu, E, o);
    if ((tid & 31) == 0) sred[tid >> 5] = E;
    __syncthreads();
    const unsigned nblk = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (tid == 0) {
        double s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += sred[w];
        partial[bid] = s;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == nblk - 1);
    }
    __syncthreads();
    if (last && tid == 0) {
        double s = 0;
        for (unsigned k = 0; k < nblk; k++) s += __ldcg(partial + k);
        *out = s;
        *ticket = 0;
    }
}

}  // namespace tsl
