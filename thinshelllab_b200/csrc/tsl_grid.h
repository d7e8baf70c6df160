// tsl_grid.h -- host-side structure of the cloth grid: the reference's mesher and the tables the owner-computes assembly kernels
// (tsl_assembly_kernels.cuh) read from constant memory.  Host C++ only (libtsl and the CPU emulation tests include it).
//
// Cloth.init_mesh (code/engine/model_fold_offset.py:929-1018) triangulates an (N+1) x (M+1) vertex grid with alternating diagonals:
// the structure around a vertex / an edge depends only on the parity of (i + j).  Instead of hand-deriving the neighbourhoods, the
// tables are READ OFF a small mesh built by the mesher itself (build_grid_tables): for a vertex of either parity, which triangles
// and hinges contribute to each block (v, v + (di, dj)) of its matrix row, and with which local vertex indices.  Elements that fall
// outside the grid simply do not exist; the kernels test that at run time, so boundary rows need no special tables.
#pragma once
#include <cstdlib>
#include <cstring>
#include <vector>

namespace tsl {

// Cloth.init_mesh: alternating-diagonal grid, neighbour tables with the reference's wiring (including the entries it never writes,
// which stay 0 -- quirk Q2).  f2v / cf / cp: [2NM][3].
inline void build_cloth_mesh(int N, int M, std::vector<int> &f2v, std::vector<int> &cf, std::vector<int> &cp)
{
    int NF = 2 * N * M;
    f2v.assign(3 * (size_t)NF, 0); cf.assign(3 * (size_t)NF, 0); cp.assign(3 * (size_t)NF, 0);
    auto F = [&](int f, int l) -> int & { return f2v[3 * (size_t)f + l]; };
    auto CF = [&](int f, int l) -> int & { return cf[3 * (size_t)f + l]; };
    auto CP = [&](int f, int l) -> int & { return cp[3 * (size_t)f + l]; };
    for (int i = 0; i < N; i++)
        for (int j = 0; j < M; j++) {
            int k = (i * M + j) * 2;
            int a = i * (M + 1) + j, b = a + 1, c = a + M + 2, d = a + M + 1;
            int up = ((i - 1) * M + j) * 2 + 1, down = ((i + 1) * M + j) * 2;
            bool even = ((i + j) % 2 == 0);
            if (even) { F(k, 0) = c; F(k, 1) = b; F(k, 2) = a; F(k + 1, 0) = a; F(k + 1, 1) = d; F(k + 1, 2) = c; }
            else { F(k, 0) = b; F(k, 1) = a; F(k, 2) = d; F(k + 1, 0) = d; F(k + 1, 1) = c; F(k + 1, 2) = b; }
            // (face, slot) <- (neighbour, opposite slot) in the reference's assignment order
            struct W { int f, l, nb, op; bool ok; };
            W even_w[4] = { { k, 0, up, 2, i > 0 }, { k, 2, k + 2, 0, j < M - 1 }, { k + 1, 0, down, 2, i < N - 1 }, { k + 1, 2, k - 2, 0, j > 0 } };
            W odd_w[4] = { { k, 2, up, 0, i > 0 }, { k + 1, 0, k + 3, 2, j < M - 1 }, { k + 1, 2, down, 0, i < N - 1 }, { k, 2, k - 2, 2, j > 0 } };
            const W *ws = even ? even_w : odd_w;
            for (int q = 0; q < 4; q++) {
                const W &w = ws[q];
                if (w.ok) { CF(w.f, w.l) = w.nb; CP(w.f, w.l) = w.op; }
                else CF(w.f, w.l) = -1;
            }
            CF(k, 1) = k + 1; CP(k, 1) = 1; CF(k + 1, 1) = k; CP(k + 1, 1) = 1;
        }
}

// edge types of the grid: horizontal (i,j)-(i,j+1), vertical (i,j)-(i+1,j), diagonal of quad (i,j)
enum { TSL_EDGE_H = 0, TSL_EDGE_V = 1, TSL_EDGE_D = 2 };

#define TSL_GT_MAX_TRI 8
#define TSL_GT_MAX_HIN 16
struct GridTables {
    // triangle t of a quad with parity q: grid offsets (dI, dJ) of its 3 vertices from the quad anchor (i, j), in f2v order
    signed char tri_v[2][2][3][2];
    // hinge on the edge (type, parity of its anchor): offsets of the hinge vertices pt0..pt3 (opposite in the owner face, shared
    // edge start / end, opposite in the neighbour face -- the order of k_hessian_hinge) from the edge anchor
    signed char hin_v[3][2][4][2];
    // owner face of that hinge (the (face i, slot l) with counter_face[i][l] > i): quad offset from the edge anchor, triangle in the
    // quad, local slot l -- ref_angle[i][l] and the side test of compute_angle (Q3) are indexed by them
    signed char hin_owner[3][2][4];
    // gather lists of the block (v, v + (di, dj)), slot = (di + 2) * 5 + (dj + 2), for a vertex v of parity p
    unsigned char n_tri[2][25], n_hin[2][25];
    struct TriE { signed char qi, qj; unsigned char t, a, b; } tri[2][25][TSL_GT_MAX_TRI];          // quad anchor = v + (qi, qj); block (local a, local b)
    struct HinE { signed char ei, ej; unsigned char type, j, k; } hin[2][25][TSL_GT_MAX_HIN];       // edge anchor = v + (ei, ej); block (pt j, pt k)
};

// classifies the edge between grid vertices u1 < u2 (ids in an (M+1)-wide grid): type and anchor vertex id; false if not a mesh edge
inline bool grid_edge_key(int u1, int u2, int M, int *type, int *anchor)
{
    if (u1 > u2) { int t = u1; u1 = u2; u2 = t; }
    int d = u2 - u1;
    if (d == 1) { *type = TSL_EDGE_H; *anchor = u1; return true; }
    if (d == M + 1) { *type = TSL_EDGE_V; *anchor = u1; return true; }
    if (d == M + 2) { *type = TSL_EDGE_D; *anchor = u1; return true; }          // even quad: a - c
    if (d == M) { *type = TSL_EDGE_D; *anchor = u1 - 1; return true; }          // odd quad: b - d, anchor a = b - 1
    return false;
}

// Reads the tables off an 8 x 8 mesh.  Returns false if the mesher's structure is not parity-periodic (it is; the check guards edits).
inline bool build_grid_tables(GridTables &T)
{
    const int N = 8, M = 8, W = M + 1;
    std::vector<int> f2v, cf, cp;
    build_cloth_mesh(N, M, f2v, cf, cp);
    const int NF = 2 * N * M;
    memset(&T, 0, sizeof(T));
    bool ok = true;
    // ---- triangles
    bool tri_set[2][2] = { { false, false }, { false, false } };
    for (int i = 0; i < N; i++)
        for (int j = 0; j < M; j++)
            for (int t = 0; t < 2; t++) {
                int q = (i + j) & 1, f = (i * M + j) * 2 + t;
                for (int l = 0; l < 3; l++) {
                    int v = f2v[3 * f + l], di = v / W - i, dj = v % W - j;
                    if (!tri_set[q][t]) { T.tri_v[q][t][l][0] = (signed char)di; T.tri_v[q][t][l][1] = (signed char)dj; }
                    else ok = ok && T.tri_v[q][t][l][0] == di && T.tri_v[q][t][l][1] == dj;
                }
                tri_set[q][t] = true;
            }
    // ---- hinges: (face i, slot l) with cf[i][l] > i (model_fold_offset.py:217 and everywhere else)
    struct Hinge { int pt[4]; int type, anchor; };
    std::vector<Hinge> hinges;
    bool hin_set[3][2] = { { false, false }, { false, false }, { false, false } };
    for (int i = 0; i < NF; i++)
        for (int l = 0; l < 3; l++) {
            int i2 = cf[3 * i + l];
            if (i2 <= i) continue;
            Hinge h;
            h.pt[0] = f2v[3 * i + l]; h.pt[1] = f2v[3 * i + (l + 1) % 3]; h.pt[2] = f2v[3 * i + (l + 2) % 3]; h.pt[3] = f2v[3 * i2 + cp[3 * i + l]];
            if (!grid_edge_key(h.pt[1], h.pt[2], M, &h.type, &h.anchor)) { ok = false; continue; }
            hinges.push_back(h);
            int ai = h.anchor / W, aj = h.anchor % W, par = (ai + aj) & 1;
            {
                int own[4] = { (i / 2) / M - ai, (i / 2) % M - aj, i & 1, l };
                for (int q = 0; q < 4; q++) {
                    if (!hin_set[h.type][par]) T.hin_owner[h.type][par][q] = (signed char)own[q];
                    else ok = ok && T.hin_owner[h.type][par][q] == own[q];
                }
            }
            for (int q = 0; q < 4; q++) {
                int di = h.pt[q] / W - ai, dj = h.pt[q] % W - aj;
                if (!hin_set[h.type][par]) { T.hin_v[h.type][par][q][0] = (signed char)di; T.hin_v[h.type][par][q][1] = (signed char)dj; }
                else ok = ok && T.hin_v[h.type][par][q][0] == di && T.hin_v[h.type][par][q][1] == dj;
            }
            hin_set[h.type][par] = true;
        }
    for (int t = 0; t < 3; t++) for (int p = 0; p < 2; p++) ok = ok && hin_set[t][p];
    // ---- gather lists of an interior vertex of each parity
    const int v0s[2][2] = { { 4, 4 }, { 4, 3 } };
    for (int p = 0; p < 2; p++) {
        const int i0 = v0s[p][0], j0 = v0s[p][1], v0 = i0 * W + j0;
        for (int f = 0; f < NF; f++) {
            int a = -1;
            for (int l = 0; l < 3; l++) if (f2v[3 * f + l] == v0) a = l;
            if (a < 0) continue;
            int qi = (f / 2) / M, qj = (f / 2) % M;
            for (int b = 0; b < 3; b++) {
                int u = f2v[3 * f + b], di = u / W - i0, dj = u % W - j0;
                if (di < -2 || di > 2 || dj < -2 || dj > 2) { ok = false; continue; }
                int slot = (di + 2) * 5 + (dj + 2);
                int n = T.n_tri[p][slot];
                if (n >= TSL_GT_MAX_TRI) { ok = false; continue; }
                T.tri[p][slot][n] = { (signed char)(qi - i0), (signed char)(qj - j0), (unsigned char)(f & 1), (unsigned char)a, (unsigned char)b };
                T.n_tri[p][slot] = (unsigned char)(n + 1);
            }
        }
        for (const Hinge &h : hinges) {
            int jl = -1;
            for (int q = 0; q < 4; q++) if (h.pt[q] == v0) jl = q;
            if (jl < 0) continue;
            int ai = h.anchor / W, aj = h.anchor % W;
            for (int k = 0; k < 4; k++) {
                int u = h.pt[k], di = u / W - i0, dj = u % W - j0;
                if (di < -2 || di > 2 || dj < -2 || dj > 2) { ok = false; continue; }
                int slot = (di + 2) * 5 + (dj + 2);
                int n = T.n_hin[p][slot];
                if (n >= TSL_GT_MAX_HIN) { ok = false; continue; }
                T.hin[p][slot][n] = { (signed char)(ai - i0), (signed char)(aj - j0), (unsigned char)h.type, (unsigned char)jl, (unsigned char)k };
                T.n_hin[p][slot] = (unsigned char)(n + 1);
            }
        }
    }
    return ok;
}

}  // namespace tsl
