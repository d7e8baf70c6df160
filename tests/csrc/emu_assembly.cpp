// emu_assembly.cpp -- runs the owner-computes cloth assembly kernels (thinshelllab_b200/csrc/tsl_assembly_kernels.cuh) on the CPU
// through cuda_emu.h.  The sliced-ELL pattern is built here exactly as tsl_finalize does (ascending columns, slices of 32 rows).
#include "cuda_emu.h"
#include <algorithm>
#include <array>
#include "../../thinshelllab_b200/csrc/tsl_assembly_kernels.cuh"

using namespace tsl;

struct Sell { std::vector<int> rowptr, colidx, slice_base, colpad, diag; int nnzb_pad; };

static Sell build_sell(int nv, const std::vector<std::pair<int, int>> &pairs)
{
    Sell A;
    std::vector<std::vector<int>> adj(nv);
    for (auto &p : pairs) adj[p.first].push_back(p.second);
    A.rowptr.assign(nv + 1, 0);
    for (int v = 0; v < nv; v++) {
        adj[v].push_back(v);
        std::sort(adj[v].begin(), adj[v].end());
        adj[v].erase(std::unique(adj[v].begin(), adj[v].end()), adj[v].end());
        A.colidx.insert(A.colidx.end(), adj[v].begin(), adj[v].end());
        A.rowptr[v + 1] = (int)A.colidx.size();
    }
    int ns = (nv + 31) / 32;
    A.slice_base.assign(ns + 1, 0);
    for (int S = 0; S < ns; S++) {
        int w = 0;
        for (int r = 32 * S; r < std::min(nv, 32 * S + 32); r++) w = std::max(w, A.rowptr[r + 1] - A.rowptr[r]);
        A.slice_base[S + 1] = A.slice_base[S] + 32 * w;
    }
    A.nnzb_pad = A.slice_base[ns];
    A.colpad.assign(A.nnzb_pad, 0);
    A.diag.assign(nv, -1);
    for (int S = 0; S < ns; S++) {
        int w = (A.slice_base[S + 1] - A.slice_base[S]) / 32;
        for (int lane = 0; lane < 32; lane++) {
            int r = 32 * S + lane;
            for (int k = 0; k < w; k++) {
                int pb = A.slice_base[S] + k * 32 + lane;
                int col = (r < nv) ? r : 0;
                bool real = r < nv && k < A.rowptr[r + 1] - A.rowptr[r];
                if (real) col = A.colidx[A.rowptr[r] + k];
                A.colpad[pb] = col;
                if (real && col == r && A.diag[r] < 0) A.diag[r] = pb;
            }
        }
    }
    return A;
}

// out_e / out_c: dense [3 nv][3 nv] row-major; cloth occupies rows [offset, offset + NV); n_extra rows follow (diag only in the pattern)
extern "C" int emu_hessian_rows(int N, int M, int offset, int n_extra, const double *pos, const int *frozen, double Kl, double Ka, double Kb,
                                double dx, double mass_dt2, double *out_e, double *out_c)
{
    GridTables T;
    if (!build_grid_tables(T)) return -1;
    c_gt = T;
    int NV = (N + 1) * (M + 1), nv = offset + NV + n_extra;
    std::vector<int> f2v, cf, cp;
    build_cloth_mesh(N, M, f2v, cf, cp);
    std::vector<std::pair<int, int>> pairs;
    for (int f = 0; f < 2 * N * M; f++) {
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) pairs.push_back({ offset + f2v[3 * f + a], offset + f2v[3 * f + b] });
        for (int l = 0; l < 3; l++) {
            int f2 = cf[3 * f + l];
            if (f2 > f) {
                int h[4] = { f2v[3 * f + l], f2v[3 * f + (l + 1) % 3], f2v[3 * f + (l + 2) % 3], f2v[3 * f2 + cp[3 * f + l]] };
                for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) pairs.push_back({ offset + h[a], offset + h[b] });
            }
        }
    }
    Sell A = build_sell(nv, pairs);
    std::vector<float> ve((size_t)A.nnzb_pad * 9, 0.f), vc((size_t)A.nnzb_pad * 9, 0.f);
    ClothGrid G = { N, M, NV, offset, (float)Kl, (float)Ka, (float)Kb, (float)dx, (float)mass_dt2 };
    dim3 grid((M + 1 + TSL_TJ - 1) / TSL_TJ, (N + 1 + TSL_TI - 1) / TSL_TI);
    emu_launch(grid, dim3(256), k_hessian_rows, G, pos, frozen, (const int *)A.slice_base.data(), (const int *)A.colpad.data(), (const int *)A.diag.data(),
               ve.data(), vc.data());
    size_t n3 = 3 * (size_t)nv;
    for (int r = 0; r < nv; r++)
        for (int k = 0; k < A.rowptr[r + 1] - A.rowptr[r]; k++) {
            int pb = A.slice_base[r >> 5] + k * 32 + (r & 31), col = A.colidx[A.rowptr[r] + k];
            long long base = (long long)(pb - (r & 31)) * 9 + (r & 31);
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) {
                    out_e[(3 * (size_t)r + a) * n3 + 3 * col + b] = ve[base + (a * 3 + b) * 32];
                    out_c[(3 * (size_t)r + a) * n3 + 3 * col + b] = vc[base + (a * 3 + b) * 32];
                }
        }
    // padding slots must be untouched (zero)
    int bad = 0;
    for (int S = 0; S < (nv + 31) / 32; S++) {
        int w = (A.slice_base[S + 1] - A.slice_base[S]) / 32;
        for (int lane = 0; lane < 32; lane++) {
            int r = 32 * S + lane;
            for (int k = (r < nv ? A.rowptr[r + 1] - A.rowptr[r] : 0); k < w; k++) {
                int pb = A.slice_base[S] + k * 32 + lane;
                long long base = (long long)(pb - lane) * 9 + lane;
                for (int q = 0; q < 9; q++) bad += ve[base + q * 32] != 0.f || vc[base + q * 32] != 0.f;
            }
        }
    }
    return bad;
}

// cloth residual (rows [offset, offset + NV) of F, [3 nv]) and cloth energy through the fp64 tile kernels
extern "C" double emu_residual_energy_rows(int N, int M, int offset, int nv, const double *pos, const double *prev_pos, const double *vel,
                                           const double *ref_angle, double Kl, double Ka, double Kb, double dx, double dt, double mass,
                                           const double *gravity, double *F)
{
    GridTables T;
    if (!build_grid_tables(T)) return -1;
    c_gt = T;
    ClothGrid64 G;
    G.N = N; G.M = M; G.NV = (N + 1) * (M + 1); G.offset = offset;
    G.Kl = Kl; G.Ka = Ka; G.Kb = Kb; G.dx = dx; G.dt = dt; G.mass = mass;
    for (int k = 0; k < 3; k++) G.g[k] = gravity[k];
    dim3 grid((M + 1 + TSL_TJ - 1) / TSL_TJ, (N + 1 + TSL_TI - 1) / TSL_TI);
    emu_launch(grid, dim3(128), k_residual_rows, G, pos, prev_pos, vel, (const double *)nullptr, ref_angle, F);
    std::vector<double> partial(grid.x * grid.y + 1, 0.0);
    unsigned int ticket = 0;
    double E = 0;
    emu_launch(grid, dim3(128), k_energy_rows, G, pos, prev_pos, vel, (const double *)nullptr, ref_angle, partial.data(), &ticket, &E);
    (void)nv;
    return ticket == 0 ? E : -1e300;
}
