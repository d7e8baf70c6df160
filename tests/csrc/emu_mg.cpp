// emu_mg.cpp -- the Galerkin kernels of the multigrid setup (thinshelllab_b200/csrc/tsl_mg_kernels.cuh) on the CPU through cuda_emu.h:
// the tiled kernels against the one-thread-per-entry kernel they replace.
#include "cuda_emu.h"
#include <array>
#include <vector>
namespace tsl { inline long long sell_addr(long long pb, int lane, int c) { return (pb - lane) * 9 + (long long)c * 32 + lane; } }
#include "../../thinshelllab_b200/csrc/tsl_mg_kernels.cuh"

using namespace tsl;

// stencil level -> next level, both ways; layouts: row-major (sv 225, se 1) or element-major (sv 1, se nvp)
extern "C" void emu_galerkin_pair(int n0f, int n1f, int elem_major_f, int elem_major_c, const float *val_f_rowmajor, float *out_ref, float *out_tiled)
{
    int nvf = n0f * n1f, nvfp = (nvf + 31) / 32 * 32;
    int n0c = (n0f - 1) / 2 + 1, n1c = (n1f - 1) / 2 + 1, nvc = n0c * n1c, nvcp = (nvc + 31) / 32 * 32;
    long long svf = elem_major_f ? 1 : 225, sef = elem_major_f ? nvfp : 1, svc = elem_major_c ? 1 : 225, sec = elem_major_c ? nvcp : 1;
    std::vector<float> vf((size_t)225 * nvfp, 0.f), c1((size_t)225 * nvcp, -7.f), c2((size_t)225 * nvcp, -7.f);
    for (int v = 0; v < nvf; v++) for (int e = 0; e < 225; e++) vf[(size_t)v * svf + (size_t)e * sef] = val_f_rowmajor[(size_t)v * 225 + e];
    long long nt = 25LL * nvc;
    emu_launch_seq(dim3((unsigned)((nt + 127) / 128)), dim3(128), k_galerkin<false>, (const float *)vf.data(), n0f, n1f, svf, sef, (const int *)nullptr,
                   c1.data(), n0c, n1c, svc, sec);
    dim3 grid((n1c + TSL_TCJ - 1) / TSL_TCJ, (n0c + TSL_TCI - 1) / TSL_TCI);
    emu_launch(grid, dim3(256), k_galerkin_tiled<float>, (const float *)vf.data(), n0f, n1f, svf, sef, c2.data(), (float *)nullptr, n0c, n1c, svc, sec,
               (const float *)nullptr);
    for (int v = 0; v < nvc; v++) for (int e = 0; e < 225; e++) {
        out_ref[(size_t)v * 225 + e] = c1[(size_t)v * svc + (size_t)e * sec];
        out_tiled[(size_t)v * 225 + e] = c2[(size_t)v * svc + (size_t)e * sec];
    }
}

// sliced-ELL fine level (cloth rows [off, off + n0f n1f) of an nv-row matrix given as block CSR) -> level 1, with a frozen mask
extern "C" void emu_galerkin_sell(int off, int n0f, int n1f, int nv, const int *rowptr, const int *colidx, const float *blocks, const int *mask_rel,
                                  float *out_ref, float *out_tiled)
{
    int ns = (nv + 31) / 32;
    std::vector<int> slice_base(ns + 1, 0), diag(nv, -1);
    for (int S = 0; S < ns; S++) {
        int w = 0;
        for (int r = 32 * S; r < std::min(nv, 32 * S + 32); r++) w = std::max(w, rowptr[r + 1] - rowptr[r]);
        slice_base[S + 1] = slice_base[S] + 32 * w;
    }
    int npad = slice_base[ns];
    std::vector<int> colpad(npad, 0);
    std::vector<float> val((size_t)npad * 9, 0.f);
    for (int S = 0; S < ns; S++) {
        int w = (slice_base[S + 1] - slice_base[S]) / 32;
        for (int lane = 0; lane < 32; lane++) {
            int r = 32 * S + lane;
            for (int k = 0; k < w; k++) {
                int pb = slice_base[S] + k * 32 + lane;
                bool real = r < nv && k < rowptr[r + 1] - rowptr[r];
                colpad[pb] = real ? colidx[rowptr[r] + k] : (r < nv ? r : 0);
                if (real) {
                    for (int c = 0; c < 9; c++) val[sell_addr(pb, lane, c)] = blocks[(size_t)(rowptr[r] + k) * 9 + c];
                    if (colpad[pb] == r && diag[r] < 0) diag[r] = pb;
                }
            }
        }
    }
    int nvf = n0f * n1f, nvfp = (nvf + 31) / 32 * 32;
    int n0c = (n0f - 1) / 2 + 1, n1c = (n1f - 1) / 2 + 1, nvc = n0c * n1c, nvcp = (nvc + 31) / 32 * 32;
    std::vector<float> st((size_t)225 * nvfp, 0.f), c1((size_t)225 * nvcp, -7.f), c2((size_t)225 * nvcp, -7.f);
    emu_launch_seq(dim3((nvf + 127) / 128), dim3(128), k_sell_to_stencil, off, nvf, n1f, (const int *)slice_base.data(), (const int *)colpad.data(),
                   (const float *)val.data(), (const int *)diag.data(), st.data(), 225LL, 1LL);
    long long nt = 25LL * nvc;
    emu_launch_seq(dim3((unsigned)((nt + 127) / 128)), dim3(128), k_galerkin<true>, (const float *)st.data(), n0f, n1f, 225LL, 1LL, mask_rel,
                   c1.data(), n0c, n1c, 1LL, (long long)nvcp);
    dim3 grid((n1c + TSL_TCJ - 1) / TSL_TCJ, (n0c + TSL_TCI - 1) / TSL_TCI);
    emu_launch(grid, dim3(256), k_galerkin_sell_tiled<float>, off, n0f, n1f, (const int *)slice_base.data(), (const int *)colpad.data(), (const float *)val.data(),
               (const int *)diag.data(), mask_rel, c2.data(), (float *)nullptr, n0c, n1c, 1LL, (long long)nvcp, (const float *)nullptr);
    for (int v = 0; v < nvc; v++) for (int e = 0; e < 225; e++) {
        out_ref[(size_t)v * 225 + e] = c1[(size_t)v + (size_t)e * nvcp];
        out_tiled[(size_t)v * 225 + e] = c2[(size_t)v + (size_t)e * nvcp];
    }
}
