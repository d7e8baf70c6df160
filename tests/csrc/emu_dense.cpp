// emu_dense.cpp -- runs the dense-LU kernels of libtsl (thinshelllab_b200/csrc/tsl_dense_kernels.cuh) on the CPU through cuda_emu.h.
// Built by tests/test_emu_cpu.py with g++; same launch sequence as tsl::dense_factor / k_lu_solve in tsl_dense.cu.
#include "cuda_emu.h"
#include <array>
#include "../../thinshelllab_b200/csrc/tsl_dense_kernels.cuh"

using namespace tsl;

extern "C" int emu_lu_solve(int n, const double *A_colmajor, const double *b, double *x, int threads_panel)
{
    int lda = (n + 31) / 32 * 32;
    std::vector<double> A((size_t)lda * n, 0.0);
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) A[(size_t)j * lda + i] = A_colmajor[(size_t)j * n + i];
    std::vector<int> ipiv(n, 0);
    int info = 0;
    for (int k0 = 0; k0 < n; k0 += LU_NB) {
        int nb = std::min(LU_NB, n - k0);
        emu_launch(dim3(1), dim3(threads_panel), k_lu_panel, A.data(), lda, n, k0, nb, ipiv.data(), &info);
        emu_launch_seq(dim3((n + 127) / 128), dim3(128), k_lu_swap, A.data(), lda, n, k0, nb, (const int *)ipiv.data());
        int rest = n - k0 - nb;
        if (rest > 0) {
            emu_launch(dim3((rest + 127) / 128), dim3(128), k_lu_trsm, A.data(), lda, n, k0, nb);
            emu_launch(dim3((rest + 63) / 64, (rest + 63) / 64), dim3(256), k_lu_gemm, A.data(), lda, n, k0, nb);
        }
    }
    for (int i = 0; i < n; i++) x[i] = b[i];
    emu_launch(dim3(1), dim3(threads_panel), k_lu_solve, (const double *)A.data(), lda, n, (const int *)ipiv.data(), x);
    return info;
}
