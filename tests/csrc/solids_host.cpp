// Host build of thinshelllab_b200/csrc/tsl_solids.cuh (the same __host__ __device__ functions the CUDA kernels call),
// exposed with a C ABI so tests/test_solids_host.py can compare them with the oracle without a GPU.
#include "../../thinshelllab_b200/csrc/tsl_solids.cuh"

using namespace tsl;

extern "C" {

// per cell: energy [nc], gradient [nc][4][3], reduced Hessian [nc][81] (projected if project), 16 blocks [nc][4][4][9]
void host_tets(int kind, double mu, double lam, double alpha, int nc, const int *tets, const double *B, const double *W, const double *pos,
               int project, double *energy, double *grad, double *H9out, double *blocks)
{
    TetParams P = { kind, mu, lam, alpha };
    for (int c = 0; c < nc; c++) {
        d3 x[4], g[4];
        for (int q = 0; q < 4; q++) x[q] = ld3(pos, tets[4 * c + q]);
        double F[9], H9[81];
        tet_F(x, B + 9 * c, F);
        energy[c] = tet_energy(P, F, W[c]);
        tet_grad(P, F, B + 9 * c, W[c], g);
        for (int q = 0; q < 4; q++) { grad[12 * c + 3 * q] = g[q].x; grad[12 * c + 3 * q + 1] = g[q].y; grad[12 * c + 3 * q + 2] = g[q].z; }
        tet_H9(P, F, B + 9 * c, W[c], H9);
        if (project) psd_clamp<9>(H9);
        for (int q = 0; q < 81; q++) H9out[81 * c + q] = H9[q];
        for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) tet_block(H9, kind, a, b, blocks + ((size_t)c * 16 + a * 4 + b) * 9);
    }
}
// Elastic.compute_deri per cell: gmu, glam [nc][4][3]
void host_tets_deri(int kind, double mu, double lam, double alpha, int nc, const int *tets, const double *B, const double *W, const double *pos,
                    double *gmu, double *glam)
{
    TetParams P = { kind, mu, lam, alpha };
    for (int c = 0; c < nc; c++) {
        d3 x[4], a[4], b[4];
        for (int q = 0; q < 4; q++) x[q] = ld3(pos, tets[4 * c + q]);
        double F[9];
        tet_F(x, B + 9 * c, F);
        tet_deri(P, F, B + 9 * c, W[c], a, b);
        for (int q = 0; q < 4; q++) {
            gmu[12 * c + 3 * q] = a[q].x; gmu[12 * c + 3 * q + 1] = a[q].y; gmu[12 * c + 3 * q + 2] = a[q].z;
            glam[12 * c + 3 * q] = b[q].x; glam[12 * c + 3 * q + 1] = b[q].y; glam[12 * c + 3 * q + 2] = b[q].z;
        }
    }
}
void host_spd9(double *M, int K) { (void)K; psd_clamp<9>(M); }
void host_spd3(double *M, int K) { (void)K; psd_project_3x3(M); }
// normal part of one constraint over (x0, x1, x2, xv): returns active, G[9], H[81] (projected if spd), blocks [4][4][9]
int host_contact(const double *x, double k_contact, double eps, int spd, double *G, double *H, double *blocks)
{
    d3 x0 = ld3(x, 0), x1 = ld3(x, 1), x2 = ld3(x, 2), xv = ld3(x, 3);
    bool act = contact_normal_full(x1 - x0, x2 - x0, xv - x0, k_contact, eps, G, H);
    if (!act) return 0;
    if (spd) psd_clamp<9>(H);
    for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) contact_block(H, a, b, blocks + (a * 4 + b) * 9);
    return 1;
}

}
