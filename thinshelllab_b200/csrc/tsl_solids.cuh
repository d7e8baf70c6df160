// tsl_solids.cuh -- per-element physics of tetrahedral bodies and of vertex-triangle contact against a moving
// triangle, as __host__ __device__ functions (element-local fp64 data, no memory traffic).
//
// Reference formulas (ThinShellLab, paths relative to code/engine):
//   box model      model_elastic_offset.py   get_force :187-208, compute_energy :315-332, compute_Hessian :95-167
//                  neo-Hookean, P = mu (F - F^-T) + lam log(J) F^-T, J clamped at 0.01
//   tactile model  model_elastic_tactile.py  get_force :158-174, compute_energy :184-201, compute_Hessian :82-124
//                  P = mu F + lam (J - alpha) J F^-T, reduced 9x9 Hessian (vertex 3 eliminated) -> SPD_Projector(9, K=20)
//   projector      linalg.py SPD_Projector :15-148 -> psd_clamp below (own algorithm: cyclic Jacobi, exact eigenvalue clamp)
//   contact        BaseScene.contact_energy :488-543 with contact_diff.det / cross (contact_diff.py:4-129)
// The host build of this header is what tests/test_solids_host.py checks against the oracle.
#pragma once
#include "tsl_elements.cuh"

namespace tsl {

// ------------------------------------------------------------------------------------------------ 3x3 helpers (row-major [9])
TSL_HD void m3mul(const double *a, const double *b, double *o)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
TSL_HD void m3mul_bt(const double *a, const double *b, double *o)   // a b^T
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
}
TSL_HD double m3det(const double *a)
{
    return a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
}
TSL_HD void m3inv(const double *a, double *o)
{
    double id = 1.0 / m3det(a);
    o[0] = (a[4] * a[8] - a[5] * a[7]) * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = (a[5] * a[6] - a[3] * a[8]) * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = (a[3] * a[7] - a[4] * a[6]) * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// ------------------------------------------------------------------------------------------------ PSD clamp
// Positive semi-definite part of a symmetric N x N matrix (row-major M, in place): eigen-decomposition by the cyclic Jacobi method,
// negative eigenvalues set to zero.  This is where the reference calls linalg.SPD_Projector (code/engine/linalg.py:15-148: Householder
// tridiagonalisation + K thresholded QR sweeps, quirk Q8) on the 9 x 9 cell / contact Hessians of the FORWARD matrix.  The projection
// only shapes the Newton path, never the fixed point or the adjoint (which uses the un-projected matrix), so the library computes the
// exact clamp with its own algorithm; the reference's approximate result differs from it by 5e-9 (median) to 1e-3 (sweeps exhausted)
// of the largest entry on the golden matrices (tests/test_solids_host.py).  A matrix that is already PSD keeps its input bits.
template <int N>
TSL_HD void psd_clamp(double *M)
{
    double a[N][N], v[N][N];
    double scale = 0;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            a[i][j] = 0.5 * (M[i * N + j] + M[j * N + i]);
            v[i][j] = (i == j) ? 1.0 : 0.0;
            scale += fabs(a[i][j]);
        }
    if (!(scale > 0)) return;                       // zero matrix (or NaN): nothing to clamp
    for (int sweep = 0; sweep < 40; sweep++) {
        double off = 0;
        for (int p = 0; p < N - 1; p++)
            for (int q = p + 1; q < N; q++) off += fabs(a[p][q]);
        if (off <= 1e-18 * scale) break;
        for (int p = 0; p < N - 1; p++)
            for (int q = p + 1; q < N; q++) {
                double apq = a[p][q];
                if (fabs(apq) <= 1e-300) continue;
                double theta = (a[q][q] - a[p][p]) / (2 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
                double cs = 1 / sqrt(t * t + 1), sn = t * cs;
                for (int k = 0; k < N; k++) {
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = cs * akp - sn * akq; a[k][q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < N; k++) {
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = cs * apk - sn * aqk; a[q][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < N; k++) {
                    double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = cs * vkp - sn * vkq; v[k][q] = sn * vkp + cs * vkq;
                }
            }
    }
    bool negative = false;
    for (int i = 0; i < N; i++) negative = negative || a[i][i] < 0;
    if (!negative) return;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++) {
            double s = 0;
            for (int k = 0; k < N; k++) { double w = a[k][k] > 0 ? a[k][k] : 0; s += w * v[i][k] * v[j][k]; }
            M[i * N + j] = s;
        }
}

// ------------------------------------------------------------------------------------------------ tetrahedra
struct TetParams { int kind; double mu, lam, alpha; };   // kind 0 box (neo-Hookean), 1 tactile

// deformation gradient F = Ds B with Ds columns x_i - x_3 (Elastic.Ds, model_elastic_offset.py:169-171)
TSL_HD void tet_F(const d3 *x, const double *B, double *F)
{
    double D[9];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        d3 e = x[i] - x[3];
        D[i] = e.x; D[3 + i] = e.y; D[6 + i] = e.z;
    }
    m3mul(D, B, F);
}
// strain energy density times rest volume
TSL_HD double tet_energy(const TetParams &t, const double *F, double W)
{
    double J = m3det(F), I = 0;
#pragma unroll
    for (int q = 0; q < 9; q++) I += F[q] * F[q];
    if (t.kind == 0) {
        double lj = log(J > 0.01 ? J : 0.01);
        return W * (t.mu / 2 * (I - 3) - t.mu * lj + t.lam / 2 * lj * lj);
    }
    return W * (t.mu / 2 * (I - 3) + t.lam / 2 * (J - t.alpha) * (J - t.alpha));
}
// gradient of the element energy w.r.t. its 4 vertices: g[i] = dE/dx_i  (= minus the reference's F_f contribution)
TSL_HD void tet_grad(const TetParams &t, const double *F, const double *B, double W, d3 *g)
{
    double Fi[9], P[9], H[9];
    m3inv(F, Fi);
    double J = m3det(F);
    if (t.kind == 0) {
        if (J < 0.01) J = 0.01;
        double lj = log(J);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) P[3 * i + j] = t.mu * (F[3 * i + j] - Fi[3 * j + i]) + t.lam * lj * Fi[3 * j + i];
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) P[3 * i + j] = t.mu * F[3 * i + j] + t.lam * (J - t.alpha) * J * Fi[3 * j + i];
    }
    m3mul_bt(P, B, H);                  // column i of W P B^T = dE/dx_i
    g[3] = mk(0, 0, 0);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        g[i] = mk(W * H[i], W * H[3 + i], W * H[6 + i]);
        g[3] = g[3] - g[i];
    }
}
// d(elastic force)/d(mu) and /d(lam) on the 4 vertices (Elastic.compute_deri, model_elastic_offset.py:423-438 /
// model_elastic_tactile.py:329-347).  The reference splits the stress as P = P1 + P2 with P1 / mu and P2 / lam taken as the
// derivatives; for the tactile model that split is mu (F - J F^-T) + lam (J - 1) J F^-T (its own, not the alpha form of the force).
TSL_HD void tet_deri(const TetParams &t, const double *F, const double *B, double W, d3 *gmu, d3 *glam)
{
    double Fi[9], P1[9], P2[9], H1[9], H2[9];
    m3inv(F, Fi);
    double J = m3det(F);
    if (t.kind == 0) {
        if (J < 0.01) J = 0.01;
        double lj = log(J);
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) { P1[3 * i + j] = t.mu * (F[3 * i + j] - Fi[3 * j + i]); P2[3 * i + j] = t.lam * lj * Fi[3 * j + i]; }
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) { P1[3 * i + j] = t.mu * (F[3 * i + j] - J * Fi[3 * j + i]); P2[3 * i + j] = t.lam * (J - 1) * J * Fi[3 * j + i]; }
    }
    m3mul_bt(P1, B, H1); m3mul_bt(P2, B, H2);
    gmu[3] = mk(0, 0, 0); glam[3] = mk(0, 0, 0);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        gmu[i] = mk(-W * H1[i] / t.mu, -W * H1[3 + i] / t.mu, -W * H1[6 + i] / t.mu);
        glam[i] = mk(-W * H2[i] / t.lam, -W * H2[3 + i] / t.lam, -W * H2[6 + i] / t.lam);
        gmu[3] = gmu[3] - gmu[i]; glam[3] = glam[3] - glam[i];
    }
}
// reduced 9x9 Hessian over (vertex n < 3, dim): H9[(n,dim)][(i,j)] = d2E / dx_{n,dim} dx_{i,j}, built from the 9 unit
// perturbations of Ds exactly as the reference does (dP for dF = e_dim e_n^T B)
TSL_HD void tet_H9(const TetParams &t, const double *F, const double *B, double W, double *H9)
{
    double Fi[9], FiT[9];
    m3inv(F, Fi);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) FiT[3 * i + j] = Fi[3 * j + i];
    double J = m3det(F);
    if (t.kind == 0 && J < 0.01) J = 0.01;
    double lj = t.kind == 0 ? log(J) : 0.0;
    for (int n = 0; n < 3; n++)
        for (int dim = 0; dim < 3; dim++) {
            double dF[9], dFT[9], tmp[9], tmp2[9], dP[9], dH[9];
#pragma unroll
            for (int q = 0; q < 9; q++) dF[q] = 0;
#pragma unroll
            for (int j = 0; j < 3; j++) dF[3 * dim + j] = B[3 * n + j];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) dFT[3 * i + j] = dF[3 * j + i];
            m3mul(Fi, dF, tmp);
            double dTr = tmp[0] + tmp[4] + tmp[8];
            m3mul(FiT, dFT, tmp); m3mul(tmp, FiT, tmp2);            // F^-T dF^T F^-T
            if (t.kind == 0) {
#pragma unroll
                for (int q = 0; q < 9; q++) dP[q] = t.mu * dF[q] + (t.mu - t.lam * lj) * tmp2[q] + t.lam * dTr * FiT[q];
            } else {
#pragma unroll
                for (int q = 0; q < 9; q++)
                    dP[q] = t.mu * dF[q] + t.lam * 2 * J * J * dTr * FiT[q] - t.lam * t.alpha * J * dTr * FiT[q]
                            - t.lam * (J - t.alpha) * J * tmp2[q];
            }
            m3mul_bt(dP, B, dH);
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int j = 0; j < 3; j++) H9[(n * 3 + dim) * 9 + i * 3 + j] = W * dH[3 * j + i];
        }
}
// 3x3 block (row vertex a, column vertex b) of the 12x12 element matrix expanded from H9 (vertex 3 = minus the sum).
//   tactile (model_elastic_tactile.py:114-124): rows are the first index of H9, M[a,j][b,j2] = H9[(a,j)][(b,j2)]
//   box     (model_elastic_offset.py:148-167): rows are the force index, M[a,r][b,dim] = H9[(b,dim)][(a,r)]
// (identical for a symmetric H9; the un-projected reference matrices are symmetric only up to rounding)
TSL_HD void tet_block(const double *H9, int kind, int a, int b, double *Bk)
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int s = 0; s < 3; s++) {
            double v = 0;
            for (int p = (a < 3 ? a : 0); p < (a < 3 ? a + 1 : 3); p++)
                for (int q = (b < 3 ? b : 0); q < (b < 3 ? b + 1 : 3); q++)
                    v += kind == 1 ? H9[(p * 3 + r) * 9 + q * 3 + s] : H9[(q * 3 + s) * 9 + p * 3 + r];
            Bk[3 * r + s] = ((a < 3) == (b < 3)) ? v : -v;
        }
}

// ------------------------------------------------------------------------------------------------ contact, normal part
// d = det(p1, p2, p) / |p1 x p2| over q = (p1, p2, p) (9 unknowns); E = k/2 (d - eps)^2 when d < eps.
// Returns false when inactive.  G[9] = dE/dq, H[81] = d2E/dq2 before projection (quotient rule, BaseScene.py:503-521).
TSL_HD bool contact_normal_full(d3 p1, d3 p2, d3 p, double k_contact, double eps, double *G, double *H)
{
    d3 cr = cross(p1, p2);
    double c = norm(cr);
    double det = dot(cr, p);
    if (!(det / c < eps)) return false;
    d3 n = (1.0 / c) * cr;
    // gradient of det and of c
    d3 gd[3] = { cross(p2, p), cross(p, p1), cr };
    double dG[9] = { gd[0].x, gd[0].y, gd[0].z, gd[1].x, gd[1].y, gd[1].z, gd[2].x, gd[2].y, gd[2].z };
    d3 Jc[6];                                   // d(p1 x p2)/dq_i, i < 6
    for (int i = 0; i < 3; i++) {
        d3 e = mk(i == 0 ? 1.0 : 0.0, i == 1 ? 1.0 : 0.0, i == 2 ? 1.0 : 0.0);
        Jc[i] = cross(e, p2); Jc[3 + i] = cross(p1, e);
    }
    double cG[9];
    for (int i = 0; i < 6; i++) cG[i] = dot(Jc[i], n);
    cG[6] = cG[7] = cG[8] = 0;
    double a[3] = { p1.x, p1.y, p1.z }, b[3] = { p2.x, p2.y, p2.z }, cc[3] = { p.x, p.y, p.z };
    double nn[3] = { n.x, n.y, n.z };
    double d = det / c, pe = k_contact * (d - eps);
    double Gd[9];
    for (int j = 0; j < 9; j++) Gd[j] = dG[j] / c - det * cG[j] / (c * c);
    for (int j = 0; j < 9; j++)
        for (int k = 0; k < 9; k++) {
            // Hessian of det: entries (block I, comp i) x (block J, comp j) = eps_{ijk} * third vector's comp k, signed
            double dH = 0;
            int bj = j / 3, ij = j % 3, bk = k / 3, ik = k % 3;
            if (bj != bk && ij != ik) {
                int kk = 3 - ij - ik;                                // remaining component
                const double *third = (bj + bk == 1) ? cc : ((bj + bk == 3) ? a : b);
                // sign: +1 if (block bj -> bk, comp ij -> ik) follow the same cyclic sense, else -1
                int sb = ((bk - bj + 3) % 3 == 1) ? 1 : -1;
                int sc = ((ik - ij + 3) % 3 == 1) ? 1 : -1;
                dH = (sb * sc) * third[kk];
            }
            // Hessian of c = |p1 x p2|
            double cH = 0;
            if (j < 6 && k < 6) {
                cH = (dot(Jc[j], Jc[k]) - cG[j] * cG[k]) / c;
                if ((j < 3) != (k < 3)) {
                    int ia = j < 3 ? j : k, ib = (j < 3 ? k : j) - 3;   // d2(a x b)/da_ia db_ib = e_ia x e_ib
                    if (ia != ib) {
                        int kk = 3 - ia - ib;
                        cH += (((ib - ia + 3) % 3 == 1) ? 1.0 : -1.0) * nn[kk];
                    }
                }
            }
            double Hd = dH / c - dG[j] * cG[k] / (c * c) - dG[k] * cG[j] / (c * c) - det * cH / (c * c) + 2 * det * cG[j] * cG[k] / (c * c * c);
            H[j * 9 + k] = k_contact * Gd[j] * Gd[k] + pe * Hd;
        }
    for (int j = 0; j < 9; j++) G[j] = pe * Gd[j];
    return true;
}
// 3x3 block (row vertex a, col vertex b; vertex order f0, f1, f2, v) of the 12x12 expansion of the 9x9 over
// (p1, p2, p) = (x_f1 - x_f0, x_f2 - x_f0, x_v - x_f0): f0 takes minus the sums (BaseScene.py:526-541)
TSL_HD void contact_block(const double *H, int a, int b, double *Bk)
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int s = 0; s < 3; s++) {
            double v = 0;
            for (int p = (a > 0 ? a - 1 : 0); p < (a > 0 ? a : 3); p++)
                for (int q = (b > 0 ? b - 1 : 0); q < (b > 0 ? b : 3); q++) v += H[(p * 3 + r) * 9 + q * 3 + s];
            Bk[3 * r + s] = ((a > 0) == (b > 0)) ? v : -v;
        }
}

}  // namespace tsl
