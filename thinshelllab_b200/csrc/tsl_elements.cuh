// tsl_elements.cuh -- per-element physics of the thin-shell step as __host__ __device__ functions.
//
// Everything here works on element-local data in registers (fp64): a kernel gathers the 3 or 4 vertex
// positions once and evaluates energy / gradient / Hessian blocks without touching memory again.
// The formulas follow the reference's energy definitions (ThinShellLab, code/engine/model_fold_offset.py,
// BaseScene.py, contact_diff.py; cited per function) including the quirks that change numbers
// (Q1-Q3, Q12, Q14, Q15), because the adjoint needs the reference's Hessian.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define TSL_HD __host__ __device__ __forceinline__
#else
#define TSL_HD inline
#endif

namespace tsl {

struct d3 { double x, y, z; };
TSL_HD d3 mk(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
TSL_HD d3 operator+(d3 a, d3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
TSL_HD d3 operator-(d3 a, d3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
TSL_HD d3 operator-(d3 a) { return mk(-a.x, -a.y, -a.z); }
TSL_HD d3 operator*(double s, d3 a) { return mk(s * a.x, s * a.y, s * a.z); }
TSL_HD double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
TSL_HD d3 cross(d3 a, d3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
TSL_HD double norm(d3 a) { return sqrt(dot(a, a)); }
TSL_HD double comp(d3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
TSL_HD d3 ld3(const double *p, int v) { return mk(p[3 * v], p[3 * v + 1], p[3 * v + 2]); }

struct ClothParams {
    double dx, dt, mass, Kl, Ka, Kb, k_angle;
};

// rest data of Cloth.init_pos* (model_fold_offset.py:780-785): l_i = (dx, dx, sqrt2 dx), V = dx^2/2
TSL_HD double rest_len(const ClothParams &c, int l) { return l == 2 ? c.dx * sqrt(2.0) : c.dx; }
TSL_HD double rest_area(const ClothParams &c) { return c.dx * c.dx * 0.5; }

// unit face normal, Cloth.compute_normal_dir (model_fold_offset.py:169-174): (b-a) x (c-b)
TSL_HD d3 face_normal(d3 a, d3 b, d3 c)
{
    d3 n = cross(b - a, c - b);
    double l = norm(n);
    return mk(n.x / l, n.y / l, n.z / l);
}

// dihedral angle magnitude, Cloth.compute_angle (model_fold_offset.py:126-134)
TSL_HD double hinge_theta_abs(d3 n1, d3 n2)
{
    double ct = dot(n1, n2);
    if (ct < 0.999999) return acos(ct);
    return 2 * sqrt(fabs(1.0 - ct)) / sqrt(1 + ct);
}

// gradient of the hinge angle w.r.t. its 4 vertices (Cloth.compute_bending_grad, :379-402), written on
// vertex identities: p0 opposite in face 1, (p1,p2) shared edge, p3 opposite in face 2.
TSL_HD void hinge_grad(d3 p0, d3 p1, d3 p2, d3 p3, d3 n1, d3 n2, d3 &ga, d3 &gb, d3 &gc, d3 &gd)
{
    double A1 = norm(cross(p1 - p0, p2 - p0));      // 2 * area of face 1
    double A2 = norm(cross(p1 - p3, p2 - p3));
    double l12 = norm(p2 - p1);
    double l02 = norm(p0 - p2), l01 = norm(p0 - p1), l32 = norm(p3 - p2), l31 = norm(p3 - p1);
    double h1_p0 = A1 / l12, h1_p1 = A1 / l02, h1_p2 = A1 / l01;
    double h2_p3 = A2 / l12, h2_p1 = A2 / l32, h2_p2 = A2 / l31;
    double c1_p1 = dot(p0 - p1, p2 - p1) / (l01 * l12), c1_p2 = dot(p0 - p2, p1 - p2) / (l02 * l12);
    double c2_p1 = dot(p3 - p1, p2 - p1) / (l31 * l12), c2_p2 = dot(p3 - p2, p1 - p2) / (l32 * l12);
    ga = (-1.0 / h1_p0) * n1;
    gd = (-1.0 / h2_p3) * n2;
    gb = (c1_p2 / h1_p1) * n1 + (c2_p2 / h2_p1) * n2;
    gc = (c1_p1 / h1_p2) * n1 + (c2_p1 / h2_p2) * n2;
}

// ---------------------------------------------------------------------------------------------
// edge spring (model_fold_offset.py:260-266, 288-294, 476-499): 3x3 block of one edge, delta = x_a - x_b.
// The off-diagonal of the "second derivative of l" carries the reference's + sign (Q15).
// H row-major [9].
TSL_HD void edge_hessian(const ClothParams &c, d3 delta, double base, double *H)
{
    double lt = norm(delta);
    double dl = -c.Kl * 2.0 * (1.0 - lt / base);
    double dl2 = c.Kl * 2.0 / base;
    double l3 = lt * lt * lt;
    double d[3] = { delta.x, delta.y, delta.z };
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double second = (j == k) ? (lt * lt - d[j] * d[j]) / l3 : d[j] * d[k] / l3;
            H[j * 3 + k] = dl * second + dl2 * (d[j] / lt) * (d[k] / lt);
        }
}
TSL_HD double edge_energy(const ClothParams &c, d3 delta, double base)
{
    double lt = norm(delta);
    return c.Kl * (1 - lt / base) * (1 - lt / base) * base;
}
// gradient w.r.t. x_a (x_b gets the negative): compute_residual :658-665
TSL_HD d3 edge_grad(const ClothParams &c, d3 delta, double base)
{
    double lt = norm(delta);
    double dl = -c.Kl * 2.0 * (1.0 - lt / base);
    return (dl / lt) * delta;
}

// exact eigenvalue clamp of a symmetric 3x3 (cyclic Jacobi).  Stands in for linalg.SPD_Projector
// (code/engine/linalg.py:15-148, K = 10 sweeps of a thresholded QR): both return the PSD part up to
// ~1e-10 relative on these blocks; the projection only shapes the Newton path, never the fixed point.
TSL_HD void psd_project_3x3(double *H)
{
    double a[3][3] = { { H[0], 0.5 * (H[1] + H[3]), 0.5 * (H[2] + H[6]) },
                       { 0, H[4], 0.5 * (H[5] + H[7]) },
                       { 0, 0, H[8] } };
    a[1][0] = a[0][1]; a[2][0] = a[0][2]; a[2][1] = a[1][2];
    double v[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
    double scale = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]) + fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (!(scale > 0)) return;
    for (int sweep = 0; sweep < 12; sweep++) {
        double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off <= 1e-17 * scale) break;
#pragma unroll
        for (int pq = 0; pq < 3; pq++) {
            const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
            double apq = a[p][q];
            if (fabs(apq) <= 1e-300) continue;
            double theta = (a[q][q] - a[p][p]) / (2 * apq);
            double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
            double cs = 1 / sqrt(t * t + 1), sn = t * cs;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                double akp = a[k][p], akq = a[k][q];
                a[k][p] = cs * akp - sn * akq; a[k][q] = sn * akp + cs * akq;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                double apk = a[p][k], aqk = a[q][k];
                a[p][k] = cs * apk - sn * aqk; a[q][k] = sn * apk + cs * aqk;
            }
#pragma unroll
            for (int k = 0; k < 3; k++) {
                double vkp = v[k][p], vkq = v[k][q];
                v[k][p] = cs * vkp - sn * vkq; v[k][q] = sn * vkp + cs * vkq;
            }
        }
    }
    if (a[0][0] >= 0 && a[1][1] >= 0 && a[2][2] >= 0) return;   // already PSD: keep the input bits
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 3; k++) { double w = a[k][k] > 0 ? a[k][k] : 0; s += w * v[i][k] * v[j][k]; }
            H[i * 3 + j] = s;
        }
}

// exact eigenvalue clamp of a symmetric 2x2; equals linalg.SPD_project_2d (linalg.py:6-12), whose SVD
// sign test (u_k . v_k < 0 <=> lambda_k < 0) zeroes the negative eigenvalues of a symmetric input.
TSL_HD void psd_project_2x2(double *h)
{
    double a = h[0], b = 0.5 * (h[1] + h[2]), d = h[3];
    double tr = a + d, df = a - d;
    double rt = sqrt(df * df + 4 * b * b);
    double l1 = 0.5 * (tr + rt), l2 = 0.5 * (tr - rt);
    if (l2 >= 0) return;
    if (l1 <= 0) { h[0] = h[1] = h[2] = h[3] = 0; return; }
    double vx, vy;
    if (fabs(b) > 0) { vx = l1 - d; vy = b; }
    else if (a >= d) { vx = 1; vy = 0; }
    else { vx = 0; vy = 1; }
    double nn = sqrt(vx * vx + vy * vy);
    vx /= nn; vy /= nn;
    h[0] = l1 * vx * vx; h[1] = h[2] = l1 * vx * vy; h[3] = l1 * vy * vy;
}

// ---------------------------------------------------------------------------------------------
// area term.  The five closed forms of model_fold_offset.py:296-377 written on the 2x2 minors
// m(x,y) = (p2-p1)_x (p3-p1)_y - (p3-p1)_x (p2-p1)_y that they share; area2 = 2*area.
// area_dxy_p12 is NOT the true mixed derivative (Q14) and is kept as the reference has it.
struct Tri {
    double p[3][3];   // p[vertex][dim]
};
TSL_HD double mnr(const double *p1, const double *p2, const double *p3, int x, int y)
{
    return (p2[x] - p1[x]) * (p3[y] - p1[y]) - (p3[x] - p1[x]) * (p2[y] - p1[y]);
}
TSL_HD double area_dx(double area2, const double *p1, const double *p2, const double *p3, int dim)
{   // compute_area_dx :312-325
    const int d1 = (dim == 0) ? 1 : 0, d2 = 3 - d1 - dim;
    double deri = 0.5 * (p1[dim] * ((p2[d1] - p3[d1]) * (p2[d1] - p3[d1]) + (p2[d2] - p3[d2]) * (p2[d2] - p3[d2]))
        - p2[dim] * (p1[d1] * (p2[d1] - p3[d1]) - p2[d1] * p3[d1] + p3[d1] * p3[d1] + p1[d2] * p2[d2] - p1[d2] * p3[d2] - p2[d2] * p3[d2] + p3[d2] * p3[d2])
        + p3[dim] * (p1[d1] * (p2[d1] - p3[d1]) - p2[d1] * p2[d1] + p2[d1] * p3[d1] + (p1[d2] - p2[d2]) * (p2[d2] - p3[d2]))) / area2;
    return deri;
}
TSL_HD double area_dx2(double area2, const double *p1, const double *p2, const double *p3, int dim)
{   // compute_area_dx2 :296-310
    const int d1 = (dim == 0) ? 1 : 0, d2 = 3 - d1 - dim;
    double q = (p2[d1] - p3[d1]) * mnr(p1, p2, p3, dim, d1) + (p2[d2] - p3[d2]) * mnr(p1, p2, p3, dim, d2);
    double deri = ((p2[d1] - p3[d1]) * (p2[d1] - p3[d1]) + (p2[d2] - p3[d2]) * (p2[d2] - p3[d2])) / area2
                - q * q / (area2 * area2 * area2);
    return deri * 0.5;
}
TSL_HD double area_dxy_p1(double area2, const double *p1, const double *p2, const double *p3, int dim, int d1)
{   // compute_area_dxy_p1 :327-341
    const int d2 = 3 - d1 - dim;
    double deri = ((p3[dim] - p2[dim]) * (p2[d1] - p3[d1])) / area2
        - (((p3[dim] - p2[dim]) * mnr(p1, p2, p3, dim, d1) + (p2[d2] - p3[d2]) * mnr(p1, p2, p3, d1, d2))
           * ((p2[d1] - p3[d1]) * mnr(p1, p2, p3, dim, d1) + (p2[d2] - p3[d2]) * mnr(p1, p2, p3, dim, d2))) / (area2 * area2 * area2);
    return deri * 0.5;
}
TSL_HD double area_dx2_p12(double area2, const double *p1, const double *p2, const double *p3, int dim)
{   // compute_area_dx2_p12 :343-361
    const int d1 = (dim == 0) ? 1 : 0, d2 = 3 - d1 - dim;
    double deri = ((p3[d1] - p1[d1]) * (p2[d1] - p3[d1]) + (p3[d2] - p1[d2]) * (p2[d2] - p3[d2])) / area2
        - (((p2[d1] - p3[d1]) * mnr(p1, p2, p3, dim, d1) + (p2[d2] - p3[d2]) * mnr(p1, p2, p3, dim, d2))
           * ((p3[d1] - p1[d1]) * mnr(p1, p2, p3, dim, d1) + (p3[d2] - p1[d2]) * mnr(p1, p2, p3, dim, d2))) / (area2 * area2 * area2);
    return deri * 0.5;
}
TSL_HD double area_dxy_p12(double area2, const double *p1, const double *p2, const double *p3, int dim, int d1)
{   // compute_area_dxy_p12 :363-377 (Q14)
    const int d2 = 3 - d1 - dim;
    double deri = (mnr(p1, p2, p3, dim, d1) + (p1[dim] - p3[dim]) * (p2[d1] - p3[d1])) / area2
        - ((2 * (p1[dim] - p3[dim]) * mnr(p1, p2, p3, dim, d1) + (p3[d2] - p1[d2]) * mnr(p1, p2, p3, d1, d2))
           * ((p2[d1] - p3[d1]) * mnr(p1, p2, p3, dim, d1) + (p2[d2] - p3[d2]) * mnr(p1, p2, p3, dim, d2))) / (area2 * area2 * area2);
    return deri * 0.5;
}
TSL_HD double tri_area(const Tri &t)
{
    d3 a = mk(t.p[0][0], t.p[0][1], t.p[0][2]), b = mk(t.p[1][0], t.p[1][1], t.p[1][2]), c = mk(t.p[2][0], t.p[2][1], t.p[2][2]);
    return 0.5 * norm(cross(b - a, c - a));
}
// one 3x3 block (vertex l, vertex m) of Cloth.compute_Hessian_ma (:526-580); fd[l][j] = area_dx of vertex l
TSL_HD void area_hessian_block(const ClothParams &c, const Tri &t, double area, const double fd[3][3], int l, int m, double *B)
{
    const double V = rest_area(c);
    const double darea2 = c.Ka * 2.0 / V;
    const double da = -c.Ka * 2.0 * (1.0 - area / V);
    const double a2 = 2.0 * area;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double h = fd[l][j] * fd[m][k] * darea2;
            if (j == k) {
                if (l == m) h += da * area_dx2(a2, t.p[l], t.p[(l + 1) % 3], t.p[(l + 2) % 3], j);
                else h += da * area_dx2_p12(a2, t.p[l], t.p[m], t.p[3 - l - m], j);
            } else {
                if (l == m) h += da * area_dxy_p1(a2, t.p[l], t.p[(l + 1) % 3], t.p[(l + 2) % 3], j, k);
                else h += da * area_dxy_p12(a2, t.p[l], t.p[m], t.p[3 - l - m], j, k);
            }
            B[j * 3 + k] = h;
        }
}

// ---------------------------------------------------------------------------------------------
// contact.  signed plane distance d = det[p1,p2,p] / |p1 x p2| (BaseScene.contact_energy :494-513).
struct Contact {
    int idx[4];
    double w[3], k, mu, dx0[3], T[6], n[3];
};
struct ContactParams {
    double k_contact, eps_contact, eps_v, h;
};
// IPC-style friction kernels f0, f1, f2 (BaseScene.py:453-478)
TSL_HD double fr_f0(const ContactParams &C, double x)
{
    double e = C.eps_v * C.h;
    if (x > e) return x;
    return -x / (3.0 * C.eps_v * C.eps_v) * x / (C.h * C.h) * x + x / e * x + e / 3.0;
}
TSL_HD double fr_f1(const ContactParams &C, double x)
{
    double e = C.eps_v * C.h;
    if (x > e) return 1.0 / x;
    return -x / (e * e) + 2.0 / e;
}
TSL_HD double fr_f2(const ContactParams &C, double x)
{
    double e = C.eps_v * C.h;
    if (x > e) return -1.0 / (x * x);
    return -1.0 / (e * e);
}

}  // namespace tsl
