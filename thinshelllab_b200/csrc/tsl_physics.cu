// tsl_physics.cu -- energy / residual / Hessian kernels of the cloth + contact implicit step (sm_100a).
//
// One thread per element (triangle, hinge, contact, vertex); the element's vertex blocks are gathered
// once into registers (fp64), all terms that share them are fused, and results leave through a warp
// shuffle reduction (energy) or red.global.add (force / Hessian blocks).  Reference formulas and quirks:
// see tsl_elements.cuh.  Reference call sites: BaseScene.compute_energy / compute_residual_and_Hessian
// (code/engine/BaseScene.py:427-451, 976-1052).
#include "tsl_internal.cuh"
#include "tsl_kernels.cuh"

namespace tsl {

// ------------------------------------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double sh[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    return v;   // valid in thread 0
}
// deterministic grid reduction: per-block partials, the last block to finish adds them in a fixed order
__device__ __forceinline__ void grid_sum_finish(double block_val, double *partial, unsigned int *ticket, double *out)
{
    __shared__ bool last;
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = block_val;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        double s = 0;
        for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += __ldcg(partial + i);
        s = block_sum(s);
        if (threadIdx.x == 0) { *out = s; *ticket = 0; }
    }
}

// ------------------------------------------------------------------------------------------------ cloth geometry helpers
struct FaceV { int v[3]; d3 p[3]; };
__device__ __forceinline__ FaceV load_face(const ClothDev &c, const double *pos, int i)
{
    FaceV f;
#pragma unroll
    for (int k = 0; k < 3; k++) { f.v[k] = c.f2v[3 * i + k]; f.p[k] = ld3(pos, c.offset + f.v[k]); }
    return f;
}
// signed dihedral angle of hinge (face i1, local l) with neighbour i2 whose unit normal is n2
// (Cloth.compute_angle, model_fold_offset.py:126-138).  `degenerate`: canonical rule D1.
__device__ __forceinline__ double signed_theta(const FaceV &f1, d3 n1, d3 n2, int l, bool degenerate)
{
    double th = hinge_theta_abs(n1, n2);
    if (!degenerate) {
        d3 e = f1.p[(l + 1) % 2] - f1.p[l];      // `% 2` is the reference's (Q3)
        if (dot(n2, e) < 0) th = -th;
    }
    return th;
}
// ------------------------------------------------------------------------------------------------ normals
__global__ void k_face_normals(ClothDev c, const double *__restrict__ pos)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.NF) return;
    FaceV f = load_face(c, pos, i);
    d3 n = face_normal(f.p[0], f.p[1], f.p[2]);
    c.norm_dir[3 * i] = n.x; c.norm_dir[3 * i + 1] = n.y; c.norm_dir[3 * i + 2] = n.z;
}

// ------------------------------------------------------------------------------------------------ energy
// BaseScene.compute_energy: Cloth.compute_energy (model_fold_offset.py:190-218) + per-vertex inertia/gravity of
// every body + contact_energy(diff=False) (BaseScene.py:487-598), fused in one launch.
__global__ void __launch_bounds__(256) k_energy(ClothSet cs, int n_verts, const double *__restrict__ pos,
                                                const double *__restrict__ prev_pos, const double *__restrict__ vel,
                                                const double *__restrict__ mass, d3 g, double dt,
                                                ContactDev con, int nc, ContactParams cp, TetSet ts, const double *__restrict__ vgrav,
                                                int own0, int own1, int nvc, int vskip0, int vskip1, const double *add_in,
                                                double *partial, unsigned int *ticket, double *out)
{
    // strip partition: a cloth vertex / triangle (by its first vertex) / constraint (by its query vertex) is counted by the rank
    // that owns it; [own0, own1) = everything when the context is not partitioned
    double E = 0;
    int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int i = tid; i < n_verts; i += nth) {
        if (i < nvc && (i < own0 || i >= own1)) continue;
        if (i >= vskip0 && i < vskip1) continue;          // cloth vertices already counted by k_energy_rows
        d3 x = ld3(pos, i), xp = ld3(prev_pos, i), v = ld3(vel, i);
        double m = mass[i];
        d3 X = x - xp - dt * v;
        d3 gi = vgrav ? ld3(vgrav, i) : g;
        E += -m * dot(x, gi) + 0.5 * m * dot(X, X) / (dt * dt);
    }
    // Elastic.compute_energy (model_elastic_offset.py:315-332, model_elastic_tactile.py:184-201): strain energy of the cells
    for (int b = 0; b < ts.n; b++) {
        const TetDev &t = ts.b[b];
        for (int c = tid; c < t.nc; c += nth) {
            d3 x[4];
#pragma unroll
            for (int q = 0; q < 4; q++) x[q] = ld3(pos, t.offset + t.tets[4 * c + q]);
            double F[9];
            tet_F(x, t.B + 9 * c, F);
            E += tet_energy(t.P, F, t.W[c]);
        }
    }
    for (int ci = 0; ci < cs.n; ci++) {
        const ClothDev &c = cs.c[ci];
        for (int i = tid; i < c.NF; i += nth) {
            FaceV f = load_face(c, pos, i);
            if (f.v[0] < own0 || f.v[0] >= own1) continue;
            double area = 0.5 * norm(cross(f.p[1] - f.p[0], f.p[2] - f.p[0]));
            double V = rest_area(c.P);
            E += c.P.Ka * (1 - area / V) * (1 - area / V) * V;
            d3 n1 = face_normal(f.p[0], f.p[1], f.p[2]);
#pragma unroll
            for (int l = 0; l < 3; l++) {
                E += edge_energy(c.P, f.p[(l + 1) % 3] - f.p[l], rest_len(c.P, l));
                int i2 = c.cf[3 * i + l];
                if (i2 > i) {
                    FaceV f2 = load_face(c, pos, i2);
                    d3 n2 = face_normal(f2.p[0], f2.p[1], f2.p[2]);
                    double th = signed_theta(f, n1, n2, l, false) - c.ref_angle[3 * i + l];
                    E += c.P.Kb * th * th * c.P.dx * c.P.dx * 1.0 / 3.0;
                }
            }
        }
    }
    for (int i = tid; i < nc; i += nth) {
        const int *idx = con.idx + 4 * i;
        if (idx[3] < nvc && (idx[3] < own0 || idx[3] >= own1)) continue;
        d3 x0 = ld3(pos, idx[0]), x1 = ld3(pos, idx[1]), x2 = ld3(pos, idx[2]), xv = ld3(pos, idx[3]);
        d3 p1 = x1 - x0, p2 = x2 - x0, p = xv - x0;
        d3 cr = cross(p1, p2);
        double d = dot(cr, p) / norm(cr);
        if (d < cp.eps_contact) E += 0.5 * cp.k_contact * (d - cp.eps_contact) * (d - cp.eps_contact);
        const double *w = con.w + 3 * i, *T = con.T + 6 * i, *dx0 = con.dx0 + 3 * i;
        d3 dx = xv - (w[0] * x0 + w[1] * x1 + w[2] * x2) - mk(dx0[0], dx0[1], dx0[2]);
        double u0 = T[0] * dx.x + T[1] * dx.y + T[2] * dx.z, u1 = T[3] * dx.x + T[4] * dx.y + T[5] * dx.z;
        E += con.k[i] * fr_f0(cp, sqrt(u0 * u0 + u1 * u1));
    }
    E = block_sum(E);
    if (add_in && blockIdx.x == 0 && threadIdx.x == 0) E += *add_in;     // the cloth part (k_energy_rows), added once
    grid_sum_finish(E, partial, ticket, out);
}

// ------------------------------------------------------------------------------------------------ residual
__device__ __forceinline__ void red_add3(double *F, int v, d3 g)
{
    atomicAdd(F + 3 * v, g.x); atomicAdd(F + 3 * v + 1, g.y); atomicAdd(F + 3 * v + 2, g.z);
}
// vertex part of the residual for every body: m (x - x_prev - v dt)/dt^2 - m g  (model_fold_offset.py:641-648,
// model_elastic_offset.py:212-214 with zero internal force for the frozen box).  Overwrites F.
__global__ void k_residual_vertex(int n_verts, const double *__restrict__ pos, const double *__restrict__ prev_pos,
                                  const double *__restrict__ vel, const double *__restrict__ mass, d3 g, const double *__restrict__ vgrav,
                                  double dt, double *F, int vskip0, int vskip1)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * n_verts) return;
    int v = i / 3, k = i - 3 * v;
    if (v >= vskip0 && v < vskip1) return;                 // cloth rows are written by k_residual_rows
    double m = mass[v];
    F[i] = m * (pos[i] - prev_pos[i] - vel[i] * dt) / (dt * dt) - m * (vgrav ? vgrav[i] : comp(g, k));
}
// elastic force of the cells of one tetrahedral body (Elastic.get_force / compute_residual,
// model_elastic_offset.py:187-213, model_elastic_tactile.py:158-182); scale as in k_residual_cloth
__global__ void __launch_bounds__(128) k_residual_tets(TetDev t, const double *__restrict__ pos, double *F, double scale)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= t.nc) return;
    d3 x[4], g[4];
    int v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { v[q] = t.offset + t.tets[4 * c + q]; x[q] = ld3(pos, v[q]); }
    double Fm[9];
    tet_F(x, t.B + 9 * c, Fm);
    tet_grad(t.P, Fm, t.B + 9 * c, t.W[c], g);
#pragma unroll
    for (int q = 0; q < 4; q++) red_add3(F, v[q], scale * g[q]);
}
__global__ void k_fill_zero(double *a, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = 0;
}
// membrane (edge + area) and bending gradient of one triangle and the hinges it owns
// (Cloth.compute_residual, model_fold_offset.py:653-687).  mask bit1 edge, bit2 area, bit3 bending.
// scale multiplies every contribution (-1/K gives Cloth.compute_deri's dF/dK).
__global__ void __launch_bounds__(128) k_residual_cloth(ClothDev c, const double *__restrict__ pos, double *F, int mask, double scale)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.NF) return;
    FaceV f = load_face(c, pos, i);
    d3 g[3] = { mk(0, 0, 0), mk(0, 0, 0), mk(0, 0, 0) };
    if (mask & 2) {
#pragma unroll
        for (int l = 0; l < 3; l++) {
            d3 ge = edge_grad(c.P, f.p[l] - f.p[(l + 1) % 3], rest_len(c.P, l));
            g[l] = g[l] + ge; g[(l + 1) % 3] = g[(l + 1) % 3] - ge;
        }
    }
    if (mask & 4) {
        Tri t;
#pragma unroll
        for (int k = 0; k < 3; k++) { t.p[k][0] = f.p[k].x; t.p[k][1] = f.p[k].y; t.p[k][2] = f.p[k].z; }
        double area = tri_area(t), V = rest_area(c.P);
        double da = -c.P.Ka * 2.0 * (1.0 - area / V);
#pragma unroll
        for (int l = 0; l < 3; l++) {
            d3 ga = mk(area_dx(2 * area, t.p[l], t.p[(l + 1) % 3], t.p[(l + 2) % 3], 0),
                       area_dx(2 * area, t.p[l], t.p[(l + 1) % 3], t.p[(l + 2) % 3], 1),
                       area_dx(2 * area, t.p[l], t.p[(l + 1) % 3], t.p[(l + 2) % 3], 2));
            g[l] = g[l] + da * ga;
        }
    }
    if (mask & 8) {
        d3 n1 = face_normal(f.p[0], f.p[1], f.p[2]);
#pragma unroll
        for (int l = 0; l < 3; l++) {
            int i2 = c.cf[3 * i + l];
            if (i2 > i) {
                FaceV f2 = load_face(c, pos, i2);
                d3 n2 = face_normal(f2.p[0], f2.p[1], f2.p[2]);
                int q = c.cp[3 * i + l];
                d3 ga, gb, gc, gd;
                hinge_grad(f.p[l], f.p[(l + 1) % 3], f.p[(l + 2) % 3], f2.p[q], n1, n2, ga, gb, gc, gd);
                double th = signed_theta(f, n1, n2, l, false);
                double dth = 2.0 * c.P.Kb * (th - c.ref_angle[3 * i + l]) * c.P.dx * c.P.dx * 1.0 / 3.0;
                g[l] = g[l] + dth * ga; g[(l + 1) % 3] = g[(l + 1) % 3] + dth * gb; g[(l + 2) % 3] = g[(l + 2) % 3] + dth * gc;
                red_add3(F, c.offset + f2.v[q], (scale * dth) * gd);
            }
        }
    }
#pragma unroll
    for (int l = 0; l < 3; l++) red_add3(F, c.offset + f.v[l], scale * g[l]);
}

// contact normal + friction gradient (BaseScene.contact_energy(diff=True), BaseScene.py:490-541, 548-588);
// the frozen mask of BaseScene.add_F is applied afterwards by k_mask_frozen.
__global__ void k_residual_contact(ContactDev con, int nc, ContactParams cp, const double *__restrict__ pos, double *F)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    const int *idx = con.idx + 4 * i;
    d3 x0 = ld3(pos, idx[0]), x1 = ld3(pos, idx[1]), x2 = ld3(pos, idx[2]), xv = ld3(pos, idx[3]);
    d3 p1 = x1 - x0, p2 = x2 - x0, p = xv - x0;
    d3 cr = cross(p1, p2);
    double c = norm(cr);
    double det = dot(cr, p);
    double d = det / c;
    if (d < cp.eps_contact) {
        double pe = cp.k_contact * (d - cp.eps_contact);
        // d(det)/d(p1,p2,p) = (p2 x p, p x p1, p1 x p2);  d(c)/d(p1,p2) = (p2 x n, n x p1) with n = cr/c ... written out:
        d3 n = (1.0 / c) * cr;
        d3 gd1 = cross(p2, p), gd2 = cross(p, p1), gd3 = cr;
        d3 gc1 = cross(p2, n), gc2 = cross(n, p1);
        d3 g1 = (pe / c) * gd1 - (pe * det / (c * c)) * gc1;
        d3 g2 = (pe / c) * gd2 - (pe * det / (c * c)) * gc2;
        d3 g3 = (pe / c) * gd3;
        red_add3(F, idx[1], g1); red_add3(F, idx[2], g2); red_add3(F, idx[3], g3);
        red_add3(F, idx[0], -(g1 + g2 + g3));
    }
    const double *w = con.w + 3 * i, *T = con.T + 6 * i, *dx0 = con.dx0 + 3 * i;
    d3 dx = xv - (w[0] * x0 + w[1] * x1 + w[2] * x2) - mk(dx0[0], dx0[1], dx0[2]);
    double u0 = T[0] * dx.x + T[1] * dx.y + T[2] * dx.z, u1 = T[3] * dx.x + T[4] * dx.y + T[5] * dx.z;
    double r = sqrt(u0 * u0 + u1 * u1);
    double kf = con.k[i] * fr_f1(cp, r);
    d3 g1 = mk(kf * (u0 * T[0] + u1 * T[3]), kf * (u0 * T[1] + u1 * T[4]), kf * (u0 * T[2] + u1 * T[5]));
    red_add3(F, idx[0], (-w[0]) * g1); red_add3(F, idx[1], (-w[1]) * g1); red_add3(F, idx[2], (-w[2]) * g1);
    red_add3(F, idx[3], g1);
}
__global__ void k_mask_frozen(int n, const int *__restrict__ frozen, double *F)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && frozen[i]) F[i] = 0;
}

// ------------------------------------------------------------------------------------------------ Hessian
template <typename T>
__device__ __forceinline__ void add_block(const Sink<T> &S, int pb, int row, int col, const int *__restrict__ frozen, const double *B)
{
    if (S.zf) {
        // counting pass of the adjoint (BaseScene.add_H :399-405): free row, frozen column
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (frozen[3 * row + a]) continue;
            double za = S.z[3 * row + a];
#pragma unroll
            for (int b = 0; b < 3; b++)
                if (frozen[3 * col + b]) atomicAdd(S.zf + 3 * col + b, -B[a * 3 + b] * za);
        }
        return;
    }
    int lane = row & 31;
    long long base = sell_addr(pb, lane, 0);
#pragma unroll
    for (int a = 0; a < 3; a++) {
        bool fr = frozen[3 * row + a] != 0;
#pragma unroll
        for (int b = 0; b < 3; b++)
            if (!fr && !frozen[3 * col + b]) atomicAdd(S.val + base + (a * 3 + b) * 32, (T)B[a * 3 + b]);
    }
}

// per-face bending preparation (Cloth.prepare_bending, model_fold_offset.py:415-448) kept in registers
struct FaceBend {
    d3 nd[3], en[3];        // mat_M[l] = nd[l] (x) en[l]
    double elen[3];         // |edge_l|  (mat_N = mat_M / elen)
    double angle[3], height[3], ci[3], di[3];
};
__device__ __forceinline__ FaceBend face_bend(const ClothDev &c, const double *pos, int i, const FaceV &f)
{
    FaceBend fb;
    d3 n = ld3(c.norm_dir, i);
    unsigned deg = c.side_deg[i], ovr = c.side_ovr[i];
#pragma unroll
    for (int l = 0; l < 3; l++) {
        d3 p = f.p[l], a = f.p[(l + 1) % 3], b = f.p[(l + 2) % 3];
        d3 edge = b - a;
        int i2 = c.cf[3 * i + l];
        bool neg = false;
        d3 n2 = mk(0, 0, 0);
        if (i2 != -1) {
            n2 = ld3(c.norm_dir, i2);
            if (!((deg >> l) & 1)) neg = dot(n2, f.p[(l + 1) % 2] - f.p[l]) < 0;
            else neg = (ovr >> l) & 1;
        }
        bool judge = !neg;                      // Cloth.judge_angle: True for borders too
        d3 nd = judge ? -n : n;
        d3 en = cross(nd, edge);
        if (dot(en, a - p) > 0) en = -en;
        fb.nd[l] = nd; fb.en[l] = en; fb.elen[l] = norm(edge);
        d3 ap = a - p, bp = b - p;
        double la = norm(ap), lb = norm(bp);
        fb.angle[l] = (ap.x / la) * (bp.x / lb) + (ap.y / la) * (bp.y / lb) + (ap.z / la) * (bp.z / lb);
        fb.height[l] = fabs(dot(p - a, en)) / norm(en);
        if (i2 != -1) {
            double th = hinge_theta_abs(n, n2);
            if (neg) th = -th;
            fb.ci[l] = 2.0 * c.P.Kb * (th - c.ref_angle[3 * i + l]) * c.P.dx * c.P.dx * 1.0 / 3.0;
        } else fb.ci[l] = 0;
    }
#pragma unroll
    for (int l = 0; l < 3; l++)
        fb.di[l] = fb.ci[(l + 1) % 3] * fb.angle[(l + 2) % 3] + fb.ci[(l + 2) % 3] * fb.angle[(l + 1) % 3] - fb.ci[l];
    return fb;
}
// c_i rows and mat_N of faces 0..2, which Hessian loop 1 reads for every face (quirk Q1)
__global__ void k_q1_prepare(ClothDev c, const double *__restrict__ pos)
{
    int i = threadIdx.x;
    if (i >= 3 || i >= c.NF) return;
    FaceV f = load_face(c, pos, i);
    FaceBend fb = face_bend(c, pos, i, f);
    for (int l = 0; l < 3; l++) {
        c.q1[3 * i + l] = fb.ci[l];
        double *N = c.q1 + 9 + (i * 3 + l) * 9;
        double nd[3] = { fb.nd[l].x, fb.nd[l].y, fb.nd[l].z }, en[3] = { fb.en[l].x, fb.en[l].y, fb.en[l].z };
        for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++) N[r * 3 + s] = nd[r] * en[s] / fb.elen[l];
    }
}

// triangle Hessian: edge springs (optionally PSD-projected), area term and bending loop 1, block by block
// (Cloth.compute_Hessian_me :466-524, _ma :526-580, _bending loop 1 :585-614).
template <typename T>
__global__ void __launch_bounds__(128) k_hessian_tri(ClothDev c, const double *__restrict__ pos, const int *__restrict__ frozen,
                                                     Sink<T> S, int spd, int sym)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.NF) return;
    FaceV f = load_face(c, pos, i);
    double He[3][9];
#pragma unroll
    for (int l = 0; l < 3; l++) {
        edge_hessian(c.P, f.p[l] - f.p[(l + 1) % 3], rest_len(c.P, l), He[l]);
        if (spd) psd_project_3x3(He[l]);
    }
    Tri t;
#pragma unroll
    for (int k = 0; k < 3; k++) { t.p[k][0] = f.p[k].x; t.p[k][1] = f.p[k].y; t.p[k][2] = f.p[k].z; }
    double area = tri_area(t);
    double fd[3][3];
#pragma unroll
    for (int l = 0; l < 3; l++)
#pragma unroll
        for (int j = 0; j < 3; j++) fd[l][j] = area_dx(2 * area, t.p[l], t.p[(l + 1) % 3], t.p[(l + 2) % 3], j);
    FaceBend fb = face_bend(c, pos, i, f);
    const double *q1c = c.q1, *q1N = c.q1 + 9;
    const int *slot = c.tri_slot + 9 * i;

    // the six block pairs (l, m): three diagonal, three (l, l+1)
#pragma unroll
    for (int l = 0; l < 3; l++) {
        // ---- diagonal block (l, l)
        {
            double B[9];
            area_hessian_block(c.P, t, area, fd, l, l, B);
            const int lp = (l + 2) % 3;   // edge lp connects (lp, l); edge l connects (l, l+1)
#pragma unroll
            for (int r = 0; r < 9; r++) B[r] += He[l][r] + He[lp][r];
            // loop 1, l == m: s (d_l M_l^T + d_l M_l) - c_i[l][i1] N[l*3+i1] - c_i[l][i2] N[l*3+i2]
            double s = 1.0 / (fb.height[l] * fb.height[l]);
            double nd[3] = { fb.nd[l].x, fb.nd[l].y, fb.nd[l].z }, en[3] = { fb.en[l].x, fb.en[l].y, fb.en[l].z };
            const int i1 = (l + 1) % 3, i2 = (l + 2) % 3;
            double c1 = q1c[3 * l + i1], c2 = q1c[3 * l + i2];
            const double *N1 = q1N + (l * 3 + i1) * 9, *N2 = q1N + (l * 3 + i2) * 9;
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int q = 0; q < 3; q++)
                    B[r * 3 + q] += s * fb.di[l] * (nd[q] * en[r] + nd[r] * en[q]) - c1 * N1[r * 3 + q] - c2 * N2[r * 3 + q];
            if (sym) {
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int q = r + 1; q < 3; q++) { double m_ = 0.5 * (B[r * 3 + q] + B[q * 3 + r]); B[r * 3 + q] = B[q * 3 + r] = m_; }
            }
            add_block(S, slot[l * 3 + l], c.offset + f.v[l], c.offset + f.v[l], frozen, B);
        }
        // ---- off-diagonal pair (l, m) and (m, l), m = l+1
        {
            const int m = (l + 1) % 3;
            double B[9], Bt[9];
            area_hessian_block(c.P, t, area, fd, l, m, B);
            area_hessian_block(c.P, t, area, fd, m, l, Bt);
#pragma unroll
            for (int r = 0; r < 9; r++) { B[r] -= He[l][r]; Bt[r] -= He[l][r]; }
            // loop 1, l != m: H_lm = s (d_l M_m^T + d_m M_l) + c_i[l][i3] N[l*3+i3]; block (l,m) += H_lm, (m,l) += H_lm^T
            double s = 1.0 / (fb.height[l] * fb.height[m]);
            double ndl[3] = { fb.nd[l].x, fb.nd[l].y, fb.nd[l].z }, enl[3] = { fb.en[l].x, fb.en[l].y, fb.en[l].z };
            double ndm[3] = { fb.nd[m].x, fb.nd[m].y, fb.nd[m].z }, enm[3] = { fb.en[m].x, fb.en[m].y, fb.en[m].z };
            const int i3 = 3 - l - m;
            double c3 = q1c[3 * l + i3];
            const double *N3 = q1N + (l * 3 + i3) * 9;
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    double h = s * (fb.di[l] * ndm[q] * enm[r] + fb.di[m] * ndl[r] * enl[q]) + c3 * N3[r * 3 + q];
                    B[r * 3 + q] += h;
                    Bt[q * 3 + r] += h;
                }
            if (sym) {
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int q = 0; q < 3; q++) { double m_ = 0.5 * (B[r * 3 + q] + Bt[q * 3 + r]); B[r * 3 + q] = m_; Bt[q * 3 + r] = m_; }
            }
            add_block(S, slot[l * 3 + m], c.offset + f.v[l], c.offset + f.v[m], frozen, B);
            add_block(S, slot[m * 3 + l], c.offset + f.v[m], c.offset + f.v[l], frozen, Bt);
        }
    }
}


// Forward Newton matrix ("Newton model") -- NOT the reference's Hessian.  Only the fixed point of
// the forward step has to match the reference, and the reference's own matrix (Q14/Q15 + per-edge projection) makes
// Newton crawl whenever the sheet carries compression; this model is the exact membrane Hessian with its indefinite
// pieces optionally clamped:
//   edge : dE/dl / l (I - d d^T) + d2E/dl2 d d^T          clamp: drop dE/dl < 0 (compressed edge)
//   area : d2E/dA2 g g^T + dE/dA * J^T (I - n n^T) J / (2|n|)      clamp: drop dE/dA < 0
//   (bending: Gauss-Newton term only = k_hessian_hinge)
// Symmetric by construction; positive definite when clamped.
template <typename T>
__global__ void __launch_bounds__(128) k_hessian_tri_newton(ClothDev c, const double *__restrict__ pos, const int *__restrict__ frozen,
                                                            Sink<T> S, int clamp)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.NF) return;
    FaceV f = load_face(c, pos, i);
    double He[3][9];
#pragma unroll
    for (int l = 0; l < 3; l++) {
        d3 dv = f.p[l] - f.p[(l + 1) % 3];
        double lt = norm(dv), base = rest_len(c.P, l);
        double dl = -c.P.Kl * 2.0 * (1.0 - lt / base), dl2 = c.P.Kl * 2.0 / base;
        double g = (dl > 0 || !clamp) ? dl / lt : 0.0;
        double d[3] = { dv.x / lt, dv.y / lt, dv.z / lt };
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int k = 0; k < 3; k++) He[l][j * 3 + k] = g * ((j == k ? 1.0 : 0.0) - d[j] * d[k]) + dl2 * d[j] * d[k];
    }
    d3 e1 = f.p[1] - f.p[0], e2 = f.p[2] - f.p[0];
    d3 n = cross(e1, e2);
    double nl = norm(n), area = 0.5 * nl, V = rest_area(c.P);
    double da = -c.P.Ka * 2.0 * (1.0 - area / V), da2 = c.P.Ka * 2.0 / V;
    double sa = (da > 0 || !clamp) ? da / (2.0 * nl) : 0.0;
    d3 nh = mk(n.x / nl, n.y / nl, n.z / nl);
    // J[a][j] = d n / d p_a,j : vertex 1: e_j x e2, vertex 2: e1 x e_j, vertex 0: minus both
    d3 J[3][3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        d3 ej = mk(j == 0 ? 1.0 : 0.0, j == 1 ? 1.0 : 0.0, j == 2 ? 1.0 : 0.0);
        J[1][j] = cross(ej, e2); J[2][j] = cross(e1, ej);
        J[0][j] = -(J[1][j] + J[2][j]);
    }
    double g[3][3];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int j = 0; j < 3; j++) g[a][j] = 0.5 * dot(J[a][j], nh);
    const int *slot = c.tri_slot + 9 * i;
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++) {
            double B[9];
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int k = 0; k < 3; k++) B[j * 3 + k] = da2 * g[a][j] * g[b][k] + sa * (dot(J[a][j], J[b][k]) - 4.0 * g[a][j] * g[b][k]);
            if (a == b) {
                const int lp = (a + 2) % 3;
#pragma unroll
                for (int r = 0; r < 9; r++) B[r] += He[a][r] + He[lp][r];
            } else {
                const int l = ((a + 1) % 3 == b) ? a : b;       // edge l joins (l, l+1)
#pragma unroll
                for (int r = 0; r < 9; r++) B[r] -= He[l][r];
            }
            add_block(S, slot[a * 3 + b], c.offset + f.v[a], c.offset + f.v[b], frozen, B);
        }
}

// hinge Hessian, loop 2 of Cloth.compute_Hessian_bending (:616-637): d2E/dtheta2 * grad(theta) grad(theta)^T
template <typename T>
__global__ void __launch_bounds__(128) k_hessian_hinge(ClothDev c, const double *__restrict__ pos, const int *__restrict__ frozen, Sink<T> S)
{
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= c.NH) return;
    int i = c.hinge_face[h], l = c.hinge_l[h];
    int i2 = c.cf[3 * i + l], q = c.cp[3 * i + l];
    FaceV f = load_face(c, pos, i), f2 = load_face(c, pos, i2);
    d3 n1 = face_normal(f.p[0], f.p[1], f.p[2]), n2 = face_normal(f2.p[0], f2.p[1], f2.p[2]);
    d3 g[4];
    hinge_grad(f.p[l], f.p[(l + 1) % 3], f.p[(l + 2) % 3], f2.p[q], n1, n2, g[0], g[1], g[2], g[3]);
    int pt[4] = { f.v[l], f.v[(l + 1) % 3], f.v[(l + 2) % 3], f2.v[q] };
    double d2 = 2.0 * c.P.Kb * c.P.dx * c.P.dx * 1.0 / 3.0;
    const int *slot = c.hinge_slot + 16 * h;
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            double gj[3] = { g[j].x, g[j].y, g[j].z }, gk[3] = { g[k].x, g[k].y, g[k].z };
            double B[9];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int s = 0; s < 3; s++) B[r * 3 + s] = d2 * gj[r] * gk[s];
            add_block(S, slot[j * 4 + k], c.offset + pt[j], c.offset + pt[k], frozen, B);
        }
}

// mass diagonal m/dt^2 on every DOF, frozen or not (Q6; model_fold_offset.py:468-470, model_elastic_offset.py:97-99)
template <typename T>
__global__ void k_hessian_mass(int n_verts, const double *__restrict__ mass, double dt, const int *__restrict__ diag_pb, T *val, int skip0, int skip1)
{
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_verts || (v >= skip0 && v < skip1)) return;      // [skip0, skip1): rows whose mass the cloth-row kernel already wrote
    long long base = sell_addr(diag_pb[v], v & 31, 0);
    T m = (T)(mass[v] / (dt * dt));
    atomicAdd(val + base + 0 * 32, m); atomicAdd(val + base + 4 * 32, m); atomicAdd(val + base + 8 * 32, m);
}

// contact Hessian restricted to the query vertex (BaseScene.contact_energy(diff=True), BaseScene.py:503-593).
// With every triangle DOF frozen (table, floor) BaseScene.add_H drops all other blocks, and the surviving
// 3x3 is k_contact * n n^T (d is linear in the vertex, so its second derivative vanishes) plus the friction
// block k T^T h T.  A non-frozen triangle DOF raises error bit 0 (TSL_ERR_UNSUPPORTED) instead of being wrong.
template <typename T>
__global__ void k_hessian_contact(ContactDev con, int nc, ContactParams cp, const double *__restrict__ pos,
                                  const int *__restrict__ frozen, const int *__restrict__ diag_pb, Sink<T> S, int spd, int *error_flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    const int *idx = con.idx + 4 * i;
    bool tri_frozen = true;
    for (int q = 0; q < 3; q++) for (int a = 0; a < 3; a++) tri_frozen = tri_frozen && frozen[3 * idx[q] + a];
    if (!tri_frozen) { atomicOr(error_flag, 1); return; }
    d3 x0 = ld3(pos, idx[0]), x1 = ld3(pos, idx[1]), x2 = ld3(pos, idx[2]), xv = ld3(pos, idx[3]);
    d3 cr = cross(x1 - x0, x2 - x0);
    double c = norm(cr);
    double d = dot(cr, xv - x0) / c;
    double B[9];
#pragma unroll
    for (int r = 0; r < 9; r++) B[r] = 0;
    if (d < cp.eps_contact) {
        double n[3] = { cr.x / c, cr.y / c, cr.z / c };
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int s = 0; s < 3; s++) B[r * 3 + s] = cp.k_contact * n[r] * n[s];
    }
    const double *w = con.w + 3 * i, *Tm = con.T + 6 * i, *dx0 = con.dx0 + 3 * i;
    d3 dx = xv - (w[0] * x0 + w[1] * x1 + w[2] * x2) - mk(dx0[0], dx0[1], dx0[2]);
    double u[2] = { Tm[0] * dx.x + Tm[1] * dx.y + Tm[2] * dx.z, Tm[3] * dx.x + Tm[4] * dx.y + Tm[5] * dx.z };
    double r_ = sqrt(u[0] * u[0] + u[1] * u[1]);
    double f1 = fr_f1(cp, r_);
    double h[4] = { f1, 0, 0, f1 };
    if (r_ > 1e-9) {
        double f2 = fr_f2(cp, r_);
        h[0] += f2 * (u[0] / r_) * u[0]; h[1] += f2 * (u[0] / r_) * u[1];
        h[2] += f2 * (u[1] / r_) * u[0]; h[3] += f2 * (u[1] / r_) * u[1];
    }
    if (spd) psd_project_2x2(h);
    double k = con.k[i];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
            B[a * 3 + b] += k * (Tm[a] * (h[0] * Tm[b] + h[1] * Tm[3 + b]) + Tm[3 + a] * (h[2] * Tm[b] + h[3] * Tm[3 + b]));
    add_block(S, diag_pb[idx[3]], idx[3], idx[3], frozen, B);
}

// Elastic.compute_deri (model_elastic_offset.py:423-438, model_elastic_tactile.py:329-347): d_mu, d_lam [n_verts][3], accumulated
__global__ void __launch_bounds__(128) k_tets_deri(TetDev t, const double *__restrict__ pos, double *d_mu, double *d_lam)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= t.nc) return;
    d3 x[4], gm[4], gl[4];
    int v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { v[q] = t.offset + t.tets[4 * c + q]; x[q] = ld3(pos, v[q]); }
    double Fm[9];
    tet_F(x, t.B + 9 * c, Fm);
    tet_deri(t.P, Fm, t.B + 9 * c, t.W[c], gm, gl);
#pragma unroll
    for (int q = 0; q < 4; q++) { red_add3(d_mu, v[q], gm[q]); red_add3(d_lam, v[q], gl[q]); }
}
// sum over the free DOFs of z * d  (Grad.get_parameters_grad, analytic_grad_system.py:69-79), two right-hand sides at once
__global__ void __launch_bounds__(256) k_masked_dot2(int n, const double *__restrict__ z, const int *__restrict__ frozen, const double *__restrict__ a,
                                                     const double *__restrict__ b, double *out2)
{
    double sa = 0, sb = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (!frozen[i]) { sa += z[i] * a[i]; sb += z[i] * b[i]; }
    sa = block_sum(sa);
    sb = block_sum(sb);
    if (threadIdx.x == 0) { out2[0] = sa; out2[1] = sb; }
}

// cell Hessian of a tetrahedral body: reduced 9x9 by the reference's nine unit perturbations, optionally through
// SPD_Projector(9, K=20), expanded to the 16 blocks of the cell (Elastic.compute_Hessian, model_elastic_offset.py:95-167,
// model_elastic_tactile.py:82-124).  One thread per cell; the 9x9 lives in local memory (bodies are a few thousand cells).
template <typename T>
__global__ void __launch_bounds__(64) k_hessian_tets(TetDev t, const double *__restrict__ pos, const int *__restrict__ frozen, Sink<T> S, int project)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= t.nc) return;
    d3 x[4];
    int v[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { v[q] = t.offset + t.tets[4 * c + q]; x[q] = ld3(pos, v[q]); }
    double Fm[9], H9[81];
    tet_F(x, t.B + 9 * c, Fm);
    tet_H9(t.P, Fm, t.B + 9 * c, t.W[c], H9);
    if (project) psd_clamp<9>(H9);
    const int *slot = t.slot + 16 * c;
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) {
            double Bk[9];
            tet_block(H9, t.P.kind, a, b, Bk);
            add_block(S, slot[a * 4 + b], v[a], v[b], frozen, Bk);
        }
}

// contact + friction Hessian of one constraint against a triangle that may move: all 16 blocks over (f0, f1, f2, v)
// (BaseScene.contact_energy(diff=True), BaseScene.py:503-593).  Diagonal blocks go to the sliced-ELL matrix, the 12
// off-diagonal ones to the side buffer side[i][a*3 + (b < a ? b : b - 1)][9] (masked by frozen, plain stores).
template <typename T>
__global__ void __launch_bounds__(64) k_hessian_contact_general(ContactDev con, int nc, ContactParams cp, const double *__restrict__ pos,
                                                                const int *__restrict__ frozen, const int *__restrict__ diag_pb,
                                                                Sink<T> S, T *side, int spd, int newton_model)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    const int *idx = con.idx + 4 * i;
    d3 x0 = ld3(pos, idx[0]), x1 = ld3(pos, idx[1]), x2 = ld3(pos, idx[2]), xv = ld3(pos, idx[3]);
    double G[9], H[81];
    bool active;
    bool tri_frozen = true;
    for (int q = 0; q < 3; q++) for (int a = 0; a < 3; a++) tri_frozen = tri_frozen && frozen[3 * idx[q] + a];
    if (newton_model && tri_frozen && !S.zf) {
        // forward Newton matrix, triangle of a frozen body (most constraints of a sheet lying on the table): only the (v, v) block
        // survives the frozen mask and d is linear in x_v, so the normal part is k n n^T -- no 9x9 needed (as k_hessian_contact)
        d3 cr = cross(x1 - x0, x2 - x0);
        double c = norm(cr);
        active = dot(cr, xv - x0) / c < cp.eps_contact;
#pragma unroll
        for (int q = 0; q < 81; q++) H[q] = 0;
        if (active) {
            double n[3] = { cr.x / c, cr.y / c, cr.z / c };
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int q = 0; q < 3; q++) H[(6 + r) * 9 + 6 + q] = cp.k_contact * n[r] * n[q];
        }
    } else {
        active = contact_normal_full(x1 - x0, x2 - x0, xv - x0, cp.k_contact, cp.eps_contact, G, H);
        if (active && spd) psd_clamp<9>(H);
    }
    const double *w = con.w + 3 * i, *Tm = con.T + 6 * i, *dx0 = con.dx0 + 3 * i;
    d3 dx = xv - (w[0] * x0 + w[1] * x1 + w[2] * x2) - mk(dx0[0], dx0[1], dx0[2]);
    double u[2] = { Tm[0] * dx.x + Tm[1] * dx.y + Tm[2] * dx.z, Tm[3] * dx.x + Tm[4] * dx.y + Tm[5] * dx.z };
    double r_ = sqrt(u[0] * u[0] + u[1] * u[1]);
    double f1 = fr_f1(cp, r_);
    double h[4] = { f1, 0, 0, f1 };
    if (r_ > 1e-9) {
        double f2 = fr_f2(cp, r_);
        h[0] += f2 * (u[0] / r_) * u[0]; h[1] += f2 * (u[0] / r_) * u[1];
        h[2] += f2 * (u[1] / r_) * u[0]; h[3] += f2 * (u[1] / r_) * u[1];
    }
    if (spd) psd_project_2x2(h);
    double k = con.k[i], h1[9];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int b = 0; b < 3; b++)
            h1[a * 3 + b] = k * (Tm[a] * (h[0] * Tm[b] + h[1] * Tm[3 + b]) + Tm[3 + a] * (h[2] * Tm[b] + h[3] * Tm[3 + b]));
    double w1[4] = { -w[0], -w[1], -w[2], 1.0 };
    for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) {
            double Bk[9];
            if (active) contact_block(H, a, b, Bk);
            else {
#pragma unroll
                for (int q = 0; q < 9; q++) Bk[q] = 0;
            }
#pragma unroll
            for (int q = 0; q < 9; q++) Bk[q] += w1[a] * w1[b] * h1[q];
            if (a == b || S.zf) add_block(S, diag_pb[idx[a]], idx[a], idx[b], frozen, Bk);
            else {
                T *dst = side + ((size_t)i * 12 + a * 3 + (b < a ? b : b - 1)) * 9;
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int q = 0; q < 3; q++)
                        dst[r * 3 + q] = (frozen[3 * idx[a] + r] || frozen[3 * idx[b] + q]) ? (T)0 : (T)Bk[r * 3 + q];
            }
        }
}

// ------------------------------------------------------------------------------------------------ misc state kernels
__global__ void k_axpy_pos(int n, const double *__restrict__ x1, const double *__restrict__ p, double alpha, double *pos)
{   // BaseScene.linesearch_step (BaseScene.py:1089-1094)
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos[i] = x1[i] - p[i] * alpha;
}
__global__ void k_axpy2_pos(int n, const double *__restrict__ x1, const double *__restrict__ u, double a, const double *__restrict__ w, double b, double *pos)
{   // pos = x1 - a u - b w  (negative-curvature move of the Newton driver)
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos[i] = x1[i] - a * u[i] - b * w[i];
}
__global__ void k_update_vel(int n, const double *__restrict__ pos, const double *__restrict__ prev_pos, double s, double *vel)
{   // BaseScene.update_vel (BaseScene.py:868-872): s = damping / dt
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vel[i] = (pos[i] - prev_pos[i]) * s;
}
__global__ void k_update_ref_angle(ClothDev c, const double *__restrict__ pos)
{   // Cloth.update_ref_angle (model_fold_offset.py:176-185)
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= c.NH) return;
    int i = c.hinge_face[h], l = c.hinge_l[h];
    int i2 = c.cf[3 * i + l];
    FaceV f = load_face(c, pos, i), f2 = load_face(c, pos, i2);
    d3 n1 = face_normal(f.p[0], f.p[1], f.p[2]), n2 = face_normal(f2.p[0], f2.p[1], f2.p[2]);
    double th = signed_theta(f, n1, n2, l, false);
    double dis = th - c.ref_angle[3 * i + l], ad = fabs(dis);
    if (ad > c.P.k_angle) c.ref_angle[3 * i + l] += (ad - c.P.k_angle) * dis / ad;
}
__global__ void k_absmax(int n, const double *__restrict__ a, double *partial, unsigned int *ticket, double *out)
{   // BaseScene.calc_p_norm (BaseScene.py:1096-1103); max is order independent, so the atomics-free ticket scheme
    // only needs a max-combine
    double m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmax(m, fabs(a[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ double sh[32];
    __shared__ bool last;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        int nw = (blockDim.x + 31) >> 5;
        for (int k = 1; k < nw; k++) m = fmax(m, sh[k]);
        partial[blockIdx.x] = m;
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
        if (last) {
            double r = 0;
            for (int k = 0; k < gridDim.x; k++) r = fmax(r, __ldcg(partial + k));
            *out = r; *ticket = 0;
        }
    }
}

// a . b with the deterministic grid reduction (Newton decrement F . p for the line search's noise test)
__global__ void __launch_bounds__(256) k_dot(int n, const double *__restrict__ a, const double *__restrict__ b, double *partial, unsigned int *ticket, double *out)
{
    double s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += a[i] * b[i];
    s = block_sum(s);
    grid_sum_finish(s, partial, ticket, out);
}

// ------------------------------------------------------------------------------------------------ adjoint element kernels
// Cloth.ref_angle_backprop_a2ax (model_fold_offset.py:1179-1206)
__global__ void k_refangle_a2ax(ClothDev c, const double *__restrict__ pos, const double *__restrict__ ag_step, double *ag_prev, double *pos_grad_step)
{
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= c.NH) return;
    int i = c.hinge_face[h], l = c.hinge_l[h];
    int i2 = c.cf[3 * i + l], q = c.cp[3 * i + l];
    FaceV f = load_face(c, pos, i), f2 = load_face(c, pos, i2);
    d3 n1 = face_normal(f.p[0], f.p[1], f.p[2]), n2 = face_normal(f2.p[0], f2.p[1], f2.p[2]);
    d3 g[4];
    hinge_grad(f.p[l], f.p[(l + 1) % 3], f.p[(l + 2) % 3], f2.p[q], n1, n2, g[0], g[1], g[2], g[3]);
    double th = signed_theta(f, n1, n2, l, false);
    double a = ag_step[3 * i + l];
    ag_prev[3 * i + l] += a;
    double dis = th - c.ref_angle[3 * i + l];
    double sign = (fabs(dis) > c.P.k_angle) ? a : 0.1 * a;
    if (sign == 0.0) return;
    int pt[4] = { f.v[l], f.v[(l + 1) % 3], f.v[(l + 2) % 3], f2.v[q] };
    for (int k = 0; k < 4; k++) red_add3(pos_grad_step, c.offset + pt[k], sign * g[k]);
}
// Cloth.ref_angle_backprop_x2a (model_fold_offset.py:1154-1168)
__global__ void k_refangle_x2a(ClothDev c, const double *__restrict__ pos, const double *__restrict__ z, double *ag_prev)
{
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= c.NH) return;
    int i = c.hinge_face[h], l = c.hinge_l[h];
    int i2 = c.cf[3 * i + l], q = c.cp[3 * i + l];
    FaceV f = load_face(c, pos, i), f2 = load_face(c, pos, i2);
    d3 n1 = face_normal(f.p[0], f.p[1], f.p[2]), n2 = face_normal(f2.p[0], f2.p[1], f2.p[2]);
    d3 g[4];
    hinge_grad(f.p[l], f.p[(l + 1) % 3], f.p[(l + 2) % 3], f2.p[q], n1, n2, g[0], g[1], g[2], g[3]);
    double d_ref = -2.0 * c.P.Kb * c.P.dx * c.P.dx * 1.0 / 3.0;
    int pt[4] = { f.v[l], f.v[(l + 1) % 3], f.v[(l + 2) % 3], f2.v[q] };
    double s = 0;
    for (int k = 0; k < 4; k++) s += -dot(ld3(z, c.offset + pt[k]), g[k]) * d_ref;
    ag_prev[3 * i + l] += s;
}
// BaseScene.contact_energy_backprop (BaseScene.py:682-730): friction lag terms into pos_grad[t-1]
__global__ void k_contact_backprop(ContactDev con, int nc, ContactParams cp, const double *__restrict__ pos,
                                   const double *__restrict__ z, double *pg_prev)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    const int *idx = con.idx + 4 * i;
    const double *w = con.w + 3 * i, *T = con.T + 6 * i, *dx0 = con.dx0 + 3 * i, *nc_ = con.n + 3 * i;
    d3 x0 = ld3(pos, idx[0]), x1 = ld3(pos, idx[1]), x2 = ld3(pos, idx[2]), xv = ld3(pos, idx[3]);
    d3 dx = xv - (w[0] * x0 + w[1] * x1 + w[2] * x2) - mk(dx0[0], dx0[1], dx0[2]);
    double u[2] = { T[0] * dx.x + T[1] * dx.y + T[2] * dx.z, T[3] * dx.x + T[4] * dx.y + T[5] * dx.z };
    double r = sqrt(u[0] * u[0] + u[1] * u[1]);
    double k = con.k[i];
    double pressure = k / con.mu[i];
    double f1 = fr_f1(cp, r);
    double g1[3];
    for (int q = 0; q < 3; q++) g1[q] = k * f1 * (u[0] * T[q] + u[1] * T[3 + q]);
    double wa[4] = { w[0], w[1], w[2], -1 };
    // sum_{i1,j1} zT * dfdp  is a scalar; the (i2, j2) loop then scatters  scalar * wa[i2] * n[j2] * k_contact
    double sc = 0;
    for (int i1 = 0; i1 < 4; i1++)
        for (int j1 = 0; j1 < 3; j1++) sc += z[idx[i1] * 3 + j1] * (wa[i1] * g1[j1] / pressure);
    double h[4] = { f1, 0, 0, f1 };
    if (r > 1e-9) {
        double f2 = fr_f2(cp, r);
        h[0] += f2 * (u[0] / r) * u[0]; h[1] += f2 * (u[0] / r) * u[1];
        h[2] += f2 * (u[1] / r) * u[0]; h[3] += f2 * (u[1] / r) * u[1];
    }
    double h1[3][3];
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++)
            h1[a][b] = k * (T[a] * (h[0] * T[b] + h[1] * T[3 + b]) + T[3 + a] * (h[2] * T[b] + h[3] * T[3 + b]));
    double w1[4] = { -w[0], -w[1], -w[2], 1 };
    // second part: pos_grad[i2][j2] += sum_{i1,j1} z[i1][j1] w1[i1] w1[i2] h1[j1][j2]
    double zw[3] = { 0, 0, 0 };
    for (int i1 = 0; i1 < 4; i1++)
        for (int j1 = 0; j1 < 3; j1++) zw[j1] += z[idx[i1] * 3 + j1] * w1[i1];
    for (int i2 = 0; i2 < 4; i2++) {
        double o[3];
        for (int j2 = 0; j2 < 3; j2++) {
            o[j2] = sc * wa[i2] * nc_[j2] * cp.k_contact;
            for (int j1 = 0; j1 < 3; j1++) o[j2] += zw[j1] * w1[i2] * h1[j1][j2];
        }
        red_add3(pg_prev, idx[i2], mk(o[0], o[1], o[2]));
    }
}
__global__ void k_clamp(int n, double lim, double *a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = fmin(fmax(a[i], -lim), lim);
}
// Grad.get_grad / get_prev_grad / get_prev_prev_grad (analytic_grad_system.py:82-102) and
// Grad.get_parameters_grad (:69-79) fused: one pass over the DOFs.
__global__ void k_adjoint_tail(int n_verts, const double *__restrict__ z, const double *__restrict__ mass, const int *__restrict__ frozen,
                               const double *__restrict__ d_kb, double dt, double damping, double *pg_tm1, double *pg_tm2,
                               double *partial, unsigned int *ticket, double *out_kb)
{
    double s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 3 * n_verts; i += gridDim.x * blockDim.x) {
        if (!frozen[i]) {
            double zi = z[i];
            double xh = zi * mass[i / 3] / (dt * dt);
            pg_tm1[i] += xh * (1 + damping);
            if (pg_tm2) pg_tm2[i] -= xh * damping;
            if (d_kb) s += zi * d_kb[i];
        }
    }
    s = block_sum(s);
    grid_sum_finish(s, partial, ticket, out_kb);
}
__global__ void k_accumulate(double *dst, const double *src) { *dst += *src; }

// ------------------------------------------------------------------------------------------------ host launchers
#define GRID(n, b) (unsigned)(((n) + (b) - 1) / (b))

void launch_face_normals(tsl_ctx *ctx, const ClothDev &c, const double *pos)
{
    k_face_normals<<<GRID(c.NF, 256), 256, 0, ctx->stream>>>(c, pos);
    ctx->launches++;
}
static TetSet tet_set(tsl_ctx *ctx)
{
    TetSet ts;
    ts.n = (int)ctx->tets.size();
    for (int b = 0; b < ts.n; b++) ts.b[b] = ctx->tets[b];
    return ts;
}
void launch_energy(tsl_ctx *ctx, const double *pos, double *out_dev)
{
    ClothSet cs;
    cs.n = (int)ctx->cloths.size();
    for (int q = 0; q < cs.n; q++) cs.c[q] = ctx->cloths[q];
    ContactParams cp = { ctx->cfg.k_contact, ctx->cfg.eps_contact, ctx->cfg.eps_v, ctx->cfg.dt };
    d3 g = mk(ctx->cfg.gravity[0], ctx->cfg.gravity[1], ctx->cfg.gravity[2]);
    const bool fast = ctx->fast_assembly >= 2 && !ctx->dist.on && cs.n == 1;
    if (fast) {
        const ClothDev &c = cs.c[0];
        ClothSet none; none.n = 0;
        // cloth part by tiles (tsl_assembly.cu); the rest (other bodies' vertices, cells, contacts) below adds it in
        double *cloth_E = ctx->egrid_partial + ctx->egrid_blocks;
        launch_energy_rows(ctx, pos, cloth_E);
        k_energy<<<ctx->red_blocks, 256, 0, ctx->stream>>>(none, ctx->cfg.n_verts, pos, ctx->prev_pos, ctx->vel, ctx->mass, g, ctx->cfg.dt, ctx->con, ctx->nc, cp,
                                                            tet_set(ctx), ctx->vgrav, 0, 0x7fffffff, 0, c.offset, c.offset + c.NV, cloth_E,
                                                            ctx->red_partial, ctx->red_ticket, out_dev);
        ctx->launches++;
        return;
    }
    k_energy<<<ctx->red_blocks, 256, 0, ctx->stream>>>(cs, ctx->cfg.n_verts, pos, ctx->prev_pos, ctx->vel,
                                                        ctx->mass, g, ctx->cfg.dt, ctx->con, ctx->nc, cp, tet_set(ctx), ctx->vgrav,
                                                        ctx->dist.own0, ctx->dist.own1, ctx->dist.on ? ctx->dist.nvc : 0, 0, 0, nullptr,
                                                        ctx->red_partial, ctx->red_ticket, out_dev);
    ctx->launches++;
}
void launch_residual(tsl_ctx *ctx, const double *pos)
{
    int n = ctx->cfg.n_verts;
    d3 g = mk(ctx->cfg.gravity[0], ctx->cfg.gravity[1], ctx->cfg.gravity[2]);
    ContactParams cp = { ctx->cfg.k_contact, ctx->cfg.eps_contact, ctx->cfg.eps_v, ctx->cfg.dt };
    const bool fast = ctx->fast_assembly >= 2 && !ctx->dist.on && ctx->cloths.size() == 1;
    if (fast) {
        const ClothDev &c = ctx->cloths[0];
        if (n > c.NV) {
            k_residual_vertex<<<GRID(3 * n, 256), 256, 0, ctx->stream>>>(n, pos, ctx->prev_pos, ctx->vel, ctx->mass, g, ctx->vgrav, ctx->cfg.dt, ctx->F, c.offset, c.offset + c.NV);
            ctx->launches++;
        }
        launch_residual_rows(ctx, pos);
    } else {
        k_residual_vertex<<<GRID(3 * n, 256), 256, 0, ctx->stream>>>(n, pos, ctx->prev_pos, ctx->vel, ctx->mass, g, ctx->vgrav, ctx->cfg.dt, ctx->F, 0, 0);
        ctx->launches++;
        for (auto &c : ctx->cloths) {
            k_residual_cloth<<<GRID(c.NF, 128), 128, 0, ctx->stream>>>(c, pos, ctx->F, 14, 1.0);
            ctx->launches++;
        }
    }
    for (auto &t : ctx->tets) {
        k_residual_tets<<<GRID(t.nc, 128), 128, 0, ctx->stream>>>(t, pos, ctx->F, 1.0);
        ctx->launches++;
    }
    if (ctx->nc > 0) {
        k_residual_contact<<<GRID(ctx->nc, 128), 128, 0, ctx->stream>>>(ctx->con, ctx->nc, cp, pos, ctx->F);
        ctx->launches++;
    }
    k_mask_frozen<<<GRID(3 * n, 256), 256, 0, ctx->stream>>>(3 * n, ctx->frozen, ctx->F);
    ctx->launches++;
    if (ctx->dist.on) launch_zero_ghost(ctx, ctx->F);      // ghost rows only saw part of their elements: the owner has them complete
}
// Cloth.compute_deri: d_kl / d_ka / d_kb = -(edge / area / bending gradient) / K on the cloth rows, zero elsewhere
void launch_cloth_deri(tsl_ctx *ctx, const ClothDev &c, const double *pos, double *d_kl, double *d_ka, double *d_kb)
{
    int n = ctx->cfg.n_verts;
    double *dst[3] = { d_kl, d_ka, d_kb };
    const int mask[3] = { 2, 4, 8 };
    const double K[3] = { c.P.Kl, c.P.Ka, c.P.Kb };
    for (int q = 0; q < 3; q++) {
        if (!dst[q]) continue;
        k_fill_zero<<<GRID(3 * n, 256), 256, 0, ctx->stream>>>(dst[q], 3LL * n);
        k_residual_cloth<<<GRID(c.NF, 128), 128, 0, ctx->stream>>>(c, pos, dst[q], mask[q], -1.0 / K[q]);
        ctx->launches += 2;
    }
}
// Scene.contact_energy_backprop_friction (Scene_sliding.py:140-177): d(friction force)/d(mu) . z over the constraints [c0, c1)
__global__ void __launch_bounds__(256) k_friction_coef_grad(ContactDev con, int c0, int c1, ContactParams cp, const double *__restrict__ pos,
                                                            const double *__restrict__ z, const int *__restrict__ frozen, double *out)
{
    double s = 0;
    for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
        const int *idx = con.idx + 4 * i;
        const double *w = con.w + 3 * i, *T = con.T + 6 * i, *dx0 = con.dx0 + 3 * i;
        d3 x0 = ld3(pos, idx[0]), x1 = ld3(pos, idx[1]), x2 = ld3(pos, idx[2]), xv = ld3(pos, idx[3]);
        d3 dx = xv - (w[0] * x0 + w[1] * x1 + w[2] * x2) - mk(dx0[0], dx0[1], dx0[2]);
        double u[2] = { T[0] * dx.x + T[1] * dx.y + T[2] * dx.z, T[3] * dx.x + T[4] * dx.y + T[5] * dx.z };
        double r = sqrt(u[0] * u[0] + u[1] * u[1]);
        double kf = con.k[i] * fr_f1(cp, r);
        double g1[3];
        for (int q = 0; q < 3; q++) g1[q] = kf * (u[0] * T[q] + u[1] * T[3 + q]);
        double w1[4] = { w[0], w[1], w[2], -1.0 };
        for (int i1 = 0; i1 < 4; i1++)
            for (int j1 = 0; j1 < 3; j1++)
                if (!frozen[3 * idx[i1] + j1]) s += z[3 * idx[i1] + j1] * (w1[i1] * g1[j1] / con.mu[i]);
    }
    s = block_sum(s);
    if (threadIdx.x == 0) *out = s;
}
void launch_friction_coef_grad(tsl_ctx *ctx, const double *pos, const double *z, int c0, int c1, double *out_dev)
{
    ContactParams cp = { ctx->cfg.k_contact, ctx->cfg.eps_contact, ctx->cfg.eps_v, ctx->cfg.dt };
    k_friction_coef_grad<<<1, 256, 0, ctx->stream>>>(ctx->con, c0, c1, cp, pos, z, ctx->frozen, out_dev);   // one block: deterministic order
    ctx->launches++;
}
// Elastic.get_force: F_f = -dE/dx + m g on the body's own vertices
__global__ void k_body_gravity(int nv, int offset, const double *__restrict__ mass, d3 g, double *Ff)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    double m = mass[offset + i];
    Ff[3 * i] = m * g.x; Ff[3 * i + 1] = m * g.y; Ff[3 * i + 2] = m * g.z;
}
void launch_elastic_force(tsl_ctx *ctx, int body, const double *pos, double *Ff)
{
    const TetDev &t = ctx->tets[body];
    d3 g = mk(ctx->h_tet_gravity[3 * body], ctx->h_tet_gravity[3 * body + 1], ctx->h_tet_gravity[3 * body + 2]);
    k_body_gravity<<<GRID(t.nv, 128), 128, 0, ctx->stream>>>(t.nv, t.offset, ctx->mass, g, Ff);
    // k_residual_tets scatters +dE/dx at global vertex ids: shift the destination so that the body's first vertex lands on row 0
    k_residual_tets<<<GRID(t.nc, 128), 128, 0, ctx->stream>>>(t, pos, Ff - 3 * (size_t)t.offset, -1.0);
    ctx->launches += 2;
}
void launch_cloth_param_deri(tsl_ctx *ctx, const ClothDev &c, const double *pos, double *d_kb, bool zero)
{   // Cloth.compute_deri_Kb: d_kb = -(bending gradient) / Kb
    int n = ctx->cfg.n_verts;
    if (zero) { k_fill_zero<<<GRID(3 * n, 256), 256, 0, ctx->stream>>>(d_kb, 3LL * n); ctx->launches++; }
    k_residual_cloth<<<GRID(c.NF, 128), 128, 0, ctx->stream>>>(c, pos, d_kb, 8, -1.0 / c.P.Kb);
    ctx->launches += 1;
}
// elements of the Hessian into the sink S (matrix values, or the counting pass of the adjoint)
template <typename T>
static void launch_hessian_elements(tsl_ctx *ctx, const double *pos, Sink<T> S, T *side, int spd, int sym, int newton_model)
{
    ContactParams cp = { ctx->cfg.k_contact, ctx->cfg.eps_contact, ctx->cfg.eps_v, ctx->cfg.dt };
    for (auto &c : ctx->cloths) {
        if (newton_model) {
            k_hessian_tri_newton<T><<<GRID(c.NF, 128), 128, 0, ctx->stream>>>(c, pos, ctx->frozen, S, spd);
            ctx->launches += 1;
        } else {
            launch_face_normals(ctx, c, pos);
            k_q1_prepare<<<1, 32, 0, ctx->stream>>>(c, pos);
            k_hessian_tri<T><<<GRID(c.NF, 128), 128, 0, ctx->stream>>>(c, pos, ctx->frozen, S, spd, sym);
            ctx->launches += 3;
        }
        k_hessian_hinge<T><<<GRID(c.NH, 128), 128, 0, ctx->stream>>>(c, pos, ctx->frozen, S);
        ctx->launches += 1;
    }
    for (auto &t : ctx->tets) {
        // reference matrices: only the tactile model is projected, and only when spd (forward); Newton model: the exact cell
        // Hessian, projected in the clamped variant for either model
        int project = newton_model ? spd : (spd && t.P.kind == 1);
        k_hessian_tets<T><<<GRID(t.nc, 64), 64, 0, ctx->stream>>>(t, pos, ctx->frozen, S, project);
        ctx->launches++;
    }
    if (ctx->nc > 0) {
        if (ctx->general_contact)
            k_hessian_contact_general<T><<<GRID(ctx->nc, 64), 64, 0, ctx->stream>>>(ctx->con, ctx->nc, cp, pos, ctx->frozen, ctx->A.diag_pb, S, side, spd | newton_model, newton_model);
        else
            k_hessian_contact<T><<<GRID(ctx->nc, 128), 128, 0, ctx->stream>>>(ctx->con, ctx->nc, cp, pos, ctx->frozen, ctx->A.diag_pb, S, spd | newton_model, ctx->error_flag);
        ctx->launches++;
    }
}
template <typename T>
static void launch_hessian_t(tsl_ctx *ctx, const double *pos, T *val, T *side, int spd, int sym, int newton_model)
{
    int n = ctx->cfg.n_verts;
    cudaMemsetAsync(val, 0, sizeof(T) * 9 * (size_t)ctx->A.nnzb_pad, ctx->stream);
    k_hessian_mass<T><<<GRID(n, 256), 256, 0, ctx->stream>>>(n, ctx->mass, ctx->cfg.dt, ctx->A.diag_pb, val, 0, 0);
    ctx->launches++;
    Sink<T> S = { val, nullptr, nullptr };
    launch_hessian_elements<T>(ctx, pos, S, side, spd, sym, newton_model);
}
// The two forward Newton matrices of one Newton iteration: exact -> A.val32, clamped -> A.val32c.
// Single-cloth scenes: the cloth rows of both leave one owner-computes pass (tsl_assembly.cu); tetrahedral bodies and contacts
// are added per matrix by the element kernels above.  Otherwise: two scatter assemblies.
void launch_hessian_newton_pair(tsl_ctx *ctx, const double *pos)
{
    if (!ctx->fast_assembly || ctx->cloths.size() != 1) {
        launch_hessian_t<float>(ctx, pos, ctx->A.val32c, ctx->cside32, 1, 0, 1);
        launch_hessian_t<float>(ctx, pos, ctx->A.val32, ctx->cside32, 0, 0, 1);
        return;
    }
    const int n = ctx->cfg.n_verts;
    const ClothDev &c = ctx->cloths[0];
    ContactParams cp = { ctx->cfg.k_contact, ctx->cfg.eps_contact, ctx->cfg.eps_v, ctx->cfg.dt };
    launch_hessian_rows(ctx, pos, ctx->A.val32, ctx->A.val32c);
    for (int pass = 0; pass < 2; pass++) {
        float *val = pass == 0 ? ctx->A.val32c : ctx->A.val32;
        const int spd = pass == 0 ? 1 : 0;
        if (n > c.NV) {
            k_hessian_mass<float><<<GRID(n, 256), 256, 0, ctx->stream>>>(n, ctx->mass, ctx->cfg.dt, ctx->A.diag_pb, val, c.offset, c.offset + c.NV);
            ctx->launches++;
        }
        Sink<float> S = { val, nullptr, nullptr };
        for (auto &t : ctx->tets) {
            k_hessian_tets<float><<<GRID(t.nc, 64), 64, 0, ctx->stream>>>(t, pos, ctx->frozen, S, spd);
            ctx->launches++;
        }
        if (ctx->nc > 0) {
            if (ctx->general_contact)
                k_hessian_contact_general<float><<<GRID(ctx->nc, 64), 64, 0, ctx->stream>>>(ctx->con, ctx->nc, cp, pos, ctx->frozen, ctx->A.diag_pb, S, ctx->cside32, 1, 1);
            else
                k_hessian_contact<float><<<GRID(ctx->nc, 128), 128, 0, ctx->stream>>>(ctx->con, ctx->nc, cp, pos, ctx->frozen, ctx->A.diag_pb, S, 1, ctx->error_flag);
            ctx->launches++;
        }
    }
}
void launch_hessian(tsl_ctx *ctx, const double *pos, bool f64, int spd, int sym, int newton_model, bool into_clamped)
{
    if (f64) launch_hessian_t<double>(ctx, pos, ctx->A.val64, ctx->cside64, spd, sym, newton_model);
    else launch_hessian_t<float>(ctx, pos, into_clamped ? ctx->A.val32c : ctx->A.val32, ctx->cside32, spd, sym, newton_model);
}
// second assembly of Grad.transfer_grad with counting_z_frozen (analytic_grad_single.py:240-243): zf[j] -= H[i][j] z[i]
// over free rows i and frozen columns j of the reference's un-projected Hessian.  zf must be zeroed by the caller.
void launch_hessian_counting(tsl_ctx *ctx, const double *pos, const double *z, double *zf)
{
    Sink<double> S = { nullptr, z, zf };
    launch_hessian_elements<double>(ctx, pos, S, nullptr, 0, 0, 0);
}
void launch_tets_param_grad(tsl_ctx *ctx, const double *pos, const double *z, double *d_mu, double *d_lam, double *out2_dev)
{
    int n = ctx->cfg.n_verts;
    k_fill_zero<<<GRID(3 * n, 256), 256, 0, ctx->stream>>>(d_mu, 3LL * n);
    k_fill_zero<<<GRID(3 * n, 256), 256, 0, ctx->stream>>>(d_lam, 3LL * n);
    ctx->launches += 2;
    for (auto &t : ctx->tets) {
        k_tets_deri<<<GRID(t.nc, 128), 128, 0, ctx->stream>>>(t, pos, d_mu, d_lam);
        ctx->launches++;
    }
    if (z) {
        // one block: deterministic order; the bodies are small and this runs once per backward step
        k_masked_dot2<<<1, 256, 0, ctx->stream>>>(3 * n, z, ctx->frozen, d_mu, d_lam, out2_dev);
        ctx->launches++;
    }
}
void launch_axpy_pos(tsl_ctx *ctx, const double *x1, const double *p, double alpha, double *pos)
{
    int n = 3 * ctx->cfg.n_verts;
    k_axpy_pos<<<GRID(n, 256), 256, 0, ctx->stream>>>(n, x1, p, alpha, pos);
    ctx->launches++;
}
void launch_axpy2_pos(tsl_ctx *ctx, const double *x1, const double *u, double a, const double *w, double b, double *pos)
{
    int n = 3 * ctx->cfg.n_verts;
    k_axpy2_pos<<<GRID(n, 256), 256, 0, ctx->stream>>>(n, x1, u, a, w, b, pos);
    ctx->launches++;
}
void launch_update_vel(tsl_ctx *ctx)
{
    int n = 3 * ctx->cfg.n_verts;
    k_update_vel<<<GRID(n, 256), 256, 0, ctx->stream>>>(n, ctx->pos, ctx->prev_pos, ctx->cfg.damping / ctx->cfg.dt, ctx->vel);
    ctx->launches++;
}
void launch_update_ref_angle(tsl_ctx *ctx, const ClothDev &c)
{
    k_update_ref_angle<<<GRID(c.NH, 128), 128, 0, ctx->stream>>>(c, ctx->pos);
    ctx->launches++;
}
void launch_absmax(tsl_ctx *ctx, const double *a, int n, double *out_dev)
{
    k_absmax<<<ctx->red_blocks, 256, 0, ctx->stream>>>(n, a, ctx->red_partial, ctx->red_ticket, out_dev);
    ctx->launches++;
}
void launch_dot(tsl_ctx *ctx, const double *a, const double *b, int n, double *out_dev)
{
    k_dot<<<ctx->red_blocks, 256, 0, ctx->stream>>>(n, a, b, ctx->red_partial, ctx->red_ticket, out_dev);
    ctx->launches++;
}
void launch_refangle_a2ax(tsl_ctx *ctx, const ClothDev &c, const double *pos, const double *ag_step, double *ag_prev, double *pg_step)
{
    k_refangle_a2ax<<<GRID(c.NH, 128), 128, 0, ctx->stream>>>(c, pos, ag_step, ag_prev, pg_step);
    ctx->launches++;
}
void launch_refangle_x2a(tsl_ctx *ctx, const ClothDev &c, const double *pos, const double *z, double *ag_prev)
{
    k_refangle_x2a<<<GRID(c.NH, 128), 128, 0, ctx->stream>>>(c, pos, z, ag_prev);
    ctx->launches++;
}
void launch_contact_backprop(tsl_ctx *ctx, const double *pos, const double *z, double *pg_prev)
{
    if (ctx->nc == 0) return;
    ContactParams cp = { ctx->cfg.k_contact, ctx->cfg.eps_contact, ctx->cfg.eps_v, ctx->cfg.dt };
    k_contact_backprop<<<GRID(ctx->nc, 128), 128, 0, ctx->stream>>>(ctx->con, ctx->nc, cp, pos, z, pg_prev);
    ctx->launches++;
}
void launch_clamp(tsl_ctx *ctx, double *a, int n, double lim)
{
    k_clamp<<<GRID(n, 256), 256, 0, ctx->stream>>>(n, lim, a);
    ctx->launches++;
}
void launch_adjoint_tail(tsl_ctx *ctx, const double *z, const double *d_kb, double *pg_tm1, double *pg_tm2, double *grad_kb_accum)
{
    k_adjoint_tail<<<ctx->red_blocks, 256, 0, ctx->stream>>>(ctx->cfg.n_verts, z, ctx->mass, ctx->frozen, d_kb, ctx->cfg.dt, 1.0, pg_tm1, pg_tm2,
                                                             ctx->red_partial, ctx->red_ticket, ctx->red_out + 2);
    ctx->launches++;
    if (grad_kb_accum) {
        k_accumulate<<<1, 1, 0, ctx->stream>>>(grad_kb_accum, ctx->red_out + 2);
        ctx->launches++;
    }
}

}  // namespace tsl
