// tsl_contact.cu -- vertex-triangle contact candidates and constraint build (sm_100a).
//
// Replaces BaseScene.calc_vn (code/engine/BaseScene.py:837-850), geometry.p2g / project_pair / pt2tri
// (code/engine/geometry.py:23-221) and BaseScene.contact_pair_analysis (BaseScene.py:778-816).
//
// Design: instead of the reference's dense 132^3 count/offset grids, the faces of a surface body are sorted
// by cell key with a stable radix sort (values = ascending face ids).  A query vertex visits its 3x3 columns of
// cells; the three cells of a column are one contiguous key range, found with two binary searches.  The
// resulting candidate order -- cells in (i,j,k) lexicographic order, faces ascending inside a cell -- is the
// serial order of the reference's scatter loop, which matters because its tie rule (|d - d_min| < 1e-5, larger
// cosine wins, first wins on equality) is order dependent.  Everything is fp64 so index sets match bit-exactly.
// Constraints are compacted with a prefix sum in ascending vertex order (deterministic; the reference's
// atomic append order is not).
#include <cub/cub.cuh>

#include "tsl_internal.cuh"
#include "tsl_kernels.cuh"

namespace tsl {

#define GRID(n, b) (unsigned)(((n) + (b) - 1) / (b))

__global__ void k_vn_scatter(int nf, const int *__restrict__ faces, const double *__restrict__ pos, double *vn)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    int a = faces[3 * i], b = faces[3 * i + 1], c = faces[3 * i + 2];
    d3 n = cross(ld3(pos, b) - ld3(pos, a), ld3(pos, c) - ld3(pos, a));
    int v[3] = { a, b, c };
    for (int q = 0; q < 3; q++) { atomicAdd(vn + 3 * v[q], n.x); atomicAdd(vn + 3 * v[q] + 1, n.y); atomicAdd(vn + 3 * v[q] + 2, n.z); }
}
__global__ void k_vn_normalize(int nv, double *vn)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    d3 n = ld3(vn, i);
    double l = norm(n);
    vn[3 * i] = n.x / l; vn[3 * i + 1] = n.y / l; vn[3 * i + 2] = n.z / l;   // 0/0 = NaN off-surface, as in the reference
}

struct GridP { double h; int n; };
__device__ __forceinline__ void grid_idx(const GridP &g, d3 x, int *o)
{   // geometry.grid_idx (geometry.py:89-94)
    double bound = g.h * (g.n - 1) / 2;
    double c[3] = { x.x, x.y, x.z };
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double v = c[k] < -bound ? -bound : (c[k] > bound ? bound : c[k]);
        o[k] = (int)floor(v / g.h) + g.n / 2;
    }
}
__global__ void k_face_cells(GridP g, int f_start, int nf, const int *__restrict__ faces, const double *__restrict__ pos,
                             unsigned int *key, int *fid)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    int f = f_start + i;
    d3 a = ld3(pos, faces[3 * f]), b = ld3(pos, faces[3 * f + 1]), c = ld3(pos, faces[3 * f + 2]);
    d3 mid = mk((a.x + b.x + c.x) / 3, (a.y + b.y + c.y) / 3, (a.z + b.z + c.z) / 3);
    int o[3];
    grid_idx(g, mid, o);
    key[i] = (unsigned int)((o[0] * g.n + o[1]) * g.n + o[2]);
    fid[i] = f;
}

// geometry.pt2tri (geometry.py:23-87): region code c, distance d, barycentric w
__device__ int pt2tri(d3 x, d3 p1, d3 p2, d3 p3, double &d, double *w)
{
    // operation order mirrors the fp64 oracle exactly (this file is compiled with --fmad=false) so that the
    // region / tie decisions, and with them the contact index sets, are reproducible bit for bit
    d3 e1 = p2 - p1, e2 = p3 - p2, e3 = p1 - p3;
    double l1 = norm(e1), l2 = norm(e2), l3 = norm(e3);
    e1 = mk(e1.x / l1, e1.y / l1, e1.z / l1); e2 = mk(e2.x / l2, e2.y / l2, e2.z / l2); e3 = mk(e3.x / l3, e3.y / l3, e3.z / l3);
    d3 t = cross(e1, e3);
    double lt = norm(t);
    d3 n = mk(-(t.x / lt), -(t.y / lt), -(t.z / lt));
    d3 x1 = x - dot(x - p1, n) * n;
    w[0] = w[1] = w[2] = 0;
    int c = 0;
    d3 a1 = x1 - p1, a2 = x1 - p2, a3 = x1 - p3;
    if (dot(cross(a1, e1), n) > 0) {
        if (dot(a1, e1) < 0) { c = 1; d = norm(x - p1); w[0] = 1; }
        else if (dot(a2, e1) > 0) { c = 2; d = norm(x - p2); w[1] = 1; }
        else {
            c = -3;
            d3 ee = p2 - p1;
            double alpha = dot(a1, e1) / dot(ee, e1);
            d = norm(x - (p1 + alpha * ee)); w[0] = 1 - alpha; w[1] = alpha;
        }
    } else if (dot(cross(a2, e2), n) > 0) {
        if (dot(a2, e2) < 0) { c = 2; d = norm(x - p2); w[1] = 1; }
        else if (dot(a3, e2) > 0) { c = 3; d = norm(x - p3); w[2] = 1; }
        else {
            c = -1;
            d3 ee = p3 - p2;
            double alpha = dot(a2, e2) / dot(ee, e2);
            d = norm(x - (p2 + alpha * ee)); w[1] = 1 - alpha; w[2] = alpha;
        }
    } else if (dot(cross(a3, e3), n) > 0) {
        if (dot(a3, e3) < 0) { c = 3; d = norm(x - p3); w[2] = 1; }
        else if (dot(a1, e3) > 0) { c = 1; d = norm(x - p1); w[0] = 1; }
        else {
            c = -2;
            d3 ee = p1 - p3;
            double alpha = dot(a3, e3) / dot(ee, e3);
            d = norm(x - (p3 + alpha * ee)); w[0] = alpha; w[2] = 1 - alpha;
        }
    } else {
        d = norm(x - x1);
        double S = norm(cross(p3 - p1, p2 - p1));
        w[0] = dot(cross(p3 - p2, x1 - p2), n) / S;
        w[1] = dot(cross(p1 - p3, x1 - p3), n) / S;
        w[2] = dot(cross(p2 - p1, x1 - p1), n) / S;
    }
    return c;
}

__device__ __forceinline__ int lower_bound_u32(const unsigned int *a, int n, unsigned int key)
{
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// geometry.project_pair (geometry.py:165-221) for the query vertices [v_start, v_end) against one sorted surface body
__global__ void __launch_bounds__(128) k_project_pair(GridP g, int nf, const unsigned int *__restrict__ key_sorted,
                                                      const int *__restrict__ fid_sorted, const int *__restrict__ faces,
                                                      const double *__restrict__ pos, const double *__restrict__ vn,
                                                      const int *__restrict__ border_flag, int v_start, int v_end,
                                                      int *proj_flag, int *proj_dir, int *proj_idx, double *proj_w)
{
    int i = v_start + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v_end) return;
    d3 xq = ld3(pos, i);
    int q[3];
    grid_idx(g, xq, q);
    double d_min = 1e6, cos_max = -1e6;
    int pflag = 0, pidx[3] = { 0, 0, 0 };
    double pw[3] = { 0, 0, 0 };
    for (int gi = max(q[0] - 1, 0); gi <= min(q[0] + 1, g.n - 1); gi++)
        for (int gj = max(q[1] - 1, 0); gj <= min(q[1] + 1, g.n - 1); gj++) {
            unsigned int k0 = (unsigned int)((gi * g.n + gj) * g.n + max(q[2] - 1, 0));
            unsigned int k1 = (unsigned int)((gi * g.n + gj) * g.n + min(q[2] + 1, g.n - 1));
            int s = lower_bound_u32(key_sorted, nf, k0);
            for (int t = s; t < nf && key_sorted[t] <= k1; t++) {
                int fi = fid_sorted[t];
                int fv[3] = { faces[3 * fi], faces[3 * fi + 1], faces[3 * fi + 2] };
                d3 v1 = ld3(pos, fv[0]), v2 = ld3(pos, fv[1]), v3 = ld3(pos, fv[2]);
                double d, w[3];
                int c = pt2tri(xq, v1, v2, v3, d, w);
                d3 vt = w[0] * v1 + w[1] * v2 + w[2] * v3;
                d3 nt = cross(v2 - v1, v3 - v1);
                double nl = norm(nt);
                d3 dd = xq - vt;
                double cs = dd.x * (nt.x / nl) + dd.y * (nt.y / nl) + dd.z * (nt.z / nl);
                if (d < d_min - 1e-5 || (d < d_min + 1e-5 && cs > cos_max)) {
                    d_min = d; cos_max = cs;
                    pidx[0] = fv[0]; pidx[1] = fv[1]; pidx[2] = fv[2];
                    pw[0] = w[0]; pw[1] = w[1]; pw[2] = w[2];
                    if (c == 0) pflag = 1;
                    else if (c > 0) pflag = !border_flag[fv[c - 1]];
                    else {
                        int p1 = (c != -3) ? fv[2] : fv[0];
                        int p2 = (c != -3) ? fv[2 + c] : fv[1];
                        pflag = !(border_flag[p1] && border_flag[p2]);
                    }
                }
            }
        }
    d3 v = pw[0] * ld3(pos, pidx[0]) + pw[1] * ld3(pos, pidx[1]) + pw[2] * ld3(pos, pidx[2]);
    d3 n = pw[0] * ld3(vn, pidx[0]) + pw[1] * ld3(vn, pidx[1]) + pw[2] * ld3(vn, pidx[2]);
    if (proj_flag[i] == 0 && pflag == 1) proj_dir[i] = dot(xq - v, n) > 0;
    proj_flag[i] = pflag;
    proj_idx[3 * i] = pidx[0]; proj_idx[3 * i + 1] = pidx[1]; proj_idx[3 * i + 2] = pidx[2];
    proj_w[3 * i] = pw[0]; proj_w[3 * i + 1] = pw[1]; proj_w[3 * i + 2] = pw[2];
}

// BaseScene.contact_pair_analysis (BaseScene.py:778-816), split into flag / compacted emit
struct Cand { int idx[3]; double w[3]; d3 n, xc, x0c; double dist; };
__device__ __forceinline__ bool candidate(int i, const double *pos, const double *prev_pos, const int *proj_flag, const int *proj_dir,
                                          const int *proj_idx, const double *proj_w, double eps, Cand &c)
{
    if (!proj_flag[i]) return false;
    for (int k = 0; k < 3; k++) { c.idx[k] = proj_idx[3 * i + k]; c.w[k] = proj_w[3 * i + k]; }
    d3 x0 = ld3(pos, c.idx[0]), x1 = ld3(pos, c.idx[1]), x2 = ld3(pos, c.idx[2]);
    c.xc = c.w[0] * x0 + c.w[1] * x1 + c.w[2] * x2;
    c.x0c = c.w[0] * ld3(prev_pos, c.idx[0]) + c.w[1] * ld3(prev_pos, c.idx[1]) + c.w[2] * ld3(prev_pos, c.idx[2]);
    d3 n = cross(x1 - x0, x2 - x0);
    double nl = norm(n);
    n = mk(n.x / nl, n.y / nl, n.z / nl);
    if (proj_dir[i] == 0) {
        n = -n;
        int t = c.idx[1]; c.idx[1] = c.idx[2]; c.idx[2] = t;
        double tw = c.w[1]; c.w[1] = c.w[2]; c.w[2] = tw;
    }
    c.n = n;
    c.dist = dot(ld3(pos, i) - c.xc, n);
    return c.dist < eps;
}
__global__ void k_contact_flag(int v_start, int v_end, const double *__restrict__ pos, const double *__restrict__ prev_pos,
                               const int *__restrict__ proj_flag, const int *__restrict__ proj_dir, const int *__restrict__ proj_idx,
                               const double *__restrict__ proj_w, double eps, int *flag)
{
    int i = v_start + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v_end) return;
    Cand c;
    flag[i - v_start] = candidate(i, pos, prev_pos, proj_flag, proj_dir, proj_idx, proj_w, eps, c) ? 1 : 0;
}
__global__ void k_contact_emit(int v_start, int v_end, const double *__restrict__ pos, const double *__restrict__ prev_pos,
                               const int *__restrict__ proj_flag, const int *__restrict__ proj_dir, const int *__restrict__ proj_idx,
                               const double *__restrict__ proj_w, double eps, double k_contact, double mu,
                               const int *__restrict__ flag, const int *__restrict__ scan, int base, ContactDev con)
{
    int i = v_start + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= v_end || !flag[i - v_start]) return;
    Cand c;
    candidate(i, pos, prev_pos, proj_flag, proj_dir, proj_idx, proj_w, eps, c);
    int s = base + scan[i - v_start];
    double cforce = k_contact * (c.dist - eps);
    con.idx[4 * s] = c.idx[0]; con.idx[4 * s + 1] = c.idx[1]; con.idx[4 * s + 2] = c.idx[2]; con.idx[4 * s + 3] = i;
    d3 dx0 = ld3(prev_pos, i) - c.x0c;
    con.w[3 * s] = c.w[0]; con.w[3 * s + 1] = c.w[1]; con.w[3 * s + 2] = c.w[2];
    con.dx0[3 * s] = dx0.x; con.dx0[3 * s + 1] = dx0.y; con.dx0[3 * s + 2] = dx0.z;
    con.n[3 * s] = c.n.x; con.n[3 * s + 1] = c.n.y; con.n[3 * s + 2] = c.n.z;
    con.k[s] = -mu * cforce; con.mu[s] = mu;
    d3 n = c.n, t1;
    if (fabs(n.x) < 0.5) t1 = mk(n.x, n.z, -n.y); else t1 = mk(n.y, -n.x, n.z);
    d3 t2 = cross(n, t1);
    t1 = cross(n, t2);                                  // Q13: orthogonal to n, not normalised
    con.T[6 * s] = t1.x; con.T[6 * s + 1] = t1.y; con.T[6 * s + 2] = t1.z;
    con.T[6 * s + 3] = t2.x; con.T[6 * s + 4] = t2.y; con.T[6 * s + 5] = t2.z;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return TSL_ERR_CUDA; } } while (0)

int contact_alloc(tsl_ctx *ctx)
{
    int nv = ctx->cfg.n_verts, nb = (int)ctx->bodies.size();
    int max_nf = 1;
    for (auto &b : ctx->bodies) max_nf = std::max(max_nf, b.f_end - b.f_start);
    CK(cudaMalloc(&ctx->vn, sizeof(double) * 3 * nv));
    size_t pb = (size_t)std::max(nb, 1) * nv;
    CK(cudaMalloc(&ctx->proj_flag, sizeof(int) * pb)); CK(cudaMalloc(&ctx->proj_dir, sizeof(int) * pb));
    CK(cudaMalloc(&ctx->proj_idx, sizeof(int) * 3 * pb)); CK(cudaMalloc(&ctx->proj_w, sizeof(double) * 3 * pb));
    CK(cudaMemset(ctx->proj_flag, 0, sizeof(int) * pb)); CK(cudaMemset(ctx->proj_dir, 0, sizeof(int) * pb));
    CK(cudaMemset(ctx->proj_idx, 0, sizeof(int) * 3 * pb)); CK(cudaMemset(ctx->proj_w, 0, sizeof(double) * 3 * pb));
    CK(cudaMalloc(&ctx->cell_key, sizeof(unsigned) * max_nf)); CK(cudaMalloc(&ctx->cell_key_sorted, sizeof(unsigned) * max_nf));
    CK(cudaMalloc(&ctx->face_id, sizeof(int) * max_nf)); CK(cudaMalloc(&ctx->face_id_sorted, sizeof(int) * max_nf));
    CK(cudaMalloc(&ctx->cflag, sizeof(int) * (nv + 1))); CK(cudaMalloc(&ctx->cscan, sizeof(int) * (nv + 1)));
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, ctx->cell_key, ctx->cell_key_sorted, ctx->face_id, ctx->face_id_sorted, max_nf);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, ctx->cflag, ctx->cscan, nv + 1);
    ctx->cub_tmp_bytes = std::max(b1, b2) + 256;
    CK(cudaMalloc(&ctx->cub_tmp, ctx->cub_tmp_bytes));
    int m = ctx->cfg.max_n_constraints;
    ctx->con.max_nc = m;
    CK(cudaMalloc(&ctx->con.idx, sizeof(int) * 4 * m)); CK(cudaMalloc(&ctx->con.w, sizeof(double) * 3 * m));
    CK(cudaMalloc(&ctx->con.k, sizeof(double) * m)); CK(cudaMalloc(&ctx->con.mu, sizeof(double) * m));
    CK(cudaMalloc(&ctx->con.dx0, sizeof(double) * 3 * m)); CK(cudaMalloc(&ctx->con.T, sizeof(double) * 6 * m));
    CK(cudaMalloc(&ctx->con.n, sizeof(double) * 3 * m));
    return TSL_OK;
}

int contact_detect(tsl_ctx *ctx, const double *pos, const double *prev_pos)
{
    int nv = ctx->cfg.n_verts;
    cudaStream_t st = ctx->stream;
    ctx->nc = 0;
    if (ctx->nc_dev) CK(cudaMemsetAsync(ctx->nc_dev, 0, sizeof(int), st));
    if (ctx->pairs.empty() || ctx->tot_nf == 0) return TSL_OK;
    GridP g = { ctx->cfg.grid_h, ctx->cfg.grid_n };
    CK(cudaMemsetAsync(ctx->vn, 0, sizeof(double) * 3 * nv, st));
    k_vn_scatter<<<GRID(ctx->tot_nf, 256), 256, 0, st>>>(ctx->tot_nf, ctx->faces, pos, ctx->vn);
    k_vn_normalize<<<GRID(nv, 256), 256, 0, st>>>(nv, ctx->vn);
    ctx->launches += 2;
    int bits = 1;
    while ((1ull << bits) < (unsigned long long)g.n * g.n * g.n) bits++;
    // geometry.projection_query: one sort per surface body, then every registered query range against it
    for (int b = 0; b < (int)ctx->bodies.size(); b++) {
        bool used = false;
        for (auto &p : ctx->pairs) used = used || p.body == b;
        if (!used) continue;
        const SurfaceBody &sb = ctx->bodies[b];
        int nf = sb.f_end - sb.f_start;
        if (nf <= 0) continue;
        k_face_cells<<<GRID(nf, 256), 256, 0, st>>>(g, sb.f_start, nf, ctx->faces, pos, ctx->cell_key, ctx->face_id);
        size_t tb = ctx->cub_tmp_bytes;
        CK(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp, tb, ctx->cell_key, ctx->cell_key_sorted, ctx->face_id, ctx->face_id_sorted, nf, 0, bits, st));
        ctx->launches += 3;
        size_t off = (size_t)b * nv;
        // the reference queries every other body's vertices against body b; only the ranges used by
        // contact_analysis are observable, so only those are evaluated (merged per body)
        for (auto &p : ctx->pairs) {
            if (p.body != b) continue;
            int n = p.v_end - p.v_start;
            if (n <= 0) continue;
            k_project_pair<<<GRID(n, 128), 128, 0, st>>>(g, nf, ctx->cell_key_sorted, ctx->face_id_sorted, ctx->faces, pos, ctx->vn,
                                                         ctx->border_flag, p.v_start, p.v_end, ctx->proj_flag + off, ctx->proj_dir + off,
                                                         ctx->proj_idx + 3 * off, ctx->proj_w + 3 * off);
            ctx->launches++;
        }
    }
    // Scene.contact_analysis: the registered pairs in order
    ctx->pair_start.assign(1, 0);
    for (auto &p : ctx->pairs) {
        int n = p.v_end - p.v_start;
        if (n <= 0) { ctx->pair_start.push_back(ctx->nc); continue; }
        size_t off = (size_t)p.body * nv;
        k_contact_flag<<<GRID(n, 128), 128, 0, st>>>(p.v_start, p.v_end, pos, prev_pos, ctx->proj_flag + off, ctx->proj_dir + off,
                                                     ctx->proj_idx + 3 * off, ctx->proj_w + 3 * off, ctx->cfg.eps_contact, ctx->cflag);
        CK(cudaMemsetAsync(ctx->cflag + n, 0, sizeof(int), st));
        size_t tb = ctx->cub_tmp_bytes;
        CK(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp, tb, ctx->cflag, ctx->cscan, n + 1, st));
        int count = 0;
        CK(cudaMemcpyAsync(&count, ctx->cscan + n, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        ctx->launches += 2;
        if (ctx->nc + count > ctx->con.max_nc) {
            ctx->err = "max_n_constraints exceeded: " + std::to_string(ctx->nc + count);
            return TSL_ERR_CAPACITY;
        }
        if (count > 0) {
            k_contact_emit<<<GRID(n, 128), 128, 0, st>>>(p.v_start, p.v_end, pos, prev_pos, ctx->proj_flag + off, ctx->proj_dir + off,
                                                         ctx->proj_idx + 3 * off, ctx->proj_w + 3 * off, ctx->cfg.eps_contact,
                                                         ctx->cfg.k_contact, p.mu, ctx->cflag, ctx->cscan, ctx->nc, ctx->con);
            ctx->launches++;
        }
        ctx->nc += count;
        ctx->pair_start.push_back(ctx->nc);
    }
    if (ctx->nc_dev) CK(cudaMemcpyAsync(ctx->nc_dev, &ctx->nc, sizeof(int), cudaMemcpyHostToDevice, st));   // pageable source: staged before return
    CK(cudaGetLastError());
    return TSL_OK;
}

}  // namespace tsl
