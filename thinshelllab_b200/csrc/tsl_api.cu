// tsl_api.cu -- the C ABI of libtsl.so (see include/tsl.h): scene description, pattern construction and the
// Newton / adjoint drivers that sequence the kernels of tsl_physics.cu, tsl_contact.cu and tsl_linalg.cu.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tsl_internal.cuh"
#include "tsl_kernels.cuh"

using namespace tsl;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return TSL_ERR_CUDA; } } while (0)
#define REQUIRE(c, msg) do { if (!(c)) { ctx->err = (msg); return TSL_ERR_INVALID; } } while (0)
#define TRY(x) do { int r_ = (x); if (r_ != TSL_OK) return r_; } while (0)

// Orders the library's stream after the caller's stream on entry and the caller's stream after the library's on exit.
struct StreamScope {
    tsl_ctx *c;
    explicit StreamScope(tsl_ctx *ctx) : c(ctx)
    {
        if (!c) return;
        cudaSetDevice(c->device);                // a second context on another device may have changed the current one
        cudaEventRecord(c->ev_in, c->user_stream);
        cudaStreamWaitEvent(c->stream, c->ev_in, 0);
    }
    ~StreamScope()
    {
        if (!c) return;
        cudaEventRecord(c->ev_out, c->stream);
        cudaStreamWaitEvent(c->user_stream, c->ev_out, 0);
    }
};

static double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

template <typename T>
static int upload(tsl_ctx *ctx, T **dst, const std::vector<T> &src)
{
    CK(cudaMalloc(dst, sizeof(T) * std::max<size_t>(src.size(), 1)));
    if (!src.empty()) CK(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
    return TSL_OK;
}

extern "C" {

const char *tsl_version(void) { return "thinshelllab_b200 libtsl 0.1 (sm_100a)"; }

int tsl_create(const tsl_config *cfg, tsl_ctx **out)
{
    if (!cfg || !out || cfg->struct_size != (int)sizeof(tsl_config)) return TSL_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return TSL_ERR_CUDA;   // no CPU fallback: fail loudly
    tsl_ctx *ctx = new tsl_ctx();
    ctx->cfg = *cfg;
    cudaGetDevice(&ctx->device);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_in, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_out, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return TSL_ERR_CUDA; }
    if (const char *e = getenv("TSL_GRAPHS")) ctx->use_graphs = atoi(e);
    if (ctx->cfg.grid_h <= 0) ctx->cfg.grid_h = 0.003;
    if (ctx->cfg.grid_n <= 0) ctx->cfg.grid_n = 132;
    *out = ctx;
    return TSL_OK;
}

int tsl_destroy(tsl_ctx *ctx)
{
    if (!ctx) return TSL_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    tsl::dist_destroy(ctx);
    tsl::dense_free(ctx);
    cudaFree(ctx->egrid_partial); cudaFree(ctx->egrid_ticket);
    cudaFree(ctx->gm_V); cudaFree(ctx->gm_Z); cudaFree(ctx->gm_h); if (ctx->gm_h_host) cudaFreeHost(ctx->gm_h_host);
    tsl::graphs_invalidate(ctx);
    tsl::mg_free(ctx);
    cudaFree(ctx->A.val32); cudaFree(ctx->A.val32c); cudaFree(ctx->A.val32m); cudaFree(ctx->A.val32t); cudaFree(ctx->cg_r64tmp); cudaFree(ctx->A.val64); cudaFree(ctx->A.colidx); cudaFree(ctx->A.slice_base); cudaFree(ctx->A.diag_pb);
    cudaFree(ctx->cg_x); cudaFree(ctx->cg_r); cudaFree(ctx->cg_z); cudaFree(ctx->cg_p); cudaFree(ctx->cg_q); cudaFree(ctx->cg_r32); cudaFree(ctx->ncdir);
    for (int i = 0; i < 8; i++) cudaFree(ctx->bi[i]);
    cudaFree(ctx->minv32); cudaFree(ctx->minv64); cudaFree(ctx->F); cudaFree(ctx->sol); cudaFree(ctx->x1);
    for (auto &c : ctx->cloths) {
        cudaFree(c.f2v); cudaFree(c.cf); cudaFree(c.cp); cudaFree(c.side_deg); cudaFree(c.side_ovr); cudaFree(c.hinge_face); cudaFree(c.hinge_l);
        cudaFree(c.tri_slot); cudaFree(c.hinge_slot); cudaFree(c.norm_dir); cudaFree(c.q1);
    }
    cudaFree(ctx->faces); cudaFree(ctx->vn); cudaFree(ctx->proj_flag); cudaFree(ctx->proj_dir); cudaFree(ctx->proj_idx); cudaFree(ctx->proj_w);
    cudaFree(ctx->cell_key); cudaFree(ctx->cell_key_sorted); cudaFree(ctx->face_id); cudaFree(ctx->face_id_sorted); cudaFree(ctx->cub_tmp);
    cudaFree(ctx->cflag); cudaFree(ctx->cscan);
    cudaFree(ctx->con.idx); cudaFree(ctx->con.w); cudaFree(ctx->con.k); cudaFree(ctx->con.mu); cudaFree(ctx->con.dx0); cudaFree(ctx->con.T); cudaFree(ctx->con.n);
    cudaFree(ctx->ks); cudaFreeHost(ctx->ks_host); for (int q = 0; q < 2; q++) if (ctx->ks_ev[q]) cudaEventDestroy(ctx->ks_ev[q]); cudaFree(ctx->red_partial); cudaFree(ctx->red_ticket); cudaFree(ctx->red_out); cudaFreeHost(ctx->red_host);
    cudaFree(ctx->d_kb); cudaFree(ctx->adj_rhs); cudaFree(ctx->adj_z); cudaFree(ctx->error_flag); cudaFree(ctx->zero_border);
    for (auto &t : ctx->tets) { cudaFree(t.tets); cudaFree(t.B); cudaFree(t.W); cudaFree(t.slot); }
    cudaFree(ctx->nc_dev); cudaFree(ctx->cside32); cudaFree(ctx->cside64); cudaFree(ctx->yc); cudaFree(ctx->vgrav);
    cudaEventDestroy(ctx->ev_in); cudaEventDestroy(ctx->ev_out); cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TSL_OK;
}

const char *tsl_last_error(tsl_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int tsl_set_stream(tsl_ctx *ctx, void *s) { if (!ctx) return TSL_ERR_INVALID; ctx->user_stream = (cudaStream_t)s; return TSL_OK; }
long long tsl_launch_count(tsl_ctx *ctx) { return ctx ? ctx->launches : -1; }

// ---------------------------------------------------------------------------------------------- scene description
// Cloth.init_mesh (code/engine/model_fold_offset.py:929-1018): alternating-diagonal grid, neighbour tables with the
// reference's wiring (including the entries it never writes, which stay 0 -- quirk Q2).
static void build_cloth_mesh(int N, int M, std::vector<int> &f2v, std::vector<int> &cf, std::vector<int> &cp)
{
    int NF = 2 * N * M;
    f2v.assign(3 * NF, 0); cf.assign(3 * NF, 0); cp.assign(3 * NF, 0);
    auto F = [&](int f, int l) -> int & { return f2v[3 * f + l]; };
    auto CF = [&](int f, int l) -> int & { return cf[3 * f + l]; };
    auto CP = [&](int f, int l) -> int & { return cp[3 * f + l]; };
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

int tsl_add_cloth(tsl_ctx *ctx, int N, int M, int v_offset, double dx, double rho, double Kl, double Ka, double Kb, double k_angle,
                  double *ref_angle_dev)
{
    if (!ctx) return TSL_ERR_INVALID;
    REQUIRE(!ctx->finalized, "tsl_add_cloth after tsl_finalize");
    REQUIRE(N >= 1 && M >= 1 && 2 * N * M >= 3 && ref_angle_dev, "tsl_add_cloth: bad arguments");
    REQUIRE((int)ctx->cloths.size() < TSL_MAX_CLOTHS, "tsl_add_cloth: too many cloths");
    ClothDev c;
    memset(&c, 0, sizeof(c));
    c.N = N; c.M = M; c.NV = (N + 1) * (M + 1); c.NF = 2 * N * M; c.offset = v_offset;
    REQUIRE(v_offset >= 0 && v_offset + c.NV <= ctx->cfg.n_verts, "tsl_add_cloth: vertex range outside n_verts");
    c.P.dx = dx; c.P.dt = ctx->cfg.dt; c.P.mass = rho * dx * dx; c.P.Kl = Kl; c.P.Ka = Ka; c.P.Kb = Kb; c.P.k_angle = k_angle;
    c.ref_angle = ref_angle_dev;
    std::vector<int> f2v, cf, cp;
    build_cloth_mesh(N, M, f2v, cf, cp);
    // hinges: (face, slot) whose neighbour has the larger index (model_fold_offset.py:217 and everywhere else)
    std::vector<int> hf, hl;
    std::vector<unsigned char> deg(c.NF, 0);
    for (int i = 0; i < c.NF; i++)
        for (int l = 0; l < 3; l++) {
            int i2 = cf[3 * i + l];
            if (i2 > i) { hf.push_back(i); hl.push_back(l); }
            if (i2 != -1) {
                // side test of compute_angle / judge_angle uses vertices f2v[i][(l+1)%2], f2v[i][l]; if both lie in face i2
                // the test is exactly 0 in exact arithmetic
                int va = f2v[3 * i + (l + 1) % 2], vb = f2v[3 * i + l];
                bool ina = false, inb = false;
                for (int q = 0; q < 3; q++) { ina |= f2v[3 * i2 + q] == va; inb |= f2v[3 * i2 + q] == vb; }
                if (ina && inb) deg[i] |= (unsigned char)(1u << l);
            }
        }
    c.NH = (int)hf.size();
    TRY(upload(ctx, &c.f2v, f2v)); TRY(upload(ctx, &c.cf, cf)); TRY(upload(ctx, &c.cp, cp));
    TRY(upload(ctx, &c.side_deg, deg)); TRY(upload(ctx, &c.side_ovr, std::vector<unsigned char>(c.NF, 0))); TRY(upload(ctx, &c.hinge_face, hf)); TRY(upload(ctx, &c.hinge_l, hl));
    CK(cudaMalloc(&c.norm_dir, sizeof(double) * 3 * c.NF));
    CK(cudaMalloc(&c.q1, sizeof(double) * 90));
    CK(cudaMemset(c.q1, 0, sizeof(double) * 90));
    ctx->cloths.push_back(c);
    ctx->h_f2v.push_back(f2v); ctx->h_cf.push_back(cf); ctx->h_cp.push_back(cp);
    return (int)ctx->cloths.size() - 1;
}

int tsl_set_cloth_params(tsl_ctx *ctx, int cloth, double Kl, double Ka, double Kb, double k_angle)
{
    if (!ctx) return TSL_ERR_INVALID;
    REQUIRE(cloth >= 0 && cloth < (int)ctx->cloths.size(), "bad cloth id");
    ClothDev &c = ctx->cloths[cloth];
    c.P.Kl = Kl; c.P.Ka = Ka; c.P.Kb = Kb; c.P.k_angle = k_angle;
    return TSL_OK;
}

int tsl_cloth_update_ref_angle(tsl_ctx *ctx, int cloth)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    REQUIRE(cloth >= 0 && cloth < (int)ctx->cloths.size(), "bad cloth id");
    StreamScope scope_(ctx);
    launch_update_ref_angle(ctx, ctx->cloths[cloth]);
    CK(cudaGetLastError());
    return TSL_OK;
}

// test hook: outcomes of the topologically degenerate side tests.  ov_host [NF][3]: 1 = negative, anything else = not
// negative (the canonical rule); NULL restores the canonical rule everywhere.
int tsl_set_side_test_override(tsl_ctx *ctx, int cloth, const signed char *ov_host)
{
    if (!ctx) return TSL_ERR_INVALID;
    REQUIRE(cloth >= 0 && cloth < (int)ctx->cloths.size(), "bad cloth id");
    ClothDev &c = ctx->cloths[cloth];
    std::vector<unsigned char> bits(c.NF, 0);
    if (ov_host)
        for (int i = 0; i < c.NF; i++)
            for (int l = 0; l < 3; l++) if (ov_host[3 * i + l] == 1) bits[i] |= (unsigned char)(1u << l);
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(c.side_ovr, bits.data(), bits.size(), cudaMemcpyHostToDevice));
    return TSL_OK;
}
int tsl_get_cloth_topology(tsl_ctx *ctx, int cloth, int *f2v, int *cf, int *cp)
{
    if (!ctx) return TSL_ERR_INVALID;
    REQUIRE(cloth >= 0 && cloth < (int)ctx->cloths.size(), "bad cloth id");
    size_t n = ctx->h_f2v[cloth].size() * sizeof(int);
    if (f2v) memcpy(f2v, ctx->h_f2v[cloth].data(), n);
    if (cf) memcpy(cf, ctx->h_cf[cloth].data(), n);
    if (cp) memcpy(cp, ctx->h_cp[cloth].data(), n);
    return TSL_OK;
}

// Elastic(...) of model_elastic_offset.py (kind 0) / model_elastic_tactile.py (kind 1): cells, inverse rest Ds, rest volumes
int tsl_add_tets(tsl_ctx *ctx, int kind, int v_offset, int n_verts, int n_cells, const int *tets_host, const double *B_host,
                 const double *W_host, double mu, double lam, double alpha, const double *gravity_host)
{
    if (!ctx) return TSL_ERR_INVALID;
    REQUIRE(!ctx->finalized, "tsl_add_tets after tsl_finalize");
    REQUIRE((kind == 0 || kind == 1) && n_verts > 0 && n_cells > 0 && tets_host && B_host && W_host, "tsl_add_tets: bad arguments");
    REQUIRE(v_offset >= 0 && v_offset + n_verts <= ctx->cfg.n_verts, "tsl_add_tets: vertex range outside n_verts");
    REQUIRE((int)ctx->tets.size() < TSL_MAX_TETS, "tsl_add_tets: too many tetrahedral bodies");
    std::vector<int> tv(tets_host, tets_host + 4 * (size_t)n_cells);
    for (int q : tv) REQUIRE(q >= 0 && q < n_verts, "tsl_add_tets: cell vertex out of range");
    TetDev t;
    memset(&t, 0, sizeof(t));
    t.nv = n_verts; t.nc = n_cells; t.offset = v_offset;
    t.P.kind = kind; t.P.mu = mu; t.P.lam = lam; t.P.alpha = alpha;
    TRY(upload(ctx, &t.tets, tv));
    TRY(upload(ctx, &t.B, std::vector<double>(B_host, B_host + 9 * (size_t)n_cells)));
    TRY(upload(ctx, &t.W, std::vector<double>(W_host, W_host + (size_t)n_cells)));
    ctx->tets.push_back(t);
    ctx->h_tets.push_back(tv);
    for (int k = 0; k < 3; k++) ctx->h_tet_gravity.push_back(gravity_host ? gravity_host[k] : ctx->cfg.gravity[k]);
    return (int)ctx->tets.size() - 1;
}
int tsl_set_tet_params(tsl_ctx *ctx, int body, double mu, double lam)
{
    if (!ctx) return TSL_ERR_INVALID;
    REQUIRE(body >= 0 && body < (int)ctx->tets.size(), "bad tet body id");
    ctx->tets[body].P.mu = mu; ctx->tets[body].P.lam = lam;
    return TSL_OK;
}

int tsl_set_surfaces(tsl_ctx *ctx, const int *faces_host, int tot_nf, const int *bodies_host, int n_bodies)
{
    if (!ctx) return TSL_ERR_INVALID;
    REQUIRE(!ctx->finalized, "tsl_set_surfaces after tsl_finalize");
    REQUIRE(faces_host && bodies_host && tot_nf > 0 && n_bodies > 0, "tsl_set_surfaces: bad arguments");
    ctx->tot_nf = tot_nf;
    CK(cudaMalloc(&ctx->faces, sizeof(int) * 3 * tot_nf));
    CK(cudaMemcpy(ctx->faces, faces_host, sizeof(int) * 3 * tot_nf, cudaMemcpyHostToDevice));
    ctx->bodies.clear();
    for (int b = 0; b < n_bodies; b++) {
        SurfaceBody sb = { bodies_host[4 * b], bodies_host[4 * b + 1], bodies_host[4 * b + 2], bodies_host[4 * b + 3] };
        REQUIRE(sb.f_start >= 0 && sb.f_end <= tot_nf && sb.v_start >= 0 && sb.v_end <= ctx->cfg.n_verts, "tsl_set_surfaces: body range");
        ctx->bodies.push_back(sb);
    }
    return TSL_OK;
}

int tsl_add_contact_pair(tsl_ctx *ctx, int surface_body, int v_start, int v_end, double mu)
{
    if (!ctx) return TSL_ERR_INVALID;
    REQUIRE(surface_body >= 0 && surface_body < (int)ctx->bodies.size(), "tsl_add_contact_pair: bad body");
    REQUIRE(v_start >= 0 && v_end <= ctx->cfg.n_verts && v_start <= v_end, "tsl_add_contact_pair: bad range");
    ContactPair p = { surface_body, v_start, v_end, mu };
    ctx->pairs.push_back(p);
    return (int)ctx->pairs.size() - 1;
}
int tsl_set_contact_mu(tsl_ctx *ctx, int pair, double mu)
{
    if (!ctx) return TSL_ERR_INVALID;
    REQUIRE(pair >= 0 && pair < (int)ctx->pairs.size(), "bad pair id");
    ctx->pairs[pair].mu = mu;
    return TSL_OK;
}

int tsl_bind_state(tsl_ctx *ctx, double *pos, double *prev_pos, double *vel, const double *mass, const int *frozen, const int *border_flag)
{
    if (!ctx) return TSL_ERR_INVALID;
    REQUIRE(pos && prev_pos && vel && mass && frozen, "tsl_bind_state: null pointer");
    ctx->pos = pos; ctx->prev_pos = prev_pos; ctx->vel = vel; ctx->mass = mass; ctx->frozen = frozen;
    graphs_invalidate(ctx);                      // captured graphs hold the old pointers
    if (border_flag) ctx->border_flag = border_flag;
    else {
        if (!ctx->zero_border) {
            CK(cudaMalloc(&ctx->zero_border, sizeof(int) * ctx->cfg.n_verts));
            CK(cudaMemset(ctx->zero_border, 0, sizeof(int) * ctx->cfg.n_verts));
        }
        ctx->border_flag = ctx->zero_border;
    }
    return TSL_OK;
}

// Block pattern = vertex adjacency through triangles and hinges + the diagonal (contacts against frozen bodies only
// touch the diagonal).  Built once on the host, laid out as sliced ELL (see SellMatrix).
int tsl_finalize(tsl_ctx *ctx)
{
    if (!ctx) return TSL_ERR_INVALID;
    REQUIRE(!ctx->finalized, "tsl_finalize called twice");
    REQUIRE(ctx->pos, "tsl_finalize before tsl_bind_state");
    int nv = ctx->cfg.n_verts;
    SellMatrix &A = ctx->A;
    // ---- adjacency lists
    std::vector<int> cnt(nv + 1, 0);
    auto for_each_pair = [&](auto &&fn) {
        for (size_t ci = 0; ci < ctx->cloths.size(); ci++) {
            const ClothDev &c = ctx->cloths[ci];
            const std::vector<int> &f2v = ctx->h_f2v[ci], &cf = ctx->h_cf[ci], &cp = ctx->h_cp[ci];
            for (int i = 0; i < c.NF; i++) {
                int v[3] = { f2v[3 * i] + c.offset, f2v[3 * i + 1] + c.offset, f2v[3 * i + 2] + c.offset };
                for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) fn(v[a], v[b]);
                for (int l = 0; l < 3; l++) {
                    int i2 = cf[3 * i + l];
                    if (i2 > i) {
                        int h[4] = { v[l], v[(l + 1) % 3], v[(l + 2) % 3], f2v[3 * i2 + cp[3 * i + l]] + c.offset };
                        for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) fn(h[a], h[b]);
                    }
                }
            }
        }
        for (size_t ti = 0; ti < ctx->tets.size(); ti++) {
            const TetDev &t = ctx->tets[ti];
            const std::vector<int> &tv = ctx->h_tets[ti];
            for (int c = 0; c < t.nc; c++)
                for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) fn(tv[4 * c + a] + t.offset, tv[4 * c + b] + t.offset);
        }
        for (int v = 0; v < nv; v++) fn(v, v);
    };
    for_each_pair([&](int r, int) { cnt[r + 1]++; });
    std::vector<long long> start(nv + 1, 0);
    for (int v = 0; v < nv; v++) start[v + 1] = start[v] + cnt[v + 1];
    std::vector<int> cols((size_t)start[nv]);
    std::vector<long long> fill(start.begin(), start.end() - 1);
    for_each_pair([&](int r, int c) { cols[(size_t)fill[r]++] = c; });
    A.h_rowptr.assign(nv + 1, 0);
    A.h_colidx.clear();
    A.h_colidx.reserve((size_t)nv * 12);
    for (int v = 0; v < nv; v++) {
        auto b = cols.begin() + start[v], e = cols.begin() + start[v + 1];
        std::sort(b, e);
        e = std::unique(b, e);
        A.h_colidx.insert(A.h_colidx.end(), b, e);
        A.h_rowptr[v + 1] = (int)A.h_colidx.size();
    }
    std::vector<int>().swap(cols);
    A.n_rows = nv;
    A.nnzb = (int)A.h_colidx.size();
    A.n_slices = (nv + 31) / 32;
    // ---- sliced ELL
    A.h_slice_base.assign(A.n_slices + 1, 0);
    for (int S = 0; S < A.n_slices; S++) {
        int w = 0;
        for (int r = 32 * S; r < std::min(nv, 32 * S + 32); r++) w = std::max(w, A.h_rowptr[r + 1] - A.h_rowptr[r]);
        long long nb = (long long)A.h_slice_base[S] + 32LL * w;
        REQUIRE(nb < (1LL << 31) - 64, "matrix too large for 32-bit block ids");
        A.h_slice_base[S + 1] = (int)nb;
    }
    A.nnzb_pad = A.h_slice_base[A.n_slices];
    A.h_colidx_pad.assign((size_t)A.nnzb_pad, 0);
    std::vector<int> diag(nv, -1);
    for (int S = 0; S < A.n_slices; S++) {
        int w = (A.h_slice_base[S + 1] - A.h_slice_base[S]) / 32;
        for (int lane = 0; lane < 32; lane++) {
            int r = 32 * S + lane;
            for (int k = 0; k < w; k++) {
                int pb = A.h_slice_base[S] + k * 32 + lane;
                int col = (r < nv) ? r : 0;                       // padding: valid column, zero value
                if (r < nv && k < A.h_rowptr[r + 1] - A.h_rowptr[r]) col = A.h_colidx[A.h_rowptr[r] + k];
                A.h_colidx_pad[pb] = col;
                if (r < nv && col == r && diag[r] < 0 && k < A.h_rowptr[r + 1] - A.h_rowptr[r]) diag[r] = pb;
            }
        }
    }
    auto slot_of = [&](int r, int c) -> int {
        const int *b = A.h_colidx.data() + A.h_rowptr[r], *e = A.h_colidx.data() + A.h_rowptr[r + 1];
        const int *it = std::lower_bound(b, e, c);
        int k = (int)(it - b);
        return A.h_slice_base[r >> 5] + k * 32 + (r & 31);
    };
    for (size_t ci = 0; ci < ctx->cloths.size(); ci++) {
        ClothDev &c = ctx->cloths[ci];
        const std::vector<int> &f2v = ctx->h_f2v[ci], &cf = ctx->h_cf[ci], &cp = ctx->h_cp[ci];
        std::vector<int> ts((size_t)c.NF * 9), hs((size_t)c.NH * 16);
        int h = 0;
        for (int i = 0; i < c.NF; i++) {
            int v[3] = { f2v[3 * i] + c.offset, f2v[3 * i + 1] + c.offset, f2v[3 * i + 2] + c.offset };
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) ts[(size_t)i * 9 + a * 3 + b] = slot_of(v[a], v[b]);
            for (int l = 0; l < 3; l++) {
                int i2 = cf[3 * i + l];
                if (i2 > i) {
                    int hv[4] = { v[l], v[(l + 1) % 3], v[(l + 2) % 3], f2v[3 * i2 + cp[3 * i + l]] + c.offset };
                    for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) hs[(size_t)h * 16 + a * 4 + b] = slot_of(hv[a], hv[b]);
                    h++;
                }
            }
        }
        TRY(upload(ctx, &c.tri_slot, ts)); TRY(upload(ctx, &c.hinge_slot, hs));
    }
    for (size_t ti = 0; ti < ctx->tets.size(); ti++) {
        TetDev &t = ctx->tets[ti];
        const std::vector<int> &tv = ctx->h_tets[ti];
        std::vector<int> sl((size_t)t.nc * 16);
        for (int c = 0; c < t.nc; c++)
            for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) sl[(size_t)c * 16 + a * 4 + b] = slot_of(tv[4 * c + a] + t.offset, tv[4 * c + b] + t.offset);
        TRY(upload(ctx, &t.slot, sl));
    }
    TRY(upload(ctx, &A.slice_base, A.h_slice_base));
    TRY(upload(ctx, &A.colidx, A.h_colidx_pad));
    TRY(upload(ctx, &A.diag_pb, diag));
    CK(cudaMalloc(&A.val32, sizeof(float) * 9 * (size_t)A.nnzb_pad));
    CK(cudaMemset(A.val32, 0, sizeof(float) * 9 * (size_t)A.nnzb_pad));
    CK(cudaMalloc(&A.val32c, sizeof(float) * 9 * (size_t)A.nnzb_pad));
    CK(cudaMemset(A.val32c, 0, sizeof(float) * 9 * (size_t)A.nnzb_pad));
    CK(cudaMalloc(&A.val32t, sizeof(float) * 9 * (size_t)A.nnzb_pad));
    CK(cudaMemset(A.val32t, 0, sizeof(float) * 9 * (size_t)A.nnzb_pad));
    CK(cudaMalloc(&A.val32m, sizeof(float) * 9 * (size_t)A.nnzb_pad));
    CK(cudaMemset(A.val32m, 0, sizeof(float) * 9 * (size_t)A.nnzb_pad));
    // ---- scratch
    CK(cudaMalloc(&ctx->F, sizeof(double) * 3 * nv));
    CK(cudaMalloc(&ctx->x1, sizeof(double) * 3 * nv));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    ctx->red_blocks = sms * 4;                                  // persistent grid-stride reductions: 4 CTAs of 256 per SM
    CK(cudaMalloc(&ctx->red_partial, sizeof(double) * ctx->red_blocks));
    CK(cudaMalloc(&ctx->red_ticket, sizeof(unsigned int)));
    CK(cudaMemset(ctx->red_ticket, 0, sizeof(unsigned int)));
    CK(cudaMalloc(&ctx->red_out, sizeof(double) * 8));
    CK(cudaMallocHost(&ctx->red_host, sizeof(double) * 8));
    CK(cudaMalloc(&ctx->error_flag, sizeof(int)));
    CK(cudaMemset(ctx->error_flag, 0, sizeof(int)));
    // trailing rows whose three DOFs are frozen (the table of every bouncing-type scene) never enter the forward solve
    {
        std::vector<int> fz(3 * (size_t)nv);
        CK(cudaMemcpy(fz.data(), ctx->frozen, sizeof(int) * fz.size(), cudaMemcpyDeviceToHost));
        int last = -1;
        for (int v = 0; v < nv; v++) if (!(fz[3 * v] && fz[3 * v + 1] && fz[3 * v + 2])) last = v;
        ctx->n_solve = std::max(last + 1, 1);
        // a contact pair whose surface body has a free DOF needs all 16 blocks of its constraints (side buffer)
        ctx->general_contact = false;
        for (auto &p : ctx->pairs) {
            const SurfaceBody &sb = ctx->bodies[p.body];
            for (int v = sb.v_start; v < sb.v_end; v++)
                if (!(fz[3 * v] && fz[3 * v + 1] && fz[3 * v + 2])) ctx->general_contact = true;
        }
    }
    if (ctx->general_contact) {
        size_t m = (size_t)std::max(ctx->cfg.max_n_constraints, 1);
        CK(cudaMalloc(&ctx->nc_dev, sizeof(int)));
        CK(cudaMemset(ctx->nc_dev, 0, sizeof(int)));
        CK(cudaMalloc(&ctx->cside32, sizeof(float) * 108 * m));
        CK(cudaMemset(ctx->cside32, 0, sizeof(float) * 108 * m));
        CK(cudaMalloc(&ctx->yc, sizeof(double) * 3 * 32 * (size_t)A.n_slices));
        CK(cudaMemset(ctx->yc, 0, sizeof(double) * 3 * 32 * (size_t)A.n_slices));
    }
    {
        // bodies whose gravity differs from the scene's (effector pads carry none, BaseScene.init_property :371-374)
        bool differs = false;
        for (size_t ti = 0; ti < ctx->tets.size(); ti++)
            for (int k = 0; k < 3; k++) differs = differs || ctx->h_tet_gravity[3 * ti + k] != ctx->cfg.gravity[k];
        if (differs) {
            std::vector<double> vg(3 * (size_t)nv);
            for (int v = 0; v < nv; v++) for (int k = 0; k < 3; k++) vg[3 * (size_t)v + k] = ctx->cfg.gravity[k];
            for (size_t ti = 0; ti < ctx->tets.size(); ti++)
                for (int v = 0; v < ctx->tets[ti].nv; v++)
                    for (int k = 0; k < 3; k++) vg[3 * (size_t)(v + ctx->tets[ti].offset) + k] = ctx->h_tet_gravity[3 * ti + k];
            TRY(upload(ctx, &ctx->vgrav, vg));
        }
    }
    TRY(contact_alloc(ctx));
    TRY(linalg_alloc(ctx));
    TRY(mg_alloc(ctx));
    TRY(assembly_init(ctx));
    if (const char *e = getenv("TSL_PRECOND")) ctx->precond = atoi(e);
    if (const char *e = getenv("TSL_NEWTON_MODE")) ctx->newton_mode = atoi(e);
    if (const char *e = getenv("TSL_PROBE")) ctx->probe = atoi(e);
    if (const char *e = getenv("TSL_THETA_BACKOFF")) ctx->theta_backoff = atoi(e);
    if (const char *e = getenv("TSL_GMRES_M")) { int v = atoi(e); if (v >= 2 && v <= 400) ctx->gmres_m = v; }
    ctx->finalized = true;
    return TSL_OK;
}

int tsl_reset_contact_state(tsl_ctx *ctx)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    StreamScope scope_(ctx);
    size_t pb = (size_t)std::max<size_t>(ctx->bodies.size(), 1) * ctx->cfg.n_verts;
    CK(cudaMemsetAsync(ctx->proj_flag, 0, sizeof(int) * pb, ctx->stream));
    return TSL_OK;
}

// ---------------------------------------------------------------------------------------------- hot path
static int check_device_flags(tsl_ctx *ctx)
{
    int f = 0;
    CK(cudaMemcpyAsync(&f, ctx->error_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (f & 1) { ctx->err = "contact against a non-frozen triangle is not implemented in this build"; return TSL_ERR_UNSUPPORTED; }
    return TSL_OK;
}

int tsl_contact_detect(tsl_ctx *ctx, int *n_out)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    StreamScope scope_(ctx);
    TRY(contact_detect(ctx, ctx->pos, ctx->prev_pos));
    if (n_out) *n_out = ctx->nc;
    return TSL_OK;
}

static int energy_sync(tsl_ctx *ctx, double *out)
{
    launch_energy(ctx, ctx->pos, ctx->red_out);
    TRY(dist_allreduce(ctx, ctx->red_out, 1));               // strip partition: every rank summed the elements it owns
    CK(cudaMemcpyAsync(ctx->red_host, ctx->red_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out = ctx->red_host[0];
    return TSL_OK;
}
int tsl_energy(tsl_ctx *ctx, double *out)
{
    if (!ctx || !ctx->finalized || !out) return TSL_ERR_INVALID;
    StreamScope scope_(ctx);
    return energy_sync(ctx, out);
}

static int ensure_f64(tsl_ctx *ctx)
{
    if (ctx->A.val64) return TSL_OK;
    size_t nr = (size_t)ctx->A.n_slices * 32;
    CK(cudaMalloc(&ctx->A.val64, sizeof(double) * 9 * (size_t)ctx->A.nnzb_pad));
    CK(cudaMalloc(&ctx->minv64, sizeof(double) * 9 * nr));
    for (int i = 0; i < 8; i++) { CK(cudaMalloc(&ctx->bi[i], sizeof(double) * 3 * nr)); CK(cudaMemset(ctx->bi[i], 0, sizeof(double) * 3 * nr)); }
    CK(cudaMalloc(&ctx->d_kb, sizeof(double) * 3 * nr));
    CK(cudaMalloc(&ctx->adj_rhs, sizeof(double) * 3 * nr));
    CK(cudaMalloc(&ctx->adj_z, sizeof(double) * 3 * nr));
    if (ctx->general_contact) {
        size_t m = (size_t)std::max(ctx->cfg.max_n_constraints, 1);
        CK(cudaMalloc(&ctx->cside64, sizeof(double) * 108 * m));
        CK(cudaMemset(ctx->cside64, 0, sizeof(double) * 108 * m));
    }
    return TSL_OK;
}

int tsl_assemble(tsl_ctx *ctx, int flags)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    StreamScope scope_(ctx);
    if (flags & TSL_ASM_RESIDUAL) launch_residual(ctx, ctx->pos);
    if (flags & TSL_ASM_HESSIAN) {
        bool f64 = (flags & TSL_ASM_F64) != 0;
        if (f64) TRY(ensure_f64(ctx));
        if (!f64 && (flags & TSL_ASM_NEWTON)) {
            // both forward Newton matrices leave one pass: A_e -> val32, A_c -> val32c; TSL_ASM_SPD asks for the clamped one in val32
            launch_hessian_newton_pair(ctx, ctx->pos);
            if (flags & TSL_ASM_SPD)
                CK(cudaMemcpyAsync(ctx->A.val32, ctx->A.val32c, sizeof(float) * 9 * (size_t)ctx->A.nnzb_pad, cudaMemcpyDeviceToDevice, ctx->stream));
        } else {
            launch_hessian(ctx, ctx->pos, f64, (flags & TSL_ASM_SPD) ? 1 : 0, (flags & TSL_ASM_SYM) ? 1 : 0, (flags & TSL_ASM_NEWTON) ? 1 : 0);
            // the preconditioner hierarchy always comes from the clamped (positive definite) Newton matrix at the same state
            launch_hessian(ctx, ctx->pos, false, 1, 0, 1, true);
        }
        TRY(mg_setup_replay(ctx));
        if (f64) launch_block_jacobi64(ctx);
        ctx->last_f64 = f64;
    }
    CK(cudaGetLastError());
    return TSL_OK;
}

// Adjoint solve H z = b on the reference's un-projected, non-symmetric fp64 Hessian.  The reference factorises (sparse QR,
// code/engine/sparse_solver.py:85-105), so its gradient never hinges on an iteration converging; here
//   * up to direct_max_dof unknowns (every task scene of the reference): dense LU with partial pivoting + iterative refinement
//     (tsl_dense.cu) -- exact, deterministic, st->iters = 0;
//   * above: FGMRES(m) with the multigrid V-cycle as flexible right preconditioner (monotone residual, true-residual restarts);
//     if that stalls, the same with the fp64 block-Jacobi preconditioner (flags bit3).
// An unconverged result is an error (TSL_ERR_NUMERIC), never a silently wrong gradient.
static int adjoint_mode(tsl_ctx *ctx)
{
    int mode = ctx->adjoint_solver;
    if (mode == 0) mode = (3LL * ctx->n_solve <= ctx->direct_max_dof) ? 1 : 2;
    return mode;
}
static int solve_adjoint64(tsl_ctx *ctx, const double *rhs, double *x, double rel_tol, int max_iters, tsl_solve_stats *st)
{
    tsl_solve_stats s0;
    memset(&s0, 0, sizeof(s0));
    const int mode = adjoint_mode(ctx);
    if (mode == 1) {
        TRY(solve_dense64(ctx, rhs, x, &s0));
    } else {
        const bool mg = ctx->precond != 0 && ctx->mg.n_levels > 0;
        auto run = [&](tsl_solve_stats *o) {
            return mode == 3 ? solve_bicgstab64(ctx, rhs, x, rel_tol, mg ? std::min(max_iters, 800) : max_iters, o)
                             : solve_fgmres64(ctx, rhs, x, rel_tol, max_iters, o);
        };
        TRY(run(&s0));
        if (mg && ((s0.flags & 3) || !(s0.rel_residual <= 10 * rel_tol)) && s0.iters < max_iters) {
            int first_iters = s0.iters;
            int saved = ctx->precond;
            ctx->precond = 0;
            int rc = mode == 3 ? solve_bicgstab64(ctx, rhs, x, rel_tol, max_iters - first_iters, &s0)
                               : solve_fgmres64(ctx, rhs, x, rel_tol, max_iters - first_iters, &s0);
            ctx->precond = saved;
            if (rc != TSL_OK) return rc;
            s0.iters += first_iters;
            s0.flags |= 8;
        }
    }
    if (st) *st = s0;
    if ((s0.flags & 3) || !(s0.rel_residual <= std::max(10 * rel_tol, mode == 1 ? 1e-9 : 1e-13))) {
        char buf[160];
        snprintf(buf, sizeof(buf), "adjoint solve did not converge: %d iterations, flags %d, relative residual %.3e (tolerance %.1e)", s0.iters, s0.flags,
                 s0.rel_residual, rel_tol);
        ctx->err = buf;
        return TSL_ERR_NUMERIC;
    }
    return TSL_OK;
}

int tsl_solve(tsl_ctx *ctx, const double *rhs, double *x, double rel_tol, int max_iters, tsl_solve_stats *st)
{
    if (!ctx || !ctx->finalized || !rhs || !x) return TSL_ERR_INVALID;
    StreamScope scope_(ctx);
    if (ctx->last_f64 && ctx->dist.on && ctx->dist.world > 1) { ctx->err = "the fp64 BiCGStab solve is not partitioned over GPUs in this build"; return TSL_ERR_UNSUPPORTED; }
    if (ctx->last_f64) return solve_adjoint64(ctx, rhs, x, rel_tol, max_iters, st);
    return solve_pcg32(ctx, ctx->A.val32, rhs, x, rel_tol, max_iters, st);
}

// BaseScene.time_step (code/engine/BaseScene.py:1327-1370) + newton_step (:1159-1230)
//
// Newton iteration of the B200 path.  The residual is the reference's exact fp64 gradient, so a
// converged state is a fixed point of the reference's iteration; matrix model, fp32 storage and Krylov tolerance only
// shape the path:
//   * A_e = exact membrane Hessian + Gauss-Newton bending (fp32): the operator of every regular solve (quadratic
//     convergence near the minimiser);
//   * A_c = the same with the indefinite pieces clamped (positive definite): source of the multigrid hierarchy, which is
//     rebuilt only every few iterations (a stale preconditioner costs Krylov iterations, never accuracy), and operator
//     of the last-resort solve;
//   * PCG meets negative curvature (buckling sheet): the step is x_k, the last iterate before it, if that lowers the
//     energy, plus a move along the direction of negative curvature p_k whose length is found by doubling on the energy;
//     the next iteration is usually positive definite again (measured: 4x fewer Newton iterations than falling back to
//     A_c, same minimiser);
//   * forcing term: Eisenstat-Walker choice 2 clipped to [1e-3, 0.1];
//   * line search: the reference's halving on E < E0 (floor 1e-8, x left at the last trial, quirk Q9), plus doubling
//     while the energy keeps falling when alpha = 1 was accepted, plus acceptance of the plain Newton step when the
//     predicted decrease is below the resolution of the fp64 energy sum.
int tsl_step_forward(tsl_ctx *ctx, int max_newton, double tol, tsl_step_stats *stats)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    StreamScope scope_(ctx);
    int n3 = 3 * ctx->cfg.n_verts;
    cudaStream_t s = ctx->stream;
    tsl_step_stats st;
    memset(&st, 0, sizeof(st));
    const bool trace = getenv("TSL_TRACE") != nullptr && atoi(getenv("TSL_TRACE")) != 0;
    double t0 = now_ms();
    // timestep_init: prev_pos <- pos
    CK(cudaMemcpyAsync(ctx->prev_pos, ctx->pos, sizeof(double) * n3, cudaMemcpyDeviceToDevice, s));
    TRY(contact_detect(ctx, ctx->pos, ctx->prev_pos));
    st.n_contacts = ctx->nc;
    st.ms_contact = now_ms() - t0;
    double E0 = 0;
    TRY(energy_sync(ctx, &E0));
    double dt = ctx->cfg.dt;
    int it = 0;
    const bool mgp = ctx->precond != 0 && ctx->mg.n_levels > 0;
    const int max_pcg = (mgp && ctx->tets.empty()) ? 200 : 4000;   // tetrahedral rows only see the level-0 smoother
    // The hierarchy may be kept only while it is demonstrably as good as a fresh one: small Newton steps (the matrix barely
    // moves: the long tail of a buckling step), Krylov count within 3 of what the fresh hierarchy needed, at most 8
    // iterations old.  (Measured: a setup costs ~4.5 PCG iterations at 1M triangles; keeping a hierarchy through the
    // first iterations of a step, while contacts and strains still change, costs 15.)
    const int refresh_every = 8;
    auto keep_hierarchy = [&](int age_, int last_, int fresh_, double delta_) {
        return age_ < refresh_every && it > 1 && delta_ < 2e-3 && last_ <= fresh_ + 3;
    };
    const double len_scale = ctx->cloths.empty() ? 1e-3 : ctx->cloths[0].P.dx;
    double eta = 0.1, fnorm_prev = -1;
    double theta = 0.0, theta_used = 0.0;     // newton_mode 2: blend factor of the next / the last solve
    int fails_prop = 0, hold = 0, backoff = 2; double floor_theta = 0.0;   // newton_mode 2: back-off of the lower-theta proposals (see below)
    bool have_ncdir = false;                  // ctx->ncdir holds a direction of negative curvature met in this step
    int skip = 0, back = 0;                   // newton_mode 0: exact attempts skipped after a failure (1, 3, 7, 8, ...)
    int age = refresh_every;                  // iterations since the hierarchy was built (forces a build at it == 1)
    int last_pcg = 0, fresh_pcg = 0;
    while (it < max_newton) {
        it++;
        t0 = now_ms();
        launch_residual(ctx, ctx->pos);
        tsl_solve_stats ss;
        double t1;
        bool fallback = false;
        std::string attempts;                 // trace only: theta:iterations of every solve attempt (x = negative curvature)
        if (ctx->newton_mode == 0) {
            // ---- projected-Newton fallback with back-off (the path closest to the reference's own iteration)
            const bool try_exact = (skip == 0);
            launch_hessian_newton_pair(ctx, ctx->pos);                               // A_e -> val32, A_c -> val32c
            // hierarchy: rebuilt every few iterations, or as soon as the Krylov count drifts away from what a fresh one gave
            if (!keep_hierarchy(age, last_pcg, fresh_pcg, st.delta)) {
                TRY(mg_setup_replay(ctx));
                age = 0;
            }
            ctx->last_f64 = false;
            if (it == 1) TRY(check_device_flags(ctx));
            t1 = now_ms();
            st.ms_assembly += t1 - t0;
            if (try_exact) {
                TRY(solve_pcg32(ctx, ctx->A.val32, ctx->F, ctx->sol, eta, max_pcg, &ss));
                st.linear_iters += ss.iters;
                if (ss.flags & 1) {
                    st.flags |= 1;
                    back = std::min(8, 2 * back + 1);
                    skip = back;
                    fallback = true;
                    TRY(solve_pcg32(ctx, ctx->A.val32c, ctx->F, ctx->sol, eta, max_pcg, &ss));
                    st.linear_iters += ss.iters;
                } else back = 0;
            } else {
                skip--;
                fallback = true;
                TRY(solve_pcg32(ctx, ctx->A.val32c, ctx->F, ctx->sol, eta, max_pcg, &ss));
                st.linear_iters += ss.iters;
            }
            last_pcg = ss.iters;
            if (age == 0) fresh_pcg = ss.iters;
            age++;
        } else if (ctx->newton_mode == 2) {
            // ---- blended operator A_theta = A_e + theta (A_c - A_e): the smallest theta in {0, 1/16, 1/8, ..., 1} for which PCG
            // meets no negative curvature (the clamped matrix over-stiffens every compressed element, the blend only as
            // much as definiteness needs)
            launch_hessian_newton_pair(ctx, ctx->pos);                               // A_e -> val32, A_c -> val32c
            if (!keep_hierarchy(age, last_pcg, fresh_pcg, st.delta)) {
                TRY(mg_setup_replay(ctx));
                age = 0;
            }
            ctx->last_f64 = false;
            if (it == 1) TRY(check_device_flags(ctx));
            t1 = now_ms();
            st.ms_assembly += t1 - t0;
            if (have_ncdir && ctx->probe) {
                // curvature probe: the direction of negative curvature met last usually still is one; two matrix passes tell,
                // for every theta, whether the blend is indefinite along it -- those solves are not attempted
                double pAe = 0, pAc = 0;
                TRY(probe_curvature(ctx, ctx->A.val32, ctx->ncdir, &pAe));
                TRY(probe_curvature(ctx, ctx->A.val32c, ctx->ncdir, &pAc));
                while (theta < 1.0 && (1.0 - theta) * pAe + theta * pAc <= 0.0) theta = std::min(1.0, std::max(2.0 * theta, 1.0 / 16));
            }
            // The driver proposes half the last successful theta (down to 0 = the exact matrix).  On a buckling step that proposal fails for
            // dozens of Newton iterations in a row, and every failed attempt costs the 15-60 PCG iterations it takes to meet the negative
            // curvature (measured at 1 M triangles: 70 % of the PCG iterations of the impact steps were spent in such attempts).  After two
            // consecutive failed proposals theta is therefore HELD at the value that worked for `backoff` iterations (2, 4, ... 16) before
            // the next lower proposal is tried; one successful proposal resets the back-off.
            const bool holding = ctx->theta_backoff && hold > 0;
            if (holding) { theta = std::max(theta, floor_theta); hold--; }
            const double theta_first = theta;
            while (true) {
                const float *op = ctx->A.val32;
                if (theta >= 1.0) op = ctx->A.val32c;
                else if (theta > 0.0) { launch_blend(ctx, ctx->A.val32, ctx->A.val32c, (float)theta, ctx->A.val32t); op = ctx->A.val32t; }
                TRY(solve_pcg32(ctx, op, ctx->F, ctx->sol, eta, max_pcg, &ss));
                st.linear_iters += ss.iters;
                if (trace) attempts += " " + std::to_string(theta).substr(0, 6) + ":" + std::to_string(ss.iters) + ((ss.flags & 1) ? "x" : "");
                if (!(ss.flags & 1) || theta >= 1.0) break;
                st.flags |= 1;
                CK(cudaMemcpyAsync(ctx->ncdir, ctx->cg_p, sizeof(double) * 3 * (size_t)ctx->n_solve, cudaMemcpyDeviceToDevice, s));
                have_ncdir = true;
                theta = std::min(1.0, std::max(2.0 * theta, 1.0 / 16));
            }
            if (ctx->theta_backoff) {
                if (holding) { if (theta > floor_theta) floor_theta = theta; }          // even the held value failed: hold the one that worked
                else if (theta > theta_first) {                                           // a proposal failed
                    if (++fails_prop >= 2) { floor_theta = theta; hold = backoff; backoff = std::min(2 * backoff, 16); }
                } else { fails_prop = 0; backoff = 2; floor_theta = 0.0; }                // a proposal worked
            }
            fallback = theta > 0.0;
            theta_used = theta;
            theta = theta > 1.0 / 16 ? 0.5 * theta : 0.0;
            last_pcg = ss.iters;
            if (age == 0) fresh_pcg = ss.iters;
            age++;
        } else {
        launch_hessian_newton_pair(ctx, ctx->pos);                               // A_e -> val32, A_c -> val32c
        // hierarchy: rebuilt every few iterations, or as soon as the Krylov count drifts away from what a fresh one gave
        if (!keep_hierarchy(age, last_pcg, fresh_pcg, st.delta)) {
            TRY(mg_setup_replay(ctx));
            age = 0;
        }
        ctx->last_f64 = false;
        if (it == 1) TRY(check_device_flags(ctx));
        t1 = now_ms();
        st.ms_assembly += t1 - t0;
        TRY(solve_pcg32(ctx, ctx->A.val32, ctx->F, ctx->sol, eta, max_pcg, &ss));
        st.linear_iters += ss.iters;
        last_pcg = ss.iters;
        if (age == 0) fresh_pcg = ss.iters;
        age++;
        }
        double fnorm = ctx->ks_host->rr0;
        double eta_used = eta;
        if (fnorm_prev > 0 && fnorm > 0) eta = std::min(0.1, std::max(1e-3, 0.9 * (fnorm / fnorm_prev) * (fnorm / fnorm_prev)));
        fnorm_prev = fnorm;
        CK(cudaMemcpyAsync(ctx->x1, ctx->pos, sizeof(double) * n3, cudaMemcpyDeviceToDevice, s));
        if (ctx->newton_mode != 0 && (ss.flags & 1) && ctx->dist.on) {
            ctx->err = "strip partition: the negative-curvature move of the Newton driver is not partitioned (use Newton mode 0 or 2)";
            return TSL_ERR_UNSUPPORTED;
        }
        if (ctx->newton_mode != 0 && (ss.flags & 1)) {
            // ---- negative curvature at PCG iteration ss.iters: x_k is in sol, p_k in cg_p
            st.flags |= 1;
            double t2 = now_ms();
            st.ms_solve += t2 - t1;
            const bool have_base = ss.iters > 1;
            double Eb = E0, base = 0.0;
            if (have_base) {
                launch_axpy_pos(ctx, ctx->x1, ctx->sol, 1.0, ctx->pos);
                TRY(energy_sync(ctx, &Eb));
                st.linesearch_evals++;
                if (Eb < E0) base = 1.0; else Eb = E0;
            }
            // orientation and scale of p_k: descent means F . p > 0
            launch_absmax(ctx, ctx->cg_p, n3, ctx->red_out + 1);
            launch_dot(ctx, ctx->F, ctx->cg_p, n3, ctx->red_out + 3);
            CK(cudaMemcpyAsync(ctx->red_host + 1, ctx->red_out + 1, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            double pmax = ctx->red_host[1], sgn = ctx->red_host[3] >= 0 ? 1.0 : -1.0;
            double tbest = 0.0, Ebest = Eb;
            if (pmax > 0 && pmax == pmax) {
                double tau = 1e-4 * len_scale / pmax;
                for (int k = 0; k < 40; k++) {
                    double E = 0;
                    launch_axpy2_pos(ctx, ctx->x1, ctx->sol, base, ctx->cg_p, sgn * tau, ctx->pos);
                    TRY(energy_sync(ctx, &E));
                    st.linesearch_evals++;
                    if (E < Ebest) { Ebest = E; tbest = tau; tau *= 2; } else break;
                }
            }
            st.ms_linesearch += now_ms() - t2;
            if (tbest > 0 || base > 0) {
                launch_axpy2_pos(ctx, ctx->x1, ctx->sol, base, ctx->cg_p, sgn * tbest, ctx->pos);
                if (trace)
                    fprintf(stderr, "[tsl] newton %3d: negcurv@pcg=%d base=%g move=%.2e dE=%.3e |F|=%.3e E=%.12e nc=%d\n", it, ss.iters, base,
                            tbest * pmax, E0 - Ebest, fnorm, Ebest, ctx->nc);
                E0 = Ebest;
                st.delta = 1.0;               // not a Newton step: no convergence test on it
                continue;
            }
            // neither the partial solve nor the curvature direction lowers the energy: positive definite fallback
            fallback = true;
            t1 = now_ms();
            if (age > 1) {                    // the stored A_c is stale as an OPERATOR: rebuild it (and the hierarchy) for this state
                launch_hessian_newton_pair(ctx, ctx->x1);
                TRY(mg_setup_replay(ctx));
                age = 1;
            }
            launch_axpy_pos(ctx, ctx->x1, ctx->sol, 0.0, ctx->pos);
            TRY(solve_pcg32(ctx, ctx->A.val32c, ctx->F, ctx->sol, eta_used, max_pcg, &ss));
            st.linear_iters += ss.iters;
        }
        st.flags |= (ss.flags & 2);
        launch_absmax(ctx, ctx->sol, n3, ctx->red_out + 1);
        launch_dot(ctx, ctx->F, ctx->sol, n3, ctx->red_out + 3);   // (F is zero on ghost rows: owned DOFs only)
        TRY(dist_allreduce(ctx, ctx->red_out + 1, 1, true));
        TRY(dist_allreduce(ctx, ctx->red_out + 3, 1));
        CK(cudaMemcpyAsync(ctx->red_host + 1, ctx->red_out + 1, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        double p_norm = ctx->red_host[1];
        double decrement = ctx->red_host[3];      // F . p: first-order energy decrease of the full step
        double t2 = now_ms();
        st.ms_solve += t2 - t1;
        if (!(p_norm == p_norm)) { ctx->err = "NaN in Newton direction"; return TSL_ERR_NUMERIC; }
        // backtracking on E < E0, floor 1e-8, x stays at the last trial (quirk Q9)
        double alpha = 1.0, E = E0;
        while (alpha > 1e-8) {
            launch_axpy_pos(ctx, ctx->x1, ctx->sol, alpha, ctx->pos);
            TRY(energy_sync(ctx, &E));
            st.linesearch_evals++;
            if (E < E0) break;
            // below the resolution of the fp64 energy sum the comparison is noise: take the plain Newton step
            if (alpha == 1.0 && decrement >= 0 && decrement <= 1e-10 * fabs(E0)) { st.flags |= 4; break; }
            alpha /= 2;
        }
        if (alpha == 1.0 && E < E0) {
            // extrapolation: keep doubling while the energy falls
            while (alpha < 64.0) {
                double E2 = 0;
                launch_axpy_pos(ctx, ctx->x1, ctx->sol, 2 * alpha, ctx->pos);
                TRY(energy_sync(ctx, &E2));
                st.linesearch_evals++;
                if (E2 < E) { alpha *= 2; E = E2; }
                else { launch_axpy_pos(ctx, ctx->x1, ctx->sol, alpha, ctx->pos); break; }
            }
        }
        st.ms_linesearch += now_ms() - t2;
        st.delta = p_norm / dt;
        if (trace)
            fprintf(stderr, "[tsl] newton %3d: %s pcg=%d age=%d |F|=%.3e eta=%.1e delta=%.3e alpha=%.3g E=%.12e nc=%d%s\n", it,
                    fallback ? (ctx->newton_mode == 2 ? (std::string("theta=") + std::to_string(theta_used)).c_str() : "clamped-fallback") : "exact", ss.iters, age,
                    fnorm, eta_used, st.delta, alpha, E, ctx->nc, attempts.c_str());
        E0 = E;                               // the reference re-evaluates the same point at the top of the loop
        if (st.delta < tol) { st.converged = 1; break; }
    }
    st.newton_iters = it;
    st.energy = E0;
    // Scene_bouncing.timestep_finish: update_vel, update_ref_angle
    launch_update_vel(ctx);
    for (auto &c : ctx->cloths) launch_update_ref_angle(ctx, c);
    CK(cudaGetLastError());
    if (stats) *stats = st;
    return TSL_OK;
}

int tsl_step_forward_host(tsl_ctx *ctx, double *pos_host, double *vel_host, int max_newton, double tol, tsl_step_stats *stats)
{
    if (!ctx || !ctx->finalized || !pos_host || !vel_host) return TSL_ERR_INVALID;
    StreamScope scope_(ctx);
    size_t nb = sizeof(double) * 3 * (size_t)ctx->cfg.n_verts;
    CK(cudaMemcpyAsync(ctx->pos, pos_host, nb, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->vel, vel_host, nb, cudaMemcpyHostToDevice, ctx->stream));
    TRY(tsl_step_forward(ctx, max_newton, tol, stats));
    CK(cudaMemcpyAsync(pos_host, ctx->pos, nb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(vel_host, ctx->vel, nb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TSL_OK;
}

// Grad.transfer_grad: analytic_grad_system.py:115-160 (system identification, grad_kb) and analytic_grad_single.py:217-255
// (trajectory optimisation: the same recurrence plus tmp_z_frozen, the sensitivity to the kinematically driven vertices)
int tsl_step_backward_ex(tsl_ctx *ctx, const double *x_t, const double *x_tm1, const double *ref_angle_tm1,
                         double *pg_t, double *pg_tm1, double *pg_tm2, double *ag_t, double *ag_tm1,
                         double *grad_kb_accum, double *z_out, double *z_frozen_out, double clamp, double clamp_angleref, double rel_tol,
                         int max_iters, tsl_solve_stats *stats)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    StreamScope scope_(ctx);
    REQUIRE(x_t && x_tm1 && ref_angle_tm1 && pg_t && pg_tm1 && ag_t && ag_tm1, "tsl_step_backward: null pointer");
    REQUIRE(!ctx->cloths.empty(), "tsl_step_backward needs a cloth");
    if (ctx->dist.on && ctx->dist.world > 1) { ctx->err = "the adjoint step is not partitioned over GPUs in this build"; return TSL_ERR_UNSUPPORTED; }
    TRY(ensure_f64(ctx));
    int nv = ctx->cfg.n_verts, n3 = 3 * nv;
    cudaStream_t s = ctx->stream;
    // several cloths (Scene_card, Scene_sliding): ref_angle_tm1 / ag_t / ag_tm1 hold the cloths one after the other, [NF_c][3] each
    size_t nb = sizeof(double) * n3;
    int nf_all = 0;
    for (auto &c : ctx->cloths) nf_all += c.NF;
    // clamp_grad(step)
    launch_clamp(ctx, pg_t, n3, clamp);
    if (clamp_angleref > 0) launch_clamp(ctx, ag_t, 3 * nf_all, clamp_angleref);   // analytic_grad_single.py:182-185
    // copy_pos_only(step-1): pos = prev_pos = x_{t-1}; contact re-detection there (quirk Q7)
    CK(cudaMemcpyAsync(ctx->pos, x_tm1, nb, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(ctx->prev_pos, x_tm1, nb, cudaMemcpyDeviceToDevice, s));
    TRY(contact_detect(ctx, ctx->pos, ctx->prev_pos));
    // copy_pos_and_refangle(step): pos = x_t, prev_pos = x_{t-1}, ref_angle = ref_angle[t-1]
    CK(cudaMemcpyAsync(ctx->pos, x_t, nb, cudaMemcpyDeviceToDevice, s));
    {
        size_t fo = 0;
        for (size_t ci = 0; ci < ctx->cloths.size(); ci++) {
            ClothDev &c = ctx->cloths[ci];
            CK(cudaMemcpyAsync(c.ref_angle, ref_angle_tm1 + fo, sizeof(double) * 3 * (size_t)c.NF, cudaMemcpyDeviceToDevice, s));
            // ref_angle_backprop_a2ax: plastic rest-angle adjoint feeds pos_grad[t] before the solve
            launch_refangle_a2ax(ctx, c, ctx->pos, ag_t + fo, ag_tm1 + fo, pg_t);
            // get_paramters_grad: d_kb = dF/dKb (every cloth adds its rows: Scene_card.get_paramters_grad)
            if (grad_kb_accum) launch_cloth_param_deri(ctx, c, ctx->pos, ctx->d_kb, ci == 0);
            fo += 3 * (size_t)c.NF;
        }
    }
    // H = reference Hessian without projection, fp64
    launch_hessian(ctx, ctx->pos, true, 0, 0, 0);
    // preconditioner of the iterative path: multigrid hierarchy of the clamped Newton matrix at x_t (the direct path needs none)
    if (adjoint_mode(ctx) != 1) {
        launch_hessian_newton_pair(ctx, ctx->pos);
        TRY(mg_setup_replay(ctx));
    }
    launch_block_jacobi64(ctx);
    ctx->last_f64 = true;
    TRY(check_device_flags(ctx));
    double *z = z_out ? z_out : ctx->adj_z;
    TRY(solve_adjoint64(ctx, pg_t, z, rel_tol, max_iters, stats));
    if (z_frozen_out) {
        // second assembly with counting_z_frozen: tmp_z_frozen[j] = -sum_{i free} H[i][j] z[i] for frozen j
        CK(cudaMemsetAsync(z_frozen_out, 0, nb, s));
        launch_hessian_counting(ctx, ctx->pos, z, z_frozen_out);
    }
    // friction lag terms and rest-angle terms into step t-1, then the time recurrence and dL/dKb
    launch_contact_backprop(ctx, ctx->pos, z, pg_tm1);
    {
        size_t fo = 0;
        for (auto &c : ctx->cloths) { launch_refangle_x2a(ctx, c, ctx->pos, z, ag_tm1 + fo); fo += 3 * (size_t)c.NF; }
    }
    launch_adjoint_tail(ctx, z, grad_kb_accum ? ctx->d_kb : nullptr, pg_tm1, pg_tm2, grad_kb_accum);
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return TSL_OK;
}
int tsl_step_backward(tsl_ctx *ctx, const double *x_t, const double *x_tm1, const double *ref_angle_tm1,
                      double *pg_t, double *pg_tm1, double *pg_tm2, const double *ag_t, double *ag_tm1,
                      double *grad_kb_accum, double *z_out, double clamp, double rel_tol, int max_iters, tsl_solve_stats *stats)
{
    if (!ctx) return TSL_ERR_INVALID;
    REQUIRE(grad_kb_accum, "tsl_step_backward: null pointer");
    return tsl_step_backward_ex(ctx, x_t, x_tm1, ref_angle_tm1, pg_t, pg_tm1, pg_tm2, const_cast<double *>(ag_t), ag_tm1, grad_kb_accum, z_out,
                                nullptr, clamp, 0.0, rel_tol, max_iters, stats);
}

// BaseScene.get_paramters_grad (elastic part, code/engine/BaseScene.py:1523-1525) + Grad.get_parameters_grad
// (code/engine/analytic_grad_system.py:69-75) at the bound positions
int tsl_elastic_param_grad(tsl_ctx *ctx, const double *z_dev, double *d_mu_dev, double *d_lam_dev, double *out2_host)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    REQUIRE(d_mu_dev && d_lam_dev, "tsl_elastic_param_grad: null pointer");
    REQUIRE(!z_dev || out2_host, "tsl_elastic_param_grad: out2_host missing");
    StreamScope scope_(ctx);
    launch_tets_param_grad(ctx, ctx->pos, z_dev, d_mu_dev, d_lam_dev, ctx->red_out + 2);
    if (z_dev) {
        CK(cudaMemcpyAsync(ctx->red_host + 2, ctx->red_out + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        out2_host[0] = ctx->red_host[2]; out2_host[1] = ctx->red_host[3];
    }
    CK(cudaGetLastError());
    return TSL_OK;
}

int tsl_cloth_param_deri(tsl_ctx *ctx, int cloth, double *d_kl_dev, double *d_ka_dev, double *d_kb_dev)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    REQUIRE(cloth >= 0 && cloth < (int)ctx->cloths.size(), "bad cloth id");
    StreamScope scope_(ctx);
    launch_cloth_deri(ctx, ctx->cloths[cloth], ctx->pos, d_kl_dev, d_ka_dev, d_kb_dev);
    CK(cudaGetLastError());
    return TSL_OK;
}
int tsl_friction_coef_grad(tsl_ctx *ctx, const double *z_dev, int pair_begin, int pair_end, double *out_host)
{
    if (!ctx || !ctx->finalized || !z_dev || !out_host) return TSL_ERR_INVALID;
    REQUIRE(pair_begin >= 0 && pair_begin <= pair_end && pair_end <= (int)ctx->pairs.size(), "tsl_friction_coef_grad: bad pair range");
    REQUIRE(ctx->pair_start.size() == ctx->pairs.size() + 1, "tsl_friction_coef_grad: no contact set (call after a step or tsl_contact_detect)");
    StreamScope scope_(ctx);
    int c0 = ctx->pair_start[pair_begin], c1 = ctx->pair_start[pair_end];
    *out_host = 0.0;
    if (c1 > c0) {
        launch_friction_coef_grad(ctx, ctx->pos, z_dev, c0, c1, ctx->red_out + 2);
        CK(cudaMemcpyAsync(ctx->red_host + 2, ctx->red_out + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        *out_host = ctx->red_host[2];
    }
    CK(cudaGetLastError());
    return TSL_OK;
}
int tsl_elastic_force(tsl_ctx *ctx, int body, double *Ff_dev)
{
    if (!ctx || !ctx->finalized || !Ff_dev) return TSL_ERR_INVALID;
    REQUIRE(body >= 0 && body < (int)ctx->tets.size(), "bad tet body id");
    StreamScope scope_(ctx);
    launch_elastic_force(ctx, body, ctx->pos, Ff_dev);
    CK(cudaGetLastError());
    return TSL_OK;
}

// ---------------------------------------------------------------------------------------------- kinematic boundary (gripper)
// gripper.get_vert_pos + update_bound + pushup (code/engine/gripper_single.py:79-83, 157-161; Scene_folding.action :213-224):
// pos[v_offset + bound_idx[i]] = p + R F_x[bound_idx[i]], R in fp32 as the reference stores it (rotmat is an f32 field)
__global__ void k_gripper_apply(int n_bound, const int *__restrict__ bound_idx, const double *__restrict__ Fx, int v_offset,
                                double px, double py, double pz, float r0, float r1, float r2, float r3, float r4, float r5, float r6, float r7, float r8,
                                double *pos)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bound) return;
    int b = bound_idx[i];
    double x = Fx[3 * b], y = Fx[3 * b + 1], z = Fx[3 * b + 2];
    double *o = pos + 3 * (size_t)(v_offset + b);
    o[0] = px + ((double)r0 * x + (double)r1 * y + (double)r2 * z);
    o[1] = py + ((double)r3 * x + (double)r4 * y + (double)r5 * z);
    o[2] = pz + ((double)r6 * x + (double)r7 * y + (double)r8 * z);
}
// gripper.gather_grad (code/engine/gripper_single.py:133-150): one block, deterministic tree sum over the bound vertices
__global__ void __launch_bounds__(256) k_gripper_gather(int n_bound, const int *__restrict__ bound_idx, const double *__restrict__ Fx, int v_offset,
                                                        const double *__restrict__ zf, float r0, float r1, float r2, float r3, float r4, float r5,
                                                        float r6, float r7, float r8, double clamp_pos, double clamp_angle, double *out6)
{
    __shared__ double sh[6][256];
    double acc[6] = { 0, 0, 0, 0, 0, 0 };
    for (int i = threadIdx.x; i < n_bound; i += blockDim.x) {
        int b = bound_idx[i];
        const double *g = zf + 3 * (size_t)(v_offset + b);
        double x = Fx[3 * b], y = Fx[3 * b + 1], z = Fx[3 * b + 2];
        double rx = (double)r0 * x + (double)r1 * y + (double)r2 * z, ry = (double)r3 * x + (double)r4 * y + (double)r5 * z,
               rz = (double)r6 * x + (double)r7 * y + (double)r8 * z;
        acc[0] += g[0]; acc[1] += g[1]; acc[2] += g[2];
        acc[3] += ry * g[2] - rz * g[1]; acc[4] += rz * g[0] - rx * g[2]; acc[5] += rx * g[1] - ry * g[0];
    }
    for (int k = 0; k < 6; k++) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) for (int k = 0; k < 6; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x < 6) {
        double v = sh[threadIdx.x][0] / (1.0 * n_bound);
        double lim = threadIdx.x < 3 ? clamp_pos : clamp_angle;
        out6[threadIdx.x] = fmin(fmax(v, -lim), lim);
    }
}
int tsl_gripper_apply(tsl_ctx *ctx, int v_offset, int n_bound, const int *bound_idx_dev, const double *Fx_dev, const double *pos3_host,
                      const float *R)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    REQUIRE(n_bound > 0 && bound_idx_dev && Fx_dev && pos3_host && R, "tsl_gripper_apply: bad arguments");
    StreamScope scope_(ctx);
    k_gripper_apply<<<(n_bound + 127) / 128, 128, 0, ctx->stream>>>(n_bound, bound_idx_dev, Fx_dev, v_offset, pos3_host[0], pos3_host[1], pos3_host[2],
                                                                   R[0], R[1], R[2], R[3], R[4], R[5], R[6], R[7], R[8], ctx->pos);
    ctx->launches++;
    CK(cudaGetLastError());
    return TSL_OK;
}
int tsl_gripper_gather(tsl_ctx *ctx, const double *z_frozen_dev, int v_offset, int n_bound, const int *bound_idx_dev, const double *Fx_dev,
                       const float *R, double clamp_pos, double clamp_angle, double *out6_host)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    REQUIRE(n_bound > 0 && z_frozen_dev && bound_idx_dev && Fx_dev && R && out6_host, "tsl_gripper_gather: bad arguments");
    StreamScope scope_(ctx);
    k_gripper_gather<<<1, 256, 0, ctx->stream>>>(n_bound, bound_idx_dev, Fx_dev, v_offset, z_frozen_dev, R[0], R[1], R[2], R[3], R[4], R[5], R[6],
                                                 R[7], R[8], clamp_pos, clamp_angle, ctx->red_out + 2);
    ctx->launches++;
    CK(cudaMemcpyAsync(ctx->red_host + 2, ctx->red_out + 2, 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 6; k++) out6_host[k] = ctx->red_host[2 + k];
    return TSL_OK;
}

// ---------------------------------------------------------------------------------------------- introspection
int tsl_get_residual(tsl_ctx *ctx, double *F_host)
{
    if (!ctx || !ctx->finalized || !F_host) return TSL_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaMemcpy(F_host, ctx->F, sizeof(double) * 3 * ctx->cfg.n_verts, cudaMemcpyDeviceToHost));
    return TSL_OK;
}
int tsl_get_matrix_nnzb(tsl_ctx *ctx, int *n) { if (!ctx || !ctx->finalized || !n) return TSL_ERR_INVALID; *n = ctx->A.nnzb; return TSL_OK; }
int tsl_get_matrix(tsl_ctx *ctx, int *rowptr, int *colidx, double *val)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    const SellMatrix &A = ctx->A;
    CK(cudaStreamSynchronize(ctx->stream));
    if (rowptr) memcpy(rowptr, A.h_rowptr.data(), sizeof(int) * (A.n_rows + 1));
    if (colidx) memcpy(colidx, A.h_colidx.data(), sizeof(int) * A.nnzb);
    if (val) {
        size_t n = 9 * (size_t)A.nnzb_pad;
        std::vector<double> pad(n);
        if (ctx->last_f64) CK(cudaMemcpy(pad.data(), A.val64, sizeof(double) * n, cudaMemcpyDeviceToHost));
        else {
            std::vector<float> t(n);
            CK(cudaMemcpy(t.data(), A.val32, sizeof(float) * n, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n; i++) pad[i] = t[i];
        }
        for (int r = 0; r < A.n_rows; r++)
            for (int k = 0; k < A.h_rowptr[r + 1] - A.h_rowptr[r]; k++) {
                long long pb = (long long)A.h_slice_base[r >> 5] + k * 32 + (r & 31);
                for (int cc = 0; cc < 9; cc++) val[9 * (size_t)(A.h_rowptr[r] + k) + cc] = pad[(size_t)sell_addr(pb, r & 31, cc)];
            }
    }
    return TSL_OK;
}
int tsl_get_projection(tsl_ctx *ctx, int body, int *flag, int *dir, int *idx, double *w)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    REQUIRE(body >= 0 && body < (int)ctx->bodies.size(), "bad body");
    size_t nv = ctx->cfg.n_verts, off = (size_t)body * nv;
    CK(cudaStreamSynchronize(ctx->stream));
    if (flag) CK(cudaMemcpy(flag, ctx->proj_flag + off, sizeof(int) * nv, cudaMemcpyDeviceToHost));
    if (dir) CK(cudaMemcpy(dir, ctx->proj_dir + off, sizeof(int) * nv, cudaMemcpyDeviceToHost));
    if (idx) CK(cudaMemcpy(idx, ctx->proj_idx + 3 * off, sizeof(int) * 3 * nv, cudaMemcpyDeviceToHost));
    if (w) CK(cudaMemcpy(w, ctx->proj_w + 3 * off, sizeof(double) * 3 * nv, cudaMemcpyDeviceToHost));
    return TSL_OK;
}
int tsl_get_constraints(tsl_ctx *ctx, int *n_out, int *idx, double *w, double *k, double *dx0, double *T, double *n)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    int nc = ctx->nc;
    if (n_out) *n_out = nc;
    if (nc == 0) return TSL_OK;
    if (idx) CK(cudaMemcpy(idx, ctx->con.idx, sizeof(int) * 4 * nc, cudaMemcpyDeviceToHost));
    if (w) CK(cudaMemcpy(w, ctx->con.w, sizeof(double) * 3 * nc, cudaMemcpyDeviceToHost));
    if (k) CK(cudaMemcpy(k, ctx->con.k, sizeof(double) * nc, cudaMemcpyDeviceToHost));
    if (dx0) CK(cudaMemcpy(dx0, ctx->con.dx0, sizeof(double) * 3 * nc, cudaMemcpyDeviceToHost));
    if (T) CK(cudaMemcpy(T, ctx->con.T, sizeof(double) * 6 * nc, cudaMemcpyDeviceToHost));
    if (n) CK(cudaMemcpy(n, ctx->con.n, sizeof(double) * 3 * nc, cudaMemcpyDeviceToHost));
    return TSL_OK;
}
int tsl_get_contact_blocks(tsl_ctx *ctx, int *n_out, int *rows, int *cols, double *val)
{
    if (!ctx || !ctx->finalized || !n_out) return TSL_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    int nc = ctx->general_contact ? ctx->nc : 0;
    *n_out = 12 * nc;
    if (nc == 0 || !rows || !cols || !val) return TSL_OK;
    std::vector<int> idx(4 * (size_t)nc);
    CK(cudaMemcpy(idx.data(), ctx->con.idx, sizeof(int) * idx.size(), cudaMemcpyDeviceToHost));
    size_t n = 108 * (size_t)nc;
    if (ctx->last_f64) CK(cudaMemcpy(val, ctx->cside64, sizeof(double) * n, cudaMemcpyDeviceToHost));
    else {
        std::vector<float> t(n);
        CK(cudaMemcpy(t.data(), ctx->cside32, sizeof(float) * n, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; i++) val[i] = t[i];
    }
    for (int i = 0; i < nc; i++)
        for (int a = 0; a < 4; a++)
            for (int k = 0; k < 3; k++) {
                int b = k < a ? k : k + 1;
                rows[12 * (size_t)i + a * 3 + k] = idx[4 * i + a];
                cols[12 * (size_t)i + a * 3 + k] = idx[4 * i + b];
            }
    return TSL_OK;
}
int tsl_get_sizes(tsl_ctx *ctx, tsl_sizes *out)
{
    if (!ctx || !ctx->finalized || !out) return TSL_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    out->n_verts = ctx->cfg.n_verts;
    for (auto &c : ctx->cloths) { out->n_tris += c.NF; out->n_hinges += c.NH; }
    out->nnzb = ctx->A.nnzb; out->nnzb_padded = (int)ctx->A.nnzb_pad; out->n_contacts = ctx->nc;
    out->bytes_matrix_f32 = (long long)ctx->A.nnzb_pad * 40; out->bytes_matrix_f64 = (long long)ctx->A.nnzb_pad * 76;
    out->n_solve = ctx->n_solve; out->nnzb_solve = ctx->A.h_rowptr[ctx->n_solve];
    return TSL_OK;
}

int tsl_bench_kernel(tsl_ctx *ctx, int what, int iters, float *ms_out)
{
    if (!ctx || !ctx->finalized || !ms_out || iters <= 0) return TSL_ERR_INVALID;
    StreamScope scope_(ctx);
    if (what == 0 || what == 1 || what == 5 || what == 6) return bench_pcg_iterations(ctx, iters, what, ms_out);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < iters; i++) {
        if (what == 2) launch_energy(ctx, ctx->pos, ctx->red_out);
        else if (what == 3) launch_residual(ctx, ctx->pos);
        else if (what == 4) launch_hessian_newton_pair(ctx, ctx->pos);       // BOTH forward Newton matrices (A_e, A_c)
        else if (what == 7) launch_hessian(ctx, ctx->pos, false, 0, 0, 1);   // one matrix through the scatter (atomics) kernels
        else { ctx->err = "tsl_bench_kernel: unknown kernel class"; return TSL_ERR_INVALID; }
    }
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return TSL_OK;
}

int tsl_set_option(tsl_ctx *ctx, int key, double value)
{
    if (!ctx) return TSL_ERR_INVALID;
    switch (key) {
    case TSL_OPT_PRECOND: ctx->precond = (int)value; break;
    case TSL_OPT_MG_DEGREE: REQUIRE(value >= 1 && value <= TSL_MG_MAX_DEGREE, "mg degree out of range"); ctx->mg.degree = (int)value; break;
    case TSL_OPT_MG_COARSE_DEGREE: REQUIRE(value >= 1 && value <= TSL_MG_MAX_DEGREE, "mg coarse degree out of range"); ctx->mg.coarse_degree = (int)value; break;
    case TSL_OPT_MG_RATIO: REQUIRE(value > 1, "mg ratio must exceed 1"); ctx->mg.ratio = (float)value; break;
    case TSL_OPT_MG_SAFETY: REQUIRE(value >= 1, "mg safety must be >= 1"); ctx->mg.safety = (float)value; break;
    case TSL_OPT_GRAPHS: ctx->use_graphs = (int)value; break;
    case TSL_OPT_NEWTON_MODE: ctx->newton_mode = (int)value; break;
    case TSL_OPT_ADJOINT_SOLVER: REQUIRE(value >= 0 && value <= 3, "adjoint solver: 0 auto, 1 dense LU, 2 FGMRES, 3 BiCGStab"); ctx->adjoint_solver = (int)value; break;
    case TSL_OPT_DIRECT_MAX_DOF: REQUIRE(value >= 0 && value <= 46000, "direct_max_dof out of range"); ctx->direct_max_dof = (int)value; break;
    case TSL_OPT_FAST_ASSEMBLY:
        if (value != 0 && !ctx->finalized) { ctx->err = "TSL_OPT_FAST_ASSEMBLY: set after tsl_finalize"; return TSL_ERR_INVALID; }
        if (value != 0) { TRY(assembly_init(ctx)); if (ctx->fast_assembly) ctx->fast_assembly = (int)value; } else ctx->fast_assembly = 0;
        break;
    case TSL_OPT_GMRES_M: REQUIRE(value >= 2 && value <= 400, "FGMRES restart length out of range"); ctx->gmres_m = (int)value; break;
    default: ctx->err = "tsl_set_option: unknown key"; return TSL_ERR_INVALID;
    }
    cudaStreamSynchronize(ctx->stream);
    graphs_invalidate(ctx);
    return TSL_OK;
}
int tsl_mg_get_level(tsl_ctx *ctx, int level, int *dims, float *lmax, float *val_host)
{
    if (!ctx || !ctx->finalized) return TSL_ERR_INVALID;
    return mg_get_level(ctx, level, dims, lmax, val_host);
}
int tsl_dense_solve_host(tsl_ctx *ctx, int n, const double *A_host, const double *b_host, double *x_host)
{
    if (!ctx || !ctx->finalized || n <= 0 || !A_host || !b_host || !x_host) return TSL_ERR_INVALID;
    StreamScope scope_(ctx);
    return dense_solve_host(ctx, n, A_host, b_host, x_host);
}
int tsl_precond_apply(tsl_ctx *ctx, const double *b_dev, double *z_dev)
{
    if (!ctx || !ctx->finalized || !b_dev || !z_dev) return TSL_ERR_INVALID;
    StreamScope scope_(ctx);
    // fp64 boundary, fp32 cycle: staged through the solver's own vectors
    return precond_apply_f64io(ctx, b_dev, z_dev);
}

}  // extern "C"
