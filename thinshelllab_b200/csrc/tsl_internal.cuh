// tsl_internal.cuh -- context and device-side data structures of libtsl (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/tsl.h"
#include "tsl_elements.cuh"
#include "tsl_solids.cuh"

namespace tsl {

// ------------------------------------------------------------------------------------------------
// Block-sparse matrix in a sliced-ELL layout tuned for one-thread-per-block-row SpMV:
//   rows are grouped in slices of 32 (one warp); slice S has width W_S = max blocks per row in it;
//   padded block id  pb = base[S] + k*32 + lane           (k-th block of row 32*S + lane)
//   value address    (pb - lane)*9 + c*32 + lane           (component c = 3*r + col of the 3x3 block)
// so for a fixed (S, k, c) the 32 lanes of a warp read 32 consecutive values: every load of the SpMV
// is a full 128-byte (fp32) / 256-byte (fp64) line.  Padding blocks carry value 0 and a valid column.
struct SellMatrix {
    int n_rows = 0;         // block rows (vertices)
    int n_slices = 0;
    int nnzb = 0;           // real blocks
    long long nnzb_pad = 0; // padded blocks
    int *slice_base = nullptr;  // [n_slices+1] device, in blocks
    int *colidx = nullptr;      // [nnzb_pad] device
    int *diag_pb = nullptr;     // [n_rows] device: padded block id of the diagonal block
    float *val32 = nullptr;     // [nnzb_pad*9]  operator of the forward solve (exact or clamped Newton matrix)
    float *val32c = nullptr;    // [nnzb_pad*9]  clamped (positive definite) Newton matrix: fallback operator, source of the hierarchy
    float *val32t = nullptr;    // [nnzb_pad*9]  blended operator val32 + theta (val32c - val32) (Newton mode 2)
    void *val16m = nullptr;     // [nnzb_pad*9]  fp16 snapshot (x MgDev::scale[0]) of val32c for the level-0 smoother (MgDev::use_half)
    float *val32m = nullptr;    // [nnzb_pad*9]  snapshot of val32c the current multigrid hierarchy was built from (level-0 smoother matrix)
    double *val64 = nullptr;    // [nnzb_pad*9], allocated on first fp64 assembly
    // host copies (pattern export, slot lookup at setup)
    std::vector<int> h_rowptr, h_colidx, h_slice_base, h_colidx_pad;
};
__host__ __device__ __forceinline__ long long sell_addr(long long pb, int lane, int c) { return (pb - lane) * 9 + (long long)c * 32 + lane; }

struct ClothDev {
    int N, M, NV, NF, NH, offset;
    ClothParams P;
    int *f2v, *cf, *cp;          // [NF][3]
    unsigned char *side_deg;     // [NF] bit l set: the side test of (face, l) is topologically degenerate
    unsigned char *side_ovr;     // [NF] bit l: outcome ("negative") of a degenerate side test; all zero = the canonical rule.
                                 // Test hook (tsl_set_side_test_override): lets a parity test inject the rounding-noise signs of a reference run
    int *hinge_face, *hinge_l;   // [NH] hinges = (i, l) with cf[i][l] > i
    int *tri_slot;               // [NF][9]  padded block ids of the (l, m) blocks of a triangle
    int *hinge_slot;             // [NH][16] padded block ids of the (j, k) blocks of a hinge
    double *ref_angle;           // [NF][3] borrowed
    double *norm_dir;            // [NF][3] scratch: unit normals
    double *q1;                  // [9 + 81] c_i rows and mat_N of faces 0..2 (quirk Q1)
};

#define TSL_MAX_CLOTHS 4
struct ClothSet { int n; ClothDev c[TSL_MAX_CLOTHS]; };      // by-value kernel argument (Scene_card / Scene_sliding stack three cloths)

// tetrahedral body (Elastic of model_elastic_offset.py / model_elastic_tactile.py): vertex range [offset, offset + nv)
#define TSL_MAX_TETS 8
struct TetDev {
    int nv, nc, offset;
    TetParams P;
    int *tets;                   // [nc][4] body-local vertex ids
    double *B, *W;               // [nc][9] inverse rest Ds, [nc] rest volume
    int *slot;                   // [nc][16] padded block ids of the (a, b) blocks of a cell
};
struct TetSet { int n; TetDev b[TSL_MAX_TETS]; };

// where Hessian blocks go: the matrix values, or -- in the "counting" pass of the adjoint (BaseScene.add_H with
// counting_z_frozen, BaseScene.py:399-405) -- tmp_z_frozen[j] -= H[i][j] z[i] for free i, frozen j
template <typename T>
struct Sink { T *val; const double *z; double *zf; };

struct SurfaceBody { int v_start, v_end, f_start, f_end; };
struct ContactPair { int body, v_start, v_end; double mu; };

struct ContactDev {
    int max_nc;
    int *idx;          // [max_nc][4]
    double *w;         // [max_nc][3]
    double *k, *mu;    // [max_nc]
    double *dx0;       // [max_nc][3]
    double *T;         // [max_nc][6]
    double *n;         // [max_nc][3]
};

// ------------------------------------------------------------------------------------------------
// Geometric multigrid over the cloth's structured vertex grid (tsl_mg.cu): the preconditioner of the forward PCG and
// of the adjoint BiCGStab.  Level 0 is the sliced-ELL matrix itself (all vertices); level l >= 1 is an n0 x n1 vertex
// grid whose operator is the Galerkin product P^T A P (bilinear P) stored as a 5x5 stencil of 3x3 blocks, SoA:
//   element e = slot*9 + comp of vertex v,  slot = (dI+2)*5 + (dJ+2),  v = I*n1 + J,  at val[v*sv + e*se]
// (row-major + one warp per vertex on small levels, element-major + one thread per vertex on large ones).
#define TSL_MG_MAX_LEVELS 12
#define TSL_MG_MAX_DEGREE 8
struct MgLevel {
    int n0 = 0, n1 = 0, nv = 0, nvp = 0;   // grid, vertices, vertices padded to 32
    long long sv = 225, se = 1;            // layout of val: element (v, e) at val[v*sv + e*se] (row-major small levels, element-major large)
    int nrows = 0;                         // rows of the level's vectors (level 0: all matrix rows; else nv)
    float *val = nullptr;                  // stencil operator [225][nvp]  (level 0: stencil copy of the cloth block, Galerkin input only)
    void *val16 = nullptr;                 // the same operator in fp16 x the level's scale (large levels when MgDev::use_half): what the smoother reads
    int half = 0;                          // 1: the level's operator lives in val16 (level 0: SellMatrix::val16m)
    float *dinv = nullptr;                 // [nrows][9] inverse diagonal blocks
    float *x[2] = { nullptr, nullptr };    // iterate ping-pong [3 nrows]
    float *b = nullptr, *r = nullptr, *d = nullptr;   // right-hand side, residual, Chebyshev direction
    float *pv[2] = { nullptr, nullptr };   // power-iteration vectors (kept across setups: warm start)
};
struct MgDev {
    int n_levels = 0;                      // 0 = multigrid unavailable (no cloth)
    int cloth_offset = 0;
    MgLevel lev[TSL_MG_MAX_LEVELS];
    float *coef = nullptr;                 // device [levels][TSL_MG_MAX_DEGREE][2]: Chebyshev (a, c) per step
    float *powc = nullptr;                 // device [levels][2][2]: (a, c) of the first / later power-iteration steps
    double *pow_acc = nullptr;             // device [levels][16]: |v_k|^2 of the power iteration
    float *lmax = nullptr;                 // device [levels]: estimate (diagnostic read-back)
    int setups = 0;
    int tiled_galerkin = 1;                // Galerkin products by shared-memory tiles (TSL_MG_TILED=0: one thread per coarse entry)
    int pair_threads = 1;                  // element-major levels: 2 threads per vertex (TSL_MG_PAIR=0: one)
    int tail_level = -1;                   // first level of the fused tail of the V-cycle (-1: none)
    // coarsest level solved exactly: dense inverse of its (<= TSL_MG_DIRECT_MAX unknowns) operator, rebuilt with the hierarchy, applied by
    // one matrix-vector kernel instead of coarse_degree Chebyshev launches (TSL_MG_DIRECT=0: the sweep)
    int coarse_direct = 1;
    float *coarse_inv = nullptr;           // device [n][n], n = 3 x vertices of the coarsest level
    int tail_cluster = 0;                  // thread blocks of the cluster that runs the tail (1: one block; 0: tail off)
    // side streams of the hierarchy build: the eigenvalue iteration of level l only needs that level's operator, so it runs beside
    // the Galerkin chain that is still producing the coarser levels (TSL_MG_FORK=0: everything on the context's stream)
    // fp16 storage of the operators only the preconditioner reads (level-0 snapshot, element-major coarse levels): TSL_MG_HALF=0 keeps fp32
    int use_half = 1;
    int sell_split = 4;                    // warps per 32-row slice in the level-0 smoother kernels (TSL_MG_SELL_SPLIT = 1 / 2 / 4)
    float *scale = nullptr;                // device [levels][2]: {s, 1 / s}, stored value = true value x s (power of two; fp32 levels: 1)
    unsigned int *maxdiag = nullptr;       // device: bits of the largest diagonal entry of the level-0 matrix (>= every |entry| of a PSD matrix)
    int fork = 1;
    cudaStream_t side[TSL_MG_MAX_LEVELS] = {};
    cudaEvent_t ev_ready[TSL_MG_MAX_LEVELS] = {}, ev_done[TSL_MG_MAX_LEVELS] = {};
    int degree = 2, coarse_degree = 8;
    float ratio = 8.f, coarse_ratio = 200.f, safety = 1.2f;
};

// Krylov scalars kept on the device so a solve never needs the host inside an iteration.  There is no parity /
// iteration index in any kernel argument: a one-thread "rotate" kernel at the end of every iteration moves the
// freshly accumulated sums into place, so one captured CUDA graph serves every iteration.
struct KrylovScalars {
    // PCG (pq | rz_new, rr_new adjacent: the strip-partitioned solve all-reduces them as one message)
    double pq;            // p.Ap of the current iteration
    double rz_new, rr_new;   // r.z and |r|^2 after the update
    double rz, rr;        // r.z and |r|^2 before the update
    // BiCGStab
    double rho, rho_old, rho_next;   // rhat.r of this / the previous / the next iteration
    double rhv, ts, tt;              // rhat.v, t.s, t.t
    double alpha, omega;             // step lengths of the previous iteration
    double rr0;           // |b|^2 (host side only)
    int flags;            // bit0 negative curvature / breakdown (iteration frozen)
    int iter;             // iterations completed
};

// Strip partition of the cloth grid over the ranks of one node: this context holds the
// grid rows [first owned row - ghost_lo, last owned row + ghost_hi] of a longer sheet; rows are contiguous in vertex numbering.
struct DistCtx {
    bool on = false;
    int rank = 0, world = 1;
    void *comm = nullptr;          // ncclComm_t
    int row_len = 0;               // vertices per grid row (M + 1)
    int ghost_lo = 0, ghost_hi = 0;   // ghost rows below / above the owned rows
    int own0 = 0, own1 = 0x7fffffff;  // owned cloth vertices [own0, own1) (local ids)
    int nvc = 0;                   // cloth vertices of this context
    long long halo_msgs = 0, allreduces = 0;
};

struct GraphSlot { cudaGraphExec_t exec = nullptr; long long launches = 0; const void *key = nullptr; };

// dense fp64 LU of the adjoint matrix (tsl_dense.cu): column-major A[i + j * lda], allocated on first use
struct DenseLU { double *A = nullptr; int *ipiv = nullptr; int *info = nullptr; int cap = 0, lda = 0; };

}  // namespace tsl

struct tsl_ctx {
    tsl_config cfg;
    std::string err;
    cudaStream_t stream = 0;                     // the library's own stream: all work runs here (graph capture needs a real stream)
    cudaStream_t user_stream = 0;                // the caller's stream (tsl_set_stream); ordered against `stream` with events
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    int use_graphs = 1;                          // TSL_GRAPHS=0 disables CUDA-graph replay of the solver iterations
    tsl::GraphSlot g_pcg[3], g_bicg, g_mgsetup, g_pcg_start;  // captured iteration bodies (PCG per operator array)
    long long launches = 0;
    bool finalized = false;

    // bound state
    double *pos = nullptr, *prev_pos = nullptr, *vel = nullptr;
    const double *mass = nullptr;
    const int *frozen = nullptr;
    const int *border_flag = nullptr;
    int *zero_border = nullptr;

    std::vector<tsl::ClothDev> cloths;
    std::vector<std::vector<int>> h_f2v, h_cf, h_cp;

    // surfaces / contact candidates
    int tot_nf = 0;
    int *faces = nullptr;                       // [tot_nf][3]
    std::vector<tsl::SurfaceBody> bodies;
    std::vector<tsl::ContactPair> pairs;
    std::vector<int> pair_start;                 // [pairs + 1] first constraint of every pair in the current set (contact_detect)
    double *vn = nullptr;                       // [n_verts][3]
    int *proj_flag = nullptr, *proj_dir = nullptr, *proj_idx = nullptr;   // [n_bodies][n_verts]([3])
    double *proj_w = nullptr;
    unsigned int *cell_key = nullptr, *cell_key_sorted = nullptr;         // per face of the current surface body
    int *face_id = nullptr, *face_id_sorted = nullptr;
    void *cub_tmp = nullptr; size_t cub_tmp_bytes = 0;
    int *cflag = nullptr, *cscan = nullptr;     // [n_verts] compaction scratch
    tsl::ContactDev con;
    int nc = 0;
    // contact against triangles that move (cloth faces, pads, ball): the 12 off-diagonal blocks of every constraint live in a
    // side buffer applied after the sliced-ELL pass (the pattern of the ELL matrix is static); diagonal blocks go into the ELL
    bool general_contact = false;                // some contact pair's surface body has a free vertex (tsl_finalize)
    int *nc_dev = nullptr;                       // [1] current constraint count for the graph-replayed side pass
    float *cside32 = nullptr; double *cside64 = nullptr;   // [max_nc][12][9]
    double *yc = nullptr;                        // [3 n_rows_pad] side-pass accumulator, consumed and cleared by the next matrix pass

    std::vector<tsl::TetDev> tets;
    std::vector<std::vector<int>> h_tets;
    double *vgrav = nullptr;                     // [n_verts][3] per-vertex gravity when a body's differs from cfg.gravity, else null
    std::vector<double> h_tet_gravity;           // [n_tets][3]

    // linear system
    tsl::SellMatrix A;
    bool last_f64 = false;
    double *F = nullptr;                         // [3 n_verts] residual
    float *minv32 = nullptr; double *minv64 = nullptr;   // [n_verts][9] block-Jacobi inverse
    double *cg_x = nullptr, *cg_r = nullptr, *cg_p = nullptr, *cg_q = nullptr;   // PCG vectors [3 n_rows_pad] (fp64)
    double *ncdir = nullptr;                     // last direction of negative curvature PCG met (curvature probe, Newton mode 2)
    float *cg_r32 = nullptr, *cg_z = nullptr;     // fp32 copy of r (preconditioner input) and z = M r
    double *bi[8] = { nullptr };                 // BiCGStab vectors: r, rhat, p, v, y, s, z, t
    double *sol = nullptr;                       // [3 n_verts] Newton direction (f64)
    double *x1 = nullptr;                        // [n_verts][3] line-search base
    int n_solve = 0;                             // rows [n_solve, n_verts) are fully frozen (decoupled, zero residual): the forward
                                                 // solve and the multigrid cycle skip them (tsl_finalize)
    tsl::MgDev mg;
    int precond = 1;                             // 0 block-Jacobi, 1 multigrid V-cycle
    int probe = 0;                               // Newton mode 2: curvature probe before the solves (TSL_PROBE=1; measured: no gain)
    int theta_backoff = 1;   // newton_mode 2: leave theta = 0 out for a while after repeated negative-curvature failures (TSL_THETA_BACKOFF=0: always try)
    int newton_mode = 2;                         // 0 projected-Newton fallback, 1 negative-curvature moves, 2 blended operator
    float *cg_r64tmp = nullptr;                  // [3 n_rows_pad] fp32 staging of fp64 vectors for the V-cycle
    tsl::KrylovScalars *ks = nullptr;            // device
    tsl::KrylovScalars *ks_host = nullptr;       // pinned [3]: [0] the solver's view, [1..2] read-back slots of the pipelined PCG checks
    cudaEvent_t ks_ev[2] = { nullptr, nullptr };
    int pcg_pipeline = 1;                        // check iteration k while k + 1 is already queued (TSL_PCG_PIPELINE=0: synchronise per iteration)
    // reductions
    double *red_partial = nullptr; unsigned int *red_ticket = nullptr; double *red_out = nullptr; double *red_host = nullptr;
    int red_blocks = 0;
    // adjoint scratch
    double *d_kb = nullptr;                      // [n_verts][3]
    double *adj_rhs = nullptr, *adj_z = nullptr; // [3 n_verts]
    int error_flag_host = 0; int *error_flag = nullptr;   // device-side "unsupported" flags
    // adjoint solve: dense LU below direct_max_dof unknowns, FGMRES(gmres_m) above
    int adjoint_solver = 0;                      // 0 auto, 1 dense LU, 2 FGMRES, 3 BiCGStab (round-1 solver, kept for comparison)
    int direct_max_dof = 12288;
    int gmres_m = 50, gm_cap_m = 0;
    tsl::DenseLU dense;
    double *gm_V = nullptr, *gm_Z = nullptr;     // FGMRES bases [gmres_m + 1] / [gmres_m] x [3 n_rows_pad]
    double *gm_h = nullptr, *gm_h_host = nullptr;   // Gram-Schmidt coefficients (device / pinned)
    int device = 0;                              // CUDA device the context lives on (tsl_create)
    // owner-computes assembly of the cloth rows (tsl_assembly.cu), single-cloth scenes: 0 off, 1 the forward Newton matrices (default),
    // 2 also the fp64 residual and energy by tiles (deterministic, but slower than the element kernels: fp64 sqrt / div / acos bound)
    int fast_assembly = 0;
    std::vector<std::pair<long long, long long>> zero_ranges;   // value ranges [a, b) of the slices holding rows of other bodies
    double *egrid_partial = nullptr; unsigned int *egrid_ticket = nullptr; int egrid_blocks = 0;   // per-tile partials of k_energy_rows (+1: result)
    tsl::DistCtx dist;
};
