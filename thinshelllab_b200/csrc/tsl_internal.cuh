// tsl_internal.cuh -- context and device-side data structures of libtsl (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/tsl.h"
#include "tsl_elements.cuh"

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
    float *val32 = nullptr;     // [nnzb_pad*9]
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
    int *hinge_face, *hinge_l;   // [NH] hinges = (i, l) with cf[i][l] > i
    int *tri_slot;               // [NF][9]  padded block ids of the (l, m) blocks of a triangle
    int *hinge_slot;             // [NH][16] padded block ids of the (j, k) blocks of a hinge
    double *ref_angle;           // [NF][3] borrowed
    double *norm_dir;            // [NF][3] scratch: unit normals
    double *q1;                  // [9 + 81] c_i rows and mat_N of faces 0..2 (quirk Q1)
};

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

// Krylov scalars kept on the device so a solve never needs the host inside the loop
struct KrylovScalars {
    double acc_pq[2];     // PCG: p.Ap   (parity buffers)
    double acc_rz[2];     // PCG: r.z
    double acc_rr[2];     // |r|^2
    double acc_rho[2];    // BiCGStab: rhat.r
    double acc_rhv;       // rhat.v
    double acc_ts, acc_tt;
    double alpha;
    double rr0;           // |b|^2
    int flags;            // bit0 negative curvature / breakdown (iteration frozen)
    int pad;
};

}  // namespace tsl

struct tsl_ctx {
    tsl_config cfg;
    std::string err;
    cudaStream_t stream = 0;
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
    double *vn = nullptr;                       // [n_verts][3]
    int *proj_flag = nullptr, *proj_dir = nullptr, *proj_idx = nullptr;   // [n_bodies][n_verts]([3])
    double *proj_w = nullptr;
    unsigned int *cell_key = nullptr, *cell_key_sorted = nullptr;         // per face of the current surface body
    int *face_id = nullptr, *face_id_sorted = nullptr;
    void *cub_tmp = nullptr; size_t cub_tmp_bytes = 0;
    int *cflag = nullptr, *cscan = nullptr;     // [n_verts] compaction scratch
    tsl::ContactDev con;
    int nc = 0;

    // linear system
    tsl::SellMatrix A;
    bool last_f64 = false;
    double *F = nullptr;                         // [3 n_verts] residual
    float *minv32 = nullptr; double *minv64 = nullptr;   // [n_verts][9] block-Jacobi inverse
    float *cg_x = nullptr, *cg_r = nullptr, *cg_z = nullptr, *cg_p = nullptr, *cg_q = nullptr;  // [3 n_rows_pad]
    double *bi[8] = { nullptr };                 // BiCGStab vectors: r, rhat, p, v, y, s, z, t
    double *sol = nullptr;                       // [3 n_verts] Newton direction (f64)
    double *x1 = nullptr;                        // [n_verts][3] line-search base
    tsl::KrylovScalars *ks = nullptr;            // device
    tsl::KrylovScalars *ks_host = nullptr;       // pinned
    // reductions
    double *red_partial = nullptr; unsigned int *red_ticket = nullptr; double *red_out = nullptr; double *red_host = nullptr;
    int red_blocks = 0;
    // adjoint scratch
    double *d_kb = nullptr;                      // [n_verts][3]
    double *adj_rhs = nullptr, *adj_z = nullptr; // [3 n_verts]
    int error_flag_host = 0; int *error_flag = nullptr;   // device-side "unsupported" flags
};
