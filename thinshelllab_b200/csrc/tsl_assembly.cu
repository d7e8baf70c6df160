// tsl_assembly.cu -- host side of the owner-computes cloth assembly (kernels: tsl_assembly_kernels.cuh; tables: tsl_grid.h).
#include "tsl_internal.cuh"
#include "tsl_kernels.cuh"
#include "tsl_assembly_kernels.cuh"

namespace tsl {

#define GRID(n, b) (unsigned)(((n) + (b) - 1) / (b))
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return TSL_ERR_CUDA; } } while (0)

// tsl_finalize: constant-memory tables of the grid structure, and the value ranges of the matrix that the cloth-row kernels do NOT
// own (slices holding rows of other bodies): those are cleared before every assembly, cloth rows are simply overwritten
int assembly_init(tsl_ctx *ctx)
{
    ctx->fast_assembly = 0;
    if (ctx->cloths.size() != 1) return TSL_OK;            // several cloths: the scatter kernels of tsl_physics.cu
    int level = 1;
    if (const char *e = getenv("TSL_FAST_ASSEMBLY")) level = atoi(e);
    if (level <= 0) return TSL_OK;
    GridTables T;
    if (!build_grid_tables(T)) { ctx->err = "grid tables: the mesher's structure is not parity-periodic"; return TSL_ERR_INVALID; }
    CK(cudaMemcpyToSymbol(c_gt, &T, sizeof(T)));
    const ClothDev &c = ctx->cloths[0];
    const SellMatrix &A = ctx->A;
    ctx->zero_ranges.clear();
    for (int S = 0; S < A.n_slices; S++) {
        bool foreign = false;
        for (int r = 32 * S; r < 32 * S + 32; r++) foreign = foreign || r < c.offset || r >= c.offset + c.NV;
        if (!foreign) continue;
        long long a = 9LL * A.h_slice_base[S], b = 9LL * A.h_slice_base[S + 1];
        if (!ctx->zero_ranges.empty() && ctx->zero_ranges.back().second == a) ctx->zero_ranges.back().second = b;
        else ctx->zero_ranges.push_back({ a, b });
    }
    // energy kernel: one partial per tile
    {
        int nblk = (int)(GRID(c.M + 1, TSL_TJ) * GRID(c.N + 1, TSL_TI));
        if (nblk > ctx->egrid_blocks) {
            cudaFree(ctx->egrid_partial); cudaFree(ctx->egrid_ticket);
            CK(cudaMalloc(&ctx->egrid_partial, sizeof(double) * (nblk + 1)));
            CK(cudaMalloc(&ctx->egrid_ticket, sizeof(unsigned int)));
            CK(cudaMemset(ctx->egrid_ticket, 0, sizeof(unsigned int)));
            ctx->egrid_blocks = nblk;
        }
    }
    CK(cudaFuncSetAttribute(k_hessian_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TSL_HESS_SMEM));
    CK(cudaFuncSetAttribute(k_residual_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 3 * TSL_HH * TSL_HW * 12)));
    ctx->fast_assembly = level;
    return TSL_OK;
}

static ClothGrid64 cloth_grid64(tsl_ctx *ctx)
{
    const ClothDev &c = ctx->cloths[0];
    ClothGrid64 G;
    G.N = c.N; G.M = c.M; G.NV = c.NV; G.offset = c.offset;
    G.Kl = c.P.Kl; G.Ka = c.P.Ka; G.Kb = c.P.Kb; G.dx = c.P.dx; G.dt = c.P.dt; G.mass = c.P.mass;
    for (int k = 0; k < 3; k++) G.g[k] = ctx->cfg.gravity[k];
    return G;
}

// F[cloth rows] = inertia / gravity + membrane + bending gradient (complete: plain stores)
void launch_residual_rows(tsl_ctx *ctx, const double *pos)
{
    ClothGrid64 G = cloth_grid64(ctx);
    dim3 grid(GRID(G.M + 1, TSL_TJ), GRID(G.N + 1, TSL_TI));
    size_t dyn = sizeof(double) * 3 * TSL_HH * TSL_HW * 12;
    k_residual_rows<<<grid, 128, dyn, ctx->stream>>>(G, pos, ctx->prev_pos, ctx->vel, ctx->vgrav, ctx->cloths[0].ref_angle, ctx->F);
    ctx->launches++;
}

// energy of the cloth (vertex terms, triangles, hinges) -> *out_dev (deterministic sum)
void launch_energy_rows(tsl_ctx *ctx, const double *pos, double *out_dev)
{
    ClothGrid64 G = cloth_grid64(ctx);
    dim3 grid(GRID(G.M + 1, TSL_TJ), GRID(G.N + 1, TSL_TI));
    k_energy_rows<<<grid, 128, 0, ctx->stream>>>(G, pos, ctx->prev_pos, ctx->vel, ctx->vgrav, ctx->cloths[0].ref_angle, ctx->egrid_partial,
                                                 ctx->egrid_ticket, out_dev);
    ctx->launches++;
}

static ClothGrid cloth_grid(tsl_ctx *ctx)
{
    const ClothDev &c = ctx->cloths[0];
    ClothGrid G;
    G.N = c.N; G.M = c.M; G.NV = c.NV; G.offset = c.offset;
    G.Kl = (float)c.P.Kl; G.Ka = (float)c.P.Ka; G.Kb = (float)c.P.Kb; G.dx = (float)c.P.dx;
    G.mass_dt2 = (float)(c.P.mass / (c.P.dt * c.P.dt));
    return G;
}

// cloth rows of the exact (val_e) and clamped (val_c) forward Newton matrices, mass diagonal included; other rows are cleared
void launch_hessian_rows(tsl_ctx *ctx, const double *pos, float *val_e, float *val_c)
{
    cudaStream_t s = ctx->stream;
    for (auto &r : ctx->zero_ranges) {
        cudaMemsetAsync(val_e + r.first, 0, sizeof(float) * (size_t)(r.second - r.first), s);
        cudaMemsetAsync(val_c + r.first, 0, sizeof(float) * (size_t)(r.second - r.first), s);
    }
    ClothGrid G = cloth_grid(ctx);
    dim3 grid(GRID(G.M + 1, TSL_TJ), GRID(G.N + 1, TSL_TI));
    k_hessian_rows<<<grid, 256, TSL_HESS_SMEM, s>>>(G, pos, ctx->frozen, ctx->A.slice_base, ctx->A.colidx, ctx->A.diag_pb, val_e, val_c);
    ctx->launches++;
}

}  // namespace tsl
