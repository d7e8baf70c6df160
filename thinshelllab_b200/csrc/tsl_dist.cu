// tsl_dist.cu -- collectives of the strip-partitioned implicit step: one process per GPU, the cloth grid cut
// into strips of rows, two ghost rows on each inner side (a hinge reaches two rows), so that every element touching an owned vertex
// is local and assembly needs no exchange.  What travels over NCCL (NVLink / NVSwitch):
//   * the ghost rows of the PCG direction p before every operator application (2 rows x (M+1) vertices x 24 B per neighbour),
//   * the PCG scalars: p.Ap, and (r.z, |r|^2) as one 16-byte message, per iteration,
//   * energy, |p|_inf and F.p of the Newton driver (one scalar each per line-search trial / iteration).
// NCCL is resolved at run time from the libnccl.so.2 the process already holds (torch's), so libtsl has no link dependency on it.
#include <dlfcn.h>
#include <algorithm>
#include <vector>
#include <nccl.h>

#include "tsl_internal.cuh"
#include "tsl_kernels.cuh"

namespace tsl {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
static NcclApi &nccl()
{
    static NcclApi api;
    if (api.ok) return api;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
#define SYM(field, name) api.field = (decltype(api.field))dlsym(h, name)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy"); SYM(CommAbort, "ncclCommAbort");
    SYM(AllReduce, "ncclAllReduce"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
    return api;
}
#define NCK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { ctx->err = std::string(#x) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r_) : "nccl error"); return TSL_ERR_CUDA; } } while (0)

int dist_allreduce(tsl_ctx *ctx, double *dev, int n, bool max_op)
{
    if (!ctx->dist.on || ctx->dist.world == 1) return TSL_OK;
    NCK(nccl().AllReduce(dev, dev, (size_t)n, ncclDouble, max_op ? ncclMax : ncclSum, (ncclComm_t)ctx->dist.comm, ctx->stream));
    ctx->dist.allreduces++;
    return TSL_OK;
}

int dist_halo(tsl_ctx *ctx, double *v)
{
    const DistCtx &d = ctx->dist;
    if (!d.on || d.world == 1) return TSL_OK;
    ncclComm_t comm = (ncclComm_t)d.comm;
    const size_t row = 3 * (size_t)d.row_len;                 // doubles per grid row
    const int rows_local = d.nvc / d.row_len;
    NCK(nccl().GroupStart());
    if (d.ghost_lo > 0) {                                     // lower neighbour: my first owned rows -> its upper ghost rows, and back
        NCK(nccl().Send(v + row * d.ghost_lo, row * d.ghost_lo, ncclDouble, d.rank - 1, comm, ctx->stream));
        NCK(nccl().Recv(v, row * d.ghost_lo, ncclDouble, d.rank - 1, comm, ctx->stream));
    }
    if (d.ghost_hi > 0) {
        NCK(nccl().Send(v + row * (rows_local - 2 * d.ghost_hi), row * d.ghost_hi, ncclDouble, d.rank + 1, comm, ctx->stream));
        NCK(nccl().Recv(v + row * (rows_local - d.ghost_hi), row * d.ghost_hi, ncclDouble, d.rank + 1, comm, ctx->stream));
    }
    NCK(nccl().GroupEnd());
    ctx->dist.halo_msgs++;
    return TSL_OK;
}

// tsl_destroy: the stream is idle (synchronised by the caller), every collective this rank enqueued has completed, so the communicator
// can be torn down without waiting for the peers.  ncclCommAbort does exactly that; ncclCommDestroy may block at interpreter exit when a
// peer process has already gone.
void dist_destroy(tsl_ctx *ctx)
{
    if (ctx->dist.comm) {
        if (nccl().CommAbort) nccl().CommAbort((ncclComm_t)ctx->dist.comm);
        else if (nccl().CommDestroy) nccl().CommDestroy((ncclComm_t)ctx->dist.comm);
    }
    ctx->dist.comm = nullptr;
    ctx->dist.on = false;
}

__global__ void k_zero_ghost(int n3, int lo3, int hi3, double *v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n3 && (i < lo3 || i >= hi3)) v[i] = 0;
}
void launch_zero_ghost(tsl_ctx *ctx, double *v)
{
    const DistCtx &d = ctx->dist;
    int n3 = 3 * d.nvc;
    k_zero_ghost<<<(n3 + 255) / 256, 256, 0, ctx->stream>>>(n3, 3 * d.own0, 3 * d.own1, v);
    ctx->launches++;
}

}  // namespace tsl

using namespace tsl;

extern "C" {

int tsl_dist_unique_id(void *out128_host)
{
    if (!out128_host || !nccl().ok) return TSL_ERR_UNSUPPORTED;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return TSL_ERR_CUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128_host, &id, 128);
    return TSL_OK;
}

int tsl_dist_init(tsl_ctx *ctx, const void *id128_host, int rank, int world, int ghost_lo_rows, int ghost_hi_rows, int first_row_global)
{
    if (!ctx || !id128_host) return TSL_ERR_INVALID;
    if (ctx->cloths.size() != 1 || ctx->cloths[0].offset != 0) { ctx->err = "tsl_dist_init: needs exactly one cloth at vertex offset 0"; return TSL_ERR_INVALID; }
    if (world < 1 || rank < 0 || rank >= world) { ctx->err = "tsl_dist_init: bad rank / world"; return TSL_ERR_INVALID; }
    // a hinge reaches two grid rows: exactly 2 ghost rows on an inner side, none on an outer side
    if ((ghost_lo_rows != 0 && ghost_lo_rows != 2) || (ghost_hi_rows != 0 && ghost_hi_rows != 2)) { ctx->err = "tsl_dist_init: ghost rows must be 0 or 2"; return TSL_ERR_INVALID; }
    if ((rank == 0) != (ghost_lo_rows == 0) || (rank == world - 1) != (ghost_hi_rows == 0)) {
        ctx->err = "tsl_dist_init: inner sides need 2 ghost rows, the outer sides of the first / last strip none"; return TSL_ERR_INVALID;
    }
    // Cloth.init_mesh picks the diagonal of quad (i, j) from the parity of i + j: the local mesh (built from LOCAL row indices) is the
    // global triangulation only if local row 0 is an even global row
    if (first_row_global < 0 || (first_row_global & 1)) { ctx->err = "tsl_dist_init: the strip's first local row must be an even global row (triangulation parity)"; return TSL_ERR_INVALID; }
    if (!nccl().ok) { ctx->err = "tsl_dist_init: libnccl.so.2 not found in this process"; return TSL_ERR_UNSUPPORTED; }
    const ClothDev &c = ctx->cloths[0];
    DistCtx &d = ctx->dist;
    d.rank = rank; d.world = world;
    d.row_len = c.M + 1;
    d.nvc = c.NV;
    d.ghost_lo = ghost_lo_rows; d.ghost_hi = ghost_hi_rows;
    int rows_local = c.N + 1, rows_owned = rows_local - ghost_lo_rows - ghost_hi_rows;
    // the halo exchange sends the first / last `ghost` OWNED rows: a strip must own at least as many rows as it lends
    if (rows_owned < std::max(2, std::max(ghost_lo_rows, ghost_hi_rows))) { ctx->err = "tsl_dist_init: a strip must own at least 2 grid rows"; return TSL_ERR_INVALID; }
    d.own0 = ghost_lo_rows * d.row_len;
    d.own1 = (rows_local - ghost_hi_rows) * d.row_len;
    if (world > 1) {
        ncclUniqueId id;
        memcpy(&id, id128_host, 128);
        ncclComm_t comm;
        NCK(nccl().CommInitRank(&comm, world, id, rank));
        d.comm = comm;
        // every rank publishes (row length, first owned global row, owned rows): the strips must tile the sheet without gaps, with
        // equal row length -- otherwise the send / recv counts of the halo exchange would not match and NCCL would hang
        std::vector<double> mine(3 * (size_t)world, 0.0);
        mine[3 * rank] = d.row_len; mine[3 * rank + 1] = first_row_global + ghost_lo_rows; mine[3 * rank + 2] = rows_owned;
        double *buf = nullptr;
        if (cudaMalloc(&buf, sizeof(double) * mine.size()) != cudaSuccess) { ctx->err = "tsl_dist_init: cudaMalloc"; return TSL_ERR_CUDA; }
        cudaMemcpyAsync(buf, mine.data(), sizeof(double) * mine.size(), cudaMemcpyHostToDevice, ctx->stream);
        ncclResult_t r = nccl().AllReduce(buf, buf, mine.size(), ncclDouble, ncclSum, comm, ctx->stream);
        cudaMemcpyAsync(mine.data(), buf, sizeof(double) * mine.size(), cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        cudaFree(buf);
        if (r != ncclSuccess) { ctx->err = "tsl_dist_init: all-reduce of the strip geometry failed"; return TSL_ERR_CUDA; }
        for (int q = 0; q < world; q++) {
            if (mine[3 * q] != d.row_len) { ctx->err = "tsl_dist_init: strips have different row lengths"; dist_destroy(ctx); return TSL_ERR_INVALID; }
            if (q > 0 && mine[3 * q + 1] != mine[3 * (q - 1) + 1] + mine[3 * (q - 1) + 2]) {
                ctx->err = "tsl_dist_init: strips do not tile the sheet (first owned row of a rank != end of the previous rank's rows)";
                dist_destroy(ctx);
                return TSL_ERR_INVALID;
            }
        }
    }
    d.on = true;
    // the NCCL calls of an iteration are captured into its CUDA graph like the kernels (measured on 2 B200: same results, 22 % faster
    // than eager launches); TSL_DIST_GRAPHS=0 falls back to eager launches
    if (const char *e = getenv("TSL_DIST_GRAPHS")) ctx->use_graphs = atoi(e);
    cudaStreamSynchronize(ctx->stream);
    graphs_invalidate(ctx);
    return TSL_OK;
}

int tsl_dist_stats(tsl_ctx *ctx, long long *halo_msgs, long long *allreduces)
{
    if (!ctx) return TSL_ERR_INVALID;
    if (halo_msgs) *halo_msgs = ctx->dist.halo_msgs;
    if (allreduces) *allreduces = ctx->dist.allreduces;
    return TSL_OK;
}

}  // extern "C"
