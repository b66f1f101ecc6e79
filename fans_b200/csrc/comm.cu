// comm.cu — the three exchanges of the slab decomposition over NCCL (NVLink / NVSwitch), replacing the reference's MPI calls:
//   halo planes        MPI_Sendrecv           include/solver.h:244-245, 264-269
//   spectrum transpose FFTW-MPI all-to-all    include/solver.h:221-222, 395, 409  (FFTW_MPI_TRANSPOSED_OUT / _IN)
//   scalars            MPI_Allreduce SUM/MAX  include/solverCG.h:57, include/solver.h:430, 733
// libnccl.so.2 is resolved lazily with dlopen, so the library loads (and the single-GPU path runs) on machines without NCCL and
// shares the NCCL instance a host process (e.g. PyTorch) has already loaded.
#include "internal.h"
#include <dlfcn.h>
#include <nccl.h>

namespace {
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi g_nccl;

bool nccl_load()
{
    if (g_nccl.h) return true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        g_nccl.err = std::string("cannot load libnccl.so.2: ") + dlerror();
        return false;
    }
#define NCCL_SYM(name)                                                          \
    g_nccl.name = reinterpret_cast<decltype(g_nccl.name)>(dlsym(h, "nccl" #name)); \
    if (!g_nccl.name) {                                                         \
        g_nccl.err = "libnccl.so.2 lacks nccl" #name;                           \
        return false;                                                           \
    }
    NCCL_SYM(GetUniqueId) NCCL_SYM(CommInitRank) NCCL_SYM(CommDestroy) NCCL_SYM(AllReduce) NCCL_SYM(Send) NCCL_SYM(Recv)
    NCCL_SYM(GroupStart) NCCL_SYM(GroupEnd) NCCL_SYM(GetErrorString)
#undef NCCL_SYM
    g_nccl.h = h;
    return true;
}
}  // namespace

#define NCCL_TRY(ctx, expr)                                                                                          \
    do {                                                                                                             \
        ncclResult_t _r = (expr);                                                                                    \
        if (_r != ncclSuccess) {                                                                                     \
            fans_set_error((ctx), FANS_ERR_NCCL, std::string(#expr) + ": " + g_nccl.GetErrorString(_r));           \
            return FANS_ERR_NCCL;                                                                                    \
        }                                                                                                            \
    } while (0)

extern "C" int fans_comm_unique_id(void *id128)
{
    if (!id128) return FANS_ERR_ARG;
    if (!nccl_load()) {
        fans_set_error(nullptr, FANS_ERR_NCCL, g_nccl.err);
        return FANS_ERR_NCCL;
    }
    ncclUniqueId id;
    NCCL_TRY(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return FANS_OK;
}

extern "C" int fans_comm_create(void **comm, int32_t n_ranks, int32_t rank, const void *id128, int32_t device)
{
    if (!comm || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) return FANS_ERR_ARG;
    if (!nccl_load()) {
        fans_set_error(nullptr, FANS_ERR_NCCL, g_nccl.err);
        return FANS_ERR_NCCL;
    }
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) {
        fans_set_error(nullptr, FANS_ERR_CUDA, "fans_comm_create: cudaSetDevice failed");
        return FANS_ERR_CUDA;
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    NCCL_TRY(nullptr, g_nccl.CommInitRank(&c, n_ranks, id, rank));
    *comm = c;
    return FANS_OK;
}

extern "C" int fans_comm_destroy(void *comm)
{
    if (!comm) return FANS_OK;
    if (!nccl_load()) return FANS_ERR_NCCL;
    g_nccl.CommDestroy((ncclComm_t)comm);
    return FANS_OK;
}

int comm_check(fans_ctx *ctx)
{
    if (ctx->P == 1) return FANS_OK;
    if (!ctx->cfg.nccl_comm) {
        fans_set_error(ctx, FANS_ERR_NCCL, "world_size > 1 needs fans_config.nccl_comm (see fans_comm_create)");
        return FANS_ERR_NCCL;
    }
    if (!nccl_load()) {
        fans_set_error(ctx, FANS_ERR_NCCL, g_nccl.err);
        return FANS_ERR_NCCL;
    }
    return FANS_OK;
}

// in-place SUM / MAX of n doubles over the slabs (out may differ from in)
int comm_allreduce(fans_ctx *ctx, const double *in, double *out, int n, bool is_max)
{
    if (ctx->P == 1) {
        if (in != out) CUDA_TRY(ctx, cudaMemcpyAsync(out, in, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->st));
        return FANS_OK;
    }
    prof_begin(ctx, PC_COMM_SCALAR);
    NCCL_TRY(ctx, g_nccl.AllReduce(in, out, (size_t)n, ncclDouble, is_max ? ncclMax : ncclSum, (ncclComm_t)ctx->cfg.nccl_comm, ctx->st));
    prof_end(ctx);
    return FANS_OK;
}

// ring exchange of one plane in each direction:
//   to_prev (may be null) is sent to rank-1 and arrives there as from_next;  to_next -> rank+1 arrives as from_prev.
int comm_halo(fans_ctx *ctx, const void *to_prev, void *from_next, const void *to_next, void *from_prev, size_t bytes)
{
    const int P = ctx->P, prev = (ctx->rank + P - 1) % P, next = (ctx->rank + 1) % P;
    ncclComm_t c = (ncclComm_t)ctx->cfg.nccl_comm;
    prof_begin(ctx, PC_COMM_HALO);
    NCCL_TRY(ctx, g_nccl.GroupStart());
    if (to_prev) {
        NCCL_TRY(ctx, g_nccl.Send(to_prev, bytes, ncclChar, prev, c, ctx->st));
        NCCL_TRY(ctx, g_nccl.Recv(from_next, bytes, ncclChar, next, c, ctx->st));
    }
    if (to_next) {
        NCCL_TRY(ctx, g_nccl.Send(to_next, bytes, ncclChar, next, c, ctx->st));
        NCCL_TRY(ctx, g_nccl.Recv(from_prev, bytes, ncclChar, prev, c, ctx->st));
    }
    NCCL_TRY(ctx, g_nccl.GroupEnd());
    prof_end(ctx);
    return FANS_OK;
}

// block transpose of the spectrum: block q of `src` goes to rank q and lands there as block `rank` of `dst`
int comm_alltoall(fans_ctx *ctx, const double2 *src, double2 *dst)
{
    const int P = ctx->P;
    const size_t blk = (size_t)ctx->h * ctx->n0 * ctx->n1 * ctx->kzp;  // double2 elements per block
    ncclComm_t c = (ncclComm_t)ctx->cfg.nccl_comm;
    prof_begin(ctx, PC_COMM_A2A);
    CUDA_TRY(ctx, cudaMemcpyAsync(dst + (size_t)ctx->rank * blk, src + (size_t)ctx->rank * blk, sizeof(double2) * blk, cudaMemcpyDeviceToDevice, ctx->st));
    NCCL_TRY(ctx, g_nccl.GroupStart());
    for (int d = 1; d < P; ++d) {
        const int to = (ctx->rank + d) % P, from = (ctx->rank + P - d) % P;
        NCCL_TRY(ctx, g_nccl.Send(src + (size_t)to * blk, 2 * blk, ncclDouble, to, c, ctx->st));
        NCCL_TRY(ctx, g_nccl.Recv(dst + (size_t)from * blk, 2 * blk, ncclDouble, from, c, ctx->st));
    }
    NCCL_TRY(ctx, g_nccl.GroupEnd());
    prof_end(ctx);
    return FANS_OK;
}
