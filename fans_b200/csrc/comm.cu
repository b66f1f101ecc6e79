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
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
    NCCL_SYM(GetUniqueId) NCCL_SYM(CommInitRank) NCCL_SYM(CommDestroy) NCCL_SYM(AllReduce) NCCL_SYM(Send) NCCL_SYM(Recv) NCCL_SYM(AllGather)
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

// MPI_Allreduce(MPI_IN_PLACE, ..., MPI_SUM) of a small HOST vector over the slabs: the averages of Solver::postprocess
// (include/solver.h:556-571).  A C++ host without MPI needs nothing but this library to run P ranks.
extern "C" int fans_allreduce_sum(fans_ctx *ctx, double *host_inout, int32_t n)
{
    if (!ctx || !host_inout || n < 0) return FANS_ERR_ARG;
    if (ctx->P == 1 || n == 0) return FANS_OK;
    cudaSetDevice(ctx->device);
    double *d = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d, sizeof(double) * n));
    int rc = FANS_OK;
    if (cudaMemcpyAsync(d, host_inout, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->st) != cudaSuccess) rc = FANS_ERR_CUDA;
    if (rc == FANS_OK) rc = comm_allreduce(ctx, d, d, n, false);
    if (rc == FANS_OK && cudaMemcpyAsync(host_inout, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->st) != cudaSuccess) rc = FANS_ERR_CUDA;
    if (cudaStreamSynchronize(ctx->st) != cudaSuccess) rc = FANS_ERR_CUDA;
    cudaFree(d);
    if (rc == FANS_ERR_CUDA) fans_set_error(ctx, rc, "fans_allreduce_sum: CUDA error");
    return rc;
}

// MAX of one int over the slabs, in place on the device: the sticky fault flag (negative Jacobian, ...) must take EVERY rank out of
// the solve in the same iteration — a rank that returned alone would leave the others hanging in the next collective (the reference
// aborts the whole MPI job on an uncaught exception)
int comm_allreduce_int_max(fans_ctx *ctx, int *d_val)
{
    if (ctx->P == 1) return FANS_OK;
    NCCL_TRY(ctx, g_nccl.AllReduce(d_val, d_val, 1, ncclInt, ncclMax, (ncclComm_t)ctx->cfg.nccl_comm, ctx->st));
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
    const size_t blk = (size_t)ctx->h * ctx->n0 * ((size_t)ctx->n1 * ctx->kzp + ctx->xpad);  // double2 elements per block
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

// ------------------------------------------------------------------------------------------------
// Peer-mapped spectrum buffers (fused transpose): every rank maps the A (x-slab) and B (y-slab) spectrum buffers of all
// other ranks through CUDA IPC, so the last FFT pass before a transpose can store its rows straight into the owner's
// memory over NVLink / NVSwitch instead of packing + ncclSend/ncclRecv.  The 64-byte handles travel with one ncclAllGather.
// ------------------------------------------------------------------------------------------------
int comm_map_peers(fans_ctx *ctx)
{
    const int P = ctx->P;
    for (int q = 0; q < 8; ++q) ctx->peerA[q] = ctx->peerB[q] = nullptr;
    ctx->peerA[ctx->rank] = ctx->spec;
    ctx->peerB[ctx->rank] = ctx->specB;
    ctx->p2p = false;
    if (P == 1 || P > 8) return FANS_OK;
    if (const char *e = getenv("FANS_P2P"))
        if (atoi(e) == 0) return FANS_OK;
    struct Pair { cudaIpcMemHandle_t a, b; };
    Pair mine;
    CUDA_TRY(ctx, cudaIpcGetMemHandle(&mine.a, ctx->spec));
    CUDA_TRY(ctx, cudaIpcGetMemHandle(&mine.b, ctx->specB));
    Pair *d_all = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d_all, sizeof(Pair) * (P + 1)));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_all + P, &mine, sizeof(Pair), cudaMemcpyHostToDevice, ctx->st));
    NCCL_TRY(ctx, g_nccl.AllGather(d_all + P, d_all, sizeof(Pair), ncclChar, (ncclComm_t)ctx->cfg.nccl_comm, ctx->st));
    std::vector<Pair> all(P);
    CUDA_TRY(ctx, cudaMemcpyAsync(all.data(), d_all, sizeof(Pair) * P, cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    cudaFree(d_all);
    bool ok = true;
    for (int q = 0; q < P && ok; ++q) {
        if (q == ctx->rank) continue;
        void *pa = nullptr, *pb = nullptr;
        if (cudaIpcOpenMemHandle(&pa, all[q].a, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
            cudaIpcOpenMemHandle(&pb, all[q].b, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = false;
            break;
        }
        ctx->peerA[q] = (double2 *)pa;
        ctx->peerB[q] = (double2 *)pb;
    }
    // all ranks must agree (a rank that cannot map its peers would deadlock the others at the first barrier)
    double flag = ok ? 0.0 : 1.0, *d_flag = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d_flag, sizeof(double)));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_flag, &flag, sizeof(double), cudaMemcpyHostToDevice, ctx->st));
    NCCL_TRY(ctx, g_nccl.AllReduce(d_flag, d_flag, 1, ncclDouble, ncclMax, (ncclComm_t)ctx->cfg.nccl_comm, ctx->st));
    CUDA_TRY(ctx, cudaMemcpyAsync(&flag, d_flag, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    cudaFree(d_flag);
    ctx->p2p = (flag == 0.0);
    if (!ctx->p2p) comm_unmap_peers(ctx);
    return FANS_OK;
}

void comm_unmap_peers(fans_ctx *ctx)
{
    for (int q = 0; q < 8; ++q) {
        if (q == ctx->rank) continue;
        if (ctx->peerA[q]) cudaIpcCloseMemHandle(ctx->peerA[q]);
        if (ctx->peerB[q]) cudaIpcCloseMemHandle(ctx->peerB[q]);
        ctx->peerA[q] = ctx->peerB[q] = nullptr;
    }
}

// stream-ordered barrier over the slabs: when it completes on this rank's stream, every rank's stream has reached it, i.e.
// the kernels (and their peer stores) enqueued before it have finished everywhere
int comm_barrier(fans_ctx *ctx)
{
    if (ctx->P == 1) return FANS_OK;
    prof_begin(ctx, PC_COMM_A2A);
    NCCL_TRY(ctx, g_nccl.AllReduce(ctx->d_red + S_BARRIER, ctx->d_red + S_BARRIER, 1, ncclDouble, ncclSum, (ncclComm_t)ctx->cfg.nccl_comm, ctx->st));
    prof_end(ctx);
    return FANS_OK;
}

// the same on another stream of this context (chunked slab pipeline, solve.cu); its own dummy operand so that it never races with
// a barrier of the main stream
int comm_barrier_on(fans_ctx *ctx, cudaStream_t st)
{
    if (ctx->P == 1) return FANS_OK;
    NCCL_TRY(ctx, g_nccl.AllReduce(ctx->d_red + S_BARRIER2, ctx->d_red + S_BARRIER2, 1, ncclDouble, ncclSum, (ncclComm_t)ctx->cfg.nccl_comm, st));
    return FANS_OK;
}
