// fft_y.cu — passes P2 / P4 of the convolution (include/solver.h:387-412): transform along y of one component (in place on one
// GPU; with slabs the forward pass PUSHES its rows into the owners' transposed spectrum and the inverse pass PULLS them back over
// NVLink, which is the x<->y transpose FFTW-MPI performs with MPI all-to-alls),
// one (component, x plane, kz tile of T columns) per CTA.  Rows are T*16-byte segments of the spectrum, loaded straight
// into registers (8 independent 16-byte loads per thread in flight), transformed with the register FFT of fft_reg.cuh and
// stored straight from registers; shared memory is only the inter-stage exchange tile.
#include "fft_reg.cuh"
#include "internal.h"

template <int T>
struct TileIdx {  // swizzled [row][t] tile: conflict-free butterflies for T = 4 (64-byte rows) and T = 8 (128-byte rows)
    int t;
    __device__ __forceinline__ int operator()(int row) const
    {
        if (T == 4) return (row * 4 + t) ^ (((row >> 3) & 1) << 2);
        return row * T + t;
    }
};

template <int N, int T, bool INV>
__global__ void __launch_bounds__((N / rp_elems(N)) * T, ((N / rp_elems(N)) * T >= 512) ? (1024 / ((N / rp_elems(N)) * T)) : 1) k_fft_y(double2 *__restrict__ spec, const double2 *__restrict__ tw, SpecGeom g, int nTiles, PeerTable peers)
{
    extern __shared__ double2 sm[];
    constexpr int E = rp_elems(N), TPC = N / E, NST = rp_nstages(N);
    const int t = threadIdx.x % T, jt = threadIdx.x / T;
    int b = blockIdx.x;
    const int tile = b % nTiles;
    b /= nTiles;
    const int xl = b % g.n0, c = b / g.n0;
    double2 *base = spec + (size_t)c * g.cStride + (size_t)xl * g.xStride + (size_t)tile * T + t;
    const TileIdx<T> idx{t};
    double2 a[1][E];
    (void)TPC;
    if (!INV) {
#pragma unroll
        for (int e = 0; e < E; ++e) a[0][e] = base[spec_row_y(g, rp_row<N, 0>(jt, e))];
        rp_forward<N, 1>(a, jt, sm, N * T, idx, tw, 1);
        if (peers.on) {  // row y belongs to rank y / n1: store it into that rank's transposed spectrum, block `me`
            const size_t off = (size_t)peers.me * g.blkStride + (size_t)c * g.cStride + (size_t)xl * g.xStride + (size_t)tile * T + t;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int row = rp_row<N, NST - 1>(jt, e);
                peer_select(peers, row >> g.l2n1)[off + (size_t)(row & (g.n1 - 1)) * g.kzp] = a[0][e];
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) base[spec_row_y(g, rp_row<N, NST - 1>(jt, e))] = a[0][e];
        }
    } else {
        if (peers.on) {  // row y lives in rank (y / n1)'s transposed spectrum, block `me`: pull it over NVLink
            const size_t off = (size_t)peers.me * g.blkStride + (size_t)c * g.cStride + (size_t)xl * g.xStride + (size_t)tile * T + t;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int row = rp_row<N, NST - 1>(jt, e);
                a[0][e] = peer_select(peers, row >> g.l2n1)[off + (size_t)(row & (g.n1 - 1)) * g.kzp];
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) a[0][e] = base[spec_row_y(g, rp_row<N, NST - 1>(jt, e))];
        }
        rp_inverse<N, 1>(a, jt, sm, N * T, idx, tw, 1);
#pragma unroll
        for (int e = 0; e < E; ++e) base[spec_row_y(g, rp_row<N, 0>(jt, e))] = a[0][e];
    }
}

template <int N, int T>
static int launch_y(fans_ctx *ctx, bool inverse, const SpecGeom &g, const PeerTable &peers)
{
    constexpr int E = rp_elems(N);
    const int nTiles = (ctx->kzc + T - 1) / T;
    const size_t smem = sizeof(double2) * N * T;
    const unsigned grid = (unsigned)((size_t)ctx->h * ctx->n0 * nTiles);
    const int nthr = (N / E) * T;
    if (!inverse) {
        if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_y<N, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_fft_y<N, T, false><<<grid, nthr, smem, ctx->st>>>(ctx->spec, ctx->plany.tw, g, nTiles, peers);
    } else {
        if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_y<N, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_fft_y<N, T, true><<<grid, nthr, smem, ctx->st>>>(ctx->spec, ctx->plany.tw, g, nTiles, peers);
    }
    return FANS_OK;
}

SpecGeom spec_geom_A(const fans_ctx *ctx)
{
    SpecGeom g;
    g.n0 = ctx->n0;
    g.n1 = ctx->n1;
    g.l2n0 = rp_log2(ctx->n0);
    g.l2n1 = rp_log2(ctx->n1);
    g.kzp = ctx->kzp;
    g.h = ctx->h;
    g.xStride = (size_t)ctx->n1 * ctx->kzp + ctx->xpad;
    g.cStride = (size_t)ctx->n0 * g.xStride;
    g.blkStride = (size_t)ctx->h * g.cStride;
    return g;
}

// y pass (forward or inverse) on the local x-slab
int fft_pass_y(fans_ctx *ctx, bool inverse)
{
    prof_begin(ctx, inverse ? PC_FFT_Y_INV : PC_FFT_Y_FWD);
    const SpecGeom g = spec_geom_A(ctx);
    int rc = FANS_ERR_ARG;
    const int T = ctx->yT;
    PeerTable peers;
    for (int q = 0; q < 8; ++q) peers.p[q] = ctx->peerB[q];
    peers.me = ctx->rank;
    peers.on = (ctx->P > 1 && ctx->p2p) ? 1 : 0;  // forward: push rows to their owners; inverse: pull them back
#define Y_CASE(N_)                                                         \
    case N_:                                                               \
        rc = (T == 8) ? launch_y<N_, 8>(ctx, inverse, g, peers) : launch_y<N_, 4>(ctx, inverse, g, peers); \
        break;
    switch (ctx->ny) {
        Y_CASE(4) Y_CASE(8) Y_CASE(16) Y_CASE(32) Y_CASE(64) Y_CASE(128) Y_CASE(256) Y_CASE(512) Y_CASE(1024)
    default:
        fans_set_error(ctx, FANS_ERR_ARG, "unsupported n_y for the y pass");
    }
#undef Y_CASE
    prof_end(ctx);
    ctx->launches++;
    if (rc != FANS_OK) return rc;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}
