// fft_y.cu — passes P2 / P4 of the convolution (include/solver.h:387-412): transform along y of one component (in place on one
// GPU; with slabs the forward pass PUSHES its rows into the owners' transposed spectrum and the inverse pass PULLS them back over
// NVLink, which is the x<->y transpose FFTW-MPI performs with MPI all-to-alls),
// one (component, x plane, kz tile of T columns) per CTA.  Rows are T*16-byte segments of the spectrum, loaded straight
// into registers (8 independent 16-byte loads per thread in flight), transformed with the register FFT of fft_reg.cuh and
// stored straight from registers; shared memory is only the inter-stage exchange tile.
#include "fft_reg.cuh"
#include "internal.h"
#include <algorithm>

template <int T>
struct TileIdx {  // swizzled [row][t] tile: conflict-free butterflies for T = 4 (64-byte rows) and T = 8 (128-byte rows)
    int t;
    __device__ __forceinline__ int operator()(int row) const
    {
        if (T == 4) return (row * 4 + t) ^ (((row >> 3) & 1) << 2);
        return row * T + t;
    }
};

// one (component, x plane, kz tile) of the y pass; b = ((c - c0) * n0 + xl) * nTiles + (tile - tile0), nTiles tiles from tile0
template <int N, int T, bool INV>
__device__ __forceinline__ void y_tile(double2 *__restrict__ spec, const double2 *__restrict__ tw, const SpecGeom &g, int nTiles, const PeerTable &peers,
                                       int b, int c0, double2 *sm, int tile0 = 0)
{
    constexpr int E = rp_elems(N), NST = rp_nstages(N);
    const int t = threadIdx.x % T, jt = threadIdx.x / T;
    const int tile = tile0 + b % nTiles;
    b /= nTiles;
    const int xl = b % g.n0, c = c0 + b / g.n0;
    double2 *base = spec + (size_t)c * g.cStride + (size_t)xl * g.xStride + (size_t)tile * T + t;
    const TileIdx<T> idx{t};
    double2 a[1][E];
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

#define Y_BOUNDS __launch_bounds__((N / rp_elems(N)) * T, ((N / rp_elems(N)) * T >= 512) ? (1024 / ((N / rp_elems(N)) * T)) : 1)

// whole spectrum, one tile per CTA
template <int N, int T, bool INV>
__global__ void Y_BOUNDS k_fft_y(double2 *__restrict__ spec, const double2 *__restrict__ tw, SpecGeom g, int nTiles, PeerTable peers)
{
    extern __shared__ double2 sm[];
    y_tile<N, T, INV>(spec, tw, g, nTiles, peers, blockIdx.x, 0, sm);
}

// Component range and persistent grid (slab pipeline, solve.cu): `c0` first component, `nWork` tiles of this launch; a grid smaller
// than nWork loops.  `gate` (optional): word set to `gate_val` by the first CTA as soon as it is resident, so a k_gate launch on
// another stream can hold back a concurrent kernel until this one owns its SMs.
template <int N, int T, bool INV>
__global__ void Y_BOUNDS k_fft_y_part(double2 *__restrict__ spec, const double2 *__restrict__ tw, SpecGeom g, int nTiles, PeerTable peers, int c0,
                                      int nWork, int *gate, int gate_val, int tile0)
{
    extern __shared__ double2 sm[];
    if (gate && blockIdx.x == 0 && threadIdx.x == 0) {
        *reinterpret_cast<volatile int *>(gate) = gate_val;
        __threadfence();
    }
    for (int w = blockIdx.x; w < nWork; w += gridDim.x) {
        y_tile<N, T, INV>(spec, tw, g, nTiles, peers, w, c0, sm, tile0);
        __syncthreads();  // the exchange tile is reused by the next tile
    }
}

template <int N, int T>
static int launch_y(fans_ctx *ctx, bool inverse, const SpecGeom &g, const PeerTable &peers, const YLaunch &yl)
{
    constexpr int E = rp_elems(N);
    const int allTiles = (ctx->kzc + T - 1) / T;
    const int tile0 = yl.ntile > 0 ? yl.tile0 : 0;
    const int nTiles = yl.ntile > 0 ? std::min(yl.ntile, allTiles - tile0) : allTiles;   // tiles of THIS launch
    const size_t smem = sizeof(double2) * N * T;
    const int nWork = (int)((size_t)yl.nc * ctx->n0 * nTiles);
    const unsigned grid = (unsigned)((yl.grid > 0 && yl.grid < nWork) ? yl.grid : nWork);
    const int nthr = (N / E) * T;
    if (grid == (unsigned)nWork && yl.c0 == 0 && !yl.gate && yl.ntile == 0) {  // whole-spectrum launch: one tile per CTA
        if (!inverse) {
            if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_y<N, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_fft_y<N, T, false><<<grid, nthr, smem, yl.st>>>(ctx->spec, ctx->plany.tw, g, nTiles, peers);
        } else {
            if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_y<N, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_fft_y<N, T, true><<<grid, nthr, smem, yl.st>>>(ctx->spec, ctx->plany.tw, g, nTiles, peers);
        }
    } else {  // component-range / persistent form (slab pipeline)
        constexpr int NP = N;
        if (!inverse) {
            if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_y_part<NP, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_fft_y_part<NP, T, false><<<grid, nthr, smem, yl.st>>>(ctx->spec, ctx->plany.tw, g, nTiles, peers, yl.c0, nWork, yl.gate, yl.gate_val, tile0);
        } else {
            if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_y_part<NP, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_fft_y_part<NP, T, true><<<grid, nthr, smem, yl.st>>>(ctx->spec, ctx->plany.tw, g, nTiles, peers, yl.c0, nWork, yl.gate, yl.gate_val, tile0);
        }
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
// (a batched solve transforms the h components of all its lanes as nb * h components of one launch)
int fft_pass_y(fans_ctx *ctx, bool inverse) { return fft_pass_y_part(ctx, inverse, YLaunch{ctx->st, 0, ctx->h * ctx->nb, 0, nullptr, 0}); }

// the same for the components [c0, c0 + nc) on stream yl.st with at most yl.grid CTAs (0: one per tile)
int fft_pass_y_part(fans_ctx *ctx, bool inverse, const YLaunch &yl)
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
        rc = (T == 8) ? launch_y<N_, 8>(ctx, inverse, g, peers, yl) : launch_y<N_, 4>(ctx, inverse, g, peers, yl); \
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
