// fft_x.cu — pass P3 of the convolution (include/solver.h:395-409): forward transform along x, the Fourier-space
// fundamental-solution multiply  r_hat <- Gamma_hat(xi) r_hat  (solver.h:399-406) and the inverse transform along x, fused:
// all `howmany` components of a (y, kz-tile) pencil are held by one CTA (registers + thread-private shared-memory slots), so the
// h x h real-symmetric multiply happens between the last forward butterfly and the first inverse butterfly, off HBM.
// Gamma_hat is streamed once per iteration in exactly the register order of the threads (gamma.cu writes it that way):
//     gamma[((cta*NG + k) * (N/E) * E + e*(N/E) + jt) * T + t],   storage row = jt*E + e,   NG = H(H+1)/2 upper triangle.
#include "fft_reg.cuh"
#include "internal.h"
#include <algorithm>

template <int T>
struct TileIdxX {
    int t;
    __device__ __forceinline__ int operator()(int row) const
    {
        if (T == 4) return (row * 4 + t) ^ (((row >> 3) & 1) << 2);
        if (T == 2) return (row * 2 + t) ^ (((row >> 3) & 3) << 1);
        return row * T + t;
    }
};

// minimum resident CTAs per SM the register budget is tuned for (threads per CTA = H*(N/E)*T)
__host__ __device__ constexpr int xg_min_blocks(int nthr) { return nthr > 512 ? 1 : (nthr > 256 ? 2 : 3); }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <int N, int H, int T>
__global__ void __launch_bounds__(H * (N / rp_elems(N)) * T, xg_min_blocks(H * (N / rp_elems(N)) * T))
    k_fft_xg(double2 *__restrict__ spec, const double *__restrict__ gamma, const double2 *__restrict__ tw, SpecGeom g, int nTiles,
             int nWork, PeerTable peers, int nb)
{
    // batched solves (nb > 1 lanes): work item wq = tile * nb + lane — the lanes of one tile run next to each other, so its Gamma_hat
    // block comes from HBM once and from L2 for the other lanes; lane l of the spectrum starts at spec + l * H * cStride
    // One thread group per component: every thread carries E points of ONE component through the stages (32 data registers),
    // the H groups share the barriers.  At the Fourier-space boundary the groups swap their values through thread-private
    // slots of the (then idle) exchange tiles and each group forms its own row of  Gamma_hat r_hat.
    // Persistent CTAs, two tile buffers: while tile w is transformed in buffer `cur` (which doubles as its exchange tile), the
    // rows of tile w + gridDim.x travel global -> shared with cp.async into the other buffer, so the strided loads of the
    // spectrum never sit on the critical path.
    extern __shared__ double2 sm[];
    constexpr int E = rp_elems(N), TPC = N / E, NST = rp_nstages(N), NG = H * (H + 1) / 2, NTC = TPC * T, BUF = H * N * T;
    const int c = threadIdx.x / NTC, tc = threadIdx.x % NTC;
    const int t = tc % T, jt = tc / T;
    const TileIdxX<T> idx{t};
    using GSync = SyncGroup<(H > 1 && NTC % 32 == 0)>;
    const GSync gsync{1 + c, NTC};  // the stage exchanges of one component only involve its own thread group
    auto tile_ptr = [&](int wq) {
        const int w = wq / nb, ln = wq - w * nb;
        return spec + (size_t)(ln * H + c) * g.cStride + (size_t)(w / nTiles) * g.kzp + (size_t)(w % nTiles) * T + t;
    };
    auto issue = [&](int w, int buf) {
        const double2 *src = tile_ptr(w);
        double2 *dst = sm + buf * BUF + c * (N * T) + tc;
#pragma unroll
        for (int e = 0; e < E; ++e) cp_async16(dst + e * NTC, src + spec_row_x(g, rp_row<N, 0>(jt, e)));
    };
    int cur = 0;
    if ((int)blockIdx.x < nWork) issue(blockIdx.x, 0);
    for (int w = blockIdx.x; w < nWork; w += gridDim.x, cur ^= 1) {
        double2 *S = sm + cur * BUF;      // all components of this tile
        double2 *X = S + c * (N * T);     // this group's exchange tile
        double2 *base = tile_ptr(w);
        const double *gam = gamma + (size_t)(w / nb) * NG * (N * T) + jt * T + t;
        double2 a[1][E];
        cp_async_wait_all();
#pragma unroll
        for (int e = 0; e < E; ++e) a[0][e] = X[e * NTC + tc];
        __syncthreads();  // everybody holds its rows (and is done with the previous tile): both buffers may be overwritten
        if (w + (int)gridDim.x < nWork) issue(w + gridDim.x, cur ^ 1);
        rp_forward<N, 1, TileIdxX<T>, 0, GSync>(a, jt, X, 0, idx, tw, 1, gsync);
        // Green operator; storage row of register e after the last stage = jt*E + e
        if (H == 1) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const double g0 = __ldg(gam + (size_t)e * NTC);
                a[0][e] = make_double2(g0 * a[0][e].x, g0 * a[0][e].y);
            }
        } else {
            constexpr size_t NT = (size_t)N * T;
            // packed upper triangle 00,01,02,11,12,22: row c of the symmetric matrix
            const int k0 = (c == 0) ? 0 : (c == 1 ? 1 : 2), k1 = (c == 0) ? 1 : (c == 1 ? 3 : 4), k2 = (c == 0) ? 2 : (c == 1 ? 4 : 5);
            if (NST > 1) gsync();  // this group is done reading its last exchange
#pragma unroll
            for (int e = 0; e < E; ++e) X[e * NTC + tc] = a[0][e];
            double gc[E][3];  // issued after the put (a[] is dead) so the loads fly while the CTA gathers at the barrier
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const double *ge = gam + (size_t)e * NTC;
                gc[e][0] = __ldg(ge + k0 * NT);
                gc[e][1] = __ldg(ge + k1 * NT);
                gc[e][2] = __ldg(ge + k2 * NT);
            }
            __syncthreads();
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const double2 r0 = S[e * NTC + tc], r1 = S[(H > 1 ? 1 : 0) * (N * T) + e * NTC + tc], r2 = S[(H > 2 ? 2 : 0) * (N * T) + e * NTC + tc];
                a[0][e] = make_double2(gc[e][0] * r0.x + gc[e][1] * r1.x + gc[e][2] * r2.x, gc[e][0] * r0.y + gc[e][1] * r1.y + gc[e][2] * r2.y);
            }
            if (NST > 1) __syncthreads();  // the other groups have read this group's slots: its tile may serve the inverse exchanges
        }
        rp_inverse<N, 1, TileIdxX<T>, NST - 1, GSync>(a, jt, X, 0, idx, tw, 1, gsync);
        if (peers.on) {  // plane x belongs to rank x / n0: store it into that rank's x-slab spectrum, block `me`
            const size_t off = (size_t)peers.me * g.blkStride + (size_t)c * g.cStride + (size_t)(w / nTiles) * g.kzp + (size_t)(w % nTiles) * T + t;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const int row = rp_row<N, 0>(jt, e);
                peer_select(peers, row >> g.l2n0)[off + (size_t)(row & (g.n0 - 1)) * g.xStride] = a[0][e];
            }
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) base[spec_row_x(g, rp_row<N, 0>(jt, e))] = a[0][e];
        }
    }
}

// Sequential-component variant (H = 3): one thread group of (N/E)*T threads carries the three components through the forward
// transform one after the other, parks each result in thread-private slots of that component's tile, applies the full symmetric
// Gamma_hat block to its own frequencies (no cross-thread traffic, no barrier) and runs the three inverse transforms.  A third of
// the threads and of the registers of k_fft_xg per tile, so TWO independent CTAs fit on an SM and one CTA's butterflies overlap the
// other's loads, stores and shared-memory exchanges.
// PF: the strided rows of the NEXT component (and, during the inverse transforms, of the next tile's first component) travel
// global -> shared with cp.async straight into that component's tile, which is idle until its own forward transform starts — the
// load latency that one CTA per SM cannot hide with other warps disappears behind the butterflies, at no register cost.
// (Parking the transformed components in registers instead — r[3][E], no park / read / write-back / re-read of the slots, 10 instead
// of 14 shared-memory accesses per element — does not fit the 128 registers a 512-thread CTA has: ptxas spills 672 + 832 bytes per
// thread and tile, as much local-memory traffic through the same L1 as the shared-memory traffic it saves.)
template <int N, int T, bool PF>
__global__ void __launch_bounds__((N / rp_elems(N)) * T, (3 * N * T * 16 <= 112 * 1024) ? 2 : 1)
    k_fft_xg_seq(double2 *__restrict__ spec, const double *__restrict__ gamma, const double2 *__restrict__ tw, SpecGeom g, int nTiles,
                 int nWork, PeerTable peers, int tile0, int ntc, int nb)
{
    // work item q of this launch = (y row q / ntc, kz tile tile0 + q % ntc): the whole spectrum (tile0 = 0, ntc = nTiles) or one
    // kz chunk of the slab pipeline; `w` below is the item's index in the full (row, tile) numbering the Gamma layout uses.
    // Batched solves (nb > 1 lanes): item = q * nb + lane, lane l of the spectrum starts at spec + l * 3 * cStride; the lanes of a tile
    // are in flight together, so its Gamma_hat block comes from HBM once and from L2 for the others.
    extern __shared__ double2 sm[];  // [3][N*T]
    constexpr int H = 3, E = rp_elems(N), TPC = N / E, NST = rp_nstages(N), NG = 6, NTC = TPC * T;
    constexpr size_t NT = (size_t)N * T;
    const int tc = threadIdx.x, t = tc % T, jt = tc / T;
    const TileIdxX<T> idx{t};
    auto tile_off = [&](int w) { return (size_t)(w / nTiles) * g.kzp + (size_t)(w % nTiles) * T + t; };
    auto prefetch = [&](int w, int c, int ln) {   // rows of component c of tile w, lane ln -> this thread's slots of tile c; one commit group
        const double2 *base = spec + (size_t)(ln * H + c) * g.cStride + tile_off(w);
        double2 *dst = sm + c * NT + tc;
#pragma unroll
        for (int e = 0; e < E; ++e) cp_async16(dst + e * NTC, base + spec_row_x(g, rp_row<N, 0>(jt, e)));
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    auto full_index = [&](int qb) {
        const int q = qb / nb;
        return (q / ntc) * nTiles + tile0 + q % ntc;
    };
    if (PF && (int)blockIdx.x < nWork) prefetch(full_index(blockIdx.x), 0, blockIdx.x % nb);
    for (int q = blockIdx.x; q < nWork; q += gridDim.x) {
        const int w = full_index(q), ln = q % nb;
        const size_t off = tile_off(w);
        double2 *lspec = spec + (size_t)ln * H * g.cStride;
        double2 a[1][E];
#pragma unroll 1
        for (int c = 0; c < H; ++c) {
            double2 *X = sm + c * NT;
            if (PF) {
                // tile c+1 is idle (its last use, the previous tile's inverse transform, lies behind barriers every thread has passed)
                if (c + 1 < H) {
                    prefetch(w, c + 1, ln);
                    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                }
#pragma unroll
                for (int e = 0; e < E; ++e) a[0][e] = X[e * NTC + tc];
                if (NST > 1) __syncthreads();  // everybody holds its rows before the first exchange overwrites the slots
            } else {
                const double2 *base = lspec + (size_t)c * g.cStride + off;
#pragma unroll
                for (int e = 0; e < E; ++e) a[0][e] = base[spec_row_x(g, rp_row<N, 0>(jt, e))];
            }
            rp_forward<N, 1>(a, jt, X, 0, idx, tw, 1);
            if (NST > 1) __syncthreads();  // everybody is done reading the last exchange of this component
#pragma unroll
            for (int e = 0; e < E; ++e) X[e * NTC + tc] = a[0][e];  // storage row of register e = jt*E + e
        }
        // Green operator on this thread's own frequencies: packed upper triangle 00,01,02,11,12,22
        const double *gam = gamma + (size_t)w * NG * NT + tc;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const double *ge = gam + (size_t)e * NTC;
            const double g00 = __ldg(ge), g01 = __ldg(ge + NT), g02 = __ldg(ge + 2 * NT), g11 = __ldg(ge + 3 * NT), g12 = __ldg(ge + 4 * NT),
                         g22 = __ldg(ge + 5 * NT);
            const double2 r0 = sm[e * NTC + tc], r1 = sm[NT + e * NTC + tc], r2 = sm[2 * NT + e * NTC + tc];
            sm[e * NTC + tc] = make_double2(g00 * r0.x + g01 * r1.x + g02 * r2.x, g00 * r0.y + g01 * r1.y + g02 * r2.y);
            sm[NT + e * NTC + tc] = make_double2(g01 * r0.x + g11 * r1.x + g12 * r2.x, g01 * r0.y + g11 * r1.y + g12 * r2.y);
            sm[2 * NT + e * NTC + tc] = make_double2(g02 * r0.x + g12 * r1.x + g22 * r2.x, g02 * r0.y + g12 * r1.y + g22 * r2.y);
        }
#pragma unroll 1
        for (int c = 0; c < H; ++c) {
            double2 *X = sm + c * NT;
#pragma unroll
            for (int e = 0; e < E; ++e) a[0][e] = X[e * NTC + tc];
            rp_inverse<N, 1>(a, jt, X, 0, idx, tw, 1);  // its first barrier also covers the slot reads above
            if (peers.on) {
                const size_t poff = (size_t)peers.me * g.blkStride + (size_t)c * g.cStride + off;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int row = rp_row<N, 0>(jt, e);
                    peer_select(peers, row >> g.l2n0)[poff + (size_t)(row & (g.n0 - 1)) * g.xStride] = a[0][e];
                }
            } else {
                double2 *base = lspec + (size_t)c * g.cStride + off;
#pragma unroll
                for (int e = 0; e < E; ++e) base[spec_row_x(g, rp_row<N, 0>(jt, e))] = a[0][e];
            }
            // component 1 is through its inverse transform => every thread has left tile 0: the next tile's first component may land
            if (PF && c == 1 && NST > 1 && q + (int)gridDim.x < nWork) prefetch(full_index(q + gridDim.x), 0, (q + gridDim.x) % nb);
        }
        if (PF && NST <= 1 && q + (int)gridDim.x < nWork) {  // single-stage transforms have no barriers to lean on
            __syncthreads();
            prefetch(full_index(q + gridDim.x), 0, (q + gridDim.x) % nb);
        }
    }
}

struct XPart {  // kz-tile range and launch of one x-pass call (whole spectrum: tile0 = 0, ntile = 0, st = ctx->st, grid_cap = 0)
    cudaStream_t st;
    int tile0, ntile, grid_cap;
};

template <int N, int T>
static int launch_xg_seq(fans_ctx *ctx, double2 *specB, const SpecGeom &g, const PeerTable &peers, const XPart &xp)
{
    constexpr int E = rp_elems(N), NTHR = (N / E) * T;
    const int nTiles = (ctx->kzc + T - 1) / T;
    const size_t smem = 3 * sizeof(double2) * N * T;
    static int resident = 0;
    if (!resident) {
        if (smem > 48 * 1024) {
            CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_xg_seq<N, T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_xg_seq<N, T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_fft_xg_seq<N, T, true>, NTHR, smem));
        if (resident < 1) resident = 1;
    }
    const int tile0 = xp.ntile > 0 ? xp.tile0 : 0, ntc = xp.ntile > 0 ? std::min(xp.ntile, nTiles - tile0) : nTiles;
    if (ntc <= 0) return FANS_OK;
    const int nWork = ctx->n1 * ntc * ctx->nb;
    int grid = FANS_SMS * resident;
    if (const char *env = getenv("FANS_XG_GRID")) grid = atoi(env);
    if (xp.grid_cap > 0 && grid > xp.grid_cap * resident) grid = xp.grid_cap * resident;
    if (grid > nWork) grid = nWork;
    // prefetch into the idle tiles pays with 128-byte rows (T = 8: 2.75 -> 2.63 ms at n_x = 512); with the 64-byte rows of n_x = 1024
    // the extra barrier per component costs more than the hidden latency (3.42 -> 3.70 ms on the 8-GPU run), so it stays off there
    const char *pf = getenv("FANS_XG_PF");   // 0 / 1: force plain loads / prefetch (A/B runs)
    const bool use_pf = pf ? atoi(pf) != 0 : (T == 8);
    if (!use_pf) k_fft_xg_seq<N, T, false><<<grid, NTHR, smem, xp.st>>>(specB, ctx->gamma, ctx->planx.tw, g, nTiles, nWork, peers, tile0, ntc, ctx->nb);
    else k_fft_xg_seq<N, T, true><<<grid, NTHR, smem, xp.st>>>(specB, ctx->gamma, ctx->planx.tw, g, nTiles, nWork, peers, tile0, ntc, ctx->nb);
    return FANS_OK;
}

template <int N, int H, int T>
static int launch_xg(fans_ctx *ctx, double2 *specB, const SpecGeom &g, const PeerTable &peers)
{
    constexpr int E = rp_elems(N), NTHR = H * (N / E) * T;
    const int nTiles = (ctx->kzc + T - 1) / T;
    const size_t smem = 2 * sizeof(double2) * N * T * H;  // two tile buffers (transform in one, next tile lands in the other)
    if (smem > 227 * 1024 || NTHR > 1024) {
        fans_set_error(ctx, FANS_ERR_ARG, "x pass tile does not fit one CTA");
        return FANS_ERR_ARG;
    }
    static int resident = 0;  // per instantiation
    if (!resident) {
        if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_xg<N, H, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_fft_xg<N, H, T>, NTHR, smem));
        if (resident < 1) resident = 1;
    }
    const int nWork = ctx->n1 * nTiles * ctx->nb;
    int grid = FANS_SMS * resident;
    if (const char *env = getenv("FANS_XG_GRID")) grid = atoi(env);
    if (grid > nWork) grid = nWork;
    k_fft_xg<N, H, T><<<grid, NTHR, smem, ctx->st>>>(specB, ctx->gamma, ctx->planx.tw, g, nTiles, nWork, peers, ctx->nb);
    return FANS_OK;
}

// tile width of the fused x pass for a given (nx, howmany): keeps the exchange tile within ~100 KB
int fft_x_tile_width(int nx, int h)
{
    int T = (h == 1 || (h == 3 && nx >= 64 && !getenv("FANS_XG_OLD"))) ? 8 : 4;
    if (const char *env = getenv("FANS_XG_T")) T = atoi(env);
    if (h == 3 && nx >= 64 && !getenv("FANS_XG_OLD")) {  // sequential-component variant: one tile buffer of all components
        while ((3 * (size_t)nx * T * sizeof(double2) > 200 * 1024 || (nx / 8) * T > 1024) && T > 2) T /= 2;
        return T;
    }
    while ((2 * (size_t)h * nx * T * sizeof(double2) > 200 * 1024 || h * (nx / 8) * T > 1024) && T > 2) T /= 2;
    return T;
}

int fft_pass_x_gamma(fans_ctx *ctx) { return fft_pass_x_gamma_part(ctx, ctx->st, 0, 0, 0); }

// kz tiles [tile0, tile0 + ntile) of the x pass (tile width ctx->gT) on stream st with at most grid_cap SMs (0: all); ntile = 0: whole
// spectrum.  Ranges are only implemented for the sequential-component kernel (h = 3, n_x >= 64), which is what the slab pipeline uses.
int fft_pass_x_gamma_part(fans_ctx *ctx, cudaStream_t st, int tile0, int ntile, int grid_cap)
{
    const XPart xp{st, tile0, ntile, grid_cap};
    prof_begin(ctx, PC_FFT_X_GAMMA);
    SpecGeom g = spec_geom_A(ctx);
    double2 *specB = ctx->P > 1 ? ctx->specB : ctx->spec;
    int rc = FANS_ERR_ARG;
    const int T = ctx->gT;
    PeerTable peers;
    for (int q = 0; q < 8; ++q) peers.p[q] = ctx->peerA[q];
    peers.me = ctx->rank;
    const bool seq = ctx->h == 3 && ctx->nx >= 64 && !getenv("FANS_XG_OLD");  // the sequential-component kernel (two CTAs per SM or 128-byte rows)
    peers.on = 0;  // the x pass stays local (in place on the transposed spectrum); the y passes carry both transposes
#define X_CASE(N_)                                                                                     \
    case N_:                                                                                           \
        if (ctx->h == 1) rc = (T == 8) ? launch_xg<N_, 1, 8>(ctx, specB, g, peers) : (T == 4 ? launch_xg<N_, 1, 4>(ctx, specB, g, peers) : launch_xg<N_, 1, 2>(ctx, specB, g, peers)); \
        else if (seq && N_ >= 64) rc = (T == 8 && N_ <= 512) ? launch_xg_seq<(N_ <= 512 ? N_ : 64), 8>(ctx, specB, g, peers, xp) : ((T == 4) ? launch_xg_seq<N_, 4>(ctx, specB, g, peers, xp) : launch_xg_seq<N_, 2>(ctx, specB, g, peers, xp));  \
        else rc = (T == 4) ? launch_xg<N_, 3, 4>(ctx, specB, g, peers) : launch_xg<N_, 3, 2>(ctx, specB, g, peers);  \
        break;
    switch (ctx->nx) {
        X_CASE(4) X_CASE(8) X_CASE(16) X_CASE(32) X_CASE(64) X_CASE(128) X_CASE(256) X_CASE(512) X_CASE(1024)
    default:
        fans_set_error(ctx, FANS_ERR_ARG, "unsupported n_x for the fused x pass");
    }
#undef X_CASE
    prof_end(ctx);
    ctx->launches++;
    if (rc != FANS_OK) return rc;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}
