// stencil.cu — the linear operator K.d of the CG fast path (include/solverCG.h:98-103) as a NODE-GATHER kernel.
//
// Reference semantics per element: res_e = phase_stiffness[phase] * (ue - u(node0)), scatter-added to the 8 nodes.
// Here every thread owns two z-adjacent nodes and gathers.  Two code paths, chosen per node pair:
//   * homogeneous neighbourhood (all adjacent elements share one phase — the bulk of any blocky microstructure):
//     the 8 element matrices collapse into a 27-point block stencil  sum_delta S_p[delta] d(n+delta)  whose
//     coefficients are compile-time offsets into __constant__ memory => DFMA with constant operands, 243 DFMA/node
//     (h = 3) instead of 576, no coefficient loads;
//   * interface nodes: the exact element form, element by element, K_phase rows again as constant operands.
// Data movement: a CTA owns an (8 x 64) node column and marches along x with a 4-slot shared-memory ring of node
// planes (one-node halo in y,z), so every d value is read from HBM once per CTA column (halo re-reads hit L2) and the
// fused update d = s + beta*d is formed on the fly while loading.  One __syncthreads per plane.
#include "internal.h"
#include "materials.cuh"
#include "stencil.h"

#define GY 8
#define GZ 64
#define GPZ (GZ + 2)
#define GTILE ((GY + 2) * GPZ)
#define GETILE ((GY + 1) * (GZ + 1))
#define G_THREADS 256
#ifndef STENCIL_MINB
#define STENCIL_MINB 2
#endif

__constant__ double c_S[STENCIL_MAXQ * 27 * 9];   // [q][delta][i][j]
__constant__ double c_KQ[STENCIL_MAXQ * 576];      // [q][row][col]  (8h x 8h, h <= 3)

struct StencilParams {
    int n0, ny, nz;
    size_t nloc;
    int xchunk;
    const double *d_old;   // direction of the previous iteration (or the plain input when s == nullptr)
    const double *s;       // nullptr: plain K.d of d_old
    double *d_new;
    const double *beta;
    double *out;
    const uint16_t *phidx;
    int nq;
    double *part;
    unsigned int *ticket;
    double *red_out;       // <d_new, K d_new>
    // slab decomposition (world_size > 1): node planes -1 and n0 and element plane -1 come from the neighbour ranks
    const double *halo_lo, *halo_hi;   // [H][ny*nz], already the UPDATED direction (no s/beta combination)
    const uint16_t *ms_lo;             // [ny*nz] phases of element plane -1
};

__device__ __forceinline__ int gwrap(int v, int n)
{
    v %= n;
    return v < 0 ? v + n : v;
}

// homogeneous path: both nodes of the pair, phase Q (compile time => constant-bank operands)
template <int H, int Q>
__device__ __forceinline__ void stencil_pair(const double *ring, int k, int ry, int rz0, double (&accA)[H], double (&accB)[H])
{
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
        const double *pl = ring + (size_t)((k + dx - 1 + 4) & 3) * H * GTILE;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            double v[H][4];
#pragma unroll
            for (int c = 0; c < H; ++c) {
                const double2 *p2 = reinterpret_cast<const double2 *>(pl + c * GTILE + (ry + dy - 1) * GPZ + rz0);
                const double2 a = p2[0], b = p2[1];
                v[c][0] = a.x, v[c][1] = a.y, v[c][2] = b.x, v[c][3] = b.y;
            }
#pragma unroll
            for (int dz = 0; dz < 3; ++dz) {
                const double *S = c_S + ((Q * 27 + (dx * 9 + dy * 3 + dz)) * H * H);
#pragma unroll
                for (int i = 0; i < H; ++i)
#pragma unroll
                    for (int j = 0; j < H; ++j) {
                        accA[i] = fma(S[i * H + j], v[j][dz], accA[i]);
                        accB[i] = fma(S[i * H + j], v[j][dz + 1], accB[i]);
                    }
            }
        }
    }
}

// exact element form for ONE node at tile position (ry, rz) of plane k; element E = (ox,oy,oz) below the node
template <int H, int Q, int A>
__device__ __forceinline__ void element_rows(const double *ring, int k, int ry, int rz, double (&acc)[H])
{
    constexpr int ox = A & 1, oy = (A >> 1) & 1, oz = (A >> 2) & 1;
    constexpr int ND = 8 * H;
    double u0[H];
#pragma unroll
    for (int c = 0; c < H; ++c) u0[c] = ring[((size_t)((k - ox + 4) & 3) * H + c) * GTILE + (ry - oy) * GPZ + (rz - oz)];
#pragma unroll
    for (int b = 1; b < 8; ++b) {
        const int bx = b & 1, by = (b >> 1) & 1, bz = (b >> 2) & 1;
        double w[H];
#pragma unroll
        for (int c = 0; c < H; ++c)
            w[c] = ring[((size_t)((k - ox + bx + 4) & 3) * H + c) * GTILE + (ry - oy + by) * GPZ + (rz - oz + bz)] - u0[c];
#pragma unroll
        for (int i = 0; i < H; ++i)
#pragma unroll
            for (int j = 0; j < H; ++j) acc[i] = fma(c_KQ[Q * (ND * ND) + (H * A + i) * ND + H * b + j], w[j], acc[i]);
    }
}

template <int H, int NQ, int A>
__device__ __forceinline__ void element_dispatch(const double *ring, int k, int ry, int rz, int ph, double (&acc)[H])
{
    if (NQ > 0 && ph == 0) element_rows<H, 0, A>(ring, k, ry, rz, acc);
    else if (NQ > 1 && ph == 1) element_rows<H, (NQ > 1 ? 1 : 0), A>(ring, k, ry, rz, acc);
    else if (NQ > 2 && ph == 2) element_rows<H, (NQ > 2 ? 2 : 0), A>(ring, k, ry, rz, acc);
    else if (NQ > 3 && ph == 3) element_rows<H, (NQ > 3 ? 3 : 0), A>(ring, k, ry, rz, acc);
}

// ms ring: element plane p in slot (p+3)%3, tile [(GY+1)][(GZ+1)], element (ry,rz) = low corner at node tile (ry,rz)
template <int H, int NQ>
__device__ __forceinline__ void node_general(const double *ring, const uint16_t *mring, int k, int ry, int rz, double (&acc)[H])
{
#define EL_PH(ox, oy, oz) ((int)mring[((k - (ox) + 3) % 3) * GETILE + (ry - (oy)) * (GZ + 1) + (rz - (oz))])
    element_dispatch<H, NQ, 0>(ring, k, ry, rz, EL_PH(0, 0, 0), acc);
    element_dispatch<H, NQ, 1>(ring, k, ry, rz, EL_PH(1, 0, 0), acc);
    element_dispatch<H, NQ, 2>(ring, k, ry, rz, EL_PH(0, 1, 0), acc);
    element_dispatch<H, NQ, 3>(ring, k, ry, rz, EL_PH(1, 1, 0), acc);
    element_dispatch<H, NQ, 4>(ring, k, ry, rz, EL_PH(0, 0, 1), acc);
    element_dispatch<H, NQ, 5>(ring, k, ry, rz, EL_PH(1, 0, 1), acc);
    element_dispatch<H, NQ, 6>(ring, k, ry, rz, EL_PH(0, 1, 1), acc);
    element_dispatch<H, NQ, 7>(ring, k, ry, rz, EL_PH(1, 1, 1), acc);
#undef EL_PH
}

template <int H, int NQ>
__global__ void __launch_bounds__(G_THREADS, STENCIL_MINB) k_stencil_linear(const StencilParams p)
{
    extern __shared__ __align__(16) double smem[];
    double *ring = smem;                                               // [4][H][GTILE]
    uint16_t *mring = reinterpret_cast<uint16_t *>(smem + 4 * H * GTILE);  // [3][GETILE]
    __shared__ double scratch[32];
    __shared__ uint16_t qlist[GY * 32];  // queued interface node pairs, per warp row
    __shared__ int qcnt[GY];

    const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
    const int z0 = blockIdx.x * GZ, y0 = blockIdx.y * GY;
    const int xs = blockIdx.z * p.xchunk, xe = min(xs + p.xchunk, p.n0);
    const int ry = wy + 1, rzA = 2 * lane + 1;  // tile coordinates of node A; node B = rzA + 1
    const int yA = y0 + wy, zA = z0 + 2 * lane;
    const bool valid = (yA < p.ny) && (zA < p.nz);  // nz is even, so A valid <=> B valid
    const double beta = p.s ? *p.beta : 0.0;

    // in-plane offsets of the tile positions this thread loads (fixed over the march): no div/mod in the x loop
    constexpr int NLD = (GTILE + G_THREADS - 1) / G_THREADS, NLM = (GETILE + G_THREADS - 1) / G_THREADS;
    int goff[NLD], moff[NLM];
    unsigned mine_mask = 0;
#pragma unroll
    for (int j = 0; j < NLD; ++j) {
        const int i = tid + j * G_THREADS;
        goff[j] = -1;
        if (i < GTILE) {
            const int r = i / GPZ, c = i % GPZ;
            goff[j] = gwrap(y0 - 1 + r, p.ny) * p.nz + gwrap(z0 - 1 + c, p.nz);
            if (r >= 1 && r <= GY && c >= 1 && c <= GZ && (y0 - 1 + r) < p.ny && (z0 - 1 + c) < p.nz) mine_mask |= 1u << j;
        }
    }
#pragma unroll
    for (int j = 0; j < NLM; ++j) {
        const int i = tid + j * G_THREADS;
        moff[j] = -1;
        if (i < GETILE) {
            const int r = i / (GZ + 1), c = i % (GZ + 1);
            moff[j] = gwrap(y0 - 1 + r, p.ny) * p.nz + gwrap(z0 - 1 + c, p.nz);
        }
    }
    const size_t plane_sz = (size_t)p.ny * p.nz;
    auto load_plane = [&](int xp, bool owned) {
        const size_t gbase = (size_t)gwrap(xp, p.n0) * plane_sz;
        double *pl = ring + (size_t)((xp + 4) & 3) * H * GTILE;
        double v[NLD][H];
        const double *hal = (p.halo_lo && xp < 0) ? p.halo_lo : ((p.halo_hi && xp >= p.n0) ? p.halo_hi : nullptr);
        if (hal) {
#pragma unroll
            for (int j = 0; j < NLD; ++j)
                if (goff[j] >= 0) {
#pragma unroll
                    for (int cc = 0; cc < H; ++cc) v[j][cc] = hal[cc * plane_sz + goff[j]];
                }
        } else {
#pragma unroll
            for (int j = 0; j < NLD; ++j)
                if (goff[j] >= 0) {
#pragma unroll
                    for (int cc = 0; cc < H; ++cc) v[j][cc] = p.d_old[cc * p.nloc + gbase + goff[j]];
                }
        }
        if (p.s && !hal) {
            double sv[NLD][H];
#pragma unroll
            for (int j = 0; j < NLD; ++j)
                if (goff[j] >= 0) {
#pragma unroll
                    for (int cc = 0; cc < H; ++cc) sv[j][cc] = p.s[cc * p.nloc + gbase + goff[j]];
                }
#pragma unroll
            for (int j = 0; j < NLD; ++j)
                if (goff[j] >= 0) {
#pragma unroll
                    for (int cc = 0; cc < H; ++cc) {
                        v[j][cc] = sv[j][cc] + beta * v[j][cc];
                        if (owned && ((mine_mask >> j) & 1u)) p.d_new[cc * p.nloc + gbase + goff[j]] = v[j][cc];
                    }
                }
        }
#pragma unroll
        for (int j = 0; j < NLD; ++j)
            if (goff[j] >= 0) {
#pragma unroll
                for (int cc = 0; cc < H; ++cc) pl[cc * GTILE + tid + j * G_THREADS] = v[j][cc];
            }
    };
    auto load_ms = [&](int xp) {
        const uint16_t *src = (p.ms_lo && xp < 0) ? p.ms_lo : p.phidx + (size_t)gwrap(xp, p.n0) * plane_sz;
        uint16_t *mp = mring + ((xp + 3) % 3) * GETILE;
#pragma unroll
        for (int j = 0; j < NLM; ++j)
            if (moff[j] >= 0) mp[tid + j * G_THREADS] = src[moff[j]];
    };

    load_plane(xs - 1, false);
    load_plane(xs, true);
    load_ms(xs - 1);
    double racc[1] = {0.0};
    for (int k = xs; k < xe; ++k) {
        load_plane(k + 1, (k + 1) < xe);
        load_ms(k);
        __syncthreads();
        // ---- phase 1: homogeneous node pairs take the 27-point stencil; interface pairs are queued
        bool homog = true;
        int ph = 0;
        if (valid) {
            // phases of the 12 elements around the node pair: planes k-1,k ; rows ry-1,ry ; columns rzA-1..rzA+1
            const uint16_t *m0 = mring + ((k - 1 + 3) % 3) * GETILE, *m1 = mring + ((k + 3) % 3) * GETILE;
            const int e00 = (ry - 1) * (GZ + 1) + rzA - 1, e10 = ry * (GZ + 1) + rzA - 1;
            ph = m1[e10 + 1];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                homog = homog && (m0[e00 + c] == ph) && (m0[e10 + c] == ph) && (m1[e00 + c] == ph) && (m1[e10 + c] == ph);
        }
        const unsigned qmask = __ballot_sync(0xffffffffu, valid && !homog);
        if (valid && homog) {
            double accA[H], accB[H];
#pragma unroll
            for (int c = 0; c < H; ++c) accA[c] = 0.0, accB[c] = 0.0;
            if (NQ > 0 && ph == 0) stencil_pair<H, 0>(ring, k, ry, rzA - 1, accA, accB);
            else if (NQ > 1 && ph == 1) stencil_pair<H, (NQ > 1 ? 1 : 0)>(ring, k, ry, rzA - 1, accA, accB);
            else if (NQ > 2 && ph == 2) stencil_pair<H, (NQ > 2 ? 2 : 0)>(ring, k, ry, rzA - 1, accA, accB);
            else if (NQ > 3 && ph == 3) stencil_pair<H, (NQ > 3 ? 3 : 0)>(ring, k, ry, rzA - 1, accA, accB);
            const size_t g = ((size_t)k * p.ny + yA) * p.nz + zA;
            const double *ctr = ring + (size_t)((k + 4) & 3) * H * GTILE + ry * GPZ + rzA;
#pragma unroll
            for (int c = 0; c < H; ++c) {
                *reinterpret_cast<double2 *>(p.out + c * p.nloc + g) = make_double2(accA[c], accB[c]);
                racc[0] += accA[c] * ctr[c * GTILE] + accB[c] * ctr[c * GTILE + 1];
            }
        }
        if (qmask) {
            if (valid && !homog) qlist[wy * 32 + __popc(qmask & ((1u << lane) - 1u))] = (uint16_t)(ry * 128 + rzA);
        }
        if (lane == 0) qcnt[wy] = __popc(qmask);
        __syncthreads();
        // ---- phase 2: the queued interface nodes of the whole CTA, densely packed onto lanes (exact element form)
        int pre[GY + 1];
        pre[0] = 0;
#pragma unroll
        for (int w = 0; w < GY; ++w) pre[w + 1] = pre[w] + qcnt[w];
        const int nitems = 2 * pre[GY];
        for (int n = tid; n < nitems; n += G_THREADS) {
            const int pair = n >> 1;
            int w = 0, base = 0;
#pragma unroll
            for (int q = 1; q < GY; ++q)
                if (pair >= pre[q]) w = q, base = pre[q];
            const int code = qlist[w * 32 + (pair - base)];
            const int qry = code >> 7, qrz = (code & 127) + (n & 1);
            double acc[H];
#pragma unroll
            for (int c = 0; c < H; ++c) acc[c] = 0.0;
            node_general<H, NQ>(ring, mring, k, qry, qrz, acc);
            const size_t g = ((size_t)k * p.ny + (y0 + qry - 1)) * p.nz + (z0 + qrz - 1);
            const double *ctr = ring + (size_t)((k + 4) & 3) * H * GTILE + qry * GPZ + qrz;
#pragma unroll
            for (int c = 0; c < H; ++c) {
                p.out[c * p.nloc + g] = acc[c];
                racc[0] += acc[c] * ctr[c * GTILE];
            }
        }
    }
    if (p.red_out) grid_reduce<1, 1>(racc, scratch, p.part, p.ticket, p.red_out);
}

// ------------------------------------------------------------------------------------------------
static uint64_t g_stencil_stamp = 0;

// 27-point block stencil of a homogeneous neighbourhood from the element matrix K (8h x 8h, node-major):
//   S[delta][i][j] = sum over local nodes a with b = a + delta inside the element of K[h a + i][h b + j]
void stencil_from_element_matrix(int h, const double *K, double *S /* [27][h][h] */)
{
    const int nd = 8 * h;
    for (int i = 0; i < 27 * h * h; ++i) S[i] = 0.0;
    for (int a = 0; a < 8; ++a)
        for (int b = 0; b < 8; ++b) {
            const int dx = (b & 1) - (a & 1), dy = ((b >> 1) & 1) - ((a >> 1) & 1), dz = ((b >> 2) & 1) - ((a >> 2) & 1);
            const int di = (dx + 1) * 9 + (dy + 1) * 3 + (dz + 1);
            for (int i = 0; i < h; ++i)
                for (int j = 0; j < h; ++j) S[(di * h + i) * h + j] += K[(h * a + i) * nd + h * b + j];
        }
}

bool stencil_supported(const fans_ctx *ctx) { return ctx->all_linear && ctx->n_k >= 1 && ctx->n_k <= STENCIL_MAXQ && ctx->n_k == ctx->n_phases; }

template <int H, int NQ>
static int launch_stencil(fans_ctx *ctx, const StencilParams &p, dim3 grid, size_t smem)
{
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_stencil_linear<H, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_stencil_linear<H, NQ><<<grid, G_THREADS, smem, ctx->st>>>(p);
    return FANS_OK;
}

int stencil_run(fans_ctx *ctx, const double *d_old, double *out, const double *s_in, double *d_new, const double *beta_dev,
                double *red_out)
{
    const int h = ctx->h, nd = 8 * h;
    if (g_stencil_stamp != ctx->const_stamp) {
        std::vector<double> S((size_t)ctx->n_k * 27 * h * h);
        for (int q = 0; q < ctx->n_k; ++q) stencil_from_element_matrix(h, ctx->K_host.data() + (size_t)q * nd * nd, S.data() + (size_t)q * 27 * h * h);
        // synchronous copies: the host vector dies at scope exit
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
        CUDA_TRY(ctx, cudaMemcpyToSymbol(c_S, S.data(), sizeof(double) * S.size()));
        CUDA_TRY(ctx, cudaMemcpyToSymbol(c_KQ, ctx->K_host.data(), sizeof(double) * (size_t)ctx->n_k * nd * nd));
        g_stencil_stamp = ctx->const_stamp;
    }
    StencilParams p;
    memset(&p, 0, sizeof(p));
    p.n0 = ctx->n0, p.ny = ctx->ny, p.nz = ctx->nz, p.nloc = ctx->nloc;
    p.d_old = d_old, p.s = s_in, p.d_new = d_new, p.beta = beta_dev, p.out = out;
    p.phidx = ctx->phidx;
    p.nq = ctx->n_k;
    p.part = ctx->d_part, p.ticket = ctx->d_ticket, p.red_out = red_out;
    if (ctx->P > 1) {
        // halo planes of the UPDATED direction: pack planes 0 / n0-1 (forming s + beta d on the fly), ring exchange
        FANS_CHECK(halo_exchange_both(ctx, d_old, s_in, beta_dev));
        p.halo_lo = ctx->halo_lo, p.halo_hi = ctx->halo_hi, p.ms_lo = ctx->ms_lo;
    }
    const int gy = (ctx->ny + GY - 1) / GY, gz = (ctx->nz + GZ - 1) / GZ;
    int xchunk = ctx->n0;
    while (xchunk > 16 && (long)gy * gz * ((ctx->n0 + xchunk - 1) / xchunk) < 48L * FANS_SMS) xchunk = (xchunk + 1) / 2;
    p.xchunk = xchunk;
    dim3 grid(gz, gy, (ctx->n0 + xchunk - 1) / xchunk);
    const size_t smem = sizeof(double) * 4 * h * GTILE + sizeof(uint16_t) * 3 * GETILE + 16;
    prof_begin(ctx, PC_SWEEP_LINEAR);
    int rc = FANS_ERR_ARG;
#define ST_CASE(H_, Q_) \
    if (h == H_ && ctx->n_k == Q_) rc = launch_stencil<H_, Q_>(ctx, p, grid, smem);
    ST_CASE(1, 1) ST_CASE(1, 2) ST_CASE(1, 3) ST_CASE(1, 4) ST_CASE(3, 1) ST_CASE(3, 2) ST_CASE(3, 3) ST_CASE(3, 4)
#undef ST_CASE
    prof_end(ctx);
    ctx->launches++;
    if (rc != FANS_OK) {
        fans_set_error(ctx, rc, "stencil_run: unsupported configuration");
        return rc;
    }
    CUDA_TRY(ctx, cudaGetLastError());
    if (red_out && ctx->P > 1) FANS_CHECK(comm_allreduce(ctx, red_out, red_out, 1, false));
    return FANS_OK;
}
