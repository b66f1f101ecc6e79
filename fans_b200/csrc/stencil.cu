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
// fused update d = s + beta*d is formed on the fly while loading.
// Latency: the raw s / d values of plane k+2 travel global -> shared with cp.async (thread-private staging slots, no
// registers held) while the CTA computes plane k; they are combined into the ring one step later, so no global-load
// latency sits on the critical path of the march.  Isotropic phases (every S[delta][i][j], i != j, vanishes unless
// delta_i != 0 and delta_j != 0 — checked numerically on the host) take a 153-DFMA variant of the 27-point stencil.
#include "internal.h"
#include "materials.cuh"
#include "stencil.h"

#define GY 8
#define GZ 64
#define GPZ (GZ + 4)                 // node tile columns z0-2 .. z0+GZ+1: every 16-byte pair of the field stays a pair in the tile
#define GTILE ((GY + 2) * GPZ)
#define GPAIRS (GTILE / 2)           // 16-byte pairs per component plane tile
#define GETILE ((GY + 1) * (GZ + 1))
#define G_CONS 256                   // consumer threads: one node pair each
#define G_PROD 128                   // producer threads (one warpgroup, runs on a reduced register budget)
#define G_THREADS (G_CONS + G_PROD)
#ifndef STENCIL_MINB
#define STENCIL_MINB 2
#endif
#ifndef STENCIL_WROWS
#define STENCIL_WROWS 1         // tile rows a consumer warp covers (1, 2, 4, 8): see the consumer branch of k_stencil_linear
#endif
#ifndef STENCIL_MIXED_TABLE
#define STENCIL_MIXED_TABLE 0   // 1: warps cut by an interface take per-lane coefficients from the table in global memory (measured slower)
#endif

// The 27-point block stencils of up to STENCIL_MAXQ phases travel as a KERNEL PARAMETER (__grid_constant__, constant bank 0): no
// process-global __constant__ table, nothing shared between contexts.  More phases (a polycrystal with one triclinic tensor per
// grain, LinearElastic.h:77-159) take the NQ = 0 instantiation, which reads the coefficients of a lane's phase from the per-context
// table in global memory (L1-resident; the lanes of a warp mostly share the address).
template <int H, int NQ>
struct StencilCoef {
    double S[(NQ > 0 ? NQ : 1) * 27 * H * H];   // [q][delta][i][j]
};

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
    const double *Ktab;    // phase stiffness table [nq][(8H)^2] in global memory (interface path)
    const double *Stab;    // 27-point block stencils [nq][27][H][H] in global memory (warps cut by an interface)
    // slab decomposition (world_size > 1): node planes -1 and n0 and element plane -1 come from the neighbour ranks
    const double *halo_lo, *halo_hi;   // [H][ny*nz], already the UPDATED direction (no s/beta combination)
    const uint16_t *ms_lo;             // [ny*nz] phases of element plane -1
    // batched solves (k_stencil_linear<..., BATCH = true>): gridDim.z = lanes * nchunk; the fields of lane l start laneF doubles
    // after those of lane l-1, its beta / red_out S_COUNT doubles, its partial sums lane_part doubles, its ticket one word
    int nchunk;
    size_t laneF, lane_part;
};

__device__ __forceinline__ int gwrap(int v, int n)
{
    v %= n;
    return v < 0 ? v + n : v;
}

// homogeneous path, x-scatter form: the contribution of ONE row (plane P, tile row ry+dy-1; v = its values at rz0..rz0+3 for
// every component) to the node pair of output plane o = P - (DXI-1), phase Q (compile time => constant-bank operands).
// Each plane row is read from shared memory once and feeds the accumulators of the three output planes it touches.
template <int H, int Q, bool ISO, int DXI, int DY>
__device__ __forceinline__ void stencil_row(const double *__restrict__ Sall, const double (&v)[H][4], double (&accA)[H], double (&accB)[H])
{
#pragma unroll
    for (int dz = 0; dz < 3; ++dz) {
        const double *S = Sall + ((Q * 27 + (DXI * 9 + DY * 3 + dz)) * H * H);
#pragma unroll
        for (int i = 0; i < H; ++i)
#pragma unroll
            for (int j = 0; j < H; ++j) {
                if (ISO && H == 3 && i != j) {  // odd in delta_i and delta_j: zero on the centre lines
                    const int dd[3] = {DXI, DY, dz};
                    if (dd[i] == 1 || dd[j] == 1) continue;
                }
                accA[i] = fma(S[i * H + j], v[j][dz], accA[i]);
                accB[i] = fma(S[i * H + j], v[j][dz + 1], accB[i]);
            }
    }
}
// the same row for a warp whose lanes sit in DIFFERENT phases (the interface cuts through the warp): coefficients per lane from the
// stencil table in global memory (L1-resident) instead of executing every phase's constant-operand copy one after the other
template <int H, bool ISO, int DXI, int DY>
__device__ __forceinline__ void stencil_row_table(const double *__restrict__ Sq, const double (&v)[H][4], double (&accA)[H], double (&accB)[H])
{
#pragma unroll
    for (int dz = 0; dz < 3; ++dz) {
        const double *S = Sq + (DXI * 9 + DY * 3 + dz) * H * H;
#pragma unroll
        for (int i = 0; i < H; ++i)
#pragma unroll
            for (int j = 0; j < H; ++j) {
                if (ISO && H == 3 && i != j) {
                    const int dd[3] = {DXI, DY, dz};
                    if (dd[i] == 1 || dd[j] == 1) continue;
                }
                const double c = __ldg(S + i * H + j);
                accA[i] = fma(c, v[j][dz], accA[i]);
                accB[i] = fma(c, v[j][dz + 1], accB[i]);
            }
    }
}
template <int H, int NQ, bool ISO, int DXI, int DY>
__device__ __forceinline__ void stencil_row_dispatch(const double *__restrict__ S, int ph, const double (&v)[H][4], double (&accA)[H], double (&accB)[H])
{
    if (NQ > 0 && ph == 0) stencil_row<H, 0, ISO, DXI, DY>(S, v, accA, accB);
    else if (NQ > 1 && ph == 1) stencil_row<H, (NQ > 1 ? 1 : 0), ISO, DXI, DY>(S, v, accA, accB);
    else if (NQ > 2 && ph == 2) stencil_row<H, (NQ > 2 ? 2 : 0), ISO, DXI, DY>(S, v, accA, accB);
    else if (NQ > 3 && ph == 3) stencil_row<H, (NQ > 3 ? 3 : 0), ISO, DXI, DY>(S, v, accA, accB);
}

// Interface nodes of plane k (queued by the consumers), exact element form (reference semantics: K_phase(element) (u_b - u_node0),
// include/solver.h:250-261, solverCG.h:98-103).  One (node, element) item per lane: the eight lanes of a node read their element's
// three rows of K from the phase table in global memory (no phase branches, no divergence) and are summed with shuffles, so a queue
// of q nodes costs ceil(8 q / 256) short passes of all consumer warps instead of one long serial chain in a few lanes.
// Not inlined on purpose: its registers must not compete with the accumulators of the homogeneous path.  (Measured, round 2: issuing
// its loads in batches — all 8 x H neighbour values, then K row by row — instead of the load -> subtract -> multiply chains below needs
// a stack frame for the callee-saved registers and slows the WHOLE kernel, homogeneous image included: 3.04 -> 3.35 ms.)
struct IfaceArgs {
    const double *ring, *Ktab;
    const uint16_t *mring, *qlist;
    const int *qcnt;
    double *out;
    size_t nloc;
    int ny, nz, y0, z0, k;
};
template <int H, int NCONS>
__device__ __noinline__ double interface_phase(const IfaceArgs a, int tid)
{
    constexpr int ND = 8 * H;
    int pre[GY + 1];
    pre[0] = 0;
#pragma unroll
    for (int w = 0; w < GY; ++w) pre[w + 1] = pre[w] + a.qcnt[w];
    const int nitems = 16 * pre[GY];  // pair x 2 nodes x 8 elements
    double dot = 0.0;
    for (int n0 = 0; n0 < nitems; n0 += NCONS) {
        const int n = n0 + tid;
        const bool on = n < nitems;
        double acc[H];
#pragma unroll
        for (int c = 0; c < H; ++c) acc[c] = 0.0;
        int qry = 0, qrz = 0;
        if (on) {
            const int pair = n >> 4, A = n & 7;
            int w = 0, base = 0;
#pragma unroll
            for (int q = 1; q < GY; ++q)
                if (pair >= pre[q]) w = q, base = pre[q];
            const int code = a.qlist[w * 32 + (pair - base)];
            qry = code >> 7, qrz = (code & 127) + ((n >> 3) & 1);
            const int ox = A & 1, oy = (A >> 1) & 1, oz = (A >> 2) & 1;
            const int ph = a.mring[((a.k - ox + 8) & 7) * GETILE + (qry - oy) * (GZ + 1) + (qrz - 1 - oz)];
            const double *Krow = a.Ktab + (size_t)ph * (ND * ND) + (size_t)(H * A) * ND;
            const double *p0 = a.ring + (size_t)((a.k - ox + 4) & 3) * H * GTILE + (qry - oy) * GPZ + (qrz - oz);
            const double *p1 = a.ring + (size_t)((a.k - ox + 5) & 3) * H * GTILE + (qry - oy) * GPZ + (qrz - oz);
            double u0[H];
#pragma unroll
            for (int c = 0; c < H; ++c) u0[c] = p0[c * GTILE];
#pragma unroll
            for (int b = 1; b < 8; ++b) {
                const double *pb = ((b & 1) ? p1 : p0) + ((b >> 1) & 1) * GPZ + ((b >> 2) & 1);
                double w3[H];
#pragma unroll
                for (int c = 0; c < H; ++c) w3[c] = pb[c * GTILE] - u0[c];
#pragma unroll
                for (int i = 0; i < H; ++i)
#pragma unroll
                    for (int j = 0; j < H; ++j) acc[i] = fma(__ldg(Krow + i * ND + H * b + j), w3[j], acc[i]);
            }
        }
#pragma unroll
        for (int c = 0; c < H; ++c) {
            acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
            acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
            acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 4);
        }
        if (on && (n & 7) == 0) {
            const size_t g = ((size_t)a.k * a.ny + (a.y0 + qry - 1)) * a.nz + (a.z0 + qrz - 2);
            const double *ctr = a.ring + (size_t)((a.k + 4) & 3) * H * GTILE + qry * GPZ + qrz;
#pragma unroll
            for (int c = 0; c < H; ++c) {
                a.out[c * a.nloc + g] = acc[c];
                dot += acc[c] * ctr[c * GTILE];
            }
        }
    }
    return dot;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
// named barriers (ids 1..15; 0 is __syncthreads): producer/consumer hand-over of ring slots without stalling the whole CTA
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(count) : "memory"); }
// warp-specialised register budget (setmaxnreg, sm_90a+): the kernel is launched with 80 registers per thread (2 CTAs x 384 threads);
// the producer warpgroup hands 40 of them back, the two consumer warpgroups take 16 more each — the consumer step (18 accumulators,
// 12 staged inputs, 3 homogeneity flags) no longer spills to local memory in the march
#ifndef STENCIL_SETMAXNREG
#define STENCIL_SETMAXNREG 1
#endif
#ifndef STENCIL_DEFER_IFACE
#define STENCIL_DEFER_IFACE 1   // 0: consumer-only barrier + interface phase at the end of every step (A/B builds)
#endif
#ifndef STENCIL_PROD_REGS
#define STENCIL_PROD_REGS 40
#endif
#ifndef STENCIL_CONS_REGS
#define STENCIL_CONS_REGS 96
#endif
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(N)); }
#define BAR_FULL 1    // +slot: plane (and the element plane after it) is in the ring
#define BAR_EMPTY 5   // +slot: the consumers are done with the plane that lived in the slot
#define BAR_CONS 9    // consumer-only barrier (interface queue)

#define G_NLP ((GPAIRS + G_PROD - 1) / G_PROD)   // pair items per producer thread and component
#define G_NLM ((GETILE + G_PROD - 1) / G_PROD)   // phase-image items per producer thread
#define G_STG (G_NLP * G_PROD)                   // staging pairs per component

// Producer warpgroup of the stencil kernels (`lane` = producer thread index 0..G_PROD-1, GT = threads of the CTA = barrier count):
// raw s / d_old pairs of plane Q+1 travel global -> staging with cp.async while plane Q is combined (d = s + beta d_old) into ring
// slot Q&3, d_new is stored and the phase-image ring refilled.
template <int H, int GT>
__device__ __forceinline__ void stencil_producer(const StencilParams &p, double *ring, double2 *stg, uint16_t *mring, int lane, int y0, int z0,
                                                 int xs, int xe, size_t plane_sz)
{
    const double beta = p.s ? *p.beta : 0.0;
    int goff[G_NLP], roff[G_NLP], moff[G_NLM];
    unsigned mine_mask = 0;
#pragma unroll
    for (int j = 0; j < G_NLP; ++j) {
        const int i = lane + j * G_PROD;
        goff[j] = -1, roff[j] = 0;
        if (i < GPAIRS) {
            const int r = i / (GPZ / 2), q = i % (GPZ / 2);
            goff[j] = gwrap(y0 - 1 + r, p.ny) * p.nz + gwrap(z0 - 2 + 2 * q, p.nz);
            roff[j] = r * GPZ + 2 * q;
            if (r >= 1 && r <= GY && q >= 1 && q <= GZ / 2 && (y0 - 1 + r) < p.ny && (z0 - 2 + 2 * q) < p.nz) mine_mask |= 1u << j;
        }
    }
#pragma unroll
    for (int j = 0; j < G_NLM; ++j) {
        const int i = lane + j * G_PROD;
        moff[j] = -1;
        if (i < GETILE) {
            const int r = i / (GZ + 1), c = i % (GZ + 1);
            moff[j] = gwrap(y0 - 1 + r, p.ny) * p.nz + gwrap(z0 - 1 + c, p.nz);
        }
    }
    auto issue = [&](int xp) {
        const bool lo = (p.halo_lo && xp < 0), hi = (p.halo_hi && xp >= p.n0);
        const double *hal = lo ? p.halo_lo : (hi ? p.halo_hi : nullptr);
        const size_t gbase = (xp < 0 ? (size_t)(xp + p.n0) : (xp >= p.n0 ? (size_t)(xp - p.n0) : (size_t)xp)) * plane_sz;
        const double *srcd = hal ? hal : p.d_old + gbase;
        const size_t cs = hal ? plane_sz : p.nloc;
#pragma unroll
        for (int cc = 0; cc < H; ++cc)
#pragma unroll
            for (int j = 0; j < G_NLP; ++j)
                if (goff[j] >= 0) cp_async16(stg + cc * G_STG + lane + j * G_PROD, srcd + cc * cs + goff[j]);
        if (p.s && !hal) {
#pragma unroll
            for (int cc = 0; cc < H; ++cc)
#pragma unroll
                for (int j = 0; j < G_NLP; ++j)
                    if (goff[j] >= 0) cp_async16(stg + (H + cc) * G_STG + lane + j * G_PROD, p.s + cc * p.nloc + gbase + goff[j]);
        }
    };
    auto combine = [&](int xp, bool owned) {
        const bool hal = (p.halo_lo && xp < 0) || (p.halo_hi && xp >= p.n0);
        const bool upd = p.s && !hal;
        const size_t gbase = (xp < 0 ? (size_t)(xp + p.n0) : (xp >= p.n0 ? (size_t)(xp - p.n0) : (size_t)xp)) * plane_sz;
        double *pl = ring + (size_t)((xp + 4) & 3) * H * GTILE;
#pragma unroll
        for (int cc = 0; cc < H; ++cc)
#pragma unroll
            for (int j = 0; j < G_NLP; ++j)
                if (goff[j] >= 0) {
                    double2 v = stg[cc * G_STG + lane + j * G_PROD];
                    if (upd) {
                        const double2 sv = stg[(H + cc) * G_STG + lane + j * G_PROD];
                        v.x = sv.x + beta * v.x, v.y = sv.y + beta * v.y;
                        if (owned && ((mine_mask >> j) & 1u)) *reinterpret_cast<double2 *>(p.d_new + cc * p.nloc + gbase + goff[j]) = v;
                    }
                    *reinterpret_cast<double2 *>(pl + cc * GTILE + roff[j]) = v;
                }
    };
    uint16_t msr[G_NLM];
    auto ms_fetch = [&](int xp) {
        const uint16_t *src = (p.ms_lo && xp < 0) ? p.ms_lo : p.phidx + (size_t)gwrap(xp, p.n0) * plane_sz;
#pragma unroll
        for (int j = 0; j < G_NLM; ++j) msr[j] = (moff[j] >= 0) ? __ldg(src + moff[j]) : (uint16_t)0;
    };
    auto ms_put = [&](int xp) {
        uint16_t *mp = mring + ((xp + 8) & 7) * GETILE;
#pragma unroll
        for (int j = 0; j < G_NLM; ++j)
            if (moff[j] >= 0) mp[lane + j * G_PROD] = msr[j];
    };
    issue(xs - 1);
    ms_fetch(xs - 1);
    ms_put(xs - 1);
    ms_fetch(xs);
    for (int Q = xs - 1; Q <= xe; ++Q) {
        cp_async_wait_all();
        if (Q >= xs + 3) bar_sync(BAR_EMPTY + (Q & 3), GT);  // consumers have left plane Q-4
        combine(Q, Q >= xs && Q < xe);
        ms_put(Q + 1);
        bar_arrive(BAR_FULL + (Q & 3), GT);
        if (Q + 1 <= xe) {
            issue(Q + 1);   // lands while the consumers work on plane Q
            ms_fetch(Q + 2);
        }
    }
}

// Warp-specialised march along x.  Producer warpgroup (warps 8..11): raw s / d_old pairs of plane Q+1 travel global -> staging with
// cp.async while it combines plane Q (d = s + beta d_old) into ring slot Q&3, stores d_new and refills the phase-image ring.
// Consumer warps (0..7): step P reads plane P ONCE from the ring and scatters it into the accumulators of the output planes
// P-1, P, P+1 (x-scatter form of the 27-point block stencil); plane P-1 is then complete.  Interface nodes (mixed-phase
// neighbourhood) are queued and evaluated in the exact element form.  Hand-over through named barriers per ring slot, so
// neither side ever waits for the other unless it is genuinely ahead.
template <int H, int NQ, bool ISO, bool BATCH>
__global__ void __launch_bounds__(G_THREADS, STENCIL_MINB) k_stencil_linear(const StencilParams p0, const __grid_constant__ StencilCoef<H, NQ> coef)
{
    // BATCH: this CTA belongs to lane blockIdx.z / nchunk; the single-problem instantiation is untouched by it (p aliases p0)
    StencilParams pl;
    int bz = blockIdx.z;
    if (BATCH) {
        const int ln = bz / p0.nchunk;
        bz -= ln * p0.nchunk;
        pl = p0;
        const size_t fo = (size_t)ln * p0.laneF;
        pl.d_old += fo, pl.d_new += fo, pl.out += fo;
        if (pl.s) pl.s += fo, pl.beta += (size_t)ln * S_COUNT;
        if (pl.red_out) pl.red_out += (size_t)ln * S_COUNT, pl.part += (size_t)ln * p0.lane_part, pl.ticket += ln;
    }
    const StencilParams &p = BATCH ? pl : p0;
    const double *__restrict__ Sc = coef.S;
    extern __shared__ __align__(16) double smem[];
    double *ring = smem;                                   // [4][H][GTILE]   combined direction planes
    double2 *stg = reinterpret_cast<double2 *>(smem + 4 * H * GTILE);  // [2][H][G_STG] raw d_old / s pairs in flight
    uint16_t *mring = reinterpret_cast<uint16_t *>(stg + 2 * H * G_STG);  // [8][GETILE]
    __shared__ double scratch[32];
    __shared__ uint16_t qlist[2][GY * 32];  // queued interface node pairs, per warp row, double-buffered by step parity
    __shared__ int qcnt[2][GY];

    const int tid = threadIdx.x, lane = tid & 31, wy = tid >> 5;
    const int z0 = blockIdx.x * GZ, y0 = blockIdx.y * GY;
    const int xs = bz * p.xchunk, xe = min(xs + p.xchunk, p.n0);
    const size_t plane_sz = (size_t)p.ny * p.nz;
    double dot_final = 0.0;   // <d_new, K d_new> of this thread; the running sum lives inside the consumer branch only (a value that
                              // is live across the producer branch would have to fit the producers' 40-register budget and gets spilled)

    if (wy >= GY) {
        // =================================== producer warpgroup ===================================
        if (STENCIL_SETMAXNREG && STENCIL_MINB == 2) reg_dealloc<STENCIL_PROD_REGS>();
        stencil_producer<H, G_THREADS>(p, ring, stg, mring, tid - G_CONS, y0, z0, xs, xe, plane_sz);
    } else {
        // =================================== consumer warps ===================================
        if (STENCIL_SETMAXNREG && STENCIL_MINB == 2) reg_alloc<STENCIL_CONS_REGS>();
        double dotacc = 0.0;
        // footprint of a warp in the (y, z) tile: STENCIL_WROWS rows x 64 / STENCIL_WROWS nodes.  A warp whose homogeneous lanes sit in
        // different phases runs the constant-operand stencil of each of them, and the whole CTA waits for it at the hand-over: the
        // squarer the footprint, the fewer warps an interface cuts (1 x 64: one per crossing of a z line)
        const int wrow = (wy % (GY / STENCIL_WROWS)) * STENCIL_WROWS + lane / (32 / STENCIL_WROWS);   // tile row 0..GY-1
        const int wzp = (wy / (GY / STENCIL_WROWS)) * (32 / STENCIL_WROWS) + lane % (32 / STENCIL_WROWS);   // node pair 0..31 along z
        const int ry = wrow + 1, rzA = 2 * wzp + 2;  // tile coordinates of node A (column = z - z0 + 2); node B = rzA + 1
        const int rzM = 2 * wzp + 1;                 // column of node A in the phase-image tile (z - z0 + 1)
        const int yA = y0 + wrow, zA = z0 + 2 * wzp;
        const bool valid = (yA < p.ny) && (zA < p.nz);  // nz is even, so A valid <=> B valid
        // accumulators of the output planes P-1, P, P+1 (slots 0,1,2) of this thread's node pair, with their homogeneity flags
        double acc[3][2][H];
        int hph[3] = {-1, -1, -1};  // phase of a homogeneous neighbourhood, -1: interface pair / outside the chunk / invalid thread
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int c = 0; c < H; ++c) acc[j][0][c] = 0.0, acc[j][1][c] = 0.0;
        // interface nodes of plane k (queued in buffer par): exact element form over planes k-1..k+1
        auto run_interface = [&](int k, int par) -> double {
            int any = 0;
#pragma unroll
            for (int w = 0; w < GY; ++w) any |= qcnt[par][w];
            if (!any) return 0.0;
            IfaceArgs ia;
            ia.ring = ring, ia.Ktab = p.Ktab, ia.mring = mring, ia.qlist = qlist[par], ia.qcnt = qcnt[par], ia.out = p.out;
            ia.nloc = p.nloc, ia.ny = p.ny, ia.nz = p.nz, ia.y0 = y0, ia.z0 = z0, ia.k = k;
            return interface_phase<H, G_CONS>(ia, tid);
        };
        for (int P = xs - 1; P <= xe; ++P) {
            bar_sync(BAR_FULL + (P & 3), G_THREADS);  // plane P and element plane P+1 are in the rings
#if STENCIL_DEFER_IFACE
            // The hand-over barrier above is also the consumers' own barrier: everybody has finished step P-1, so the interface queue
            // of plane P-2 (filled during step P-1) is complete.  Its nodes are evaluated now, over the planes P-3..P-1 that are still
            // in the ring; only then does the slot of plane P-3 go back to the producer.  No consumer-only barrier is left in the march.
            if (P - 2 >= xs) dotacc += run_interface(P - 2, (P - 1) & 1);
            if (P - 1 >= xs + 1 && P + 1 <= xe) bar_arrive(BAR_EMPTY + ((P + 1) & 3), G_THREADS);
#endif
            // ---- phase 1
            const int o2 = P + 1;  // newest output plane: classify its node pair from the 12 elements around it (planes o2-1, o2)
            hph[2] = -1;
            if (valid && o2 >= xs && o2 < xe) {
                const uint16_t *m0 = mring + ((o2 - 1 + 8) & 7) * GETILE, *m1 = mring + ((o2 + 8) & 7) * GETILE;
                const int e00 = (ry - 1) * (GZ + 1) + rzM - 1, e10 = ry * (GZ + 1) + rzM - 1;
                const int ph = m1[e10 + 1];
                bool homog = true;
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    homog = homog && (m0[e00 + c] == ph) && (m0[e10 + c] == ph) && (m1[e00 + c] == ph) && (m1[e10 + c] == ph);
                hph[2] = homog ? ph : -1;
            }
            const double *pl = ring + (size_t)((P + 4) & 3) * H * GTILE;
            // bit j: the homogeneous lanes of this warp do not all share one phase for output plane j (warp-uniform decision)
            int mixed = 0;
            if (STENCIL_MIXED_TABLE && NQ > 1) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const unsigned act = __ballot_sync(0xffffffffu, hph[j] >= 0);
                    const int first = __shfl_sync(0xffffffffu, hph[j], act ? (__ffs(act) - 1) : 0);
                    if (act && !__all_sync(0xffffffffu, hph[j] < 0 || hph[j] == first)) mixed |= 1 << j;
                }
            }
            if (hph[0] >= 0 || hph[1] >= 0 || hph[2] >= 0) {
                double v[H][4];
#define ST_ROW(DY)                                                                                              \
    {                                                                                                           \
        _Pragma("unroll") for (int c = 0; c < H; ++c)                                                           \
        {                                                                                                       \
            const double2 *p2 = reinterpret_cast<const double2 *>(pl + c * GTILE + (ry + DY - 1) * GPZ + rzA - 2); \
            const double2 a = p2[0], b = p2[1], d = p2[2];                                                      \
            v[c][0] = a.y, v[c][1] = b.x, v[c][2] = b.y, v[c][3] = d.x;                                         \
        }                                                                                                       \
        if (hph[0] >= 0) {                                                                                      \
            if (NQ == 0 || (STENCIL_MIXED_TABLE && (mixed & 1))) stencil_row_table<H, ISO, 2, DY>(p.Stab + hph[0] * (27 * H * H), v, acc[0][0], acc[0][1]); \
            else stencil_row_dispatch<H, NQ, ISO, 2, DY>(Sc, hph[0], v, acc[0][0], acc[0][1]);                       \
        }                                                                                                       \
        if (hph[1] >= 0) {                                                                                      \
            if (NQ == 0 || (STENCIL_MIXED_TABLE && (mixed & 2))) stencil_row_table<H, ISO, 1, DY>(p.Stab + hph[1] * (27 * H * H), v, acc[1][0], acc[1][1]); \
            else stencil_row_dispatch<H, NQ, ISO, 1, DY>(Sc, hph[1], v, acc[1][0], acc[1][1]);                       \
        }                                                                                                       \
        if (hph[2] >= 0) {                                                                                      \
            if (NQ == 0 || (STENCIL_MIXED_TABLE && (mixed & 4))) stencil_row_table<H, ISO, 0, DY>(p.Stab + hph[2] * (27 * H * H), v, acc[2][0], acc[2][1]); \
            else stencil_row_dispatch<H, NQ, ISO, 0, DY>(Sc, hph[2], v, acc[2][0], acc[2][1]);                       \
        }                                                                                                       \
    }
                ST_ROW(0)
                ST_ROW(1)
                ST_ROW(2)
#undef ST_ROW
            }
            const int k = P - 1;  // the output plane completed by this step
            const bool kin = (k >= xs);
            const int par = P & 1;
            const unsigned qmask = __ballot_sync(0xffffffffu, valid && kin && hph[0] < 0);
            if (valid && kin && hph[0] >= 0) {
                const size_t g = ((size_t)k * p.ny + yA) * p.nz + zA;
                const double *ctr = ring + (size_t)((k + 4) & 3) * H * GTILE + ry * GPZ + rzA;
#pragma unroll
                for (int c = 0; c < H; ++c) {
                    *reinterpret_cast<double2 *>(p.out + c * p.nloc + g) = make_double2(acc[0][0][c], acc[0][1][c]);
                    const double2 cv = *reinterpret_cast<const double2 *>(ctr + c * GTILE);
                    dotacc += acc[0][0][c] * cv.x + acc[0][1][c] * cv.y;
                }
            }
            if (qmask) {
                if (valid && kin && hph[0] < 0) qlist[par][wy * 32 + __popc(qmask & ((1u << lane) - 1u))] = (uint16_t)(ry * 128 + rzA);
            }
            if (lane == 0) qcnt[par][wy] = __popc(qmask);
            // rotate the output window
#pragma unroll
            for (int c = 0; c < H; ++c) {
                acc[0][0][c] = acc[1][0][c], acc[0][1][c] = acc[1][1][c];
                acc[1][0][c] = acc[2][0][c], acc[1][1][c] = acc[2][1][c];
                acc[2][0][c] = 0.0, acc[2][1][c] = 0.0;
            }
            hph[0] = hph[1], hph[1] = hph[2];
#if !STENCIL_DEFER_IFACE
            bar_sync(BAR_CONS, G_CONS);
            dotacc += run_interface(k, par);
            // plane P-2 is not needed any more: its slot may take plane P+2
            if (P >= xs + 1 && P + 2 <= xe) bar_arrive(BAR_EMPTY + ((P + 2) & 3), G_THREADS);
#endif
        }
#if STENCIL_DEFER_IFACE
        bar_sync(BAR_CONS, G_CONS);          // the queue of the last plane is complete
        dotacc += run_interface(xe - 1, xe & 1);
#endif
        dot_final = dotacc;
    }
    if (p.red_out) {
        if (BATCH)
            grid_reduce_sum1_part(dot_final, scratch, p.part, p.ticket, p.red_out, gridDim.x * gridDim.y * p.nchunk,
                                  blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * bz));
        else grid_reduce_sum1(dot_final, scratch, p.part, p.ticket, p.red_out);
    }
}

// ------------------------------------------------------------------------------------------------
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

// every all-linear problem: up to STENCIL_MAXQ phases with constant-bank coefficients, more through the coefficient table
// (the node pairs of a thread and the 16-byte row copies need an even n_z; odd n_z takes the element sweep)
bool stencil_supported(const fans_ctx *ctx) { return ctx->all_linear && ctx->n_k >= 1 && ctx->n_k == ctx->n_phases && ctx->nz % 2 == 0; }

template <int H, int NQ>
static int launch_stencil(fans_ctx *ctx, const StencilParams &p, dim3 grid, size_t smem, bool iso, const std::vector<double> &S)
{
    StencilCoef<H, NQ> coef;
    for (size_t i = 0; i < sizeof(coef.S) / sizeof(double); ++i) coef.S[i] = i < S.size() ? S[i] : 0.0;
#define ST_LAUNCH(ISO_, B_)                                                                                                          \
    do {                                                                                                                             \
        CUDA_TRY(ctx, cudaFuncSetAttribute(k_stencil_linear<H, NQ, ISO_, B_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_stencil_linear<H, NQ, ISO_, B_><<<grid, G_THREADS, smem, ctx->st>>>(p, coef);                                                 \
    } while (0)
    const bool batch = ctx->nb > 1;
    if (iso && H == 3) {
        if (batch) ST_LAUNCH((H == 3), true);
        else ST_LAUNCH((H == 3), false);
    } else {
        if (batch) ST_LAUNCH(false, true);
        else ST_LAUNCH(false, false);
    }
#undef ST_LAUNCH
    return FANS_OK;
}

// true when every off-diagonal coefficient S[delta][i][j] (i != j) with delta_i == 0 or delta_j == 0 vanishes (to rounding) for all
// phases: holds for isotropic / cubic / orthotropic-aligned stiffness with any of the element types; the kernel then skips them.
static bool stencil_iso_pattern(int h, int nq, const double *S)
{
    if (h != 3) return false;
    double mx = 0.0, off = 0.0;
    for (int q = 0; q < nq; ++q)
        for (int di = 0; di < 27; ++di) {
            const int dd[3] = {di / 9, (di / 3) % 3, di % 3};
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    const double v = fabs(S[((q * 27 + di) * 3 + i) * 3 + j]);
                    mx = fmax(mx, v);
                    if (i != j && (dd[i] == 1 || dd[j] == 1)) off = fmax(off, v);
                }
        }
    return off <= 1e-13 * mx;
}

int stencil_run(fans_ctx *ctx, const double *d_old, double *out, const double *s_in, double *d_new, const double *beta_dev,
                double *red_out)
{
    const int h = ctx->h, nd = 8 * h;
    if (ctx->stencil_stamp != ctx->const_stamp) {   // (re)build this context's tables after fans_set_materials
        std::vector<double> &S = ctx->S_host;
        S.assign((size_t)ctx->n_k * 27 * h * h, 0.0);
        for (int q = 0; q < ctx->n_k; ++q) stencil_from_element_matrix(h, ctx->K_host.data() + (size_t)q * nd * nd, S.data() + (size_t)q * 27 * h * h);
        if (ctx->d_Stab) cudaFree(ctx->d_Stab), ctx->d_Stab = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_Stab, sizeof(double) * S.size()));
        CUDA_TRY(ctx, cudaMemcpy(ctx->d_Stab, S.data(), sizeof(double) * S.size(), cudaMemcpyHostToDevice));
        const char *env = getenv("FANS_STENCIL_ISO");
        ctx->stencil_iso = stencil_iso_pattern(h, ctx->n_k, S.data()) && !(env && env[0] == '0');
        ctx->stencil_stamp = ctx->const_stamp;
    }
    StencilParams p;
    memset(&p, 0, sizeof(p));
    p.n0 = ctx->n0, p.ny = ctx->ny, p.nz = ctx->nz, p.nloc = ctx->nloc;
    p.d_old = d_old, p.s = s_in, p.d_new = d_new, p.beta = beta_dev, p.out = out;
    p.phidx = ctx->phidx;
    p.Ktab = ctx->d_K;
    p.Stab = ctx->d_Stab;
    p.nq = ctx->n_k;
    p.part = ctx->d_part, p.ticket = ctx->d_ticket, p.red_out = red_out;
    if (ctx->P > 1) {
        // halo planes of the UPDATED direction: pack planes 0 / n0-1 (forming s + beta d on the fly), ring exchange
        FANS_CHECK(halo_exchange_both(ctx, d_old, s_in, beta_dev));
        p.halo_lo = ctx->halo_lo, p.halo_hi = ctx->halo_hi, p.ms_lo = ctx->ms_lo;
    }
    const int gy = (ctx->ny + GY - 1) / GY, gz = (ctx->nz + GZ - 1) / GZ;
    int xchunk = pick_xchunk(ctx->n0, (long)gy * gz * ctx->nb, 2L * FANS_SMS, 2);   // 2 resident CTAs per SM, 2 run-in planes per march
    if (const char *env = getenv("FANS_STENCIL_CTAS")) {   // experiments: at least this many CTAs
        const long want = atol(env);
        xchunk = ctx->n0;
        while (xchunk > 2 && (long)gy * gz * ((ctx->n0 + xchunk - 1) / xchunk) < want) xchunk = (xchunk + 1) / 2;
    }
    p.xchunk = xchunk;
    p.nchunk = (ctx->n0 + xchunk - 1) / xchunk;
    p.laneF = (size_t)ctx->h * ctx->nloc;
    p.lane_part = (size_t)gy * gz * p.nchunk;
    dim3 grid(gz, gy, p.nchunk * ctx->nb);
    const size_t smem = sizeof(double) * 4 * h * GTILE + sizeof(double2) * 2 * h * G_STG + sizeof(uint16_t) * 8 * GETILE + 16;
    prof_begin(ctx, PC_SWEEP_LINEAR);
    int rc = FANS_ERR_ARG;
#define ST_CASE(H_, Q_) \
    if (h == H_ && (Q_ > 0 ? ctx->n_k == Q_ : ctx->n_k > STENCIL_MAXQ)) rc = launch_stencil<H_, Q_>(ctx, p, grid, smem, ctx->stencil_iso, ctx->S_host);
    ST_CASE(1, 1) ST_CASE(1, 2) ST_CASE(1, 3) ST_CASE(1, 4) ST_CASE(3, 1) ST_CASE(3, 2) ST_CASE(3, 3) ST_CASE(3, 4) ST_CASE(1, 0) ST_CASE(3, 0)
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
