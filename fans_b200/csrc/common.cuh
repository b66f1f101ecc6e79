// common.cuh — internal context, error handling and device reduction helpers of libfans_gpu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>
#include "../../include/fans_gpu.h"

#define FANS_PROF_CLASSES 16
#define FANS_SMS 148  // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of this

#define CUDA_TRY(ctx, expr)                                                                                    \
    do {                                                                                                       \
        cudaError_t _e = (expr);                                                                               \
        if (_e != cudaSuccess) {                                                                               \
            fans_set_error((ctx), FANS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
            return FANS_ERR_CUDA;                                                                              \
        }                                                                                                      \
    } while (0)

#define FANS_CHECK(call)            \
    do {                            \
        int _rc = (call);           \
        if (_rc != FANS_OK) return _rc; \
    } while (0)

// ---- mixed-radix power-of-two FFT plan (DIF forward: natural -> digit-reversed; DIT inverse back) ----
struct FftPlan {
    int      N      = 0;  // transform length (complex points)
    int      nst    = 0;  // number of stages
    int      radix[8];    // radix per stage (8,4,2), forward order
    int      ntab   = 0;  // length of the twiddle table tw (>= N, multiple of N)
    double2 *tw     = nullptr;  // device: tw[i] = exp(-2 pi i * i / ntab)
    int     *pos    = nullptr;  // device: pos[f] = storage row of frequency f after the forward DIF
    std::vector<int> pos_host;
};

// Bluestein plan of one axis for grids that are not powers of two (fft_any.cu)
struct AnyPlan {
    int n = 0, M = 0, logM = 0;   // transform length, power-of-two convolution length >= 2n-1
    double2 *w = nullptr;         // device: chirp exp(-i pi k^2 / n), k < n
    double2 *Bhat = nullptr;      // device: FFT_M of the wrapped conjugate chirp / M, bit-reversed order
    double2 *tw = nullptr;        // device: exp(-2 pi i j / M), j < M/2
};

struct PhaseDev {  // device copy of one fans_phase_desc (params trimmed)
    int    model, local_mat, group_n_mat, k_index;  // k_index: slot of the phase stiffness in the K table (linear) or -1
    const double *tangent;                          // linear phases: device pointer to C (n_str x n_str, row-major)
    int    lin_iso, has_hist;                       // lin_iso: linear phase whose tangent is exactly isotropic (params = lambda, 2 mu / conductivity)
    double params[12];
};

// One linear CG iteration as a CUDA graph (solve.cu): on small grids (the micro problems of a two-scale simulation) the seven
// kernels of an iteration take a few microseconds each and the iteration is bound by launch latency; replayed as a graph the
// whole iteration — scalars' read-back included — is one launch.  Two graphs, one per orientation of the d / d_alt ping-pong;
// rebuilt when any captured pointer or coefficient table changes.
struct IterGraph {
    cudaGraphExec_t exec[2] = {nullptr, nullptr};
    double *dA = nullptr, *dB = nullptr;      // exec[0]: d_old = dA, d_new = dB; exec[1]: the other way round
    const void *key[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint64_t cstamp = 0, sstamp = 0;
    int nb = 1;                               // lanes the graphs were captured for (batched solves)
    int launches = 0;                         // kernels per replay
    int replays = 0;                          // iterations of the last solve that were graph replays
    double fft_per_iter = 0.0;                // device time inside the convolution per iteration, measured on plain iterations
    bool valid = false, failed = false;
};

struct fans_ctx {
    fans_config cfg;
    int nx, ny, nz, n0, x0, n1, y1, h, nstr, ngp, fe, P, rank;
    size_t nloc;  // n0*ny*nz local voxels
    double L[3], le[3], ve;
    int device;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_loop0 = nullptr, ev_loop1 = nullptr;
    // component pipeline of the convolution over slabs (solve.cu, conv_run): the NVLink-bound y passes of component c run on `st2`
    // on a limited number of SMs while the HBM-bound z pass of the neighbouring component runs on `st`
    cudaStream_t st2 = nullptr;
    cudaEvent_t ev_pipe[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int *d_gate = nullptr;      // device word a y pass sets when its first CTA is resident (k_gate on `st` waits for it)
    int gate_seq = 0;
    int pipe = 0;               // 1: pipelined convolution (P > 1, fused transposes, h > 1)
    int y_grid = 0;             // CTAs of the persistent y pass in the pipeline (0: one CTA per tile)
    int chunks = 0;             // > 0: kz-chunked y <-> x pipeline (solve.cu, conv_run_chunked) with this many chunks

    double *field[FANS_N_FIELDS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double *d_alt = nullptr;   // ping-pong partner of D (fused d = s + beta d must not update in place)
    double *stage_io = nullptr; // AoS staging buffer for upload/download
    uint16_t *ms = nullptr;    // [n0][ny][nz]
    uint16_t *phidx = nullptr;  // phase id per voxel on the device (== ms values, validated < n_phases), [n0][ny][nz]

    // slab halos (world_size > 1): one node plane [h][ny][nz] each; ms_lo = phases of element plane -1
    double *halo_send_lo = nullptr, *halo_send_hi = nullptr, *halo_lo = nullptr, *halo_hi = nullptr;
    uint16_t *ms_lo = nullptr;

    // spectrum + Green operator
    int kzc = 0, kzp = 0;      // nz/2+1 and padded pitch (complex elements)
    int xpad = 0;              // complex elements appended to every x plane of the spectrum (FANS_XPAD; measured: no effect on B200)
    double2 *spec = nullptr;   // [h][n0][ny][kzp]   (P==1)   /   transposed [h][n1][nx][kzp] (P>1)
    double *gamma = nullptr;   // tile-major layout, see gamma.cu
    double2 *specB = nullptr;  // P > 1: the transposed spectrum (this rank's y rows, all x), blocks [p][h][n0][n1][kzp]
    double2 *peerA[8], *peerB[8];  // peer-mapped spectrum buffers of every rank (own entry = spec / specB), see comm.cu
    bool p2p = false;          // fused transposes: FFT passes store straight into the owner's buffer over NVLink
    int gT = 4;                // kz tile width of the fused x pass
    int yT = 8;                // kz tile width of the y passes
    FftPlan planx, plany, planz;  // planz: half-length complex plan of the r2c/c2r transform (N = nz/2)
    bool any_fft = false;         // some dimension is not a power of two: Bluestein passes of fft_any.cu, natural frequency order
    AnyPlan anyx, anyy, anyz;
    int gE = 8;                   // rows of the x transform a thread of the fused x pass holds (Gamma layout, gamma.cu); 1 for any_fft
    bool gamma_ready = false;

    // materials
    int n_phases = 0;
    std::vector<fans_phase_desc> phases;
    PhaseDev *d_phase = nullptr;
    double *d_K = nullptr;      // phase stiffness table [n_k][(8h)^2]
    double *d_C = nullptr;      // phase tangent table   [n_k][n_str^2]
    std::vector<double> K_host; // host copy of the phase stiffness table
    std::vector<double> S_host; // 27-point block stencils of the phases [q][delta][i][j] (stencil.cu), passed as a kernel parameter
    double *d_Stab = nullptr;   // the same table on the device (more than STENCIL_MAXQ phases, warps cut by an interface)
    bool stencil_iso = false;   // every phase has the isotropic sparsity pattern (stencil.cu: stencil_iso_pattern)
    uint64_t stencil_stamp = 0; // const_stamp the stencil tables were built for
    uint16_t ms_max = 0;
    int n_k = 0;
    bool all_linear = false, any_history = false, any_flag = false;
    int *phase_lut = nullptr;   // device: phase id (ms value) -> dense index, size 65536 only when needed
    double g0[9];
    double kapparef[81];
    std::vector<double> Bgp;    // basic gradient at the GPs: [ngp][3][8] then centre [3][8]
    bool materials_ready = false, ms_ready = false;
    uint64_t const_stamp = 0;   // identifies this ctx's content of the __constant__ tables

    // history: compact over the elements whose phase carries history, SoA [var][gp][compact element]
    double *hist = nullptr, *hist_t = nullptr;
    int n_hist = 0;             // doubles per GP
    unsigned *hidx = nullptr;   // [nloc] compact history index of an element, 0xffffffff: its phase has no history
    size_t nh = 0;              // number of history-bearing elements of this slab
    bool hist_ready = false;    // hidx / hist / hist_t match the current microstructure + materials
    std::vector<uint8_t> has_hist_host;  // per phase: its model carries history
    int *pflag = nullptr;       // plastic_flag [gp][element]

    // mixed BC
    bool mixed = false;
    fans_mixed_bc mbc;

    // reductions / scalars
    double *d_part = nullptr;   // per-block partial sums
    double *d_red = nullptr;    // reduced scalars (device)
    double *h_red = nullptr;    // pinned host mirror
    double *h_stage = nullptr;  // pinned host->device staging scalar
    unsigned int *d_ticket = nullptr;
    int neg_jac_flag_host = 0;
    int *h_fault = nullptr;     // pinned mirror of d_flag, refreshed by read_scalars (one synchronisation for scalars and fault)
    int *d_flag = nullptr;      // sticky device fault flag (J <= 0)

    // optional per-kernel-class device timing (bench.py roofline): event pairs resolved at the next stream sync
    bool prof = false;
    std::vector<cudaEvent_t> prof_pool;
    struct ProfRec { cudaEvent_t a, b; int cls; };
    std::vector<ProfRec> prof_pending;
    double prof_ms[FANS_PROF_CLASSES] = {0};
    int64_t prof_n[FANS_PROF_CLASSES] = {0};

    // device time inside convolution() (the reference's "FFT Time per iteration", solver.h:293): one event pair per call, resolved by fans_solve
    std::vector<cudaEvent_t> conv_pool;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> conv_pending;

    // batched linear solves (solve.cu, fans_solve_batch): `nb` right-hand sides ("lanes") travel through every pass of the
    // iteration in ONE launch.  Lane l of a field lives at base + l * h * nloc, of the spectrum at base + l * h * cStride, its scalar
    // block at d_red + l * S_COUNT.  nb == 1 outside a batched solve.
    int nb = 1;
    struct BatchArena {
        int lanes = 0;
        double *field[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // U, R, S, D, D_alt, K.d
        double2 *spec = nullptr;
        double *red = nullptr, *h_red = nullptr, *part = nullptr, *h_flag = nullptr;
        unsigned int *ticket = nullptr;
        IterGraph graph;                          // one sweep over all lanes as a CUDA graph
    } arena;

    IterGraph igraph;                             // the linear CG iteration as a CUDA graph (small grids)
    bool capturing = false;                       // launchers run under stream capture: no synchronisation, no timing events

    std::string err;
    int64_t launches = 0;
    int n_residual_evals = 0;
};

void fans_set_error(fans_ctx *ctx, int code, const std::string &msg);

// x-chunk length of a marching kernel (stencil, element sweeps): `cols` CTA columns march `n0` planes in chunks; every chunk pays
// `runin` extra planes, `slots` CTAs are resident at a time.  Minimises waves x steps per chunk over the power-of-two splits — long
// marches on big grids (few redundant run-in planes), many short ones on small grids (a 32^3 micro problem has 4 to 8 columns and
// would otherwise leave most of the 148 SMs idle).
static inline int pick_xchunk(int n0, long cols, long slots, int runin)
{
    auto cost_of = [&](int xc) {
        const long chunks = (n0 + xc - 1) / xc, ctas = cols * chunks;
        return ((ctas + slots - 1) / slots) * (long)(xc + runin);
    };
    long best_cost = -1;
    for (int xc = n0;; xc = (xc + 1) / 2) {
        const long c = cost_of(xc);
        if (best_cost < 0 || c < best_cost) best_cost = c;
        if (xc <= 2) break;
    }
    int best = n0;   // the shortest march within 3 % of the optimum: more CTAs balance the last wave better (measured at 512^3)
    for (int xc = n0;; xc = (xc + 1) / 2) {
        if (cost_of(xc) * 100 <= best_cost * 103) best = xc;
        if (xc <= 2) break;
    }
    return best;
}

// kernel classes for profiling (index into prof_ms / prof_n); names in api.cu
enum { PC_FFT_Z_FWD = 0, PC_FFT_Y_FWD, PC_FFT_X_GAMMA, PC_FFT_Y_INV, PC_FFT_Z_INV, PC_SWEEP_LINEAR, PC_SWEEP_RESIDUAL,
       PC_SWEEP_STRAINSTRESS, PC_CG_UPDATE, PC_REDUCE, PC_AXPY, PC_OTHER, PC_COMM_A2A, PC_COMM_HALO, PC_COMM_SCALAR };
void prof_begin(fans_ctx *ctx, int cls);
void prof_end(fans_ctx *ctx);
void prof_resolve(fans_ctx *ctx);  // call after a stream synchronisation

// ---------------- device helpers ----------------
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block reduction of NV values (sum for the first NSUM, max for the rest). Result valid in thread 0.
// scratch must hold NV*32 doubles. All threads of the block must call.
template <int NV, int NSUM>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double *scratch)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = (i < NSUM) ? warp_sum(v[i]) : warp_max(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) scratch[i * 32 + wid] = v[i];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = (lane < nw) ? scratch[i * 32 + lane] : ((i < NSUM) ? 0.0 : -1.0e300);
            v[i] = (i < NSUM) ? warp_sum(x) : warp_max(x);
        }
    }
}

// Grid-level deterministic reduce: every block deposits its NV partials, the last block to arrive
// (ticket counter) folds them in fixed block order and writes out[0..NV). "one grid-level reduce".
// grid_reduce_part: the same over a SUBSET of the grid (one lane of a batched launch): `nb` CTAs numbered `bid` share part / ticket / out.
template <int NV, int NSUM>
__device__ __forceinline__ void grid_reduce_part(double (&v)[NV], double *scratch, double *part, unsigned int *ticket,
                                                 double *out, bool accumulate, const unsigned nb, const unsigned bid)
{
    block_reduce<NV, NSUM>(v, scratch);
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) part[(size_t)i * nb + bid] = v[i];
        __threadfence();
        unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == nb - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double acc[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = (i < NSUM) ? 0.0 : -1.0e300;
        // fixed order: thread t sums blocks t, t+blockDim, ... then a block tree
        for (unsigned b = threadIdx.x; b < nb; b += blockDim.x) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                double x = __ldcg(&part[(size_t)i * nb + b]);
                acc[i] = (i < NSUM) ? acc[i] + x : fmax(acc[i], x);
            }
        }
        block_reduce<NV, NSUM>(acc, scratch);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) out[i] = (accumulate && i < NSUM) ? out[i] + acc[i] : acc[i];
            *ticket = 0u;
        }
    }
}
template <int NV, int NSUM>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], double *scratch, double *part, unsigned int *ticket,
                                            double *out, bool accumulate = false)
{
    grid_reduce_part<NV, NSUM>(v, scratch, part, ticket, out, accumulate, gridDim.x * gridDim.y * gridDim.z,
                               blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
}

// scalar form of grid_reduce<1, 1>: the same deterministic two-level sum without an array argument (an array handed over by
// reference pins the caller's accumulator to a local-memory slot for its whole lifetime)
__device__ __forceinline__ void grid_reduce_sum1_part(double v, double *scratch, double *part, unsigned int *ticket, double *out,
                                                      const unsigned nb, const unsigned bid)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) v = warp_sum(lane < nw ? scratch[lane] : 0.0);
    __shared__ bool is_last1;
    if (threadIdx.x == 0) {
        part[bid] = v;
        __threadfence();
        is_last1 = (atomicAdd(ticket, 1u) == nb - 1);
    }
    __syncthreads();
    if (is_last1) {
        __threadfence();
        double a = 0.0;
        for (unsigned b = threadIdx.x; b < nb; b += blockDim.x) a += __ldcg(&part[b]);   // fixed order: thread t sums blocks t, t + blockDim, ...
        a = warp_sum(a);
        __syncthreads();
        if (lane == 0) scratch[wid] = a;
        __syncthreads();
        if (wid == 0) {
            a = warp_sum(lane < nw ? scratch[lane] : 0.0);
            if (lane == 0) {
                *out = a;
                *ticket = 0u;
            }
        }
    }
}
__device__ __forceinline__ void grid_reduce_sum1(double v, double *scratch, double *part, unsigned int *ticket, double *out)
{
    grid_reduce_sum1_part(v, scratch, part, ticket, out, gridDim.x * gridDim.y * gridDim.z,
                          blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
}
#endif
