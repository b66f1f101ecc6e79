// fft_z.cu — passes P1 / P5 of the convolution (include/solver.h:387-412): r2c and c2r along z, the contiguous axis.
//   P1  a real line of nz doubles is read as Nh = nz/2 complex numbers, transformed by the register FFT and untangled:
//         E = (Z_k + conj Z_{Nh-k})/2, D = (Z_k - conj Z_{Nh-k})/2, w = exp(-2 pi i k/nz):  X_k = E - i w D,  X_{Nh-k} = conj(E + i w D)
//   P5  the mirror image (unnormalised like FFTW's c2r) with a fused epilogue:  out = scale * x  and  red = <dotw, out>
//       (s = -Gamma r and delta = <r, s> of SolverCG, include/solverCG.h:88-92, cost no extra pass over HBM).
// One line is carried by Nh/8 threads (one warp for nz = 512); a CTA holds 256/(Nh/8) lines.  Global loads/stores go
// straight from/to registers, fully coalesced along z.
#include "fft_reg.cuh"
#include "internal.h"

struct LineIdx {  // padded line: +1 slot per 8 rows, +1 per 64 rows -> conflict-free stage exchanges and untangle
    __device__ __forceinline__ int operator()(int row) const { return row + (row >> 3) + (row >> 6); }
};
__host__ __device__ constexpr int zline_pitch(int NH) { return NH + NH / 8 + NH / 64 + 2; }
__host__ __device__ constexpr int z_tpl(int NH) { return NH / rp_elems(NH); }
__host__ __device__ constexpr int z_lpb(int NH) { return z_tpl(NH) >= 256 ? 1 : 256 / z_tpl(NH); }
// a line is carried by z_tpl threads: up to 32 they sit in one warp and the stage exchanges only need __syncwarp
template <int NH> struct ZSyncSel { using type = SyncCta; };
template <> struct ZSyncSel<2> { using type = SyncWarp; };
template <> struct ZSyncSel<4> { using type = SyncWarp; };
template <> struct ZSyncSel<8> { using type = SyncWarp; };
template <> struct ZSyncSel<16> { using type = SyncWarp; };
template <> struct ZSyncSel<32> { using type = SyncWarp; };
template <> struct ZSyncSel<64> { using type = SyncWarp; };
template <> struct ZSyncSel<128> { using type = SyncWarp; };
template <> struct ZSyncSel<256> { using type = SyncWarp; };

__device__ __forceinline__ size_t spec_line(const SpecGeom &g, int ny, size_t line)
{
    const int y = (int)(line % ny);
    const size_t r = line / ny;
    const int xl = (int)(r % g.n0), c = (int)(r / g.n0);
    return (size_t)c * g.cStride + (size_t)(y >> g.l2n1) * g.blkStride + (size_t)xl * g.xStride + (size_t)(y & (g.n1 - 1)) * g.kzp;
}

// (forcing 4 resident CTAs per SM through a 64-register budget measured slower: 1.32 ms vs 1.27 ms at 512^3)
template <int NH>
__global__ void __launch_bounds__(z_tpl(NH) * z_lpb(NH)) k_fft_zf(const double *__restrict__ real, double2 *__restrict__ spec,
                                                                  const double2 *__restrict__ tw, const int *__restrict__ pos,
                                                                  SpecGeom g, int ny, size_t line0, size_t nlines)
{
    extern __shared__ double2 sm[];
    constexpr int E = rp_elems(NH), TPL = z_tpl(NH), LPB = z_lpb(NH), NST = rp_nstages(NH), PITCH = zline_pitch(NH);
    const int l = threadIdx.x / TPL, jt = threadIdx.x % TPL;
    const size_t line = line0 + (size_t)blockIdx.x * LPB + l;  // lines [line0, nlines) belong to this launch
    const bool valid = line < nlines;
    double2 *sml = sm + l * PITCH;
    const LineIdx idx;
    double2 a[1][E];
    const double2 *in = reinterpret_cast<const double2 *>(real) + line * NH;
#pragma unroll
    for (int e = 0; e < E; ++e) a[0][e] = valid ? in[rp_row<NH, 0>(jt, e)] : make_double2(0.0, 0.0);
    using ZSync = typename ZSyncSel<NH>::type;
    const ZSync zsync;
    rp_forward<NH, 1, LineIdx, 0, ZSync>(a, jt, sml, 0, idx, tw, 2, zsync);  // table length nz = 2 Nh
    if (NST > 1) zsync();
    rp_put<NH, NST - 1>(a[0], jt, sml, idx);
    zsync();
    if (!valid) return;
    double2 *out = spec + spec_line(g, ny, line);
    for (int k = jt; k <= NH / 2; k += TPL) {
        const double2 A = sml[idx(__ldg(&pos[k]))];
        const double2 Bc = sml[idx(__ldg(&pos[(NH - k) & (NH - 1)]))];
        const double2 B = make_double2(Bc.x, -Bc.y);
        const double2 Ev = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y + B.y));
        const double2 D = make_double2(0.5 * (A.x - B.x), 0.5 * (A.y - B.y));
        const double2 wd = rc_mul(__ldg(&tw[k]), D);  // w D ;  i w D = (-wd.y, wd.x)
        out[NH - k] = make_double2(Ev.x - wd.y, -(Ev.y + wd.x));  // conj(E + i w D)
        out[k] = make_double2(Ev.x + wd.y, Ev.y - wd.x);          // E - i w D   (for k = Nh/2 both coincide; this one wins)
    }
}

// ---- bulk asynchronous copies (TMA, cp.async.bulk -> UBLKCP) with transaction barriers: a warp's next line travels global -> shared
// on the copy engine while the warp transforms the current one; no registers, no address arithmetic, no load instructions in the
// warps.  Built as the alternative to the plain kernel above and kept as an opt-in (FANS_Z_TMA=1): see z_use_tma() for the measurement ----
__device__ __forceinline__ unsigned z_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(z_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(z_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(z_smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(z_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ZWAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ZDONE;\n"
        "bra ZWAIT;\n"
        "ZDONE:\n"
        "}\n" ::"r"(z_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__host__ __device__ constexpr int ztma_pitch(int NH) { return NH + NH / 8 + NH / 64; }   // LineIdx of row NH-1 is below this
#define ZTMA_WARPS 8

// Forward z pass, persistent warps (lines of nz <= 512: a line, or 32 / z_tpl lines, per warp).  Every warp owns two line buffers and
// two transaction barriers: lane 0 starts the bulk copy of the warp's NEXT lines into the idle buffer, the warp waits for the current
// ones, takes them into registers and then uses the same buffer as its exchange tile.  No CTA-wide barrier anywhere.
template <int NH>
__global__ void __launch_bounds__(ZTMA_WARPS * 32, 3) k_fft_zf_tma(const double *__restrict__ real, double2 *__restrict__ spec,
                                                                const double2 *__restrict__ tw, const int *__restrict__ pos, SpecGeom g, int ny,
                                                                size_t line0, size_t nlines)
{
    extern __shared__ __align__(16) double2 sm[];
    constexpr int E = rp_elems(NH), TPL = z_tpl(NH), LPW = 32 / TPL, NST = rp_nstages(NH), PITCH = ztma_pitch(NH);
    static_assert(TPL <= 32, "one line must fit a warp");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l = lane / TPL, jt = lane % TPL;
    double2 *wbuf = sm + (size_t)warp * (2 * LPW * PITCH);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + (size_t)ZTMA_WARPS * 2 * LPW * PITCH) + 2 * warp;
    const size_t ntiles = (nlines - line0 + LPW - 1) / LPW;               // a tile = the LPW lines a warp carries at a time
    const size_t nwarps = (size_t)gridDim.x * ZTMA_WARPS, gw = (size_t)blockIdx.x * ZTMA_WARPS + warp;
    const double2 *in2 = reinterpret_cast<const double2 *>(real);
    if (lane == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        mbar_init_fence();
    }
    __syncwarp();
    auto issue = [&](size_t tile, int buf) {   // lane 0: the lines of `tile` -> buffer `buf`
        const size_t first = line0 + tile * LPW;
        const int nval = (int)((nlines - first) < (size_t)LPW ? (nlines - first) : (size_t)LPW);
        mbar_expect_tx(bars + buf, (unsigned)(nval * NH * sizeof(double2)));
        for (int q = 0; q < nval; ++q)
            bulk_g2s(wbuf + (size_t)(buf * LPW + q) * PITCH, in2 + (first + q) * NH, (unsigned)(NH * sizeof(double2)), bars + buf);
    };
    if (lane == 0 && gw < ntiles) issue(gw, 0);
    const LineIdx idx;
    const SyncWarp zsync;
    int it = 0;
    for (size_t tile = gw; tile < ntiles; tile += nwarps, ++it) {
        const int buf = it & 1;
        if (lane == 0 && tile + nwarps < ntiles) {   // the other buffer was left behind a __syncwarp by every lane
            proxy_fence_async();
            issue(tile + nwarps, buf ^ 1);
        }
        mbar_wait(bars + buf, (unsigned)((it >> 1) & 1));
        const size_t line = line0 + tile * LPW + l;
        const bool valid = line < nlines;
        double2 *sml = wbuf + (size_t)(buf * LPW + l) * PITCH;
        double2 a[1][E];
#pragma unroll
        for (int e = 0; e < E; ++e) a[0][e] = valid ? sml[rp_row<NH, 0>(jt, e)] : make_double2(0.0, 0.0);
        __syncwarp();   // everybody holds its rows: the line buffer becomes the exchange tile
        rp_forward<NH, 1, LineIdx, 0, SyncWarp>(a, jt, sml, 0, idx, tw, 2, zsync);
        if (NST > 1) zsync();
        rp_put<NH, NST - 1>(a[0], jt, sml, idx);
        zsync();
        if (valid) {
            double2 *out = spec + spec_line(g, ny, line);
            for (int k = jt; k <= NH / 2; k += TPL) {
                const double2 A = sml[idx(__ldg(&pos[k]))];
                const double2 Bc = sml[idx(__ldg(&pos[(NH - k) & (NH - 1)]))];
                const double2 B = make_double2(Bc.x, -Bc.y);
                const double2 Ev = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y + B.y));
                const double2 D = make_double2(0.5 * (A.x - B.x), 0.5 * (A.y - B.y));
                const double2 wd = rc_mul(__ldg(&tw[k]), D);
                out[NH - k] = make_double2(Ev.x - wd.y, -(Ev.y + wd.x));
                out[k] = make_double2(Ev.x + wd.y, Ev.y - wd.x);
            }
        }
        __syncwarp();   // the buffer is free for the copy after next
    }
}

template <int NH>
__global__ void __launch_bounds__(z_tpl(NH) * z_lpb(NH)) k_fft_zi(const double2 *__restrict__ spec, double *__restrict__ real,
                                                                  const double2 *__restrict__ tw, const int *__restrict__ pos,
                                                                  SpecGeom g, int ny, size_t line0, size_t nlines, double scale,
                                                                  const double *__restrict__ dotw, double *part,
                                                                  unsigned int *ticket, double *red_out, int accumulate)
{
    // batched solves: blockIdx.y = lane; a lane owns the lines [lane * nlines, (lane + 1) * nlines) (line0 = 0) and its own
    // partial sums, ticket and scalar block
    extern __shared__ double2 sm[];
    __shared__ double scratch[32];
    constexpr int E = rp_elems(NH), TPL = z_tpl(NH), LPB = z_lpb(NH), NST = rp_nstages(NH), PITCH = zline_pitch(NH);
    const int l = threadIdx.x / TPL, jt = threadIdx.x % TPL;
    const size_t lline = line0 + (size_t)blockIdx.x * LPB + l;  // lines [line0, nlines) belong to this launch
    const bool valid = lline < nlines;
    const size_t line = lline + (size_t)blockIdx.y * nlines;
    double2 *sml = sm + l * PITCH;
    const LineIdx idx;
    if (valid) {
        const double2 *in = spec + spec_line(g, ny, line);
        for (int k = jt; k <= NH / 2; k += TPL) {
            const double2 A = in[k];
            const double2 Bc = in[NH - k];
            const double2 B = make_double2(Bc.x, -Bc.y);
            const double2 Ev = rc_add(A, B), D = rc_sub(A, B);
            const double2 wd = rc_mulc(D, __ldg(&tw[k]));  // cw D with cw = conj(w_k)
            if (k > 0) sml[idx(__ldg(&pos[NH - k]))] = make_double2(Ev.x + wd.y, -(Ev.y - wd.x));  // conj(E - i cw D)
            sml[idx(__ldg(&pos[k]))] = make_double2(Ev.x - wd.y, Ev.y + wd.x);                      // E + i cw D
        }
    } else {
        for (int k = jt; k < NH; k += TPL) sml[idx(k)] = make_double2(0.0, 0.0);
    }
    using ZSync = typename ZSyncSel<NH>::type;
    const ZSync zsync;
    zsync();
    double2 a[1][E];
    rp_get<NH, NST - 1>(a[0], jt, sml, idx);
    rp_inverse<NH, 1, LineIdx, NST - 1, ZSync>(a, jt, sml, 0, idx, tw, 2, zsync);
    double acc[1] = {0.0};
    if (valid) {
        double2 *out = reinterpret_cast<double2 *>(real) + line * NH;
        const double2 *dw = dotw ? reinterpret_cast<const double2 *>(dotw) + line * NH : nullptr;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int row = rp_row<NH, 0>(jt, e);
            double2 v = a[0][e];
            v.x *= scale;
            v.y *= scale;
            if (dw) {
                const double2 r = dw[row];
                acc[0] += r.x * v.x + r.y * v.y;
            }
            out[row] = v;
        }
    }
    if (red_out)
        grid_reduce_part<1, 1>(acc, scratch, part + (size_t)blockIdx.y * gridDim.x, ticket + blockIdx.y, red_out + (size_t)blockIdx.y * S_COUNT,
                               accumulate != 0, gridDim.x, blockIdx.x);
}

static bool z_use_tma()
{
    static const bool on = [] {
        // opt-in: measured slower than the plain kernel on B200 (512^3: 1.40 vs 1.32 ms, 256^3: 0.177 vs 0.161 ms,
        // profiles/r2tma_zfwd_bulk_copy.txt).  ncu of the plain kernel: 175 M shared-memory wavefronts = 49 % of the wavefront
        // rate, stalls mio_throttle 3.5 / short scoreboard 3.6 next to long scoreboard 4.3 — the shared-memory queue is as much the limit
        // as the global loads, and landing the line in shared memory first adds two accesses per element to the seven of the transform
        const char *e = getenv("FANS_Z_TMA");
        return e && e[0] == '1';
    }();
    return on;
}

template <int NH>
static int launch_zf(fans_ctx *ctx, const double *in, const SpecGeom &g, size_t line0, size_t nlines, cudaStream_t st)
{
    if constexpr (z_tpl(NH) <= 32) {
        if (z_use_tma()) {
            constexpr int LPW = 32 / z_tpl(NH);
            const size_t smem = sizeof(double2) * ZTMA_WARPS * 2 * LPW * ztma_pitch(NH) + sizeof(uint64_t) * 2 * ZTMA_WARPS;
            static int resident = 0;
            if (!resident) {
                CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_zf_tma<NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k_fft_zf_tma<NH>, ZTMA_WARPS * 32, smem));
                if (resident < 1) resident = 1;
            }
            const size_t ntiles = (nlines - line0 + LPW - 1) / LPW;
            size_t grid = (ntiles + ZTMA_WARPS - 1) / ZTMA_WARPS;
            if (grid > (size_t)FANS_SMS * resident) grid = (size_t)FANS_SMS * resident;
            k_fft_zf_tma<NH><<<(unsigned)grid, ZTMA_WARPS * 32, smem, st>>>(in, ctx->spec, ctx->planz.tw, ctx->planz.pos, g, ctx->ny, line0, nlines);
            return FANS_OK;
        }
    }
    constexpr int LPB = z_lpb(NH), NTHR = z_tpl(NH) * LPB;
    const size_t smem = sizeof(double2) * zline_pitch(NH) * LPB;
    if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_zf<NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fft_zf<NH><<<(unsigned)((nlines - line0 + LPB - 1) / LPB), NTHR, smem, st>>>(in, ctx->spec, ctx->planz.tw, ctx->planz.pos, g, ctx->ny, line0, nlines);
    return FANS_OK;
}
template <int NH>
static int launch_zi(fans_ctx *ctx, double *out, const SpecGeom &g, size_t line0, size_t nlines, double scale, const double *dotw, double *red_out,
                     bool accumulate, cudaStream_t st)
{
    constexpr int LPB = z_lpb(NH), NTHR = z_tpl(NH) * LPB;
    const size_t smem = sizeof(double2) * zline_pitch(NH) * LPB;
    if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_zi<NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid((unsigned)((nlines - line0 + LPB - 1) / LPB), (unsigned)ctx->nb);
    k_fft_zi<NH><<<grid, NTHR, smem, st>>>(ctx->spec, out, ctx->planz.tw, ctx->planz.pos, g, ctx->ny, line0,
                                                                                 nlines, scale, dotw, ctx->d_part, ctx->d_ticket, red_out, accumulate ? 1 : 0);
    return FANS_OK;
}

#define Z_SWITCH(CALL)                                                                                          \
    switch (ctx->nz / 2) {                                                                                      \
    case 2: rc = CALL(2); break;                                                                                \
    case 4: rc = CALL(4); break;                                                                                \
    case 8: rc = CALL(8); break;                                                                                \
    case 16: rc = CALL(16); break;                                                                              \
    case 32: rc = CALL(32); break;                                                                              \
    case 64: rc = CALL(64); break;                                                                              \
    case 128: rc = CALL(128); break;                                                                            \
    case 256: rc = CALL(256); break;                                                                            \
    case 512: rc = CALL(512); break;                                                                            \
    default: fans_set_error(ctx, FANS_ERR_ARG, "unsupported n_z for the z pass");                              \
    }

// (a batched solve transforms the h components of all its lanes as nb * h components of one launch)
int fft_pass_z_fwd(fans_ctx *ctx, const double *in) { return fft_pass_z_fwd_part(ctx, in, 0, ctx->h * ctx->nb, ctx->st); }

// components [c0, c0 + nc) only, on stream st (slab pipeline, solve.cu)
int fft_pass_z_fwd_part(fans_ctx *ctx, const double *in, int c0, int nc, cudaStream_t st)
{
    prof_begin(ctx, PC_FFT_Z_FWD);
    const SpecGeom g = spec_geom_A(ctx);
    const size_t line0 = (size_t)c0 * ctx->n0 * ctx->ny, nlines = (size_t)(c0 + nc) * ctx->n0 * ctx->ny;
    int rc = FANS_ERR_ARG;
#define ZF(N_) launch_zf<N_>(ctx, in, g, line0, nlines, st)
    Z_SWITCH(ZF)
#undef ZF
    prof_end(ctx);
    ctx->launches++;
    if (rc != FANS_OK) return rc;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

int fft_pass_z_inv(fans_ctx *ctx, double *out, double scale, const double *dotw, double *red_out)
{
    return fft_pass_z_inv_part(ctx, out, scale, dotw, red_out, 0, ctx->h, false, ctx->st);
}

// components [c0, c0 + nc) only; accumulate: red_out[0] += <dotw, out> of these components instead of overwriting it
int fft_pass_z_inv_part(fans_ctx *ctx, double *out, double scale, const double *dotw, double *red_out, int c0, int nc, bool accumulate,
                        cudaStream_t st)
{
    prof_begin(ctx, PC_FFT_Z_INV);
    const SpecGeom g = spec_geom_A(ctx);
    const size_t line0 = (size_t)c0 * ctx->n0 * ctx->ny, nlines = (size_t)(c0 + nc) * ctx->n0 * ctx->ny;
    int rc = FANS_ERR_ARG;
#define ZI(N_) launch_zi<N_>(ctx, out, g, line0, nlines, scale, dotw, red_out, accumulate, st)
    Z_SWITCH(ZI)
#undef ZI
    prof_end(ctx);
    ctx->launches++;
    if (rc != FANS_OK) return rc;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}
