// fft.cu — the five passes of FANS's convolution (include/solver.h:387-412) without cuFFT/FFTW:
//   P1  z r2c            real [line][nz]            -> spec [line][kzp]      (half-length complex FFT + untangle)
//   P2  y forward        spec, tiles of T kz-columns, in place
//   P3  x forward . Gamma_hat . x inverse   (all `howmany` components of a frequency in one CTA, in place)
//   P4  y inverse        in place
//   P5  z c2r            spec -> real, fused epilogue: scale (s = -Gamma r), dot product <r, s>
// Layout of the spectrum: component-major planes [c][x][y][kz], kz contiguous with padded pitch kzp.
#include "fft.cuh"
#include <cmath>

// ------------------------------------------------------------------------------------------------
// P2 / P4: strided in-place transform of one component, one (outer, kz-tile) per CTA
// ------------------------------------------------------------------------------------------------
template <int T, bool INV>
__global__ void __launch_bounds__(256, 3) k_fft_strided(double2 *__restrict__ spec, FftStages st,
                                                      const double2 *__restrict__ tw, size_t cStride, size_t oStride,
                                                      int nO, int nTiles, size_t rowStride)
{
    extern __shared__ double2 sm[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int b = blockIdx.x;
    const int tile = b % nTiles;
    const int o = (b / nTiles) % nO;
    const int c = b / (nTiles * nO);
    double2 *g = spec + (size_t)c * cStride + (size_t)o * oStride + (size_t)tile * T;
    const int NT = st.N * T;
    for (int i = tid; i < NT; i += nthr) {
        const int row = i / T, t = i % T;
        sm[tix<T>(row, t)] = g[(size_t)row * rowStride + t];
    }
    __syncthreads();
    fft_tile<T, INV>(sm, st, tw, 1, tid, nthr);
    for (int i = tid; i < NT; i += nthr) {
        const int row = i / T, t = i % T;
        g[(size_t)row * rowStride + t] = sm[tix<T>(row, t)];
    }
}

// ------------------------------------------------------------------------------------------------
// P3: x forward, Green operator, x inverse.  gamma is stored tile-major:
//     gamma[(((o*nTiles + tile)*NG + k) * N + row) * T + t],  NG = H(H+1)/2 (upper triangle, row-major)
// ------------------------------------------------------------------------------------------------
template <int H, int T>
__global__ void __launch_bounds__(256, 2) k_fft_x_gamma(double2 *__restrict__ spec, const double *__restrict__ gamma,
                                                      FftStages st, const double2 *__restrict__ tw, size_t cStride,
                                                      size_t oStride, int nTiles, size_t rowStride)
{
    extern __shared__ double2 sm[];
    constexpr int NG = H * (H + 1) / 2;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int tile = blockIdx.x % nTiles;
    const int o = blockIdx.x / nTiles;
    const int NT = st.N * T;
    double2 *g = spec + (size_t)o * oStride + (size_t)tile * T;
    const double *gam = gamma + (size_t)blockIdx.x * NG * NT;
    // pull the Green-operator tile towards L2 while the forward transform runs
    for (int i = tid * 16; i < NG * NT; i += nthr * 16)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(gam + i));
#pragma unroll
    for (int c = 0; c < H; ++c)
        for (int i = tid; i < NT; i += nthr) {
            const int row = i / T, t = i % T;
            sm[c * NT + tix<T>(row, t)] = g[(size_t)c * cStride + (size_t)row * rowStride + t];
        }
    __syncthreads();
    fft_tile<T, false>(sm, st, tw, H, tid, nthr);
    for (int i = tid; i < NT; i += nthr) {
        const int row = i / T, t = i % T;
        const int p = tix<T>(row, t);
        if (H == 1) {
            const double g0 = __ldg(&gam[i]);
            double2 r = sm[p];
            sm[p] = make_double2(g0 * r.x, g0 * r.y);
        } else {
            const double g00 = __ldg(&gam[0 * NT + i]), g01 = __ldg(&gam[1 * NT + i]), g02 = __ldg(&gam[2 * NT + i]);
            const double g11 = __ldg(&gam[3 * NT + i]), g12 = __ldg(&gam[4 * NT + i]), g22 = __ldg(&gam[5 * NT + i]);
            const double2 r0 = sm[p], r1 = sm[NT + p], r2 = sm[2 * NT + p];
            sm[p] = make_double2(g00 * r0.x + g01 * r1.x + g02 * r2.x, g00 * r0.y + g01 * r1.y + g02 * r2.y);
            sm[NT + p] = make_double2(g01 * r0.x + g11 * r1.x + g12 * r2.x, g01 * r0.y + g11 * r1.y + g12 * r2.y);
            sm[2 * NT + p] = make_double2(g02 * r0.x + g12 * r1.x + g22 * r2.x, g02 * r0.y + g12 * r1.y + g22 * r2.y);
        }
    }
    __syncthreads();
    fft_tile<T, true>(sm, st, tw, H, tid, nthr);
#pragma unroll
    for (int c = 0; c < H; ++c)
        for (int i = tid; i < NT; i += nthr) {
            const int row = i / T, t = i % T;
            g[(size_t)c * cStride + (size_t)row * rowStride + t] = sm[c * NT + tix<T>(row, t)];
        }
}

// ------------------------------------------------------------------------------------------------
// P1: z r2c.  8 lines per CTA; a real line of nz doubles is read as nz/2 complex, transformed with the
// half-length plan and untangled:  E = (Zk + conj Z(Nh-k))/2, D = (Zk - conj Z(Nh-k))/2, w = exp(-2 pi i k/nz)
//     X[k] = E - i w D ,   X[Nh-k] = conj(E + i w D)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_fft_z_fwd(const double *__restrict__ real, double2 *__restrict__ spec,
                                                     FftStages st, const double2 *__restrict__ tw,
                                                     const int *__restrict__ pos, size_t nlines, int kzp)
{
    extern __shared__ double2 sm[];
    constexpr int T = 8;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int Nh = st.N;
    const size_t line0 = (size_t)blockIdx.x * T;
    const int nl = (int)min((size_t)T, nlines - line0);
    const double2 *in = reinterpret_cast<const double2 *>(real) + line0 * Nh;
    for (int t = 0; t < nl; ++t)
        for (int m = tid; m < Nh; m += nthr) sm[tix<T>(m, t)] = in[(size_t)t * Nh + m];
    for (int t = nl; t < T; ++t)
        for (int m = tid; m < Nh; m += nthr) sm[tix<T>(m, t)] = make_double2(0.0, 0.0);
    __syncthreads();
    fft_tile<T, false>(sm, st, tw, 1, tid, nthr);
    // untangle into registers (pairs k, Nh-k), then write back in natural kz order
    const int npair = (Nh / 2 + 1) * T;
    double2 xa[5], xb[5];
#pragma unroll
    for (int it = 0; it < 5; ++it) {
        const int w = tid + it * nthr;
        if (w < npair) {
            const int t = w & 7, k = w >> 3;
            const double2 A = sm[tix<T>(__ldg(&pos[k]), t)];
            const double2 B = cconj(sm[tix<T>(__ldg(&pos[(Nh - k) & (Nh - 1)]), t)]);
            const double2 E = make_double2(0.5 * (A.x + B.x), 0.5 * (A.y + B.y));
            const double2 D = make_double2(0.5 * (A.x - B.x), 0.5 * (A.y - B.y));
            const double2 wk = __ldg(&tw[k]);  // table length nz: tw[k] = exp(-2 pi i k / nz)
            const double2 wd = cmul(wk, D);    // w D ;  i w D = (-wd.y, wd.x)
            xa[it] = make_double2(E.x + wd.y, E.y - wd.x);       // E - i w D
            xb[it] = make_double2(E.x - wd.y, -(E.y + wd.x));    // conj(E + i w D)
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 5; ++it) {
        const int w = tid + it * nthr;
        if (w < npair) {
            const int t = w & 7, k = w >> 3;
            sm[tix<T>(Nh - k, t)] = xb[it];
            sm[tix<T>(k, t)] = xa[it];  // for k == Nh/2 both coincide (xa == xb up to rounding); xa wins
        }
    }
    __syncthreads();
    double2 *out = spec + line0 * kzp;
    for (int t = 0; t < nl; ++t)
        for (int k = tid; k <= Nh; k += nthr) out[(size_t)t * kzp + k] = sm[tix<T>(k, t)];
}

// ------------------------------------------------------------------------------------------------
// P5: z c2r (unnormalised, like FFTW's c2r) with fused epilogue
//     Z[k] = E + i cw D,  Z[Nh-k] = conj(E - i cw D),  E = Xk + conj X(Nh-k), D = Xk - conj X(Nh-k), cw = exp(+2 pi i k/nz)
//     out = scale * x ;  optional  red[slot] = sum(out * dotw)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_fft_z_inv(const double2 *__restrict__ spec, double *__restrict__ real,
                                                     FftStages st, const double2 *__restrict__ tw,
                                                     const int *__restrict__ pos, size_t nlines, int kzp, double scale,
                                                     const double *__restrict__ dotw, double *part,
                                                     unsigned int *ticket, double *red_out)
{
    extern __shared__ double2 sm[];
    __shared__ double scratch[32];
    constexpr int T = 8;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int Nh = st.N;
    const size_t line0 = (size_t)blockIdx.x * T;
    const int nl = (int)min((size_t)T, nlines - line0);
    const double2 *in = spec + line0 * kzp;
    for (int t = 0; t < nl; ++t)
        for (int k = tid; k <= Nh; k += nthr) sm[tix<T>(k, t)] = in[(size_t)t * kzp + k];
    for (int t = nl; t < T; ++t)
        for (int k = tid; k <= Nh; k += nthr) sm[tix<T>(k, t)] = make_double2(0.0, 0.0);
    __syncthreads();
    const int npair = (Nh / 2 + 1) * T;
    double2 za[5], zb[5];
#pragma unroll
    for (int it = 0; it < 5; ++it) {
        const int w = tid + it * nthr;
        if (w < npair) {
            const int t = w & 7, k = w >> 3;
            const double2 A = sm[tix<T>(k, t)];
            const double2 B = cconj(sm[tix<T>(Nh - k, t)]);
            const double2 E = cadd(A, B), D = csub(A, B);
            const double2 wk = __ldg(&tw[k]);
            const double2 wd = cmulc(D, wk);  // cw D with cw = conj(w_k)
            za[it] = make_double2(E.x - wd.y, E.y + wd.x);     // E + i cw D
            zb[it] = make_double2(E.x + wd.y, -(E.y - wd.x));  // conj(E - i cw D)
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 5; ++it) {
        const int w = tid + it * nthr;
        if (w < npair) {
            const int t = w & 7, k = w >> 3;
            if (k > 0) sm[tix<T>(__ldg(&pos[Nh - k]), t)] = zb[it];
            sm[tix<T>(__ldg(&pos[k]), t)] = za[it];
        }
    }
    __syncthreads();
    fft_tile<T, true>(sm, st, tw, 1, tid, nthr);
    double2 *out = reinterpret_cast<double2 *>(real) + line0 * Nh;
    const double2 *dw = dotw ? reinterpret_cast<const double2 *>(dotw) + line0 * Nh : nullptr;
    double acc[1] = {0.0};
    for (int t = 0; t < nl; ++t)
        for (int m = tid; m < Nh; m += nthr) {
            double2 v = sm[tix<T>(m, t)];
            v.x *= scale;
            v.y *= scale;
            if (dw) {
                const double2 r = dw[(size_t)t * Nh + m];
                acc[0] += r.x * v.x + r.y * v.y;
            }
            out[(size_t)t * Nh + m] = v;
        }
    if (red_out) grid_reduce<1, 1>(acc, scratch, part, ticket, red_out);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int ilog2(int n)
{
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}

int fft_plan_init(fans_ctx *ctx, FftPlan &p, int N, int ntab)
{
    if (N < 1 || (N & (N - 1)) != 0) {
        fans_set_error(ctx, FANS_ERR_ARG, "FFT length " + std::to_string(N) + " is not a power of two (only 2^k grids are supported)");
        return FANS_ERR_ARG;
    }
    p.N = N;
    p.ntab = ntab;
    int k = ilog2(N);
    p.nst = 0;
    for (int i = 0; i < k / 3; ++i) p.radix[p.nst++] = 8;
    if (k % 3 == 1) p.radix[p.nst++] = 2;
    if (k % 3 == 2) p.radix[p.nst++] = 4;
    p.pos_host.assign(N, 0);
    for (int f = 0; f < N; ++f) {
        int ff = f, pp = 0, Lr = N;
        for (int s = 0; s < p.nst; ++s) {
            int r = p.radix[s];
            int m = ff % r;
            ff /= r;
            Lr /= r;
            pp += m * Lr;
        }
        p.pos_host[f] = pp;
    }
    std::vector<double2> tw(ntab);
    for (int i = 0; i < ntab; ++i) {
        // exact octant reduction keeps the table accurate to the last bit or so
        long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)i / (long double)ntab;
        tw[i].x = (double)cosl(a);
        tw[i].y = (double)sinl(a);
    }
    // exact values on the axes
    tw[0] = make_double2(1.0, 0.0);
    if (ntab % 2 == 0) tw[ntab / 2] = make_double2(-1.0, 0.0);
    if (ntab % 4 == 0) {
        tw[ntab / 4] = make_double2(0.0, -1.0);
        tw[3 * ntab / 4] = make_double2(0.0, 1.0);
    }
    CUDA_TRY(ctx, cudaMalloc(&p.tw, sizeof(double2) * ntab));
    CUDA_TRY(ctx, cudaMemcpy(p.tw, tw.data(), sizeof(double2) * ntab, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMalloc(&p.pos, sizeof(int) * N));
    CUDA_TRY(ctx, cudaMemcpy(p.pos, p.pos_host.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
    return FANS_OK;
}

void fft_plan_free(FftPlan &p)
{
    if (p.tw) cudaFree(p.tw);
    if (p.pos) cudaFree(p.pos);
    p.tw = nullptr;
    p.pos = nullptr;
}

static FftStages make_stages(const FftPlan &p)
{
    FftStages s;
    s.N = p.N;
    s.logN = ilog2(p.N);
    s.nst = p.nst;
    s.twmul = p.ntab / p.N;
    for (int i = 0; i < 8; ++i) s.radix[i] = i < p.nst ? p.radix[i] : 0;
    return s;
}

static int pick_threads(int N, int T, int nbatch)
{
    long n = (long)N * T * nbatch / 8;
    int t = 64;
    while (t < n && t < 256) t <<= 1;
    return t;
}

template <typename K>
static int set_smem(fans_ctx *ctx, K kernel, size_t bytes)
{
    if (bytes > 227 * 1024) {
        fans_set_error(ctx, FANS_ERR_ARG, "FFT tile needs " + std::to_string(bytes) + " B of shared memory (> 227 KB)");
        return FANS_ERR_ARG;
    }
    if (bytes > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return FANS_OK;
}

// y pass (forward or inverse) on the local slab: lines along y for every (c, x, kz tile)
int fft_pass_y(fans_ctx *ctx, bool inverse)
{
    prof_begin(ctx, inverse ? PC_FFT_Y_INV : PC_FFT_Y_FWD);
    const FftStages st = make_stages(ctx->plany);
    const int ny = ctx->ny;
    // T = 8 keeps 128 B segments; very long lines fall back to T = 4 to keep >= 2 CTAs per SM
    const int T = (ny >= 1024) ? 4 : 8;
    const int nTiles = (ctx->kzc + T - 1) / T;
    const size_t smem = sizeof(double2) * ny * T;
    const int nthr = pick_threads(ny, T, 1);
    const size_t cStride = (size_t)ctx->n0 * ny * ctx->kzp;
    const size_t oStride = (size_t)ny * ctx->kzp;
    const unsigned grid = (unsigned)((size_t)ctx->h * ctx->n0 * nTiles);
    if (T == 8) {
        if (!inverse) {
            FANS_CHECK(set_smem(ctx, k_fft_strided<8, false>, smem));
            k_fft_strided<8, false><<<grid, nthr, smem, ctx->st>>>(ctx->spec, st, ctx->plany.tw, cStride, oStride, ctx->n0, nTiles, ctx->kzp);
        } else {
            FANS_CHECK(set_smem(ctx, k_fft_strided<8, true>, smem));
            k_fft_strided<8, true><<<grid, nthr, smem, ctx->st>>>(ctx->spec, st, ctx->plany.tw, cStride, oStride, ctx->n0, nTiles, ctx->kzp);
        }
    } else {
        if (!inverse) {
            FANS_CHECK(set_smem(ctx, k_fft_strided<4, false>, smem));
            k_fft_strided<4, false><<<grid, nthr, smem, ctx->st>>>(ctx->spec, st, ctx->plany.tw, cStride, oStride, ctx->n0, nTiles, ctx->kzp);
        } else {
            FANS_CHECK(set_smem(ctx, k_fft_strided<4, true>, smem));
            k_fft_strided<4, true><<<grid, nthr, smem, ctx->st>>>(ctx->spec, st, ctx->plany.tw, cStride, oStride, ctx->n0, nTiles, ctx->kzp);
        }
    }
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

// fused x pass with the Green operator (single-GPU layout [c][x][y][kz]: x stride = ny*kzp)
int fft_pass_x_gamma(fans_ctx *ctx)
{
    prof_begin(ctx, PC_FFT_X_GAMMA);
    const FftStages st = make_stages(ctx->planx);
    const int nx = ctx->nx, T = ctx->gT;
    const int nTiles = (ctx->kzc + T - 1) / T;
    const size_t smem = sizeof(double2) * nx * T * ctx->h;
    const int nthr = pick_threads(nx, T, ctx->h);
    const size_t cStride = (size_t)ctx->n0 * ctx->ny * ctx->kzp;
    const size_t oStride = ctx->kzp;                    // next y line
    const size_t rowStride = (size_t)ctx->ny * ctx->kzp;  // next x plane
    const unsigned grid = (unsigned)((size_t)ctx->ny * nTiles);
#define LAUNCH_XG(H_, T_)                                                                                             \
    do {                                                                                                              \
        FANS_CHECK(set_smem(ctx, k_fft_x_gamma<H_, T_>, smem));                                                       \
        k_fft_x_gamma<H_, T_><<<grid, nthr, smem, ctx->st>>>(ctx->spec, ctx->gamma, st, ctx->planx.tw, cStride, oStride, nTiles, rowStride); \
    } while (0)
    if (ctx->h == 1 && T == 8) LAUNCH_XG(1, 8);
    else if (ctx->h == 1 && T == 4) LAUNCH_XG(1, 4);
    else if (ctx->h == 3 && T == 8) LAUNCH_XG(3, 8);
    else if (ctx->h == 3 && T == 4) LAUNCH_XG(3, 4);
    else {
        fans_set_error(ctx, FANS_ERR_ARG, "unsupported howmany / tile width in the fused x pass");
        return FANS_ERR_ARG;
    }
#undef LAUNCH_XG
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

int fft_pass_z_fwd(fans_ctx *ctx, const double *in)
{
    prof_begin(ctx, PC_FFT_Z_FWD);
    const FftStages st = make_stages(ctx->planz);
    const int Nh = ctx->nz / 2;
    const size_t nlines = (size_t)ctx->h * ctx->n0 * ctx->ny;
    const size_t smem = sizeof(double2) * (Nh + 1) * 8;
    int nthr = Nh < 32 ? 32 : (Nh > 1024 ? 1024 : Nh);
    FANS_CHECK(set_smem(ctx, k_fft_z_fwd, smem));
    const unsigned grid = (unsigned)((nlines + 7) / 8);
    k_fft_z_fwd<<<grid, nthr, smem, ctx->st>>>(in, ctx->spec, st, ctx->planz.tw, ctx->planz.pos, nlines, ctx->kzp);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

int fft_pass_z_inv(fans_ctx *ctx, double *out, double scale, const double *dotw, double *red_out)
{
    prof_begin(ctx, PC_FFT_Z_INV);
    const FftStages st = make_stages(ctx->planz);
    const int Nh = ctx->nz / 2;
    const size_t nlines = (size_t)ctx->h * ctx->n0 * ctx->ny;
    const size_t smem = sizeof(double2) * (Nh + 1) * 8;
    int nthr = Nh < 32 ? 32 : (Nh > 1024 ? 1024 : Nh);
    FANS_CHECK(set_smem(ctx, k_fft_z_inv, smem));
    const unsigned grid = (unsigned)((nlines + 7) / 8);
    k_fft_z_inv<<<grid, nthr, smem, ctx->st>>>(ctx->spec, out, st, ctx->planz.tw, ctx->planz.pos, nlines, ctx->kzp, scale,
                                               dotw, ctx->d_part, ctx->d_ticket, red_out);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}
