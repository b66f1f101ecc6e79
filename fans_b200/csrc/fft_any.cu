// fft_any.cu — the convolution (include/solver.h:387-412) for grids whose dimensions are NOT all powers of two.
// The reference takes any n_x, n_y, n_z (FFTW; src/reader.cpp:300-305).  The register FFTs of fft_x/y/z.cu are radix-2^k only, so
// every other size goes through this file: Bluestein's chirp-z algorithm turns a length-n DFT into a cyclic convolution of
// power-of-two length M >= 2n-1, evaluated with a radix-2 FFT in shared memory:
//     X_k = w_k * sum_j (a_j w_j) conj(w)_{k-j},   w_k = exp(-i pi k^2 / n)
//     A = FFT_M(a w, zero padded)  (decimation in frequency: natural -> bit-reversed order)
//     C = A .* Bhat                (Bhat = FFT_M of the wrapped chirp / M, stored in bit-reversed order)
//     c = IFFT_M(C)                (decimation in time: bit-reversed -> natural order),  X_k = w_k c_k
// so no reordering pass exists.  The inverse transform is conj(forward(conj(.))), unnormalised like FFTW's.  Five passes as in
// the power-of-two path (z r2c, y, x | Gamma | x, y, z c2r) on the same spectrum buffer, natural frequency order along every axis;
// Gamma_hat uses the same builder (gamma.cu) with identity frequency maps.  A correctness path (about 8x the flops of a direct FFT),
// single GPU; the BASELINE grids (2^k) never come here.
#include "internal.h"
#include <cmath>
#include <complex>

struct AnyDev {
    int n, M, logM;
    const double2 *w, *Bhat, *tw;
};

static AnyDev any_dev(const AnyPlan &p) { return AnyDev{p.n, p.M, p.logM, p.w, p.Bhat, p.tw}; }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a conj(b)

// L lines of length pl.n sit in sm[l*M + k] (k < n, rest zero), already multiplied by the chirp; on exit sm[l*M + k] = c_k (before
// the final chirp multiply).  All threads of the CTA must call.
__device__ __forceinline__ void bluestein_core(double2 *sm, const AnyDev &pl, int L)
{
    const int M = pl.M, hM = M >> 1, work = L * hM;
    for (int s = pl.logM - 1; s >= 0; --s) {   // forward FFT_M, decimation in frequency
        const int half = 1 << s, tstep = hM >> s;
        __syncthreads();
        for (int idx = threadIdx.x; idx < work; idx += blockDim.x) {
            const int l = idx / hM, j = idx - l * hM;
            const int pos = j & (half - 1), i0 = l * M + ((j >> s) << (s + 1)) + pos, i1 = i0 + half;
            const double2 a = sm[i0], b = sm[i1];
            sm[i0] = make_double2(a.x + b.x, a.y + b.y);
            sm[i1] = cmul(make_double2(a.x - b.x, a.y - b.y), __ldg(&pl.tw[pos * tstep]));
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < L * M; idx += blockDim.x) sm[idx] = cmul(sm[idx], __ldg(&pl.Bhat[idx & (M - 1)]));
    for (int s = 0; s < pl.logM; ++s) {        // inverse FFT_M, decimation in time
        const int half = 1 << s, tstep = hM >> s;
        __syncthreads();
        for (int idx = threadIdx.x; idx < work; idx += blockDim.x) {
            const int l = idx / hM, j = idx - l * hM;
            const int pos = j & (half - 1), i0 = l * M + ((j >> s) << (s + 1)) + pos, i1 = i0 + half;
            const double2 a = sm[i0], b = cmulc(sm[i1], __ldg(&pl.tw[pos * tstep]));
            sm[i0] = make_double2(a.x + b.x, a.y + b.y);
            sm[i1] = make_double2(a.x - b.x, a.y - b.y);
        }
    }
    __syncthreads();
}

// complex lines along y or x, in place on the spectrum: CTA = L adjacent kz columns of one (component, other-axis) pair
__global__ void k_any_lines(AnyDev pl, double2 *spec, size_t strideA, size_t strideB, int nB, size_t estride, int kzc, int L, int inverse)
{
    extern __shared__ double2 sm[];
    const int a = blockIdx.x / nB, b = blockIdx.x % nB, kz0 = blockIdx.y * L, n = pl.n, M = pl.M;
    double2 *base = spec + (size_t)a * strideA + (size_t)b * strideB + kz0;
    for (int idx = threadIdx.x; idx < L * M; idx += blockDim.x) {
        const int k = idx / L, l = idx - k * L;   // l fastest: adjacent kz columns are adjacent in memory
        double2 v = make_double2(0.0, 0.0);
        if (k < n && kz0 + l < kzc) {
            v = base[(size_t)k * estride + l];
            if (inverse) v.y = -v.y;
            v = cmul(v, __ldg(&pl.w[k]));
        }
        sm[l * M + k] = v;
    }
    bluestein_core(sm, pl, L);
    for (int idx = threadIdx.x; idx < L * n; idx += blockDim.x) {
        const int k = idx / L, l = idx - k * L;
        if (kz0 + l < kzc) {
            double2 v = cmul(sm[l * M + k], __ldg(&pl.w[k]));
            if (inverse) v.y = -v.y;
            base[(size_t)k * estride + l] = v;
        }
    }
}

// r2c along z: L adjacent y rows of one (component, x) plane; real input [c][x][y][z], half spectrum out
__global__ void k_any_z_fwd(AnyDev pl, const double *__restrict__ real, double2 *__restrict__ spec, size_t nloc, int ny, int nz, size_t cStride,
                            size_t xStride, int kzp, int kzc, int nx, int L)
{
    extern __shared__ double2 sm[];
    const int c = blockIdx.x / nx, x = blockIdx.x % nx, y0 = blockIdx.y * L, M = pl.M;
    for (int idx = threadIdx.x; idx < L * M; idx += blockDim.x) {
        const int l = idx / M, k = idx - l * M;
        double2 v = make_double2(0.0, 0.0);
        if (k < nz && y0 + l < ny) {
            const double r = real[(size_t)c * nloc + ((size_t)x * ny + (y0 + l)) * nz + k];
            const double2 wk = __ldg(&pl.w[k]);
            v = make_double2(r * wk.x, r * wk.y);
        }
        sm[idx] = v;
    }
    bluestein_core(sm, pl, L);
    for (int idx = threadIdx.x; idx < L * kzc; idx += blockDim.x) {
        const int l = idx / kzc, k = idx - l * kzc;
        if (y0 + l < ny) spec[(size_t)c * cStride + (size_t)x * xStride + (size_t)(y0 + l) * kzp + k] = cmul(sm[l * M + k], __ldg(&pl.w[k]));
    }
}

// c2r along z (unnormalised, like FFTW): Hermitian extension, inverse transform, out = scale * Re
__global__ void k_any_z_inv(AnyDev pl, const double2 *__restrict__ spec, double *__restrict__ real, size_t nloc, int ny, int nz, size_t cStride,
                            size_t xStride, int kzp, int kzc, int nx, int L, double scale)
{
    extern __shared__ double2 sm[];
    const int c = blockIdx.x / nx, x = blockIdx.x % nx, y0 = blockIdx.y * L, M = pl.M;
    for (int idx = threadIdx.x; idx < L * M; idx += blockDim.x) {
        const int l = idx / M, k = idx - l * M;
        double2 v = make_double2(0.0, 0.0);
        if (k < nz && y0 + l < ny) {
            const double2 *line = spec + (size_t)c * cStride + (size_t)x * xStride + (size_t)(y0 + l) * kzp;
            // X_k for k <= nz/2, conj(X_{nz-k}) above; the inverse is conj(forward(conj X)), so conj once more on the way in
            if (k < kzc) v = make_double2(line[k].x, -line[k].y);
            else v = line[nz - k];
            if (k == 0 || 2 * k == nz) v.y = 0.0;   // c2r ignores the imaginary parts of the self-conjugate bins
            v = cmul(v, __ldg(&pl.w[k]));
        }
        sm[idx] = v;
    }
    bluestein_core(sm, pl, L);
    for (int idx = threadIdx.x; idx < L * nz; idx += blockDim.x) {
        const int l = idx / nz, k = idx - l * nz;
        if (y0 + l < ny) real[(size_t)c * nloc + ((size_t)x * ny + (y0 + l)) * nz + k] = scale * cmul(sm[l * M + k], __ldg(&pl.w[k])).x;
    }
}

// r_hat <- Gamma_hat r_hat per frequency (solver.h:398-407); gamma in the tile layout of gamma.cu with E = 1 (row = x)
template <int H>
__global__ void k_any_gamma(double2 *__restrict__ spec, const double *__restrict__ gamma, int nx, int ny, int kzc, int kzp, size_t cStride,
                            size_t xStride, int T, int nTiles)
{
    constexpr int NG = H * (H + 1) / 2;
    const size_t total = (size_t)nx * ny * kzc, NT = (size_t)nx * T;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int kz = (int)(i % kzc);
        const size_t r = i / kzc;
        const int y = (int)(r % ny), x = (int)(r / ny);
        const double *g = gamma + (((size_t)y * nTiles + kz / T) * NG) * NT + (size_t)x * T + kz % T;
        double2 *s = spec + (size_t)x * xStride + (size_t)y * kzp + kz;
        if (H == 1) {
            const double g0 = g[0];
            s[0] = make_double2(g0 * s[0].x, g0 * s[0].y);
        } else {
            const double g00 = g[0], g01 = g[NT], g02 = g[2 * NT], g11 = g[3 * NT], g12 = g[4 * NT], g22 = g[5 * NT];
            const double2 a = s[0], b = s[cStride], c = s[2 * cStride];
            s[0] = make_double2(g00 * a.x + g01 * b.x + g02 * c.x, g00 * a.y + g01 * b.y + g02 * c.y);
            s[cStride] = make_double2(g01 * a.x + g11 * b.x + g12 * c.x, g01 * a.y + g11 * b.y + g12 * c.y);
            s[2 * cStride] = make_double2(g02 * a.x + g12 * b.x + g22 * c.x, g02 * a.y + g12 * b.y + g22 * c.y);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef std::complex<long double> lcplx;

static void host_fft(std::vector<lcplx> &a)   // radix-2, forward, in place (tables only)
{
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    const long double pi = 3.14159265358979323846264338327950288L;
    for (size_t len = 2; len <= n; len <<= 1)
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const lcplx w = std::polar(1.0L, -2.0L * pi * (long double)k / (long double)len);
                const lcplx u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
}

int any_plan_init(fans_ctx *ctx, AnyPlan &p, int n)
{
    p.n = n;
    p.M = 1, p.logM = 0;
    while (p.M < 2 * n - 1) p.M <<= 1, p.logM++;
    if (p.M < 2) p.M = 2, p.logM = 1;
    const long double pi = 3.14159265358979323846264338327950288L;
    std::vector<double2> w(n), tw(p.M / 2), Bh(p.M);
    std::vector<lcplx> b(p.M, lcplx(0, 0)), wl(n);
    for (int k = 0; k < n; ++k) {
        const long long k2 = ((long long)k * k) % (2LL * n);   // the phase only matters modulo 2 pi
        wl[k] = std::polar(1.0L, -pi * (long double)k2 / (long double)n);
        w[k] = make_double2((double)wl[k].real(), (double)wl[k].imag());
        b[k] = std::conj(wl[k]);
        if (k > 0) b[p.M - k] = std::conj(wl[k]);
    }
    host_fft(b);
    for (int i = 0; i < p.M; ++i) {
        int r = 0;
        for (int q = 0; q < p.logM; ++q)
            if (i & (1 << q)) r |= 1 << (p.logM - 1 - q);
        const lcplx v = b[r] / (long double)p.M;   // position i of the DIF output holds frequency bitrev(i)
        Bh[i] = make_double2((double)v.real(), (double)v.imag());
    }
    for (int j = 0; j < p.M / 2; ++j) {
        const lcplx t = std::polar(1.0L, -2.0L * pi * (long double)j / (long double)p.M);
        tw[j] = make_double2((double)t.real(), (double)t.imag());
    }
    CUDA_TRY(ctx, cudaMalloc(&p.w, sizeof(double2) * n));
    CUDA_TRY(ctx, cudaMalloc(&p.Bhat, sizeof(double2) * p.M));
    CUDA_TRY(ctx, cudaMalloc(&p.tw, sizeof(double2) * (p.M / 2)));
    CUDA_TRY(ctx, cudaMemcpy(p.w, w.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(p.Bhat, Bh.data(), sizeof(double2) * p.M, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(p.tw, tw.data(), sizeof(double2) * (p.M / 2), cudaMemcpyHostToDevice));
    return FANS_OK;
}

void any_plan_free(AnyPlan &p)
{
    if (p.w) cudaFree(p.w);
    if (p.Bhat) cudaFree(p.Bhat);
    if (p.tw) cudaFree(p.tw);
    p.w = p.Bhat = p.tw = nullptr;
}

static int lines_per_cta(int M, int count) { return std::max(1, std::min(std::min(count, 8), (int)(96 * 1024 / (sizeof(double2) * M)))); }

template <class K, class... Args>
static int launch_any(fans_ctx *ctx, K kernel, dim3 grid, size_t smem, Args... args)
{
    if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, 256, smem, ctx->st>>>(args...);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

// out = scale * Gamma * in ;  optional red_out[0] = <dotw, out>
int conv_run_any(fans_ctx *ctx, const double *in, double *out, double scale, const double *dotw, double *red_out)
{
    const int h = ctx->h, nx = ctx->nx, ny = ctx->ny, nz = ctx->nz, kzc = ctx->kzc, kzp = ctx->kzp;
    const size_t xStride = (size_t)ny * kzp + ctx->xpad, cStride = (size_t)nx * xStride;
    const AnyDev px = any_dev(ctx->anyx), py = any_dev(ctx->anyy), pz = any_dev(ctx->anyz);
    const int Lz = lines_per_cta(pz.M, ny), Ly = lines_per_cta(py.M, kzc), Lx = lines_per_cta(px.M, kzc);
    prof_begin(ctx, PC_FFT_Z_FWD);
    FANS_CHECK(launch_any(ctx, k_any_z_fwd, dim3(h * nx, (ny + Lz - 1) / Lz), sizeof(double2) * pz.M * Lz, pz, in, ctx->spec, ctx->nloc, ny, nz,
                          cStride, xStride, kzp, kzc, nx, Lz));
    prof_end(ctx);
    prof_begin(ctx, PC_FFT_Y_FWD);
    FANS_CHECK(launch_any(ctx, k_any_lines, dim3(h * nx, (kzc + Ly - 1) / Ly), sizeof(double2) * py.M * Ly, py, ctx->spec, cStride, xStride, nx,
                          (size_t)kzp, kzc, Ly, 0));
    prof_end(ctx);
    prof_begin(ctx, PC_FFT_X_GAMMA);
    FANS_CHECK(launch_any(ctx, k_any_lines, dim3(h * ny, (kzc + Lx - 1) / Lx), sizeof(double2) * px.M * Lx, px, ctx->spec, cStride, (size_t)kzp, ny,
                          xStride, kzc, Lx, 0));
    {
        const int T = ctx->gT, nTiles = (kzc + T - 1) / T;
        const unsigned nb = (unsigned)std::min<size_t>(((size_t)nx * ny * kzc + 255) / 256, (size_t)FANS_SMS * 16);
        if (h == 1) k_any_gamma<1><<<nb, 256, 0, ctx->st>>>(ctx->spec, ctx->gamma, nx, ny, kzc, kzp, cStride, xStride, T, nTiles);
        else k_any_gamma<3><<<nb, 256, 0, ctx->st>>>(ctx->spec, ctx->gamma, nx, ny, kzc, kzp, cStride, xStride, T, nTiles);
        ctx->launches++;
        CUDA_TRY(ctx, cudaGetLastError());
    }
    FANS_CHECK(launch_any(ctx, k_any_lines, dim3(h * ny, (kzc + Lx - 1) / Lx), sizeof(double2) * px.M * Lx, px, ctx->spec, cStride, (size_t)kzp, ny,
                          xStride, kzc, Lx, 1));
    prof_end(ctx);
    prof_begin(ctx, PC_FFT_Y_INV);
    FANS_CHECK(launch_any(ctx, k_any_lines, dim3(h * nx, (kzc + Ly - 1) / Ly), sizeof(double2) * py.M * Ly, py, ctx->spec, cStride, xStride, nx,
                          (size_t)kzp, kzc, Ly, 1));
    prof_end(ctx);
    prof_begin(ctx, PC_FFT_Z_INV);
    FANS_CHECK(launch_any(ctx, k_any_z_inv, dim3(h * nx, (ny + Lz - 1) / Lz), sizeof(double2) * pz.M * Lz, pz, ctx->spec, out, ctx->nloc, ny, nz,
                          cStride, xStride, kzp, kzc, nx, Lz, scale));
    prof_end(ctx);
    if (red_out) {   // <dotw, out> as a separate reduction (the power-of-two path fuses it into the c2r epilogue)
        FANS_CHECK(vec_reduce4(ctx, dotw, out, ctx->d_red + S_GEN));
        CUDA_TRY(ctx, cudaMemcpyAsync(red_out, ctx->d_red + S_GEN + 2, sizeof(double), cudaMemcpyDeviceToDevice, ctx->st));
    }
    return FANS_OK;
}
