// fft_x.cu — pass P3 of the convolution (include/solver.h:395-409): forward transform along x, the Fourier-space
// fundamental-solution multiply  r_hat <- Gamma_hat(xi) r_hat  (solver.h:399-406) and the inverse transform along x, fused:
// all `howmany` components of a (y, kz-tile) pencil live in the registers of one CTA, so the h x h real-symmetric multiply
// happens between the last forward butterfly and the first inverse butterfly without touching shared or global memory.
// Gamma_hat is streamed once per iteration in exactly the register order of the threads (gamma.cu writes it that way):
//     gamma[((cta*NG + k) * (N/E) * E + e*(N/E) + jt) * T + t],   storage row = jt*E + e,   NG = H(H+1)/2 upper triangle.
#include "fft_reg.cuh"
#include "internal.h"

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

template <int N, int H, int T>
__global__ void __launch_bounds__((N / rp_elems(N)) * T, ((N / rp_elems(N)) * T <= 256 && H == 3) ? 2 : 1)
    k_fft_xg(double2 *__restrict__ spec, const double *__restrict__ gamma, const double2 *__restrict__ tw, SpecGeom g, int nTiles)
{
    extern __shared__ double2 sm[];
    constexpr int E = rp_elems(N), TPC = N / E, NST = rp_nstages(N), NG = H * (H + 1) / 2;
    const int t = threadIdx.x % T, jt = threadIdx.x / T;
    const int tile = blockIdx.x % nTiles, yl = blockIdx.x / nTiles;
    double2 *base = spec + (size_t)yl * g.kzp + (size_t)tile * T + t;
    const double *gam = gamma + (size_t)blockIdx.x * NG * (N * T) + jt * T + t;
    const TileIdxX<T> idx{t};
    double2 a[H][E];
#pragma unroll
    for (int c = 0; c < H; ++c)
#pragma unroll
        for (int e = 0; e < E; ++e) a[c][e] = base[(size_t)c * g.cStride + spec_row_x(g, rp_row<N, 0>(jt, e))];
    rp_forward<N, H>(a, jt, sm, N * T, idx, tw, 1);
    // Green operator on the registers (storage row of register e after the last stage = jt*E + e)
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const double *ge = gam + (size_t)e * TPC * T;
        if (H == 1) {
            const double g0 = __ldg(ge);
            a[0][e] = make_double2(g0 * a[0][e].x, g0 * a[0][e].y);
        } else {
            constexpr size_t NT = (size_t)N * T;
            const double g00 = __ldg(ge), g01 = __ldg(ge + NT), g02 = __ldg(ge + 2 * NT);
            const double g11 = __ldg(ge + 3 * NT), g12 = __ldg(ge + 4 * NT), g22 = __ldg(ge + 5 * NT);
            const double2 r0 = a[0][e], r1 = a[H > 1 ? 1 : 0][e], r2 = a[H > 2 ? 2 : 0][e];
            a[0][e] = make_double2(g00 * r0.x + g01 * r1.x + g02 * r2.x, g00 * r0.y + g01 * r1.y + g02 * r2.y);
            a[H > 1 ? 1 : 0][e] = make_double2(g01 * r0.x + g11 * r1.x + g12 * r2.x, g01 * r0.y + g11 * r1.y + g12 * r2.y);
            a[H > 2 ? 2 : 0][e] = make_double2(g02 * r0.x + g12 * r1.x + g22 * r2.x, g02 * r0.y + g12 * r1.y + g22 * r2.y);
        }
    }
    rp_inverse<N, H>(a, jt, sm, N * T, idx, tw, 1);
#pragma unroll
    for (int c = 0; c < H; ++c)
#pragma unroll
        for (int e = 0; e < E; ++e) base[(size_t)c * g.cStride + spec_row_x(g, rp_row<N, 0>(jt, e))] = a[c][e];
    (void)NST;
}

template <int N, int H, int T>
static int launch_xg(fans_ctx *ctx, double2 *specB, const SpecGeom &g)
{
    constexpr int E = rp_elems(N);
    const int nTiles = (ctx->kzc + T - 1) / T;
    const size_t smem = sizeof(double2) * N * T * H;
    if (smem > 227 * 1024) {
        fans_set_error(ctx, FANS_ERR_ARG, "x pass tile does not fit shared memory");
        return FANS_ERR_ARG;
    }
    if (smem > 48 * 1024) CUDA_TRY(ctx, cudaFuncSetAttribute(k_fft_xg<N, H, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((size_t)ctx->n1 * nTiles);
    k_fft_xg<N, H, T><<<grid, (N / E) * T, smem, ctx->st>>>(specB, ctx->gamma, ctx->planx.tw, g, nTiles);
    return FANS_OK;
}

// tile width of the fused x pass for a given (nx, howmany): keeps the exchange tile within ~100 KB
int fft_x_tile_width(int nx, int h)
{
    int T = (h == 1) ? 8 : 4;
    while ((size_t)h * nx * T * sizeof(double2) > 100 * 1024 && T > 2) T /= 2;
    return T;
}

int fft_pass_x_gamma(fans_ctx *ctx)
{
    prof_begin(ctx, PC_FFT_X_GAMMA);
    SpecGeom g = spec_geom_A(ctx);
    double2 *specB = ctx->P > 1 ? ctx->specB : ctx->spec;
    int rc = FANS_ERR_ARG;
    const int T = ctx->gT;
#define X_CASE(N_)                                                                                     \
    case N_:                                                                                           \
        if (ctx->h == 1) rc = (T == 8) ? launch_xg<N_, 1, 8>(ctx, specB, g) : (T == 4 ? launch_xg<N_, 1, 4>(ctx, specB, g) : launch_xg<N_, 1, 2>(ctx, specB, g)); \
        else rc = (T == 4) ? launch_xg<N_, 3, 4>(ctx, specB, g) : launch_xg<N_, 3, 2>(ctx, specB, g);  \
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
