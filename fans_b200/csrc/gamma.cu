// gamma.cu — fundamental solution (Green operator) of the reference medium, built on the device.
// Follows Solver::computeFundamentalSolution (include/solver.h:144-204):
//   per frequency xi=(kx,ky,kz<=nz/2):  A = [1, ex, ey, ex ey, ez, ex ez, ez ey, ex ey ez], e. = exp(2 pi i k./n.)
//   AA = Re A Re A^T + Im A Im A^T ;  block(i,j) = sum(Ker0[8i:8i+8, 8j:8j+8] o AA)
//   Gamma_hat = pinv(block) with ABSOLUTE singular-value cut 1e-14 (solver.h:89-96,189-191), xi = 0 left zero (:169),
//   scaled by 1/(nx ny nz) (:198).
// Stored in the tile-major, register-ordered layout the fused x pass streams (fft_x.cu, k_fft_xg) and in the
// digit-reversed frequency order the DIF transforms produce along x and y.
#include "common.cuh"

__device__ __forceinline__ void jacobi_pinv3(double a[3][3], double tol, double out[6])
{
    double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double dia = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-34 * dia || off == 0.0) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
            const double apq = a[p][q];
            if (apq == 0.0) continue;
            const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
            for (int k = 0; k < 3; ++k) {  // A <- A J
                const double akp = a[k][p], akq = a[k][q];
                a[k][p] = c * akp - s * akq;
                a[k][q] = s * akp + c * akq;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {  // A <- J^T A
                const double apk = a[p][k], aqk = a[q][k];
                a[p][k] = c * apk - s * aqk;
                a[q][k] = s * apk + c * aqk;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double vkp = v[k][p], vkq = v[k][q];
                v[k][p] = c * vkp - s * vkq;
                v[k][q] = s * vkp + c * vkq;
            }
        }
    }
    double inv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) inv[i] = (fabs(a[i][i]) > tol) ? 1.0 / a[i][i] : 0.0;
    int n = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i; j < 3; ++j)
            out[n++] = inv[0] * v[i][0] * v[j][0] + inv[1] * v[i][1] * v[j][1] + inv[2] * v[i][2] * v[j][2];
}

template <int H>
__global__ void k_build_gamma(double *__restrict__ gamma, const double *__restrict__ Ker0, const int *__restrict__ frqx,
                              const int *__restrict__ frqy, int nx, int ny, int nz, int n1, int y1, int kzc, int T,
                              int nTiles, double invN, int E)
{
    constexpr int NG = H * (H + 1) / 2;
    const size_t NT = (size_t)nx * T;
    const size_t total = (size_t)n1 * nTiles * NT;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx % NT);
        const size_t ot = idx / NT;
        const int tile = (int)(ot % nTiles);
        const int o = (int)(ot / nTiles);
        // register order of the fused x pass (fft_x.cu): i = (e*(nx/E) + jt)*T + t, storage row = jt*E + e
        const int t = i % T, qq = i / T, TPC = nx / E;
        const int row = (qq % TPC) * E + qq / TPC;
        const int kz = tile * T + t;
        double g[NG];
#pragma unroll
        for (int k = 0; k < NG; ++k) g[k] = 0.0;
        const int kx = frqx[row], ky = frqy[y1 + o];
        if (kz < kzc && !(kx == 0 && ky == 0 && kz == 0)) {
            double sx, cx, sy, cy, sz, cz;
            sincospi(2.0 * (double)kx / (double)nx, &sx, &cx);
            sincospi(2.0 * (double)ky / (double)ny, &sy, &cy);
            sincospi(2.0 * (double)kz / (double)nz, &sz, &cz);
            double ar[8], ai[8];
            ar[0] = 1.0, ai[0] = 0.0;
            ar[1] = cx, ai[1] = sx;
            ar[2] = cy, ai[2] = sy;
            ar[3] = cx * cy - sx * sy, ai[3] = cx * sy + sx * cy;
            ar[4] = cz, ai[4] = sz;
            ar[5] = cx * cz - sx * sz, ai[5] = cx * sz + sx * cz;
            ar[6] = cz * cy - sz * sy, ai[6] = cz * sy + sz * cy;
            ar[7] = ar[3] * cz - ai[3] * sz, ai[7] = ar[3] * sz + ai[3] * cz;
            double blk[H][H];
#pragma unroll
            for (int bi = 0; bi < H; ++bi)
#pragma unroll
                for (int bj = bi; bj < H; ++bj) {
                    double s = 0.0;
                    for (int a = 0; a < 8; ++a)
                        for (int b = 0; b < 8; ++b)
                            s += __ldg(&Ker0[(8 * bi + a) * (8 * H) + 8 * bj + b]) * (ar[a] * ar[b] + ai[a] * ai[b]);
                    blk[bi][bj] = s;
                    blk[bj][bi] = s;
                }
            if (H == 1) {
                g[0] = (fabs(blk[0][0]) > 1e-14) ? invN / blk[0][0] : 0.0;
            } else {
                double a3[3][3];
#pragma unroll
                for (int bi = 0; bi < 3; ++bi)
#pragma unroll
                    for (int bj = 0; bj < 3; ++bj) a3[bi][bj] = blk[bi % H][bj % H];
                double o6[6];
                jacobi_pinv3(a3, 1e-14, o6);
#pragma unroll
                for (int k = 0; k < NG; ++k) g[k] = o6[k % 6] * invN;
            }
        }
#pragma unroll
        for (int k = 0; k < NG; ++k) gamma[(ot * NG + k) * NT + i] = g[k];
    }
}

int gamma_build(fans_ctx *ctx, const double *Ker0_dev, const int *frqx, const int *frqy)
{
    prof_begin(ctx, PC_OTHER);
    const int T = ctx->gT;
    const int nTiles = (ctx->kzc + T - 1) / T;
    const size_t total = (size_t)ctx->n1 * nTiles * ctx->nx * T;
    const int nthr = 128;
    size_t nb = (total + nthr - 1) / nthr;
    if (nb > (size_t)FANS_SMS * 64) nb = (size_t)FANS_SMS * 64;
    const double invN = 1.0 / ((double)ctx->nx * (double)ctx->ny * (double)ctx->nz);
    if (ctx->h == 1)
        k_build_gamma<1><<<(unsigned)nb, nthr, 0, ctx->st>>>(ctx->gamma, Ker0_dev, frqx, frqy, ctx->nx, ctx->ny, ctx->nz, ctx->n1, ctx->y1,
                                                           ctx->kzc, T, nTiles, invN, ctx->gE);
    else
        k_build_gamma<3><<<(unsigned)nb, nthr, 0, ctx->st>>>(ctx->gamma, Ker0_dev, frqx, frqy, ctx->nx, ctx->ny, ctx->nz, ctx->n1, ctx->y1,
                                                           ctx->kzc, T, nTiles, invN, ctx->gE);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}
