// fft.cu — twiddle / digit-reversal tables of the power-of-two FFT plans (the kernels live in fft_y.cu, fft_x.cu, fft_z.cu).
#include "internal.h"
#include <cmath>

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
    tw.resize(2 * (size_t)ntab);  // second half: conjugates, read by the inverse transforms (fft_reg.cuh, rp_stage)
    for (int i = 0; i < ntab; ++i) tw[ntab + i] = make_double2(tw[i].x, -tw[i].y);
    CUDA_TRY(ctx, cudaMalloc(&p.tw, sizeof(double2) * 2 * ntab));
    CUDA_TRY(ctx, cudaMemcpy(p.tw, tw.data(), sizeof(double2) * 2 * ntab, cudaMemcpyHostToDevice));
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

