// fft.cuh — shared-memory tile FFT core (FP64, mixed radix 8/4/2, power-of-two lengths).
//
// Every pass of the 3-D transform works on a shared-memory tile  sm[row][t]:  `row` runs along the
// transform axis (N points), `t` over T independent columns that are CONTIGUOUS in global memory
// (T complex = T*16 B segments).  The forward transform is an in-place decimation-in-frequency
// (natural order in -> digit-reversed order out); the inverse is the matching decimation-in-time
// (digit-reversed in -> natural out, unnormalised).  The spectrum therefore lives in digit-reversed
// order along x and y and is never reordered: the Green operator is built in that same order
// (gamma.cu), which is all a convolution needs.  Replaces FFTW's r2c/c2r plans (include/solver.h:206-226).
#pragma once
#include "common.cuh"

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b)  // a * conj(b)
{
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// swizzled tile index: every quarter-warp access (8 x 16 B) stays bank-conflict free for
//   (a) butterflies (8 columns of one row / 2 rows x 4 columns),
//   (b) the transposing load/store of the z pass (8 consecutive rows of one column, T = 8).
template <int T>
__device__ __forceinline__ int tix(int row, int t)
{
    if (T == 8) return row * 8 + (t ^ (row & 7));
    if (T == 4) return (row * 4 + t) ^ (((row >> 3) & 1) << 2);
    return row * T + t;
}

template <bool INV>
__device__ __forceinline__ void dft4(double2 &a0, double2 &a1, double2 &a2, double2 &a3)
{
    double2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    if (!INV) {  // w4 = -i
        a1 = make_double2(t1.x + t3.y, t1.y - t3.x);
        a3 = make_double2(t1.x - t3.y, t1.y + t3.x);
    } else {
        a1 = make_double2(t1.x - t3.y, t1.y + t3.x);
        a3 = make_double2(t1.x + t3.y, t1.y - t3.x);
    }
}

template <int R, bool INV>
__device__ __forceinline__ void dft_r(double2 (&a)[R])
{
    if (R == 2) {
        double2 t = a[0];
        a[0] = cadd(t, a[1]);
        a[1] = csub(t, a[1]);
    } else if (R == 4) {
        dft4<INV>(a[0], a[1], a[2], a[3]);
    } else if (R == 8) {
        const double c = 0.70710678118654752440;
        // even / odd 4-point transforms
        dft4<INV>(a[0], a[2], a[4], a[6]);  // E0..E3 in a0,a2,a4,a6
        dft4<INV>(a[1], a[3], a[5], a[7]);  // O0..O3 in a1,a3,a5,a7
        double2 o1, o2, o3;
        if (!INV) {
            o1 = make_double2(c * (a[3].x + a[3].y), c * (a[3].y - a[3].x));
            o2 = make_double2(a[5].y, -a[5].x);
            o3 = make_double2(c * (a[7].y - a[7].x), -c * (a[7].x + a[7].y));
        } else {
            o1 = make_double2(c * (a[3].x - a[3].y), c * (a[3].x + a[3].y));
            o2 = make_double2(-a[5].y, a[5].x);
            o3 = make_double2(-c * (a[7].x + a[7].y), c * (a[7].x - a[7].y));
        }
        double2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1];
        a[0] = cadd(e0, o0);
        a[4] = csub(e0, o0);
        a[1] = cadd(e1, o1);
        a[5] = csub(e1, o1);
        a[2] = cadd(e2, o2);
        a[6] = csub(e2, o2);
        a[3] = cadd(e3, o3);
        a[7] = csub(e3, o3);
    }
}

// One radix-R stage over sub-blocks of length L = R*M (M = 1<<logM) on NB stacked tiles of N rows each.
template <int R, int T, bool INV>
__device__ __forceinline__ void fft_stage(double2 *sm, int N, int logN, int logM, const double2 *__restrict__ tw,
                                          int tws, int nbatch, int tid, int nthr)
{
    constexpr int logT = (T == 8) ? 3 : (T == 4 ? 2 : (T == 2 ? 1 : 0));
    constexpr int logR = (R == 8) ? 3 : (R == 4 ? 2 : 1);
    const int M = 1 << logM;
    const int lognbf = logN - logR + logT;  // butterflies x columns per tile
    const int total = nbatch << lognbf;
    for (int w = tid; w < total; w += nthr) {
        const int bi = w >> lognbf;
        const int wb = w & ((1 << lognbf) - 1);
        const int t = wb & (T - 1);
        const int q = wb >> logT;
        const int j = q & (M - 1);
        const int row0 = ((q >> logM) << (logM + logR)) + j;
        double2 *base = sm + ((size_t)bi << (logN + logT));
        double2 a[R];
#pragma unroll
        for (int k = 0; k < R; ++k) a[k] = base[tix<T>(row0 + (k << logM), t)];
        if (!INV) {
            dft_r<R, false>(a);
            if (M > 1) {
#pragma unroll
                for (int m = 1; m < R; ++m) a[m] = cmul(a[m], __ldg(&tw[(j * m) * tws]));
            }
        } else {
            if (M > 1) {
#pragma unroll
                for (int m = 1; m < R; ++m) a[m] = cmulc(a[m], __ldg(&tw[(j * m) * tws]));
            }
            dft_r<R, true>(a);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) base[tix<T>(row0 + (k << logM), t)] = a[k];
    }
}

// Full in-place transform of `nbatch` stacked tiles. Ends with a __syncthreads().
template <int T, bool INV>
__device__ __forceinline__ void fft_tile(double2 *sm, const FftStages &st, const double2 *__restrict__ tw, int nbatch,
                                         int tid, int nthr)
{
    if (!INV) {
        int logL = st.logN;
        for (int s = 0; s < st.nst; ++s) {
            const int R = st.radix[s];
            const int logR = (R == 8) ? 3 : (R == 4 ? 2 : 1);
            const int logM = logL - logR;
            const int tws = st.twmul << (st.logN - logL);  // omega_L = omega_ntab^(ntab/L)
            if (R == 8) fft_stage<8, T, false>(sm, st.N, st.logN, logM, tw, tws, nbatch, tid, nthr);
            else if (R == 4) fft_stage<4, T, false>(sm, st.N, st.logN, logM, tw, tws, nbatch, tid, nthr);
            else fft_stage<2, T, false>(sm, st.N, st.logN, logM, tw, tws, nbatch, tid, nthr);
            logL = logM;
            __syncthreads();
        }
    } else {
        int logM = 0;
        for (int s = st.nst - 1; s >= 0; --s) {
            const int R = st.radix[s];
            const int logR = (R == 8) ? 3 : (R == 4 ? 2 : 1);
            const int logL = logM + logR;
            const int tws = st.twmul << (st.logN - logL);
            if (R == 8) fft_stage<8, T, true>(sm, st.N, st.logN, logM, tw, tws, nbatch, tid, nthr);
            else if (R == 4) fft_stage<4, T, true>(sm, st.N, st.logN, logM, tw, tws, nbatch, tid, nthr);
            else fft_stage<2, T, true>(sm, st.N, st.logN, logM, tw, tws, nbatch, tid, nthr);
            logM = logL;
            __syncthreads();
        }
    }
}
