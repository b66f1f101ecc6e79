// fft_reg.cuh — register-resident power-of-two FFT building blocks (FP64, radix 8/4/2).
//
// A transform of length N is carried by N/E threads (E = min(8, N) elements per thread).  Every stage is one radix-R
// butterfly set on registers; between two stages the elements are exchanged through a shared-memory tile.  Forward =
// in-place decimation in frequency (natural order in, digit-reversed order out), inverse = the exact mirror (decimation in
// time, digit-reversed in, natural out, unnormalised like FFTW).  The spectrum therefore stays in digit-reversed order along
// x and y for its whole life; the Green operator is built in that order (gamma.cu).
//
// Stage s of the plan (radices r_0..r_{S-1}, r = 8,...,8[,4|2]):  L_s = N / (r_0..r_{s-1}),  M_s = L_s / r_s.
// Butterfly q = tid*(E/r_s) + i  ->  j = q mod M_s, b = q div M_s,  rows  b*L_s + j + k*M_s  (k < r_s),  register a[i*r_s + k].
// Replaces FFTW's codelets behind the plans of include/solver.h:206-226.
#pragma once
#include "common.cuh"

__host__ __device__ constexpr int rp_log2(int n) { return n <= 1 ? 0 : 1 + rp_log2(n / 2); }
__host__ __device__ constexpr int rp_elems(int N) { return N < 8 ? N : 8; }
__host__ __device__ constexpr int rp_nstages(int N) { return rp_log2(N) / 3 + (rp_log2(N) % 3 != 0 ? 1 : 0); }
__host__ __device__ constexpr int rp_radix(int N, int s)
{
    return s < rp_log2(N) / 3 ? 8 : (rp_log2(N) % 3 == 1 ? 2 : 4);
}
__host__ __device__ constexpr int rp_L(int N, int s)
{
    int L = N;
    for (int q = 0; q < s; ++q) L /= rp_radix(N, q);
    return L;
}

__device__ __forceinline__ double2 rc_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 rc_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 rc_mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 rc_mulc(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a conj(b)

template <bool INV>
__device__ __forceinline__ void rdft2(double2 &a0, double2 &a1)
{
    const double2 t = a0;
    a0 = rc_add(t, a1);
    a1 = rc_sub(t, a1);
}
template <bool INV>
__device__ __forceinline__ void rdft4(double2 &a0, double2 &a1, double2 &a2, double2 &a3)
{
    const double2 t0 = rc_add(a0, a2), t1 = rc_sub(a0, a2), t2 = rc_add(a1, a3), t3 = rc_sub(a1, a3);
    a0 = rc_add(t0, t2);
    a2 = rc_sub(t0, t2);
    if (!INV) {  // w4 = -i
        a1 = make_double2(t1.x + t3.y, t1.y - t3.x);
        a3 = make_double2(t1.x - t3.y, t1.y + t3.x);
    } else {
        a1 = make_double2(t1.x - t3.y, t1.y + t3.x);
        a3 = make_double2(t1.x + t3.y, t1.y - t3.x);
    }
}
template <bool INV>
__device__ __forceinline__ void rdft8(double2 *a)
{
    const double c = 0.70710678118654752440;
    rdft4<INV>(a[0], a[2], a[4], a[6]);
    rdft4<INV>(a[1], a[3], a[5], a[7]);
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
    const double2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1];
    a[0] = rc_add(e0, o0);
    a[4] = rc_sub(e0, o0);
    a[1] = rc_add(e1, o1);
    a[5] = rc_sub(e1, o1);
    a[2] = rc_add(e2, o2);
    a[6] = rc_sub(e2, o2);
    a[3] = rc_add(e3, o3);
    a[7] = rc_sub(e3, o3);
}

// row owned by register e = i*R + k of thread `tid` in stage S
template <int N, int S>
__device__ __forceinline__ int rp_row(int tid, int e)
{
    constexpr int R = rp_radix(N, S), L = rp_L(N, S), M = L / R, NB = rp_elems(N) / R;
    constexpr int logM = rp_log2(M), logL = rp_log2(L), logR = rp_log2(R);
    const int i = e >> logR, k = e & (R - 1);
    const int q = tid * NB + i;
    return ((q >> logM) << logL) + (q & (M - 1)) + (k << logM);
}

// butterflies + twiddles of stage S on the registers.  tw: table of exp(-2 pi i n / ntab) followed by its conjugate, twmul = ntab / N.
template <int N, int S, bool INV>
__device__ __forceinline__ void rp_stage(double2 *a, int tid, const double2 *__restrict__ tw, int twmul)
{
    constexpr int R = rp_radix(N, S), L = rp_L(N, S), M = L / R, NB = rp_elems(N) / R;
    const int tws = twmul * (N / L);
    // The inverse reads the CONJUGATE half of the table (tw[ntab + n] = conj(tw[n])): were both directions to share the
    // same loads, the compiler would keep every twiddle of the forward transform alive (spilled) until the inverse needs it.
    if (INV) tw += twmul * N;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        double2 *b = a + i * R;
        const int j = (tid * NB + i) & (M - 1);
        double2 w[R];
        if (M > 1) {
            w[1] = __ldg(&tw[j * tws]);
            if (R >= 4) {
                w[2] = __ldg(&tw[2 * j * tws]);
                w[3] = rc_mul(w[1], w[2]);
            }
            if (R == 8) {
                w[4] = __ldg(&tw[4 * j * tws]);
                w[5] = rc_mul(w[1], w[4]);
                w[6] = rc_mul(w[2], w[4]);
                w[7] = rc_mul(w[3], w[4]);
            }
        }
        if (INV && M > 1) {
#pragma unroll
            for (int m = 1; m < R; ++m) b[m] = rc_mul(b[m], w[m]);
        }
        if (R == 8) rdft8<INV>(b);
        else if (R == 4) rdft4<INV>(b[0], b[1], b[2], b[3]);
        else rdft2<INV>(b[0], b[1]);
        if (!INV && M > 1) {
#pragma unroll
            for (int m = 1; m < R; ++m) b[m] = rc_mul(b[m], w[m]);
        }
    }
}

// shared-memory exchange: IDX(row) -> element index inside the caller's tile
template <int N, int S, class IDX>
__device__ __forceinline__ void rp_put(const double2 *a, int tid, double2 *sm, IDX idx)
{
#pragma unroll
    for (int e = 0; e < rp_elems(N); ++e) sm[idx(rp_row<N, S>(tid, e))] = a[e];
}
template <int N, int S, class IDX>
__device__ __forceinline__ void rp_get(double2 *a, int tid, const double2 *sm, IDX idx)
{
#pragma unroll
    for (int e = 0; e < rp_elems(N); ++e) a[e] = sm[idx(rp_row<N, S>(tid, e))];
}

// barrier policy of the stage exchanges: the whole CTA (default) or one named barrier per thread group (x pass: the H component
// groups run their transforms independently, so one group's butterflies overlap another group's shared-memory exchange)
struct SyncCta {
    __device__ __forceinline__ void operator()() const { __syncthreads(); }
};
struct SyncWarp {  // every transform lives inside one warp (z pass, nz <= 512): no CTA barrier at all
    __device__ __forceinline__ void operator()() const { __syncwarp(); }
};
template <bool NAMED>  // NAMED only when the group is a whole number of warps and not the whole CTA
struct SyncGroup {
    int id, count;
    __device__ __forceinline__ void operator()() const
    {
        if (NAMED) asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory");
        else __syncthreads();
    }
};

// Full forward transform of the registers that were loaded with the stage-0 mapping; leaves the last-stage mapping.
// NC independent transforms (components) share the barriers; sm holds NC tiles of `tile_elems` double2.
template <int N, int NC, class IDX, int S = 0, class SYNC = SyncCta>
__device__ __forceinline__ void rp_forward(double2 (*a)[rp_elems(N)], int tid, double2 *sm, int tile_elems, IDX idx,
                                           const double2 *__restrict__ tw, int twmul, SYNC sync = SYNC())
{
    if constexpr (S < rp_nstages(N)) {
        if constexpr (S > 0) {
#pragma unroll
            for (int c = 0; c < NC; ++c) rp_put<N, S - 1>(a[c], tid, sm + c * tile_elems, idx);
            sync();
#pragma unroll
            for (int c = 0; c < NC; ++c) rp_get<N, S>(a[c], tid, sm + c * tile_elems, idx);
            if constexpr (S + 1 < rp_nstages(N)) sync();
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) rp_stage<N, S, false>(a[c], tid, tw, twmul);
        rp_forward<N, NC, IDX, S + 1, SYNC>(a, tid, sm, tile_elems, idx, tw, twmul, sync);
    }
}

// Full inverse transform: registers hold the last-stage mapping on entry, the stage-0 mapping (natural rows) on exit.
// `first_sync`: the caller has used sm before (needs a barrier before the first put).
template <int N, int NC, class IDX, int S = rp_nstages(N) - 1, class SYNC = SyncCta>
__device__ __forceinline__ void rp_inverse(double2 (*a)[rp_elems(N)], int tid, double2 *sm, int tile_elems, IDX idx,
                                           const double2 *__restrict__ tw, int twmul, SYNC sync = SYNC())
{
    if constexpr (S >= 0) {
#pragma unroll
        for (int c = 0; c < NC; ++c) rp_stage<N, S, true>(a[c], tid, tw, twmul);
        if constexpr (S > 0) {
            sync();  // previous readers of sm are done
#pragma unroll
            for (int c = 0; c < NC; ++c) rp_put<N, S>(a[c], tid, sm + c * tile_elems, idx);
            sync();
#pragma unroll
            for (int c = 0; c < NC; ++c) rp_get<N, S - 1>(a[c], tid, sm + c * tile_elems, idx);
        }
        rp_inverse<N, NC, IDX, S - 1, SYNC>(a, tid, sm, tile_elems, idx, tw, twmul, sync);
    }
}

// where the spectrum lives (single GPU: one block; P ranks: P blocks of [h][n0][n1][kzp])
struct SpecGeom {
    int n0, n1, l2n0, l2n1, kzp, h;
    size_t xStride;    // n1*kzp + pad       (one x plane; the pad keeps the 2^k-strided rows of the x pass off the same DRAM channels)
    size_t cStride;    // n0*xStride         (one component inside a block)
    size_t blkStride;  // h*cStride          (one rank block)
};
__device__ __forceinline__ size_t spec_row_y(const SpecGeom &g, int y) { return (size_t)(y >> g.l2n1) * g.blkStride + (size_t)(y & (g.n1 - 1)) * g.kzp; }
__device__ __forceinline__ size_t spec_row_x(const SpecGeom &g, int x)
{
    return (size_t)(x >> g.l2n0) * g.blkStride + (size_t)(x & (g.n0 - 1)) * g.xStride;
}

// Fused transpose: where the last pass before a transpose stores its rows.  on == 0: locally (blocks of this rank's buffer,
// exchanged afterwards with ncclSend/ncclRecv); on == 1: straight into the owning rank's buffer (peer-mapped through CUDA
// IPC, comm.cu), block `me` there -- the data crosses NVLink once, tile by tile, while the other CTAs keep computing.
struct PeerTable {
    double2 *p[8];
    int me, on;
};
__device__ __forceinline__ double2 *peer_select(const PeerTable &pt, int q)
{
    double2 *sel = pt.p[0];
#pragma unroll
    for (int i = 1; i < 8; ++i)
        if (q == i) sel = pt.p[i];
    return sel;
}
