// sweep.cu — the fused element sweep: gradient stencil -> material law -> B^T sigma -> nodal assembly in ONE pass.
// Replaces Solver::compute_residual_basic / iterateCubes (include/solver.h:229-385) with Matmodel::element_residual
// (include/matmodel.h:190-201), the linear-operator lambda of SolverCG (include/solverCG.h:98-103) and
// Matmodel::getStrainStress (include/matmodel.h:202-225, driven from solver.h:707-737).
//
// Decomposition: a CTA owns a (TY x TZ) column of nodes and marches along x.  Per x step it
//   (a) loads ONE new node plane (with a one-node halo in y,z) into a 2-slot shared-memory ring
//       [optionally forming d = s + beta*d on the fly and writing the owned part back],
//   (b) evaluates its (TY+1)x(TZ+1) elements (the low-side halo elements are recomputed by two extra warps so
//       that no inter-CTA exchange and no atomics are needed) and deposits the 8h element forces in a staging tile,
//   (c) each node thread gathers its 4 contributions of this element plane, adds the 4 carried in registers from the
//       previous plane, writes the node force (coalesced along z) and keeps the next carry.
// The assembly order is fixed => bitwise reproducible results.  Halo voxels are re-read from shared memory only.
#include "internal.h"
#include "materials.cuh"
#include <cmath>
#include <type_traits>

#define TY 8
#define TZ 32
#define NTILE ((TY + 2) * (TZ + 2))       // node tile incl. halo
#define NELT ((TY + 1) * (TZ + 1))        // elements evaluated per plane
#define SWEEP_THREADS (TY * TZ + 64)

// No __constant__ tables: the basic gradients travel inside the kernel parameters (SweepParams::bg) and the phase stiffness table of
// the fallback K.d sweep is read from the context's global-memory table, so two contexts never share mutable device state.

enum { SW_LINEAR = 0, SW_RESIDUAL = 1, SW_STRAINSTRESS = 2 };

struct SweepParams {
    int n0, ny, nz;          // local planes, global ny, nz
    size_t nloc;             // n0*ny*nz = component stride of the SoA fields
    int xchunk;              // node planes per CTA
    const double *in;        // u (or d_old when in2 != nullptr)
    const double *in2;       // s : input becomes in2 + beta*in  (fused CG direction update)
    double *in_out;          // where the fused input is written (d_new)
    const double *beta;      // device scalar
    double *out;             // nodal result
    const uint16_t *phidx;
    const PhaseDev *phases;
    int n_phases;
    const double *Kglob;     // phase stiffness table in global memory (used when it does not fit __constant__)
    int n_k;
    double g0[9];
    double bg[9 * 3 * 8];    // basic gradient dN_a/dx_j at Gauss point g: [(g*3 + j)*8 + a]; g = 8: element centre
    double vw;               // v_e / n_gp
    int ngp, bbar;
    double *hist, *hist_t;
    const unsigned *hidx;    // compact history index per element (0xffffffff: none)
    size_t nh;               // history-bearing elements (stride of the compact history arrays)
    int *pflag;
    double gm, gp, il[3];    // sum-factorised path: Gauss coordinates 0.5 -/+ sqrt(3)/6 and 1 / element length per axis
    int stg2;                // k_sweep_sf: staging tile double-buffered (one barrier per plane)
    int one_point;           // strain/stress sweep of an all-linear problem: element averages = values at the element centre
    int hstage;              // 1: history of Gauss point g+1 is staged in shared memory (cp.async) while g is evaluated
    int *fault;
    // reductions
    double *part;
    unsigned int *ticket;
    double *red_out;         // SW_LINEAR: <in_new, out>;  SW_STRAINSTRESS: sum of element stress (n_str values)
    // strain/stress output (optional)
    double *eps_out, *sig_out;   // [n_str][nloc] element averages
    double *eps_gp, *sig_gp;     // [n_str][ngp][nloc] every Gauss point (strain_gp / stress_gp, solver.h:534-542), optional
    // slab decomposition (world_size > 1), scatter form across the slab boundary like the reference (solver.h:244-269):
    const double *in_hi;         // node plane n0 of the input (= plane 0 of the next rank), [H][ny*nz], final values
    double *out_hi;              // contributions of this rank's last element plane to node plane n0 (sent to the next rank)
};

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_0() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ int wrapi(int v, int n)
{
    v %= n;
    return v < 0 ? v + n : v;
}

template <int H, int NSTR, int MODE>
__global__ void __launch_bounds__(SWEEP_THREADS, (MODE == SW_LINEAR) ? 2 : 1) k_sweep(const SweepParams p)
{
    extern __shared__ double smem[];
    constexpr int ND = 8 * H;
    double *ring = smem;                   // [2][H][NTILE]
    double *stg = smem + 2 * H * NTILE;    // [ND][NELT]
    double *hstg = stg + ND * NELT;        // [2][FANS_HIST_STAGE_SLOTS][SWEEP_THREADS] thread-private history staging (only if p.hstage)
    __shared__ double scratch[32 * (MODE == SW_STRAINSTRESS ? NSTR : 1)];

    const int tid = threadIdx.x;
    const int z0 = blockIdx.x * TZ, y0 = blockIdx.y * TY;
    const int xs = blockIdx.z * p.xchunk;
    const int xe = min(xs + p.xchunk, p.n0);
    const bool own = tid < TY * TZ;
    int ely, elz;  // element handled by this thread, tile coordinates in [-1,TY) x [-1,TZ)
    bool has_el = true;
    if (own) {
        ely = tid / TZ;
        elz = tid % TZ;
    } else {
        const int hh = tid - TY * TZ;
        if (hh < TZ + 1) {
            ely = -1;
            elz = hh - 1;
        } else if (hh < TZ + 1 + TY) {
            ely = hh - (TZ + 1);
            elz = -1;
        } else {
            ely = 0;
            elz = 0;
            has_el = false;
        }
    }
    const int ey = wrapi(y0 + ely, p.ny), ez = wrapi(z0 + elz, p.nz);  // global element y,z (periodic)
    const bool own_valid = own && (y0 + ely < p.ny) && (z0 + elz < p.nz);
    const int eidx = (ely + 1) * (TZ + 1) + (elz + 1);
    const int n00 = (ely + 1) * (TZ + 2) + (elz + 1);  // tile index of the element's node 0

    double beta = 0.0;
    if (MODE == SW_LINEAR && p.in2) beta = *p.beta;

    double carry[H];
#pragma unroll
    for (int c = 0; c < H; ++c) carry[c] = 0.0;
    double racc[(MODE == SW_STRAINSTRESS) ? NSTR : 1];
#pragma unroll
    for (int i = 0; i < ((MODE == SW_STRAINSTRESS) ? NSTR : 1); ++i) racc[i] = 0.0;

    // ---- plane loader: node plane xp (periodic in x) into ring slot `slot`
    auto load_plane = [&](int xp, int slot, bool owned_plane) {
        const int xg = wrapi(xp, p.n0);
        const bool hal = p.in_hi && xp >= p.n0;
        for (int i = tid; i < NTILE; i += SWEEP_THREADS) {
            const int ry = i / (TZ + 2), rz = i % (TZ + 2);
            const int y = wrapi(y0 - 1 + ry, p.ny), z = wrapi(z0 - 1 + rz, p.nz);
            const size_t g = ((size_t)xg * p.ny + y) * p.nz + z;
#pragma unroll
            for (int c = 0; c < H; ++c) {
                double v = hal ? p.in_hi[(size_t)c * p.ny * p.nz + (size_t)y * p.nz + z] : p.in[c * p.nloc + g];
                if (MODE == SW_LINEAR && p.in2 && !hal) {
                    v = p.in2[c * p.nloc + g] + beta * v;
                    if (owned_plane && ry >= 1 && ry <= TY && rz >= 1 && rz <= TZ && (y0 - 1 + ry) < p.ny && (z0 - 1 + rz) < p.nz)
                        p.in_out[c * p.nloc + g] = v;
                }
                ring[(slot * H + c) * NTILE + i] = v;
            }
        }
    };

    int lo = 0;  // ring slot of the lower node plane of the current element plane
    load_plane(xs - 1, 0, false);
    for (int x = xs - 1; x < xe; ++x) {  // element plane x uses node planes x (slot lo) and x+1 (slot lo^1)
        load_plane(x + 1, lo ^ 1, (x + 1) < xe);
        __syncthreads();
        double u0[H];
#pragma unroll
        for (int c = 0; c < H; ++c) u0[c] = 0.0;
        // strain/stress sweeps only evaluate owned elements (no assembly => no halo elements, no plane xs-1)
        // world_size > 1: element plane -1 belongs to the previous rank, which sends its contributions instead
        const bool skip_lo = p.in_hi && x < 0;
        const bool do_el = has_el && !skip_lo && (MODE != SW_STRAINSTRESS || (own_valid && x >= xs));
        if (skip_lo && has_el && MODE != SW_STRAINSTRESS) {
#pragma unroll
            for (int i = 0; i < ND; ++i) stg[i * NELT + eidx] = 0.0;
        }
        if (do_el) {
            // gather the 8 nodes: local node i = bx + 2 by + 4 bz  (include/solver.h:333-340)
            double ue[ND];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int slot = (i & 1) ? (lo ^ 1) : lo;
                const int ti = n00 + ((i >> 1) & 1) * (TZ + 2) + ((i >> 2) & 1);
#pragma unroll
                for (int c = 0; c < H; ++c) ue[H * i + c] = ring[(slot * H + c) * NTILE + ti];
            }
#pragma unroll
            for (int c = 0; c < H; ++c) u0[c] = ue[c];
            const int xg = wrapi(x, p.n0);
            const size_t e = ((size_t)xg * p.ny + ey) * p.nz + ez;
            const int ph = p.phidx[e];
            if (MODE == SW_LINEAR) {
                // ue <- ue - u(node 0)   (solver.h:250-254);  res_e = K_phase ue   (solverCG.h:98-103)
#pragma unroll
                for (int i = 1; i < 8; ++i)
#pragma unroll
                    for (int c = 0; c < H; ++c) ue[H * i + c] -= u0[c];
                const int kidx = p.phases[ph].k_index;
                const double *Kq = p.Kglob + (size_t)kidx * (ND * ND);
#pragma unroll 1
                for (int i = 0; i < ND; ++i) {
                    double a = 0.0;
#pragma unroll
                    for (int j = H; j < ND; ++j) a = fma(__ldg(&Kq[i * ND + j]), ue[j], a);
                    stg[i * NELT + eidx] = a;
                }
            } else {
                if (MODE == SW_RESIDUAL) {
#pragma unroll
                    for (int i = 1; i < 8; ++i)
#pragma unroll
                        for (int c = 0; c < H; ++c) ue[H * i + c] -= u0[c];
#pragma unroll
                    for (int c = 0; c < H; ++c) ue[c] = 0.0;
                }  // SW_STRAINSTRESS keeps the absolute ue (solver.h:507,723)
                const PhaseDev &pd = p.phases[ph];
                const size_t he = (pd.has_hist && p.hidx) ? (size_t)p.hidx[e] : 0;
                double res[ND];
#pragma unroll
                for (int i = 0; i < ND; ++i) res[i] = 0.0;
                double esum[NSTR], ssum[NSTR];
#pragma unroll
                for (int i = 0; i < NSTR; ++i) esum[i] = 0.0, ssum[i] = 0.0;
                // B-bar: centre value of the "volumetric" row (include/matmodel.h:113-140)
                double mc = 0.0, Qsum = 0.0;
                if (p.bbar && NSTR > 3) {
                    double t = 0.0;
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        if (NSTR == 6) {
                            t = fma(p.bg[(8 * 3 + 0) * 8 + a], ue[H * a + 0], t);
                            t = fma(p.bg[(8 * 3 + 1) * 8 + a], ue[H * a + (H > 1 ? 1 : 0)], t);
                            t = fma(p.bg[(8 * 3 + 2) * 8 + a], ue[H * a + (H > 2 ? 2 : 0)], t);
                        } else {
                            t = fma(p.bg[(8 * 3 + 0) * 8 + a], ue[H * a], t);
                            t = fma(p.bg[(8 * 3 + 1) * 8 + a], ue[H * a], t);
                            t = fma(p.bg[(8 * 3 + 2) * 8 + a], ue[H * a], t);
                        }
                    }
                    mc = t * (1.0 / 3.0);
                }
                const bool wr = own_valid && x >= xs;  // only the owner of an element updates its history / flags
                // History staging: a thread evaluates its 8 Gauss points one after the other and would pay one DRAM round trip per
                // point (a single 320-thread CTA per SM cannot hide it); instead the values of point g+1 travel global -> shared
                // with cp.async while point g is evaluated.  nT committed values (hist_t), current values of variables 6..12 for J2.
                const bool st_j2 = (pd.model == FANS_MAT_J2_LINEAR_ISO || pd.model == FANS_MAT_J2_NONLIN_ISO);
                const int nT = (p.hstage && MODE != SW_LINEAR && NSTR == 6) ? (st_j2 ? 13 : (pd.model == FANS_MAT_J2NEW_LINEAR_ISO ? 7 : 0)) : 0;
                double *hmine = hstg + tid;
                auto hist_issue = [&](int g, int buf) {
                    double *dst = hmine + (size_t)buf * FANS_HIST_STAGE_SLOTS * SWEEP_THREADS;
                    const size_t vs = (size_t)p.ngp * p.nh, go = (size_t)g * p.nh + he;   // variable v of Gauss point g: [v * vs + go]
                    const double *bt = p.hist_t + go;
#pragma unroll
                    for (int v = 0; v < 13; ++v)
                        if (v < nT) cp_async8(dst + v * SWEEP_THREADS, bt + v * vs);
                    cp_async_commit();
                };
                if (nT) hist_issue(0, 0);
#pragma unroll 1
                for (int g = 0; g < p.ngp; ++g) {
                    HistStage hs{nullptr, SWEEP_THREADS};
                    if (nT) {
                        if (g + 1 < p.ngp) {
                            hist_issue(g + 1, (g + 1) & 1);
                            cp_async_wait_1();
                        } else {
                            cp_async_wait_0();
                        }
                        hs.s = hmine + (size_t)(g & 1) * FANS_HIST_STAGE_SLOTS * SWEEP_THREADS;
                    }
                    const double *bg = p.bg + g * 24;
                    double Hm[H][3];
#pragma unroll
                    for (int c = 0; c < H; ++c)
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            double s = 0.0;
#pragma unroll
                            for (int a = 0; a < 8; ++a) s = fma(bg[j * 8 + a], ue[H * a + c], s);
                            Hm[c][j] = s;
                        }
                    double eps[NSTR], sig[NSTR];
                    strain_from_grad<H, NSTR>(Hm, eps);
                    if (p.bbar && NSTR > 3) {
                        const double m = (eps[0] + eps[1] + eps[2]) * (1.0 / 3.0);
                        eps[0] += mc - m;
                        eps[1] += mc - m;
                        eps[2] += mc - m;
                    }
#pragma unroll
                    for (int i = 0; i < NSTR; ++i) eps[i] += p.g0[i];
                    material_law<NSTR>(pd, eps, sig, p.hist, p.hist_t, p.pflag, p.nloc, p.nh, p.ngp, g, e, he, wr, p.fault, hs);
                    if (MODE == SW_STRAINSTRESS) {
#pragma unroll
                        for (int i = 0; i < NSTR; ++i) {
                            esum[i] += eps[i], ssum[i] += sig[i];
                            if (wr && p.eps_gp) p.eps_gp[((size_t)i * p.ngp + g) * p.nloc + e] = eps[i];
                            if (wr && p.sig_gp) p.sig_gp[((size_t)i * p.ngp + g) * p.nloc + e] = sig[i];
                        }
                    } else {
                        if (p.bbar && NSTR > 3) {
                            const double qv = (sig[0] + sig[1] + sig[2]) * (1.0 / 3.0);
                            sig[0] -= qv;
                            sig[1] -= qv;
                            sig[2] -= qv;
                            Qsum += qv;
                        }
                        double Tm[H][3];
                        stress_tensor<H, NSTR>(sig, Tm);
#pragma unroll
                        for (int a = 0; a < 8; ++a)
#pragma unroll
                            for (int c = 0; c < H; ++c) {
                                double s = res[H * a + c];
#pragma unroll
                                for (int j = 0; j < 3; ++j) s = fma(bg[j * 8 + a], Tm[c][j], s);
                                res[H * a + c] = s;
                            }
                    }
                }
                if (MODE == SW_RESIDUAL) {
                    if (p.bbar && NSTR > 3) {
                        double sq[NSTR];
#pragma unroll
                        for (int i = 0; i < NSTR; ++i) sq[i] = (i < 3) ? Qsum : 0.0;
                        double Tm[H][3];
                        stress_tensor<H, NSTR>(sq, Tm);
                        const double *bc = p.bg + 8 * 24;
#pragma unroll
                        for (int a = 0; a < 8; ++a)
#pragma unroll
                            for (int c = 0; c < H; ++c) {
                                double s = res[H * a + c];
#pragma unroll
                                for (int j = 0; j < 3; ++j) s = fma(bc[j * 8 + a], Tm[c][j], s);
                                res[H * a + c] = s;
                            }
                    }
#pragma unroll
                    for (int i = 0; i < ND; ++i) stg[i * NELT + eidx] = res[i] * p.vw;
                } else {  // SW_STRAINSTRESS: element averages, only for owned elements
                    if (wr) {
                        const double inv = 1.0 / (double)p.ngp;
#pragma unroll
                        for (int i = 0; i < NSTR; ++i) {
                            const double ev = esum[i] * inv, sv = ssum[i] * inv;
                            if (p.eps_out) p.eps_out[i * p.nloc + e] = ev;
                            if (p.sig_out) p.sig_out[i * p.nloc + e] = sv;
                            racc[i] += sv;
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (MODE != SW_STRAINSTRESS) {
            if (own) {
                // assemble: node (ly,lz) is local node (bx, by, bz) of element (ly-by, lz-bz) of plane x (bx=0) / x-1 (bx=1)
                double outv[H], nxt[H];
#pragma unroll
                for (int c = 0; c < H; ++c) outv[c] = carry[c], nxt[c] = 0.0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int by = b & 1, bz = b >> 1;
                    const int ee = eidx - by * (TZ + 1) - bz;
                    const int i0 = 2 * by + 4 * bz;
#pragma unroll
                    for (int c = 0; c < H; ++c) {
                        outv[c] += stg[(H * i0 + c) * NELT + ee];
                        nxt[c] += stg[(H * (i0 + 1) + c) * NELT + ee];
                    }
                }
#pragma unroll
                for (int c = 0; c < H; ++c) carry[c] = nxt[c];
                if (own_valid && x >= xs) {
                    const size_t g = ((size_t)x * p.ny + ey) * p.nz + ez;
#pragma unroll
                    for (int c = 0; c < H; ++c) {
                        p.out[c * p.nloc + g] = outv[c];
                        if (MODE == SW_LINEAR) racc[0] += outv[c] * u0[c];
                    }
                }
            }
        }
        lo ^= 1;
    }
    if (MODE != SW_STRAINSTRESS && p.out_hi && xe == p.n0 && own_valid) {
#pragma unroll
        for (int c = 0; c < H; ++c) p.out_hi[(size_t)c * p.ny * p.nz + (size_t)ey * p.nz + ez] = carry[c];
    }
    if (p.red_out) {
        if constexpr (MODE == SW_STRAINSTRESS) grid_reduce<NSTR, NSTR>(racc, scratch, p.part, p.ticket, p.red_out);
        else if constexpr (MODE == SW_LINEAR) grid_reduce<1, 1>(racc, scratch, p.part, p.ticket, p.red_out);
    }
}

// ------------------------------------------------------------------------------------------------
// Sum-factorised sweep for the 8-point elements (HEX8, BBAR), residual and strain/stress modes.
//
// The trilinear shape functions are tensor products, so  eps = B ue  and  res = B^T sigma  (include/matmodel.h:190-201) factor per
// axis: d/dx of the interpolant does not depend on the x coordinate of the Gauss point, etc.  Per element and displacement
// component the 8 Gauss-point gradients are 3 x 4 distinct values, obtained from nodal differences by two 2-term interpolations
// (weights N(0.5 -/+ sqrt(3)/6)); the transpose runs the same steps backwards on Gauss-point sums of the stress tensor.  That is
// ~400 FP64 operations per element for both directions instead of the 1152 FMAs of the dense 48x24 products.
// The Gauss points are visited in two halves (gz = 0, 1; a real loop, so the law code exists 4 times, not 8); what survives a half
// is 12 + 12 + 12 doubles (z-derivative values, their stress sums, the first half's nodal x/y forces).
// Node planes travel global -> shared with cp.async into a 3-slot ring, one plane ahead of the element plane being evaluated.
// Tiling, halo-element recomputation and the fixed-order nodal assembly are those of k_sweep above.
// ------------------------------------------------------------------------------------------------
// Tile of the sum-factorised kernel: 10 x 32 owned nodes + 43 halo elements = 363 of 384 threads.  Twelve warps are three per SM
// sub-partition, i.e. the same 168-register budget as ten, and the 16 KB register file of every sub-partition is used.
#define FY 10
#define FZ 32
#define FNTILE ((FY + 2) * (FZ + 2))
#define FNELT ((FY + 1) * (FZ + 1))
#define SF_THREADS 384
#define SF_NPH 8   // phase descriptors kept in shared memory (more phases: read from global memory)

template <int H, int NSTR, int MODE, bool BBAR>
__global__ void __launch_bounds__(SF_THREADS, 1) k_sweep_sf(const SweepParams p)
{
    extern __shared__ double smem[];
    constexpr int ND = 8 * H;
    constexpr int NR = (MODE == SW_STRAINSTRESS) ? NSTR : 1;
    double *ring = smem;                    // [3][H][FNTILE]
    double *stg0 = smem + 3 * H * FNTILE;   // [nstg][ND][FNELT]: element forces of plane x live in buffer x & (nstg - 1)
    const int nstg = p.stg2 ? 2 : 1;
    double *hstg = stg0 + (MODE == SW_STRAINSTRESS ? 0 : nstg * ND * FNELT);   // [2][FANS_HIST_STAGE_SLOTS][SF_THREADS] (only if p.hstage)
    __shared__ double scratch[32 * NR];
    __shared__ PhaseDev sph[SF_NPH];

    const int tid = threadIdx.x;
    const int z0 = blockIdx.x * FZ, y0 = blockIdx.y * FY;
    const int xs = blockIdx.z * p.xchunk;
    const int xe = min(xs + p.xchunk, p.n0);
    const bool own = tid < FY * FZ;
    int ely, elz;
    bool has_el = true;
    if (own) {
        ely = tid / FZ;
        elz = tid % FZ;
    } else {
        const int hh = tid - FY * FZ;
        if (hh < FZ + 1) {
            ely = -1;
            elz = hh - 1;
        } else if (hh < FZ + 1 + FY) {
            ely = hh - (FZ + 1);
            elz = -1;
        } else {
            ely = 0;
            elz = 0;
            has_el = false;
        }
    }
    const int ey = wrapi(y0 + ely, p.ny), ez = wrapi(z0 + elz, p.nz);
    const bool own_valid = own && (y0 + ely < p.ny) && (z0 + elz < p.nz);
    const int eidx = (ely + 1) * (FZ + 1) + (elz + 1);
    const int n00 = (ely + 1) * (FZ + 2) + (elz + 1);
    for (int i = tid; i < (int)(sizeof(PhaseDev) / sizeof(double)) * min(p.n_phases, SF_NPH); i += SF_THREADS)
        reinterpret_cast<double *>(sph)[i] = reinterpret_cast<const double *>(p.phases)[i];

    double carry[H];
#pragma unroll
    for (int c = 0; c < H; ++c) carry[c] = 0.0;
    double racc[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) racc[i] = 0.0;

    // plane loader: node plane xp (periodic in x / upper slab halo) -> ring slot, asynchronous; one commit group per call
    auto issue_plane = [&](int xp, int slot) {
        if (xp <= xe) {
            const int xg = wrapi(xp, p.n0);
            const bool hal = p.in_hi && xp >= p.n0;
            for (int i = tid; i < FNTILE; i += SF_THREADS) {
                const int ry = i / (FZ + 2), rz = i % (FZ + 2);
                const int y = wrapi(y0 - 1 + ry, p.ny), z = wrapi(z0 - 1 + rz, p.nz);
                const size_t g = ((size_t)xg * p.ny + y) * p.nz + z;
#pragma unroll
                for (int c = 0; c < H; ++c)
                    cp_async8(&ring[(slot * H + c) * FNTILE + i],
                              hal ? &p.in_hi[(size_t)c * p.ny * p.nz + (size_t)y * p.nz + z] : &p.in[c * p.nloc + g]);
            }
        }
        cp_async_commit();
    };

    // interpolation weights at the Gauss coordinates 0.5 -/+ sqrt(3)/6: N_0(m) = N_1(p) = wp, N_1(m) = N_0(p) = wm; the derivative
    // factors 1 / l_e are folded into a second copy of the weights
    const double wm = p.gm, wp = p.gp;
    const double wmx = wm * p.il[0], wpx = wp * p.il[0], wmy = wm * p.il[1], wpy = wp * p.il[1], wmz = wm * p.il[2], wpz = wp * p.il[2];
    // transposed direction: the quadrature weight v_e / n_gp (matmodel.h:199) rides on the same factors
    const double vmx = wmx * p.vw, vpx = wpx * p.vw, vmy = wmy * p.vw, vpy = wpy * p.vw, vmz = wmz * p.vw, vpz = wpz * p.vw;

    // nodal assembly of element plane xq (forces staged in buffer `stq`): node (ly,lz) is local node (bx,by,bz) of element
    // (ly-by, lz-bz) of plane xq (bx = 0) / xq-1 (bx = 1, carried in registers from the previous call); fixed order
    auto assemble = [&](int xq, const double *stq) {
        if (!own) return;
        double outv[H], nxt[H];
#pragma unroll
        for (int c = 0; c < H; ++c) outv[c] = carry[c], nxt[c] = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int by = b & 1, bz = b >> 1;
            const int ee = eidx - by * (FZ + 1) - bz;
            const int i0 = 2 * by + 4 * bz;
#pragma unroll
            for (int c = 0; c < H; ++c) {
                outv[c] += stq[(H * i0 + c) * FNELT + ee];
                nxt[c] += stq[(H * (i0 + 1) + c) * FNELT + ee];
            }
        }
#pragma unroll
        for (int c = 0; c < H; ++c) carry[c] = nxt[c];
        if (own_valid && xq >= xs) {
            const size_t g = ((size_t)xq * p.ny + ey) * p.nz + ez;
#pragma unroll
            for (int c = 0; c < H; ++c) p.out[c * p.nloc + g] = outv[c];
        }
    };

    // One barrier per element plane when the staging tile is double-buffered (p.stg2): the barrier that publishes node plane x+1
    // also publishes the element forces of plane x-1, whose assembly is deferred to the start of iteration x.
    int sa = 0, sb = 1, sc = 2;  // ring slots of node planes x, x+1, x+2
    issue_plane(xs - 1, sa);
    issue_plane(xs, sb);
    for (int x = xs - 1; x < xe; ++x) {
        cp_async_wait_0();       // plane x+1 has landed (this thread's copies) ...
        __syncthreads();         // ... and everybody else's; everybody has left element plane x-1: its forces are complete and
                                 // the ring slot of node plane x-1 (sc) is free
        issue_plane(x + 2, sc);
        double *stg = stg0 + (size_t)((x - (xs - 1)) & (nstg - 1)) * ND * FNELT;
        if (MODE != SW_STRAINSTRESS && p.stg2 && x > xs - 1) assemble(x - 1, stg0 + (size_t)((x - 1 - (xs - 1)) & 1) * ND * FNELT);
        const bool skip_lo = p.in_hi && x < 0;
        const bool do_el = has_el && !skip_lo && (MODE != SW_STRAINSTRESS || (own_valid && x >= xs));
        if (skip_lo && has_el && MODE != SW_STRAINSTRESS) {
#pragma unroll
            for (int i = 0; i < ND; ++i) stg[i * FNELT + eidx] = 0.0;
        }
        if (do_el) {
            const double *rA = ring + (size_t)sa * H * FNTILE + n00, *rB = ring + (size_t)sb * H * FNTILE + n00;
            // nodal value of component c at local node (bx, by, bz): plane bx, tile offset by*(FZ+2) + bz  (include/solver.h:333-340)
#define UN(c, bx, by, bz) (((bx) ? rB : rA)[(c) * FNTILE + (by) * (FZ + 2) + (bz)])
            const int xg = wrapi(x, p.n0);
            const size_t e = ((size_t)xg * p.ny + ey) * p.nz + ez;
            const int ph = p.phidx[e];
            const PhaseDev &pd = (ph < SF_NPH) ? sph[ph] : p.phases[ph];
            const size_t he = (pd.has_hist && p.hidx) ? (size_t)p.hidx[e] : 0;
            const bool wr = own_valid && x >= xs;

            // the whole element for ONE material law known at compile time: no model branch inside the Gauss-point loop
            auto element = [&](auto law_tag) {
                constexpr int LAW = decltype(law_tag)::value;
                if constexpr (MODE == SW_STRAINSTRESS && LAW == FANS_MAT_LINEAR) {
                    if (p.one_point) {
                        // linear law on a trilinear element: the Gauss-point averages of strain and stress (matmodel.h:202-225) are
                        // exactly the values at the element centre (symmetric points; B-bar only moves the volumetric part, whose
                        // average is the centre value by construction) — one law evaluation instead of eight
                        double Hm[H][3];
#pragma unroll
                        for (int c = 0; c < H; ++c) {
                            const double u000 = UN(c, 0, 0, 0), u100 = UN(c, 1, 0, 0), u010 = UN(c, 0, 1, 0), u110 = UN(c, 1, 1, 0);
                            const double u001 = UN(c, 0, 0, 1), u101 = UN(c, 1, 0, 1), u011 = UN(c, 0, 1, 1), u111 = UN(c, 1, 1, 1);
                            Hm[c][0] = 0.25 * p.il[0] * ((u100 - u000) + (u110 - u010) + (u101 - u001) + (u111 - u011));
                            Hm[c][1] = 0.25 * p.il[1] * ((u010 - u000) + (u110 - u100) + (u011 - u001) + (u111 - u101));
                            Hm[c][2] = 0.25 * p.il[2] * ((u001 - u000) + (u101 - u100) + (u011 - u010) + (u111 - u110));
                        }
                        double eps[NSTR], sig[NSTR];
                        strain_from_grad<H, NSTR>(Hm, eps);
#pragma unroll
                        for (int i = 0; i < NSTR; ++i) eps[i] += p.g0[i];
                        material_law<NSTR, LAW>(pd, eps, sig, p.hist, p.hist_t, p.pflag, p.nloc, p.nh, 8, 0, e, he, false, p.fault);
                        if (wr) {
#pragma unroll
                            for (int i = 0; i < NSTR; ++i) {
                                if (p.eps_out) p.eps_out[i * p.nloc + e] = eps[i];
                                if (p.sig_out) p.sig_out[i * p.nloc + e] = sig[i];
                                racc[i] += sig[i];
                            }
                        }
                        return;
                    }
                }
                // ---- z derivative at (gx, gy), the same for both gz: GZ[c][gx][gy]
                double GZ[H][2][2];
#pragma unroll
                for (int c = 0; c < H; ++c) {
                    const double d00 = UN(c, 0, 0, 1) - UN(c, 0, 0, 0), d01 = UN(c, 0, 1, 1) - UN(c, 0, 1, 0);
                    const double d10 = UN(c, 1, 0, 1) - UN(c, 1, 0, 0), d11 = UN(c, 1, 1, 1) - UN(c, 1, 1, 0);
                    const double e00 = wpz * d00 + wmz * d10, e01 = wpz * d01 + wmz * d11;   // gx = 0, by = 0 / 1
                    const double e10 = wmz * d00 + wpz * d10, e11 = wmz * d01 + wpz * d11;   // gx = 1
                    GZ[c][0][0] = wp * e00 + wm * e01;
                    GZ[c][0][1] = wm * e00 + wp * e01;
                    GZ[c][1][0] = wp * e10 + wm * e11;
                    GZ[c][1][1] = wm * e10 + wp * e11;
                }
                // B-bar: centre value of the volumetric row (include/matmodel.h:113-140), as in k_sweep
                double mc = 0.0, Qsum = 0.0;
                if (BBAR && NSTR > 3) {
                    double t = 0.0;
#pragma unroll
                    for (int a = 0; a < 8; ++a) {
                        const int bx = a & 1, by = (a >> 1) & 1, bz = (a >> 2) & 1;
                        if (NSTR == 6) {
                            t = fma(p.bg[(8 * 3 + 0) * 8 + a], UN(0, bx, by, bz), t);
                            t = fma(p.bg[(8 * 3 + 1) * 8 + a], UN((H > 1 ? 1 : 0), bx, by, bz), t);
                            t = fma(p.bg[(8 * 3 + 2) * 8 + a], UN((H > 2 ? 2 : 0), bx, by, bz), t);
                        } else {
                            t = fma(p.bg[(8 * 3 + 0) * 8 + a], UN(0, bx, by, bz), t);
                            t = fma(p.bg[(8 * 3 + 1) * 8 + a], UN(0, bx, by, bz), t);
                            t = fma(p.bg[(8 * 3 + 2) * 8 + a], UN(0, bx, by, bz), t);
                        }
                    }
                    mc = t * (1.0 / 3.0);
                }
                // history staging (see k_sweep): the values of Gauss point g+1 travel global -> shared while g is evaluated
                constexpr bool st_j2 = (LAW == FANS_MAT_J2_LINEAR_ISO || LAW == FANS_MAT_J2_NONLIN_ISO);
                constexpr int nTc = (NSTR == 6) ? (st_j2 ? 13 : (LAW == FANS_MAT_J2NEW_LINEAR_ISO ? 7 : 0)) : 0;
                const int nT = p.hstage ? nTc : 0;
                double *hmine = hstg + tid;
                auto hist_issue = [&](int g, int buf) {
                    double *dst = hmine + (size_t)buf * FANS_HIST_STAGE_SLOTS * SF_THREADS;
                    const size_t vs = (size_t)8 * p.nh, go = (size_t)g * p.nh + he;
                    const double *bt = p.hist_t + go;
#pragma unroll
                    for (int v = 0; v < nTc; ++v) cp_async8(dst + v * SF_THREADS, bt + v * vs);
                    cp_async_commit();
                };
                if (nT) hist_issue(0, 0);

                double TZs[H][2][2];   // sum over gz of T[c][z] at (gx, gy)
                double esum[NSTR], ssum[NSTR];
#pragma unroll
                for (int i = 0; i < NSTR; ++i) esum[i] = 0.0, ssum[i] = 0.0;
#pragma unroll
                for (int c = 0; c < H; ++c) TZs[c][0][0] = TZs[c][0][1] = TZs[c][1][0] = TZs[c][1][1] = 0.0;

#pragma unroll 1
                for (int gzi = 0; gzi < 2; ++gzi) {
                    const double wz0 = gzi ? wm : wp, wz1 = gzi ? wp : wm;   // N_0(gz), N_1(gz)
                    double GX[H][2], GY[H][2];   // d/dx at gy, d/dy at gx (this gz)
#pragma unroll
                    for (int c = 0; c < H; ++c) {
                        const double a00 = wz0 * UN(c, 0, 0, 0) + wz1 * UN(c, 0, 0, 1), a01 = wz0 * UN(c, 0, 1, 0) + wz1 * UN(c, 0, 1, 1);
                        const double a10 = wz0 * UN(c, 1, 0, 0) + wz1 * UN(c, 1, 0, 1), a11 = wz0 * UN(c, 1, 1, 0) + wz1 * UN(c, 1, 1, 1);
                        const double dx0 = a10 - a00, dx1 = a11 - a01;   // by = 0 / 1
                        const double dy0 = a01 - a00, dy1 = a11 - a10;   // bx = 0 / 1
                        GX[c][0] = wpx * dx0 + wmx * dx1;
                        GX[c][1] = wmx * dx0 + wpx * dx1;
                        GY[c][0] = wpy * dy0 + wmy * dy1;
                        GY[c][1] = wmy * dy0 + wpy * dy1;
                    }
                    double TX[H][2], TY_[H][2];   // sums over gx of T[c][x] at gy ; over gy of T[c][y] at gx
#pragma unroll
                    for (int c = 0; c < H; ++c) TX[c][0] = TX[c][1] = TY_[c][0] = TY_[c][1] = 0.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int gx = k & 1, gy = k >> 1;
                        const int g = 4 * gzi + k;   // Gauss point number gx + 2 gy + 4 gz (include/matmodel.h:109-111)
                        HistStage hs{nullptr, SF_THREADS};
                        if (nT) {
                            if (g + 1 < 8) {
                                hist_issue(g + 1, (g + 1) & 1);
                                cp_async_wait_1();
                            } else {
                                cp_async_wait_0();
                            }
                            hs.s = hmine + (size_t)(g & 1) * FANS_HIST_STAGE_SLOTS * SF_THREADS;
                        }
                        double Hm[H][3];
#pragma unroll
                        for (int c = 0; c < H; ++c) Hm[c][0] = GX[c][gy], Hm[c][1] = GY[c][gx], Hm[c][2] = GZ[c][gx][gy];
                        double eps[NSTR], sig[NSTR];
                        strain_from_grad<H, NSTR>(Hm, eps);
                        if (BBAR && NSTR > 3) {
                            const double m = (eps[0] + eps[1] + eps[2]) * (1.0 / 3.0);
                            eps[0] += mc - m;
                            eps[1] += mc - m;
                            eps[2] += mc - m;
                        }
#pragma unroll
                        for (int i = 0; i < NSTR; ++i) eps[i] += p.g0[i];
                        material_law<NSTR, LAW>(pd, eps, sig, p.hist, p.hist_t, p.pflag, p.nloc, p.nh, 8, g, e, he, wr, p.fault, hs);
                        if (MODE == SW_STRAINSTRESS) {
#pragma unroll
                            for (int i = 0; i < NSTR; ++i) {
                                esum[i] += eps[i], ssum[i] += sig[i];
                                if (wr && p.eps_gp) p.eps_gp[((size_t)i * 8 + g) * p.nloc + e] = eps[i];
                                if (wr && p.sig_gp) p.sig_gp[((size_t)i * 8 + g) * p.nloc + e] = sig[i];
                            }
                        } else {
                            if (BBAR && NSTR > 3) {
                                const double qv = (sig[0] + sig[1] + sig[2]) * (1.0 / 3.0);
                                sig[0] -= qv;
                                sig[1] -= qv;
                                sig[2] -= qv;
                                Qsum += qv;
                            }
                            double Tm[H][3];
                            stress_tensor<H, NSTR>(sig, Tm);
#pragma unroll
                            for (int c = 0; c < H; ++c) {
                                TX[c][gy] += Tm[c][0];
                                TY_[c][gx] += Tm[c][1];
                                TZs[c][gx][gy] += Tm[c][2];
                            }
                        }
                    }
                    if (MODE != SW_STRAINSTRESS) {
                        // transpose of the x / y interpolations of this half: Q[c][bx][by] = -/+ RX[c][by] -/+ RY[c][bx], then the z
                        // interpolation N_bz(gz) Q to the two node layers.  The first half parks its result in the element's own
                        // column of the staging tile (thread-private until the barrier), the second half adds to it.
#pragma unroll
                        for (int c = 0; c < H; ++c) {
                            const double rx0 = vpx * TX[c][0] + vmx * TX[c][1], rx1 = vmx * TX[c][0] + vpx * TX[c][1];       // by = 0 / 1
                            const double ry0 = vpy * TY_[c][0] + vmy * TY_[c][1], ry1 = vmy * TY_[c][0] + vpy * TY_[c][1];   // bx = 0 / 1
                            const double q[2][2] = {{-rx0 - ry0, -rx1 + ry0}, {rx0 - ry1, rx1 + ry1}};
#pragma unroll
                            for (int bx = 0; bx < 2; ++bx)
#pragma unroll
                                for (int by = 0; by < 2; ++by) {
                                    double *s0 = &stg[(H * (bx + 2 * by) + c) * FNELT + eidx], *s1 = &stg[(H * (bx + 2 * by + 4) + c) * FNELT + eidx];
                                    if (gzi == 0) {
                                        *s0 = wz0 * q[bx][by];
                                        *s1 = wz1 * q[bx][by];
                                    } else {
                                        *s0 = fma(wz0, q[bx][by], *s0);
                                        *s1 = fma(wz1, q[bx][by], *s1);
                                    }
                                }
                        }
                    }
                }
                if (MODE == SW_RESIDUAL) {
                    // z part: RZ[c][bx][by] = sum_gx,gy N_bx(gx) N_by(gy) TZs / l_z ; node layer bz = 0 gets -RZ, bz = 1 +RZ
#pragma unroll
                    for (int c = 0; c < H; ++c) {
                        const double f00 = vpz * TZs[c][0][0] + vmz * TZs[c][1][0], f01 = vpz * TZs[c][0][1] + vmz * TZs[c][1][1];  // bx = 0, gy = 0 / 1
                        const double f10 = vmz * TZs[c][0][0] + vpz * TZs[c][1][0], f11 = vmz * TZs[c][0][1] + vpz * TZs[c][1][1];  // bx = 1
                        const double rz[2][2] = {{wp * f00 + wm * f01, wm * f00 + wp * f01}, {wp * f10 + wm * f11, wm * f10 + wp * f11}};
#pragma unroll
                        for (int bx = 0; bx < 2; ++bx)
#pragma unroll
                            for (int by = 0; by < 2; ++by) {
                                double *s0 = &stg[(H * (bx + 2 * by) + c) * FNELT + eidx], *s1 = &stg[(H * (bx + 2 * by + 4) + c) * FNELT + eidx];
                                *s0 -= rz[bx][by];
                                *s1 += rz[bx][by];
                            }
                    }
                    if (BBAR && NSTR > 3) {
                        double sq[NSTR];
#pragma unroll
                        for (int i = 0; i < NSTR; ++i) sq[i] = (i < 3) ? Qsum : 0.0;
                        double Tm[H][3];
                        stress_tensor<H, NSTR>(sq, Tm);
                        const double *bc = p.bg + 8 * 24;
#pragma unroll
                        for (int a = 0; a < 8; ++a)
#pragma unroll
                            for (int c = 0; c < H; ++c) {
                                double sacc = 0.0;
#pragma unroll
                                for (int j = 0; j < 3; ++j) sacc = fma(bc[j * 8 + a], Tm[c][j], sacc);
                                stg[(H * a + c) * FNELT + eidx] += sacc * p.vw;
                            }
                    }
                } else if (wr) {  // SW_STRAINSTRESS: element averages of the owned elements
#pragma unroll
                    for (int i = 0; i < NSTR; ++i) {
                        const double ev = esum[i] * 0.125, sv = ssum[i] * 0.125;
                        if (p.eps_out) p.eps_out[i * p.nloc + e] = ev;
                        if (p.sig_out) p.sig_out[i * p.nloc + e] = sv;
                        racc[i] += sv;
                    }
                }
            };
#ifdef SF_ONLY_LAW
            if constexpr (NSTR == 6 || NSTR == 9) { element(std::integral_constant<int, SF_ONLY_LAW>{}); } else
#endif
            if constexpr (NSTR == 6) {
                switch (pd.model) {
                case FANS_MAT_LINEAR: element(std::integral_constant<int, FANS_MAT_LINEAR>{}); break;
                case FANS_MAT_PSEUDOPLASTIC_LINEAR: element(std::integral_constant<int, FANS_MAT_PSEUDOPLASTIC_LINEAR>{}); break;
                case FANS_MAT_PSEUDOPLASTIC_NONLIN: element(std::integral_constant<int, FANS_MAT_PSEUDOPLASTIC_NONLIN>{}); break;
                case FANS_MAT_J2_LINEAR_ISO: element(std::integral_constant<int, FANS_MAT_J2_LINEAR_ISO>{}); break;
                case FANS_MAT_J2_NONLIN_ISO: element(std::integral_constant<int, FANS_MAT_J2_NONLIN_ISO>{}); break;
                default: element(std::integral_constant<int, FANS_MAT_J2NEW_LINEAR_ISO>{}); break;
                }
            } else if constexpr (NSTR == 9) {
                if (pd.model == FANS_MAT_SVK) element(std::integral_constant<int, FANS_MAT_SVK>{});
                else element(std::integral_constant<int, FANS_MAT_NEOHOOKE>{});
            } else {
                element(std::integral_constant<int, FANS_MAT_LINEAR>{});
            }
#undef UN
        }
        if (MODE != SW_STRAINSTRESS && !p.stg2) {   // single staging buffer: assemble right away, second barrier per plane
            __syncthreads();
            assemble(x, stg);
        }
        const int t = sa;
        sa = sb, sb = sc, sc = t;
    }
    if (MODE != SW_STRAINSTRESS && p.stg2) {
        __syncthreads();
        assemble(xe - 1, stg0 + (size_t)((xe - 1 - (xs - 1)) & 1) * ND * FNELT);
    }
    cp_async_wait_0();
    if (MODE != SW_STRAINSTRESS && p.out_hi && xe == p.n0 && own_valid) {
#pragma unroll
        for (int c = 0; c < H; ++c) p.out_hi[(size_t)c * p.ny * p.nz + (size_t)ey * p.nz + ez] = carry[c];
    }
    if (p.red_out) {
        if constexpr (MODE == SW_STRAINSTRESS) grid_reduce<NSTR, NSTR>(racc, scratch, p.part, p.ticket, p.red_out);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static uint64_t g_stamp_counter = 0;

uint64_t sweep_new_stamp() { return ++g_stamp_counter; }   // identifies a context's material tables (stencil.cu rebuilds on change)

template <int H, int NSTR, int MODE>
static int launch_sweep(fans_ctx *ctx, const SweepParams &p, dim3 grid, size_t smem, bool sf)
{
    prof_begin(ctx, MODE == SW_LINEAR ? PC_SWEEP_LINEAR : (MODE == SW_RESIDUAL ? PC_SWEEP_RESIDUAL : PC_SWEEP_STRAINSTRESS));
    if constexpr (MODE != SW_LINEAR) {
        if (sf) {  // 8-point elements: sum-factorised gradient / divergence
            if (p.bbar) {
                CUDA_TRY(ctx, cudaFuncSetAttribute(k_sweep_sf<H, NSTR, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_sweep_sf<H, NSTR, MODE, true><<<grid, SF_THREADS, smem, ctx->st>>>(p);
            } else {
                CUDA_TRY(ctx, cudaFuncSetAttribute(k_sweep_sf<H, NSTR, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_sweep_sf<H, NSTR, MODE, false><<<grid, SF_THREADS, smem, ctx->st>>>(p);
            }
            prof_end(ctx);
            ctx->launches++;
            CUDA_TRY(ctx, cudaGetLastError());
            return FANS_OK;
        }
    }
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_sweep<H, NSTR, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_sweep<H, NSTR, MODE><<<grid, SWEEP_THREADS, smem, ctx->st>>>(p);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

// mode: SW_*; in/out device SoA fields. s_in/d_new/beta only for the fused CG direction update.
int sweep_run(fans_ctx *ctx, int mode, const double *in, double *out, const double *s_in, double *d_new,
              const double *beta_dev, double *red_out, double *eps_out, double *sig_out, double *eps_gp, double *sig_gp)
{
    if (!ctx->ms_ready || !ctx->materials_ready) {
        fans_set_error(ctx, FANS_ERR_STATE, "microstructure and materials must be set before an element sweep");
        return FANS_ERR_STATE;
    }
    SweepParams p;
    memset(&p, 0, sizeof(p));
    double *red_final = nullptr;
    if (ctx->P > 1) {
        // upper halo of the (updated) input; the K.d dot product is formed after the boundary contributions have been added
        FANS_CHECK(halo_exchange_up(ctx, in, mode == SW_LINEAR ? s_in : nullptr, beta_dev));
        p.in_hi = ctx->halo_hi;
        p.out_hi = ctx->halo_send_hi;
        if (mode == SW_LINEAR) red_final = red_out, red_out = nullptr;
    }
    p.n0 = ctx->n0;
    p.ny = ctx->ny;
    p.nz = ctx->nz;
    p.nloc = ctx->nloc;
    p.in = in;
    p.in2 = s_in;
    p.in_out = d_new;
    p.beta = beta_dev;
    p.out = out;
    p.phidx = ctx->phidx;
    p.phases = ctx->d_phase;
    p.n_phases = ctx->n_phases;
    p.Kglob = ctx->d_K;
    p.n_k = ctx->n_k;
    for (int i = 0; i < 9; ++i) p.g0[i] = ctx->g0[i];
    for (int i = 0; i < 9 * 24; ++i) p.bg[i] = ctx->Bgp[i];
    p.vw = ctx->ve / ctx->ngp;
    p.ngp = ctx->ngp;
    p.bbar = (ctx->fe == FANS_FE_BBAR);
    if (ctx->any_history && mode != SW_LINEAR) FANS_CHECK(history_prepare(ctx));
    p.hist = ctx->hist;
    p.hist_t = ctx->hist_t;
    p.hidx = ctx->hidx;
    p.nh = ctx->nh;
    p.pflag = ctx->pflag;
    p.gm = 0.5 - sqrt(3.0) / 6.0;   // include/matmodel.h:109-111
    p.gp = 0.5 + sqrt(3.0) / 6.0;
    for (int d = 0; d < 3; ++d) p.il[d] = 1.0 / ctx->le[d];
    p.fault = ctx->d_flag;
    p.part = ctx->d_part;
    p.ticket = ctx->d_ticket;
    p.red_out = red_out;
    p.eps_out = eps_out;
    p.sig_out = sig_out;
    p.eps_gp = eps_gp;
    p.sig_gp = sig_gp;
    // x chunking: enough CTAs to fill 148 SMs a few times over, at most ~6% redundant plane loads
    // the 8-point elements take the sum-factorised kernel (FANS_SWEEP_DENSE=1: the dense B products of k_sweep, for A/B runs)
    bool sf = mode != SW_LINEAR && ctx->ngp == 8 && !(getenv("FANS_SWEEP_DENSE") && atoi(getenv("FANS_SWEEP_DENSE")) != 0);
    // Element averages of a LINEAR law on an 8-point element are exactly the values at the element centre: the Gauss points are
    // symmetric about it, the gradient is trilinear, the stress is linear in it (and B-bar only moves the volumetric part, whose
    // average is the centre value by construction).  The strain/stress sweep of an all-linear problem (homogenized stress,
    // postprocess averages) therefore evaluates one point per element — 1/8 of the law and gradient work — unless the
    // Gauss-point fields themselves are asked for.  (FANS_SS_FULL=1 keeps the 8-point evaluation, for A/B runs and tests.)
    p.one_point = (sf && mode == SW_STRAINSTRESS && ctx->all_linear && !eps_gp && !sig_gp && !(getenv("FANS_SS_FULL") && atoi(getenv("FANS_SS_FULL")))) ? 1 : 0;
    const int ty = sf ? FY : TY, tz = sf ? FZ : TZ;
    const int gy = (ctx->ny + ty - 1) / ty, gz = (ctx->nz + tz - 1) / tz;
    const int xchunk = pick_xchunk(ctx->n0, (long)gy * gz, (long)FANS_SMS * ((!sf && mode == SW_LINEAR) ? 2 : 1), 1);
    p.xchunk = xchunk;
    dim3 grid(gz, gy, (ctx->n0 + xchunk - 1) / xchunk);
    p.hstage = (ctx->any_history && mode != SW_LINEAR && !(getenv("FANS_HIST_STAGE") && atoi(getenv("FANS_HIST_STAGE")) == 0)) ? 1 : 0;
    // sum-factorised kernel: double-buffered staging tile (one barrier per plane) whenever it fits beside the history staging
    const size_t sf_base = sizeof(double) * (3 * ctx->h * FNTILE + (p.hstage ? 2 * FANS_HIST_STAGE_SLOTS * SF_THREADS : 0));
    const size_t sf_stg = (mode == SW_STRAINSTRESS) ? 0 : sizeof(double) * 8 * ctx->h * FNELT;
    p.stg2 = (sf && sf_base + 2 * sf_stg <= 200 * 1024 && !(getenv("FANS_SWEEP_STG1") && atoi(getenv("FANS_SWEEP_STG1")))) ? 1 : 0;
    const size_t smem = sf ? sf_base + (p.stg2 ? 2 : 1) * sf_stg
                           : sizeof(double) * (2 * ctx->h * NTILE + 8 * ctx->h * NELT + (p.hstage ? 2 * FANS_HIST_STAGE_SLOTS * SWEEP_THREADS : 0));
    int rc = FANS_ERR_ARG;
#define SW_DISPATCH(H_, N_)                                                                           \
    do {                                                                                              \
        if (mode == SW_LINEAR) rc = launch_sweep<H_, N_, SW_LINEAR>(ctx, p, grid, smem, sf);          \
        else if (mode == SW_RESIDUAL) rc = launch_sweep<H_, N_, SW_RESIDUAL>(ctx, p, grid, smem, sf); \
        else rc = launch_sweep<H_, N_, SW_STRAINSTRESS>(ctx, p, grid, smem, sf);                      \
    } while (0)
    if (ctx->h == 1 && ctx->nstr == 3) SW_DISPATCH(1, 3);
    else if (ctx->h == 3 && ctx->nstr == 6) SW_DISPATCH(3, 6);
    else if (ctx->h == 3 && ctx->nstr == 9) SW_DISPATCH(3, 9);
    else fans_set_error(ctx, FANS_ERR_ARG, "unsupported (howmany, n_str) combination");
#undef SW_DISPATCH
    if (rc != FANS_OK) return rc;
    if (ctx->P > 1 && mode != SW_STRAINSTRESS) {
        FANS_CHECK(halo_add_down(ctx, out));
        if (red_final) {  // <d_new, K d_new> over the completed field, summed over the slabs
            FANS_CHECK(vec_reduce4(ctx, d_new ? d_new : in, out, ctx->d_red + S_GEN));
            CUDA_TRY(ctx, cudaMemcpyAsync(red_final, ctx->d_red + S_GEN + 2, sizeof(double), cudaMemcpyDeviceToDevice, ctx->st));
        }
    }
    return FANS_OK;
}
