// vecops.cu — CG / fixed-point vector updates, dot products and error norms as fused HBM passes.
// Replaces the Eigen array expressions of SolverCG::internalSolve / LineSearchSecant (include/solverCG.h:86-107,
// 119-160), SolverFP::internalSolve (include/solverFP.h:46), Solver::compute_error (include/solver.h:414-431) and
// Solver::extrapolateDisplacement (include/solver.h:302-311).  All fields are flat SoA arrays of h*nloc doubles.
#include "internal.h"
#include <algorithm>

#define VEC_THREADS 256

static inline unsigned vec_grid(size_t n2)
{
    size_t nb = (n2 + VEC_THREADS - 1) / VEC_THREADS;
    const size_t cap = (size_t)FANS_SMS * 16;  // 16 resident CTAs of 256 threads per SM would exceed the SM; 8 fit, x2 waves
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    return (unsigned)nb;
}

// r -= alpha*Kd ; u -= alpha*d ; norms of the new r ; deltamid = <r_new, s>      (solverCG.h:105-107, :86, solver.h:419-425)
// alpha = delta / <d,Kd> is formed from the device scalars (no host round trip).
__global__ void __launch_bounds__(VEC_THREADS) k_cg_update(double2 *__restrict__ r, const double2 *__restrict__ kd,
                                                            double2 *__restrict__ u, const double2 *__restrict__ d,
                                                            const double2 *__restrict__ s, size_t n2, double *S,
                                                            double *part, unsigned int *ticket)
{
    // batched solves: blockIdx.y = lane (fields n2 double2 apart, own scalar block / partial sums / ticket); a lane that has
    // converged (S_FREEZE) is left exactly as it is
    __shared__ double scratch[4 * 32];
    {
        const size_t lo = (size_t)blockIdx.y * n2;
        r += lo, kd += lo, u += lo, d += lo, s += lo;
        S += (size_t)blockIdx.y * S_COUNT, part += (size_t)blockIdx.y * 4 * gridDim.x, ticket += blockIdx.y;
    }
    if (S[S_FREEZE] != 0.0) return;
    const double alpha = S[S_DELTA] / S[S_DKD];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};  // L1, L2^2, <r,s>, Linf
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        double2 rv = r[i];
        const double2 kv = kd[i], dv = d[i], sv = s[i];
        double2 uv = u[i];
        rv.x -= alpha * kv.x;
        rv.y -= alpha * kv.y;
        uv.x -= alpha * dv.x;
        uv.y -= alpha * dv.y;
        r[i] = rv;
        u[i] = uv;
        acc[0] += fabs(rv.x) + fabs(rv.y);
        acc[1] += rv.x * rv.x + rv.y * rv.y;
        acc[2] += rv.x * sv.x + rv.y * sv.y;
        acc[3] = fmax(acc[3], fmax(fabs(rv.x), fabs(rv.y)));
    }
    grid_reduce_part<4, 3>(acc, scratch, part, ticket, S + S_L1, false, gridDim.x, blockIdx.x);  // -> S_L1, S_L2SQ, S_DELTAMID, S_LINF
}

// generic fused reductions: out[0]=sum|a|, out[1]=sum a^2, out[2]=sum a*b (b may be null), out[3]=max|a|
__global__ void __launch_bounds__(VEC_THREADS) k_reduce4(const double2 *__restrict__ a, const double2 *__restrict__ b, size_t n2,
                                                          double *part, unsigned int *ticket, double *out)
{
    __shared__ double scratch[4 * 32];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        const double2 av = a[i];
        acc[0] += fabs(av.x) + fabs(av.y);
        acc[1] += av.x * av.x + av.y * av.y;
        if (b) {
            const double2 bv = b[i];
            acc[2] += av.x * bv.x + av.y * bv.y;
        }
        acc[3] = fmax(acc[3], fmax(fabs(av.x), fabs(av.y)));
    }
    grid_reduce<4, 3>(acc, scratch, part, ticket, out);
}

__global__ void __launch_bounds__(VEC_THREADS) k_axpy(double2 *__restrict__ y, double alpha, const double2 *__restrict__ x, size_t n2)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        double2 yv = y[i];
        const double2 xv = x[i];
        yv.x += alpha * xv.x;
        yv.y += alpha * xv.y;
        y[i] = yv;
    }
}

// the same with beta read from the device scalar block (nonlinear CG: beta is formed on the device like in the linear path)
__global__ void __launch_bounds__(VEC_THREADS) k_xpby_dev(double2 *__restrict__ y, const double *__restrict__ beta_dev, const double2 *__restrict__ x, size_t n2)
{
    const double beta = *beta_dev;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        double2 yv = y[i];
        const double2 xv = x[i];
        yv.x = xv.x + beta * yv.x;
        yv.y = xv.y + beta * yv.y;
        y[i] = yv;
    }
}

// y = x + beta*y     (d = s + beta d, solverCG.h:94)
__global__ void __launch_bounds__(VEC_THREADS) k_xpby(double2 *__restrict__ y, double beta, const double2 *__restrict__ x, size_t n2)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        double2 yv = y[i];
        const double2 xv = x[i];
        yv.x = xv.x + beta * yv.x;
        yv.y = xv.y + beta * yv.y;
        y[i] = yv;
    }
}

// delta = u - u_prev; u_prev = u; u += delta          (solver.h:302-311)
__global__ void __launch_bounds__(VEC_THREADS) k_extrapolate(double2 *__restrict__ u, double2 *__restrict__ up, size_t n2)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        const double2 uv = u[i], pv = up[i];
        up[i] = uv;
        u[i] = make_double2(uv.x + (uv.x - pv.x), uv.y + (uv.y - pv.y));
    }
}

// interleaved host order [voxel][h]  <->  component planes [h][voxel]
__global__ void k_aos_to_soa(const double *__restrict__ aos, double *__restrict__ soa, size_t nloc, int h)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nloc * h; i += (size_t)gridDim.x * blockDim.x) {
        const size_t v = i / h;
        const int c = (int)(i % h);
        soa[c * nloc + v] = aos[i];
    }
}
__global__ void k_soa_to_aos(const double *__restrict__ soa, double *__restrict__ aos, size_t nloc, int h)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nloc * h; i += (size_t)gridDim.x * blockDim.x) {
        const size_t v = i / h;
        const int c = (int)(i % h);
        aos[i] = soa[c * nloc + v];
    }
}

// scalar bookkeeping of one CG iteration after the convolution (solverCG.h:91-94):
//   delta0 = delta ; delta = <r,s> ; beta = fmax(0, (delta - deltamid)/delta0)
__global__ void k_scalars_after_conv(double *S)
{
    S += (size_t)blockIdx.x * S_COUNT;   // one block per lane of a batched solve
    const double delta0 = S[S_DELTA];
    const double delta = S[S_RS];
    S[S_DELTA0] = delta0;
    S[S_DELTA] = delta;
    S[S_BETA] = fmax(0.0, (delta - S[S_DELTAMID]) / delta0);
}

// slab halos: planes 0 and n0-1 of every component of  v = s ? s + beta*in : in  into contiguous send buffers [h][plane]
__global__ void k_pack_planes(const double *__restrict__ in, const double *__restrict__ s, const double *__restrict__ beta_dev,
                              size_t nloc, size_t plane, int n0, int h, double *__restrict__ out_lo, double *__restrict__ out_hi)
{
    const double beta = s ? *beta_dev : 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < plane * h; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / plane, o = i % plane;
        const size_t g0 = c * nloc + o, g1 = c * nloc + (size_t)(n0 - 1) * plane + o;
        if (out_lo) out_lo[i] = s ? s[g0] + beta * in[g0] : in[g0];
        if (out_hi) out_hi[i] = s ? s[g1] + beta * in[g1] : in[g1];
    }
}
// r(plane 0) += contribution of the previous rank's last element plane (solver.h:267-269)
__global__ void k_add_plane0(double *__restrict__ r, const double *__restrict__ add, size_t nloc, size_t plane, int h)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < plane * h; i += (size_t)gridDim.x * blockDim.x) {
        const size_t c = i / plane, o = i % plane;
        r[c * nloc + o] += add[i];
    }
}

// ------------------------------------------------------------------------------------------------
static int ensure_halo_buffers(fans_ctx *ctx)
{
    if (ctx->halo_send_lo) return FANS_OK;
    const size_t bytes = sizeof(double) * ctx->h * ctx->ny * ctx->nz;
    CUDA_TRY(ctx, cudaMalloc(&ctx->halo_send_lo, bytes));
    CUDA_TRY(ctx, cudaMalloc(&ctx->halo_send_hi, bytes));
    CUDA_TRY(ctx, cudaMalloc(&ctx->halo_lo, bytes));
    CUDA_TRY(ctx, cudaMalloc(&ctx->halo_hi, bytes));
    return FANS_OK;
}

// both neighbours get a plane of v = s ? s + beta*in : in  (gather-form stencil): fills ctx->halo_lo / halo_hi
int halo_exchange_both(fans_ctx *ctx, const double *in, const double *s, const double *beta_dev)
{
    FANS_CHECK(ensure_halo_buffers(ctx));
    const size_t plane = (size_t)ctx->ny * ctx->nz;
    k_pack_planes<<<vec_grid(plane * ctx->h), VEC_THREADS, 0, ctx->st>>>(in, s, beta_dev, ctx->nloc, plane, ctx->n0, ctx->h, ctx->halo_send_lo,
                                                                         ctx->halo_send_hi);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return comm_halo(ctx, ctx->halo_send_lo, ctx->halo_hi, ctx->halo_send_hi, ctx->halo_lo, sizeof(double) * ctx->h * plane);
}

// only the upper halo (node plane n0 = the next rank's plane 0), like the reference's u exchange (solver.h:244-245)
int halo_exchange_up(fans_ctx *ctx, const double *in, const double *s, const double *beta_dev)
{
    FANS_CHECK(ensure_halo_buffers(ctx));
    const size_t plane = (size_t)ctx->ny * ctx->nz;
    k_pack_planes<<<vec_grid(plane * ctx->h), VEC_THREADS, 0, ctx->st>>>(in, s, beta_dev, ctx->nloc, plane, ctx->n0, ctx->h, ctx->halo_send_lo, nullptr);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return comm_halo(ctx, ctx->halo_send_lo, ctx->halo_hi, nullptr, nullptr, sizeof(double) * ctx->h * plane);
}

// scatter-form assembly across the slab boundary: send the contributions to node plane n0 (ctx->halo_send_hi, written by the
// sweep) to the next rank and add what the previous rank sent into plane 0 (solver.h:264-269)
int halo_add_down(fans_ctx *ctx, double *r)
{
    const size_t plane = (size_t)ctx->ny * ctx->nz;
    FANS_CHECK(comm_halo(ctx, nullptr, nullptr, ctx->halo_send_hi, ctx->halo_lo, sizeof(double) * ctx->h * plane));
    k_add_plane0<<<vec_grid(plane * ctx->h), VEC_THREADS, 0, ctx->st>>>(r, ctx->halo_lo, ctx->nloc, plane, ctx->h);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

int vec_cg_update(fans_ctx *ctx, double *r, const double *kd, double *u, const double *d, const double *s)
{
    prof_begin(ctx, PC_CG_UPDATE);
    const size_t n2 = ((size_t)ctx->h * ctx->nloc + 1) / 2;   // fields carry one zero pad value (api.cu: ensure_field)
    // batched solve: lanes are exactly h*nloc doubles apart (even, checked by fans_solve_batch), every lane reads its own scalar block
    const size_t n2l = ctx->nb > 1 ? (size_t)ctx->h * ctx->nloc / 2 : n2;
    const unsigned gx = std::min(vec_grid(n2l), (unsigned)std::max(FANS_SMS, FANS_SMS * 16 / ctx->nb));
    k_cg_update<<<dim3(gx, ctx->nb), VEC_THREADS, 0, ctx->st>>>((double2 *)r, (const double2 *)kd, (double2 *)u, (const double2 *)d,
                                                               (const double2 *)s, n2l, ctx->d_red, ctx->d_part, ctx->d_ticket);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    if (ctx->nb > 1) return FANS_OK;   // single GPU: the host reads S_L1.. of every lane directly
    // error norms: MAX over the slabs for every measure (solver.h:430);  deltamid = <r,s>: SUM (solverCG.h:57)
    FANS_CHECK(comm_allreduce(ctx, ctx->d_red + S_L1, ctx->d_red + S_ERRMAX, 4, true));
    if (ctx->P > 1) FANS_CHECK(comm_allreduce(ctx, ctx->d_red + S_DELTAMID, ctx->d_red + S_DELTAMID, 1, false));
    return FANS_OK;
}

int vec_reduce4(fans_ctx *ctx, const double *a, const double *b, double *out_dev)
{
    prof_begin(ctx, PC_REDUCE);
    const size_t n2 = ((size_t)ctx->h * ctx->nloc + 1) / 2;   // fields carry one zero pad value (api.cu: ensure_field)
    k_reduce4<<<vec_grid(n2), VEC_THREADS, 0, ctx->st>>>((const double2 *)a, (const double2 *)b, n2, ctx->d_part, ctx->d_ticket, out_dev);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    if (out_dev == ctx->d_red + S_GEN) {  // norms: MAX over the slabs (solver.h:430); dot product: SUM (solverCG.h:57)
        FANS_CHECK(comm_allreduce(ctx, out_dev, ctx->d_red + S_GENMAX, 4, true));
        if (ctx->P > 1) FANS_CHECK(comm_allreduce(ctx, out_dev + 2, out_dev + 2, 1, false));
    } else if (out_dev == ctx->d_red + S_LS && ctx->P > 1) {  // only the dot product of this block is used
        FANS_CHECK(comm_allreduce(ctx, out_dev + 2, out_dev + 2, 1, false));
    }
    return FANS_OK;
}

int vec_axpy(fans_ctx *ctx, double *y, double alpha, const double *x)
{
    prof_begin(ctx, PC_AXPY);
    const size_t n2 = ((size_t)ctx->h * ctx->nloc + 1) / 2;   // fields carry one zero pad value (api.cu: ensure_field)
    k_axpy<<<vec_grid(n2), VEC_THREADS, 0, ctx->st>>>((double2 *)y, alpha, (const double2 *)x, n2);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

int vec_xpby(fans_ctx *ctx, double *y, double beta, const double *x)
{
    prof_begin(ctx, PC_AXPY);
    const size_t n2 = ((size_t)ctx->h * ctx->nloc + 1) / 2;   // fields carry one zero pad value (api.cu: ensure_field)
    k_xpby<<<vec_grid(n2), VEC_THREADS, 0, ctx->st>>>((double2 *)y, beta, (const double2 *)x, n2);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

int vec_xpby_dev(fans_ctx *ctx, double *y, const double *beta_dev, const double *x)
{
    prof_begin(ctx, PC_AXPY);
    const size_t n2 = ((size_t)ctx->h * ctx->nloc + 1) / 2;
    k_xpby_dev<<<vec_grid(n2), VEC_THREADS, 0, ctx->st>>>((double2 *)y, beta_dev, (const double2 *)x, n2);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

int vec_extrapolate(fans_ctx *ctx, double *u, double *up)
{
    prof_begin(ctx, PC_OTHER);
    const size_t n2 = ((size_t)ctx->h * ctx->nloc + 1) / 2;   // fields carry one zero pad value (api.cu: ensure_field)
    k_extrapolate<<<vec_grid(n2), VEC_THREADS, 0, ctx->st>>>((double2 *)u, (double2 *)up, n2);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

int vec_aos_to_soa(fans_ctx *ctx, const double *aos, double *soa)
{
    prof_begin(ctx, PC_OTHER);
    const size_t n = (size_t)ctx->h * ctx->nloc;
    k_aos_to_soa<<<vec_grid(n), VEC_THREADS, 0, ctx->st>>>(aos, soa, ctx->nloc, ctx->h);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

int vec_soa_to_aos(fans_ctx *ctx, const double *soa, double *aos)
{
    prof_begin(ctx, PC_OTHER);
    const size_t n = (size_t)ctx->h * ctx->nloc;
    k_soa_to_aos<<<vec_grid(n), VEC_THREADS, 0, ctx->st>>>(soa, aos, ctx->nloc, ctx->h);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

int vec_scalars_after_conv(fans_ctx *ctx)
{
    prof_begin(ctx, PC_OTHER);
    k_scalars_after_conv<<<ctx->nb, 1, 0, ctx->st>>>(ctx->d_red);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}
