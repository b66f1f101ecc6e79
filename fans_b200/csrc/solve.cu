// solve.cu — drivers: convolution, SolverCG / SolverFP loops, homogenized stress, mixed-BC update.
//   Solver::convolution            include/solver.h:387-412
//   Solver::solve / compute_error  include/solver.h:282-300, 414-452
//   SolverCG::internalSolve / LineSearchSecant   include/solverCG.h:61-160
//   SolverFP::internalSolve        include/solverFP.h:32-55
//   Solver::get_homogenized_stress include/solver.h:707-737
//   MixedBCController::update      include/mixedBCs.h:160-178
#include "internal.h"
#include "stencil.h"
#include <cstdlib>
#include <cmath>
#include <algorithm>

int ensure_fields(fans_ctx *ctx, std::initializer_list<int> ids);
int ensure_dalt(fans_ctx *ctx);
int check_fault(fans_ctx *ctx);
int check_fault_cached(fans_ctx *ctx);

int read_scalars(fans_ctx *ctx)
{
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_red, ctx->d_red, sizeof(double) * S_COUNT * ctx->nb, cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_fault, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));   // see check_fault_cached
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    prof_resolve(ctx);
    return FANS_OK;
}

static int write_scalar(fans_ctx *ctx, int slot, double v)
{
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    ctx->h_stage[0] = v;  // pinned staging slot
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_red + slot, ctx->h_stage, sizeof(double), cudaMemcpyHostToDevice, ctx->st));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    return FANS_OK;
}

// sums (and recycles) the event pairs conv_run left since the last call; the stream must have been synchronised past them
static double conv_time_resolve(fans_ctx *ctx)
{
    double total = 0.0;
    for (auto &pr : ctx->conv_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) total += ms;
        else cudaGetLastError();
        ctx->conv_pool.push_back(pr.first);
        ctx->conv_pool.push_back(pr.second);
    }
    ctx->conv_pending.clear();
    return total;
}

// Holds back the stream until the device word `gate` reaches gate_val (set by the first resident CTA of a y pass on the other
// stream), bounded by a ~10 ms timeout so a failed launch can never hang the device.
__global__ void k_gate(const int *gate, int gate_val)
{
    const long long t0 = clock64();
    while (*reinterpret_cast<const volatile int *>(gate) < gate_val && clock64() - t0 < 20000000LL) __nanosleep(200);
}
int conv_gate(fans_ctx *ctx, int gate_val, cudaStream_t st)
{
    k_gate<<<1, 1, 0, st>>>(ctx->d_gate, gate_val);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return FANS_OK;
}

// Slab convolution as a component pipeline (world_size > 1, fused transposes).  The y passes are NVLink-bound (7/8 of the spectrum
// leaves the GPU at 8 ranks) and the z passes HBM-bound, and the `howmany` components are independent up to the Green operator:
//     st :  zf(0)        | gate zf(1)          | gate zf(2)          |      barrier  x.Gamma  barrier |         zi(0)        | zi(1)        | zi(2)
//     st2:        yf(0) ------------- yf(1) ------------- yf(2) ----/                                  yi(0) -- yi(1) ------- yi(2) ------/
// A y pass runs persistent on y_grid SMs (its 1024-thread CTAs own the whole register file of an SM); k_gate releases the next z
// pass only once that y pass is resident, so the block scheduler hands the z pass the REMAINING SMs instead of starving the y pass.
static int conv_run_pipelined(fans_ctx *ctx, const double *in, double *out, double scale, const double *dotw, double *red_out)
{
    const int h = ctx->h;
    cudaStream_t s1 = ctx->st, s2 = ctx->st2;
    cudaEvent_t *ev = ctx->ev_pipe;
    for (int c = 0; c < h; ++c) {
        if (c > 0) FANS_CHECK(conv_gate(ctx, ctx->gate_seq, s1));  // y pass of component c-1 is resident
        FANS_CHECK(fft_pass_z_fwd_part(ctx, in, c, 1, s1));
        CUDA_TRY(ctx, cudaEventRecord(ev[c], s1));
        CUDA_TRY(ctx, cudaStreamWaitEvent(s2, ev[c], 0));
        FANS_CHECK(fft_pass_y_part(ctx, false, YLaunch{s2, c, 1, c + 1 < h ? ctx->y_grid : 0, ctx->d_gate, ++ctx->gate_seq}));
    }
    CUDA_TRY(ctx, cudaEventRecord(ev[3], s2));
    CUDA_TRY(ctx, cudaStreamWaitEvent(s1, ev[3], 0));
    FANS_CHECK(comm_barrier(ctx));
    FANS_CHECK(fft_pass_x_gamma(ctx));
    FANS_CHECK(comm_barrier(ctx));
    CUDA_TRY(ctx, cudaEventRecord(ev[4], s1));
    CUDA_TRY(ctx, cudaStreamWaitEvent(s2, ev[4], 0));
    for (int c = 0; c < h; ++c) {
        FANS_CHECK(fft_pass_y_part(ctx, true, YLaunch{s2, c, 1, c > 0 ? ctx->y_grid : 0, ctx->d_gate, ++ctx->gate_seq}));
        CUDA_TRY(ctx, cudaEventRecord(ev[5 + c], s2));
    }
    for (int c = 0; c < h; ++c) {
        CUDA_TRY(ctx, cudaStreamWaitEvent(s1, ev[5 + c], 0));
        if (c + 1 < h) FANS_CHECK(conv_gate(ctx, ctx->gate_seq - (h - 2 - c), s1));  // y pass of component c+1 is resident
        FANS_CHECK(fft_pass_z_inv_part(ctx, out, scale, dotw, red_out, c, 1, c > 0, s1));
    }
    if (red_out) FANS_CHECK(comm_allreduce(ctx, red_out, red_out, 1, false));
    else FANS_CHECK(comm_barrier(ctx));
    return FANS_OK;
}

// Slab convolution as a kz-CHUNKED pipeline (world_size > 1, fused transposes, h = 3).  The two transposes are NVLink-bound, the
// x pass with the Green operator is HBM / shared-memory bound, and a kz range of the spectrum is independent of every other one
// from the y pass to the inverse y pass:
//     st :  zf(all) | yf(0) yf(1) yf(2) yf(3)            | yi(0) yi(1) yi(2) yi(3) | zi(all)
//     st2:                 b x(0) b  b x(1) b  b x(2) b   b x(3) b
// yf(q) pushes chunk q of this rank's rows into the owners' transposed spectra; after a slab barrier (b) on the second stream the x
// pass of chunk q runs there, on the SMs the persistent y pass leaves free, while the first stream already pushes chunk q+1; a
// second barrier releases chunk q for the pulls of the inverse y pass.  Only the first y chunk and the last x chunk are exposed.
static int conv_run_chunked(fans_ctx *ctx, const double *in, double *out, double scale, const double *dotw, double *red_out)
{
    const int Q = ctx->chunks;
    cudaStream_t sA = ctx->st, sB = ctx->st2;
    cudaEvent_t *ev = ctx->ev_pipe;   // [0, Q): y chunk pushed ; [4, 4 + Q): x chunk done everywhere
    const int yTiles = (ctx->kzc + ctx->yT - 1) / ctx->yT, per = (yTiles + Q - 1) / Q;
    const int ratio = ctx->yT / ctx->gT;   // x tiles per y tile (1 or 2)
    const int ygrid = ctx->y_grid > 0 ? ctx->y_grid : 96;
    const int xcap = std::max(16, FANS_SMS - ygrid);
    FANS_CHECK(fft_pass_z_fwd(ctx, in));
    for (int q = 0; q < Q; ++q) {
        const int t0 = q * per, nt = std::min(per, yTiles - t0);
        if (nt <= 0) {
            CUDA_TRY(ctx, cudaEventRecord(ev[4 + q], sB));
            continue;
        }
        FANS_CHECK(fft_pass_y_part(ctx, false, YLaunch{sA, 0, ctx->h, ygrid, nullptr, 0, t0, nt}));
        CUDA_TRY(ctx, cudaEventRecord(ev[q], sA));
        CUDA_TRY(ctx, cudaStreamWaitEvent(sB, ev[q], 0));
        FANS_CHECK(comm_barrier_on(ctx, sB));   // every rank has pushed chunk q
        FANS_CHECK(fft_pass_x_gamma_part(ctx, sB, t0 * ratio, nt * ratio, xcap));
        FANS_CHECK(comm_barrier_on(ctx, sB));   // every rank is done with chunk q: it may be pulled
        CUDA_TRY(ctx, cudaEventRecord(ev[4 + q], sB));
    }
    for (int q = 0; q < Q; ++q) {
        const int t0 = q * per, nt = std::min(per, yTiles - t0);
        CUDA_TRY(ctx, cudaStreamWaitEvent(sA, ev[4 + q], 0));
        if (nt > 0) FANS_CHECK(fft_pass_y_part(ctx, true, YLaunch{sA, 0, ctx->h, q + 1 < Q ? ygrid : 0, nullptr, 0, t0, nt}));
    }
    FANS_CHECK(fft_pass_z_inv(ctx, out, scale, dotw, red_out));
    if (red_out) FANS_CHECK(comm_allreduce(ctx, red_out, red_out, 1, false));
    else FANS_CHECK(comm_barrier(ctx));
    return FANS_OK;
}

// out = scale * Gamma * in ;  optional red_out[0] = <dotw, out>
int conv_run(fans_ctx *ctx, const double *in, double *out, double scale, const double *dotw, double *red_out)
{
    if (!ctx->gamma_ready) {
        fans_set_error(ctx, FANS_ERR_STATE, "fundamental solution not built: call fans_set_reference_stiffness first");
        return FANS_ERR_STATE;
    }
    // event pair around the whole convolution (on ctx->st: the pipelined form joins its second stream before it returns)
    auto conv_event = [&]() {
        cudaEvent_t e = nullptr;
        if (!ctx->conv_pool.empty()) {
            e = ctx->conv_pool.back();
            ctx->conv_pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        cudaEventRecord(e, ctx->st);
        return e;
    };
    if (!ctx->capturing && ctx->conv_pending.size() > 1024) {  // many convolutions outside a solve: do not let the event pairs pile up
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
        conv_time_resolve(ctx);
    }
    const cudaEvent_t ev_a = ctx->capturing ? nullptr : conv_event();
    struct ConvTimer {
        fans_ctx *c;
        cudaEvent_t a;
        decltype(conv_event) &mk;
        ~ConvTimer()
        {
            if (a) c->conv_pending.emplace_back(a, mk());
        }
    } conv_timer{ctx, ev_a, conv_event};
    if (ctx->any_fft) return conv_run_any(ctx, in, out, scale, dotw, red_out);
    if (ctx->pipe && ctx->chunks > 0 && !ctx->prof) return conv_run_chunked(ctx, in, out, scale, dotw, red_out);
    if (ctx->pipe && !ctx->prof) return conv_run_pipelined(ctx, in, out, scale, dotw, red_out);
    FANS_CHECK(fft_pass_z_fwd(ctx, in));
    FANS_CHECK(fft_pass_y(ctx, false));
    // x-slabs -> y-slabs (FFTW_MPI_TRANSPOSED_OUT) and back (FFTW_MPI_TRANSPOSED_IN).  Fused form: the y / x pass has already
    // stored its rows into the owners' buffers over NVLink, only a stream-ordered barrier is left; otherwise block all-to-all.
    if (ctx->P > 1) FANS_CHECK(ctx->p2p ? comm_barrier(ctx) : comm_alltoall(ctx, ctx->spec, ctx->specB));
    FANS_CHECK(fft_pass_x_gamma(ctx));
    if (ctx->P > 1) FANS_CHECK(ctx->p2p ? comm_barrier(ctx) : comm_alltoall(ctx, ctx->specB, ctx->spec));
    FANS_CHECK(fft_pass_y(ctx, true));
    FANS_CHECK(fft_pass_z_inv(ctx, out, scale, dotw, red_out));
    if (red_out && ctx->P > 1) FANS_CHECK(comm_allreduce(ctx, red_out, red_out, 1, false));
    // fused form: nobody may push the next spectrum into a buffer a peer is still pulling from (the all-reduce above is a barrier)
    else if (ctx->P > 1 && ctx->p2p) FANS_CHECK(comm_barrier(ctx));
    return FANS_OK;
}

// ---- compute_error (solver.h:414-452). Reads the norms the last fused pass left in the scalar block. ----
struct ErrState {
    int measure, err_type;
    double err0;
    double *hist;
    int iter;
};

static double error_from_scalars(fans_ctx *ctx, ErrState &es, int base)
{
    double err;
    if (es.measure == FANS_MEASURE_L1) err = ctx->h_red[base + 0];
    else if (es.measure == FANS_MEASURE_L2) err = std::sqrt(ctx->h_red[base + 1]);
    else err = ctx->h_red[base + 3];
    if (es.hist) es.hist[es.iter] = err;
    if (es.iter == 0) es.err0 = err;
    const double err_rel = (es.iter == 0) ? 100.0 : err / es.err0;
    return es.err_type == FANS_ERR_ABSOLUTE ? err : err_rel;
}

static int compute_error(fans_ctx *ctx, const double *r, ErrState &es, double *err_out)
{
    FANS_CHECK(vec_reduce4(ctx, r, nullptr, ctx->d_red + S_GEN));
    FANS_CHECK(read_scalars(ctx));
    *err_out = error_from_scalars(ctx, es, S_GENMAX);
    return FANS_OK;
}

// ---- homogenized stress: one strain/stress sweep + sum (solver.h:707-737) ----
static int homogenized_stress(fans_ctx *ctx, double *out)
{
    FANS_CHECK(ensure_fields(ctx, {FANS_FIELD_U}));
    FANS_CHECK(sweep_run(ctx, SWEEP_STRAINSTRESS, ctx->field[FANS_FIELD_U], nullptr, nullptr, nullptr, nullptr, ctx->d_red + S_STRESS,
                         nullptr, nullptr));
    if (ctx->P > 1) FANS_CHECK(comm_allreduce(ctx, ctx->d_red + S_STRESS, ctx->d_red + S_STRESS, ctx->nstr, false));  // solver.h:733
    FANS_CHECK(read_scalars(ctx));
    const double N = (double)ctx->nx * ctx->ny * ctx->nz;
    for (int i = 0; i < ctx->nstr; ++i) out[i] = ctx->h_red[S_STRESS + i] / N;
    return check_fault(ctx);
}

extern "C" int fans_homogenized_stress(fans_ctx *ctx, double *out)
{
    if (!ctx || !out) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    return homogenized_stress(ctx, out);
}

extern "C" int fans_set_mixed_bc(fans_ctx *ctx, const fans_mixed_bc *mbc)
{
    if (!ctx) return FANS_ERR_ARG;
    if (!mbc) {
        ctx->mixed = false;
        return FANS_OK;
    }
    if (mbc->n_F < 0 || mbc->n_F > ctx->nstr) {
        fans_set_error(ctx, FANS_ERR_ARG, "mixed BC: invalid number of stress-controlled components");
        return FANS_ERR_ARG;
    }
    ctx->mbc = *mbc;
    ctx->mixed = true;
    return FANS_OK;
}

// g0 += Q_F M (P_target - Q_F^T Pbar)    (mixedBCs.h:160-178)
extern "C" int fans_update_mixed_bc(fans_ctx *ctx)
{
    if (!ctx) return FANS_ERR_ARG;
    if (!ctx->mixed) return FANS_OK;
    cudaSetDevice(ctx->device);
    double Pbar[9];
    FANS_CHECK(homogenized_stress(ctx, Pbar));
    const fans_mixed_bc &m = ctx->mbc;
    if (m.n_F > 0) {
        double rhs[9], dE[9];
        for (int i = 0; i < m.n_F; ++i) rhs[i] = m.P_target[i] - Pbar[m.idx_F[i]];
        for (int i = 0; i < m.n_F; ++i) {
            double s = 0.0;
            for (int j = 0; j < m.n_F; ++j) s += m.M[i * m.n_F + j] * rhs[j];
            dE[i] = s;
        }
        for (int i = 0; i < m.n_F; ++i) ctx->g0[m.idx_F[i]] += dE[i];
    }
    return FANS_OK;
}

// ---- one linear CG iteration as a CUDA graph (see IterGraph, common.cuh) ----
static void iter_graph_free(IterGraph &G)
{
    for (cudaGraphExec_t &e : G.exec)
        if (e) cudaGraphExecDestroy(e), e = nullptr;
    G.valid = false;
}
void iter_graph_free(fans_ctx *ctx)
{
    iter_graph_free(ctx->igraph);
    iter_graph_free(ctx->arena.graph);
}

static bool iter_graph_wanted(const fans_ctx *ctx, int lanes)
{
    if (const char *e = getenv("FANS_GRAPH")) return e[0] == '1';
    return ctx->nloc * (size_t)lanes <= (size_t)128 * 128 * 128;   // above this an iteration is bound by HBM, not by launches
}

static void iter_graph_key(const fans_ctx *ctx, const double *r, const double *s, const double *u, const double *rnew, const void *(&key)[8])
{
    key[0] = r, key[1] = s, key[2] = u, key[3] = rnew, key[4] = ctx->spec, key[5] = ctx->gamma, key[6] = ctx->phidx, key[7] = ctx->d_red;
}

// graphs exist for exactly these fields / tables / lanes and the current direction D is one of the two buffers they alternate between
static bool iter_graph_ready(const fans_ctx *ctx, const IterGraph &G, const double *r, const double *s, const double *u, const double *rnew,
                             const double *D, const double *Dalt)
{
    if (!G.valid || G.cstamp != ctx->const_stamp || G.sstamp != ctx->stencil_stamp || G.nb != ctx->nb) return false;
    const void *key[8];
    iter_graph_key(ctx, r, s, u, rnew, key);
    for (int i = 0; i < 8; ++i)
        if (key[i] != G.key[i]) return false;
    return (D == G.dA && Dalt == G.dB) || (D == G.dB && Dalt == G.dA);
}

// the body of the linear iteration (all lanes of a batched solve at once); under capture the trailing read-back has no synchronisation
static int linear_iteration(fans_ctx *ctx, double *r, double *s, double *u, double *rnew, const double *d_old, double *d_new, bool use_stencil)
{
    FANS_CHECK(conv_run(ctx, r, s, -1.0, r, ctx->d_red + S_RS));         // s = -Gamma r ; S_RS = <r,s>
    FANS_CHECK(vec_scalars_after_conv(ctx));                             // delta0, delta, beta
    if (use_stencil) FANS_CHECK(stencil_run(ctx, d_old, rnew, s, d_new, ctx->d_red + S_BETA, ctx->d_red + S_DKD));
    else FANS_CHECK(sweep_run(ctx, SWEEP_LINEAR, d_old, rnew, s, d_new, ctx->d_red + S_BETA, ctx->d_red + S_DKD, nullptr, nullptr));
    FANS_CHECK(vec_cg_update(ctx, r, rnew, u, d_new, s));                // r,u update + norms + deltamid
    if (ctx->capturing) {
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_red, ctx->d_red, sizeof(double) * S_COUNT * ctx->nb, cudaMemcpyDeviceToHost, ctx->st));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_fault, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
        return FANS_OK;
    }
    return read_scalars(ctx);
}

// Captures the iteration for both orientations of the direction ping-pong: exec[0] reads dA and writes dB.  Called after one plain
// iteration of the same solve, so every lazy initialisation of the launchers (coefficient tables, function attributes) lies behind.
// A failed capture only switches the graphs off for this context: the plain launches remain.
static int iter_graph_build(fans_ctx *ctx, IterGraph &G, double *r, double *s, double *u, double *rnew, double *dA, double *dB)
{
    iter_graph_free(G);
    G.dA = dA, G.dB = dB;
    for (int par = 0; par < 2; ++par) {
        const double *d_old = par ? G.dB : G.dA;
        double *d_new = par ? G.dA : G.dB;
        const int64_t l0 = ctx->launches;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(ctx->st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            G.failed = true;
            return FANS_OK;
        }
        ctx->capturing = true;
        const int rc = linear_iteration(ctx, r, s, u, rnew, d_old, d_new, true);
        ctx->capturing = false;
        const cudaError_t ce = cudaStreamEndCapture(ctx->st, &graph);
        G.launches = (int)(ctx->launches - l0);
        ctx->launches = l0;   // nothing ran
        if (rc != FANS_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&G.exec[par], graph, 0) != cudaSuccess) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            iter_graph_free(G);
            G.failed = true;
            return FANS_OK;
        }
        cudaGraphDestroy(graph);
    }
    iter_graph_key(ctx, r, s, u, rnew, G.key);
    G.cstamp = ctx->const_stamp, G.sstamp = ctx->stencil_stamp, G.nb = ctx->nb;
    G.valid = true;
    return FANS_OK;
}

// ------------------------------------------------------------------------------------------------
// SolverCG::internalSolve, linear fast path (solverCG.h:96-107): everything between two error checks stays on
// the device; alpha/beta are formed from device scalars. One host poll per iteration (the error).
// ------------------------------------------------------------------------------------------------
static int solve_cg(fans_ctx *ctx, const fans_solve_params *p, fans_solve_result *res, ErrState &es)
{
    FANS_CHECK(ensure_fields(ctx, {FANS_FIELD_U, FANS_FIELD_R, FANS_FIELD_S, FANS_FIELD_D, FANS_FIELD_RNEW}));
    double *u = ctx->field[FANS_FIELD_U], *r = ctx->field[FANS_FIELD_R], *s = ctx->field[FANS_FIELD_S];
    double *rnew = ctx->field[FANS_FIELD_RNEW];
    const size_t fbytes = sizeof(double) * ctx->h * ctx->nloc;
    const bool linear = ctx->all_linear && !ctx->mixed && !p->force_nonlinear;
    if (linear) FANS_CHECK(ensure_dalt(ctx));
    const bool use_stencil = stencil_supported(ctx) && !getenv("FANS_LINEAR_SWEEP");
    // s = 0, d = 0 (solverCG.h:70-74), alpha_warm = 0.1, delta = 1 (solverCG.h:68,82)
    CUDA_TRY(ctx, cudaMemsetAsync(s, 0, fbytes, ctx->st));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->field[FANS_FIELD_D], 0, fbytes, ctx->st));
    double alpha_warm = 0.1;
    ctx->n_residual_evals++;
    FANS_CHECK(sweep_run(ctx, SWEEP_RESIDUAL, u, r, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
    es.iter = 0;
    double err_rel;
    FANS_CHECK(compute_error(ctx, r, es, &err_rel));
    FANS_CHECK(check_fault(ctx));
    if (p->verbose) printf("it %3d .... err %16.8e\n", es.iter, es.hist ? es.hist[es.iter] : err_rel);
    FANS_CHECK(write_scalar(ctx, S_DELTA, 1.0));
    FANS_CHECK(write_scalar(ctx, S_DELTAMID, 0.0));  // <r, s> with s = 0
    // small grids: iterations after the first replay a CUDA graph (single GPU, stencil form, no per-kernel profiling)
    const bool graph_ok = linear && use_stencil && ctx->P == 1 && ctx->nb == 1 && !ctx->prof && !ctx->any_fft && iter_graph_wanted(ctx, 1);
    bool graph_tried = false;
    int graph_iters = 0;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_loop0, ctx->st));
    while (es.iter < p->n_it && err_rel > p->tol) {
        if (linear) {
            // deltamid already sits in S_DELTAMID (left by k_cg_update, 0 at iter 0)
            double *d_old = ctx->field[FANS_FIELD_D], *d_new = ctx->d_alt;
            ctx->n_residual_evals++;
            if (graph_ok && iter_graph_ready(ctx, ctx->igraph, r, s, u, rnew, d_old, d_new)) {   // the whole iteration as one graph launch
                const auto &G = ctx->igraph;
                CUDA_TRY(ctx, cudaGraphLaunch(G.exec[d_old == G.dA ? 0 : 1], ctx->st));
                ctx->launches += G.launches;
                CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
                graph_iters++;
            } else {
                FANS_CHECK(linear_iteration(ctx, r, s, u, rnew, d_old, d_new, use_stencil));
                if (graph_ok && !ctx->igraph.failed && !graph_tried) {
                    graph_tried = true;
                    FANS_CHECK(iter_graph_build(ctx, ctx->igraph, r, s, u, rnew, d_new, d_old));   // the next iteration reads d_new
                }
            }
            ctx->field[FANS_FIELD_D] = d_new;
            ctx->d_alt = d_old;
            es.iter++;
            err_rel = error_from_scalars(ctx, es, S_ERRMAX);
        } else {
            double *d = ctx->field[FANS_FIELD_D];
            // Host round trips per iteration: one per residual evaluation of the line search (the secant decisions are taken on the
            // host) and one for the error — delta, beta and the direction update stay on the device like in the linear path.
            // deltamid = <r,s> (solverCG.h:86): left in S_GEN+2 by the reduction that produced the error of the previous iteration
            // (s = 0 before the first iteration, the initial compute_error leaves 0 there)
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_red + S_DELTAMID, ctx->d_red + S_GEN + 2, sizeof(double), cudaMemcpyDeviceToDevice, ctx->st));
            FANS_CHECK(conv_run(ctx, r, s, -1.0, r, ctx->d_red + S_RS));        // s = -Gamma r ; S_RS = <r,s>
            FANS_CHECK(vec_scalars_after_conv(ctx));                            // delta0 = delta ; delta = <r,s> ; beta (solverCG.h:91-94)
            FANS_CHECK(vec_xpby_dev(ctx, d, ctx->d_red + S_BETA, s));           // d = s + beta d
            // ---- LineSearchSecant (solverCG.h:119-160) ----
            double err = 10.0;
            int it = 0;
            double alpha_prev = 0.0, alpha_curr = alpha_warm;
            FANS_CHECK(vec_reduce4(ctx, r, d, ctx->d_red + S_LS));              // <r,d>, read below together with <rnew,d>
            FANS_CHECK(vec_axpy(ctx, u, alpha_curr, d));
            FANS_CHECK(fans_update_mixed_bc(ctx));
            ctx->n_residual_evals++;
            FANS_CHECK(sweep_run(ctx, SWEEP_RESIDUAL, u, rnew, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
            FANS_CHECK(vec_reduce4(ctx, rnew, d, ctx->d_red + S_GEN));
            FANS_CHECK(read_scalars(ctx));
            double rpd = ctx->h_red[S_LS + 2];
            double r1pd = ctx->h_red[S_GEN + 2];
            while (it < p->ls_max_iter && err > p->ls_tol) {
                const double denom = r1pd - rpd;
                if (std::fabs(denom) < 1e-14 * (std::fabs(r1pd) + std::fabs(rpd))) break;
                double alpha_next = alpha_curr - r1pd * (alpha_curr - alpha_prev) / denom;
                if (alpha_next <= 0.0) alpha_next = 0.5 * (alpha_prev + alpha_curr);
                err = std::fabs(alpha_next - alpha_curr);
                FANS_CHECK(vec_axpy(ctx, u, alpha_next - alpha_curr, d));
                alpha_prev = alpha_curr;
                rpd = r1pd;
                alpha_curr = alpha_next;
                it++;
                FANS_CHECK(fans_update_mixed_bc(ctx));
                ctx->n_residual_evals++;
                FANS_CHECK(sweep_run(ctx, SWEEP_RESIDUAL, u, rnew, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
                FANS_CHECK(vec_reduce4(ctx, rnew, d, ctx->d_red + S_GEN));
                FANS_CHECK(read_scalars(ctx));
                r1pd = ctx->h_red[S_GEN + 2];
            }
            alpha_warm = (it == p->ls_max_iter && err > p->ls_tol) ? 0.1 : alpha_curr;
            // v_r = rnew (solverCG.h:157): swap the buffers instead of copying
            ctx->field[FANS_FIELD_R] = rnew;
            ctx->field[FANS_FIELD_RNEW] = r;
            r = ctx->field[FANS_FIELD_R];
            rnew = ctx->field[FANS_FIELD_RNEW];
            if (p->verbose) printf("line search iter %i, alpha %f - error %e - ", it, alpha_curr, err);
            es.iter++;
            // compute_error (solver.h:414-452) and the next iteration's deltamid = <r,s> in ONE pass over r; the fault flag rides along
            FANS_CHECK(vec_reduce4(ctx, r, s, ctx->d_red + S_GEN));
            FANS_CHECK(read_scalars(ctx));
            FANS_CHECK(check_fault_cached(ctx));
            err_rel = error_from_scalars(ctx, es, S_GENMAX);
        }
        if (p->verbose) printf("it %3d .... err %16.8e\n", es.iter, es.hist ? es.hist[es.iter] : err_rel);
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_loop1, ctx->st));
    res->err_last = err_rel;
    ctx->igraph.replays = graph_iters;   // fans_solve: the convolutions inside graph replays carry no event pair
    return FANS_OK;
}

// SolverFP::internalSolve (solverFP.h:32-55): convolution in place on v_r, u -= r
static int solve_fp(fans_ctx *ctx, const fans_solve_params *p, fans_solve_result *res, ErrState &es)
{
    FANS_CHECK(ensure_fields(ctx, {FANS_FIELD_U, FANS_FIELD_R}));
    double *u = ctx->field[FANS_FIELD_U], *r = ctx->field[FANS_FIELD_R];
    ctx->n_residual_evals++;
    FANS_CHECK(sweep_run(ctx, SWEEP_RESIDUAL, u, r, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
    es.iter = 0;
    double err_rel;
    FANS_CHECK(compute_error(ctx, r, es, &err_rel));
    FANS_CHECK(check_fault(ctx));
    if (p->verbose) printf("it %3d .... err %16.8e\n", es.iter, es.hist ? es.hist[es.iter] : err_rel);
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_loop0, ctx->st));
    while (es.iter < p->n_it && err_rel > p->tol) {
        FANS_CHECK(conv_run(ctx, r, r, 1.0, nullptr, nullptr));
        FANS_CHECK(vec_axpy(ctx, u, -1.0, r));
        FANS_CHECK(fans_update_mixed_bc(ctx));
        ctx->n_residual_evals++;
        FANS_CHECK(sweep_run(ctx, SWEEP_RESIDUAL, u, r, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
        es.iter++;
        FANS_CHECK(compute_error(ctx, r, es, &err_rel));
        FANS_CHECK(check_fault_cached(ctx));   // the flag came back with the scalars of compute_error
        if (p->verbose) printf("it %3d .... err %16.8e\n", es.iter, es.hist ? es.hist[es.iter] : err_rel);
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_loop1, ctx->st));
    res->err_last = err_rel;
    return FANS_OK;
}

// Solver::solve (solver.h:282-300): err_all = 0, internalSolve, update_internal_variables
extern "C" int fans_solve(fans_ctx *ctx, const fans_solve_params *p, fans_solve_result *res, double *err_hist)
{
    if (!ctx || !p || !res) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (p->measure < FANS_MEASURE_L1 || p->measure > FANS_MEASURE_LINF) {
        fans_set_error(ctx, FANS_ERR_ARG, "Unknown measure type");
        return FANS_ERR_ARG;
    }
    if (p->err_type != FANS_ERR_ABSOLUTE && p->err_type != FANS_ERR_RELATIVE) {
        fans_set_error(ctx, FANS_ERR_ARG, "Unknown error type");
        return FANS_ERR_ARG;
    }
    memset(res, 0, sizeof(*res));
    ctx->igraph.replays = 0;
    if (err_hist)
        for (int i = 0; i <= p->n_it; ++i) err_hist[i] = 0.0;
    ErrState es{p->measure, p->err_type, 0.0, err_hist, 0};
    const int evals0 = ctx->n_residual_evals;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    conv_time_resolve(ctx);  // drop the pairs of convolutions issued outside a solve (fans_convolution)
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->st));
    int rc;
    if (p->method == FANS_METHOD_CG) rc = solve_cg(ctx, p, res, es);
    else if (p->method == FANS_METHOD_FP) rc = solve_fp(ctx, p, res, es);
    else {
        fans_set_error(ctx, FANS_ERR_ARG, "not a valid method");
        return FANS_ERR_ARG;
    }
    if (rc != FANS_OK) return rc;
    FANS_CHECK(fans_commit_history(ctx));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->st));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    res->elapsed_ms = ms;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_loop0, ctx->ev_loop1));
    res->loop_ms = ms;
    prof_resolve(ctx);
    res->fft_ms = conv_time_resolve(ctx);
    if (p->method == FANS_METHOD_CG && ctx->igraph.replays > 0) {   // replayed iterations: the convolution time of the plain ones stands in
        const int plain = es.iter - ctx->igraph.replays;
        if (plain > 0) ctx->igraph.fft_per_iter = res->fft_ms / plain;
        res->fft_ms = ctx->igraph.fft_per_iter * es.iter;
    }
    res->iters = es.iter;
    res->n_residual_evals = ctx->n_residual_evals - evals0;
    return FANS_OK;
}

// ------------------------------------------------------------------------------------------------
// Batched linear solves.  Solver::get_homogenized_tangent (solver.h:739-778) runs n_str solves one after the other over the same
// microstructure and the same Gamma_hat; config 2 of the reference's benchmark set is exactly these "6 load cases".  Here the n_b
// right-hand sides are LANES of one CG loop: every pass of the iteration (z/y/x transforms, Gamma_hat multiply, stencil, update)
// is ONE launch over all lanes, each lane with its own scalar block (alpha, beta, delta, norms) on the device.  What the lanes
// share — Gamma_hat in the x pass (24 of 554 B/voxel), the phase image in the stencil (2 B) — is read from HBM once per tile and
// from L2 for the other lanes; what a small grid cannot fill by itself (a 32^3 micro problem keeps 8..30 of 148 SMs busy and is
// launch-latency bound) the lanes fill together.  A lane that has converged is frozen (S_FREEZE) and keeps its u and r bit for bit.
// Every lane runs the arithmetic of the single solve on its own data, so lane l reproduces fans_solve with g0 = macro[l], u = 0.
// ------------------------------------------------------------------------------------------------
void batch_arena_free(fans_ctx *ctx)
{
    auto &A = ctx->arena;
    iter_graph_free(A.graph);
    for (double *&f : A.field)
        if (f) cudaFree(f), f = nullptr;
    if (A.spec) cudaFree(A.spec), A.spec = nullptr;
    if (A.red) cudaFree(A.red), A.red = nullptr;
    if (A.part) cudaFree(A.part), A.part = nullptr;
    if (A.ticket) cudaFree(A.ticket), A.ticket = nullptr;
    if (A.h_red) cudaFreeHost(A.h_red), A.h_red = nullptr;
    if (A.h_flag) cudaFreeHost(A.h_flag), A.h_flag = nullptr;
    A.lanes = 0;
}

static size_t batch_lane_parts(const fans_ctx *ctx)
{
    // partial sums one lane needs: z inverse pass (one per CTA, at most one CTA per line), update (4 per CTA), stencil (one per CTA)
    return std::max<size_t>((size_t)ctx->h * ctx->n0 * ctx->ny + 64, 4 * (size_t)FANS_SMS * 16 + 64);
}

static int batch_arena_ensure(fans_ctx *ctx, int lanes)
{
    auto &A = ctx->arena;
    if (A.lanes >= lanes) return FANS_OK;
    batch_arena_free(ctx);
    const size_t fN = (size_t)ctx->h * ctx->nloc;
    const size_t specN = (size_t)ctx->h * ctx->n0 * ((size_t)ctx->n1 * ctx->kzp + ctx->xpad);
    auto fail = [&](const char *what) {
        cudaGetLastError();
        batch_arena_free(ctx);
        fans_set_error(ctx, FANS_ERR_CUDA, std::string("fans_solve_batch: out of device memory for ") + std::to_string(lanes) + " lanes (" + what + ")");
        return FANS_ERR_CUDA;
    };
    for (double *&f : A.field) {
        if (cudaMalloc(&f, sizeof(double) * (fN * lanes + 2)) != cudaSuccess) return fail("fields");
        if (cudaMemsetAsync(f, 0, sizeof(double) * (fN * lanes + 2), ctx->st) != cudaSuccess) return fail("fields");
    }
    if (cudaMalloc(&A.spec, sizeof(double2) * specN * lanes) != cudaSuccess) return fail("spectrum");
    if (cudaMemsetAsync(A.spec, 0, sizeof(double2) * specN * lanes, ctx->st) != cudaSuccess) return fail("spectrum");
    const size_t nparts = std::max<size_t>((size_t)1 << 20, batch_lane_parts(ctx) * lanes);
    if (cudaMalloc(&A.part, sizeof(double) * nparts) != cudaSuccess) return fail("partial sums");
    if (cudaMalloc(&A.red, sizeof(double) * S_COUNT * lanes) != cudaSuccess) return fail("scalars");
    if (cudaMalloc(&A.ticket, sizeof(unsigned int) * (lanes + 2)) != cudaSuccess) return fail("tickets");
    if (cudaMemsetAsync(A.ticket, 0, sizeof(unsigned int) * (lanes + 2), ctx->st) != cudaSuccess) return fail("tickets");
    if (cudaMallocHost(&A.h_red, sizeof(double) * S_COUNT * lanes) != cudaSuccess) return fail("pinned scalars");
    if (cudaMallocHost(&A.h_flag, sizeof(double) * lanes) != cudaSuccess) return fail("pinned flags");
    A.lanes = lanes;
    return FANS_OK;
}

// why this context cannot run a batched solve (nullptr: it can)
static const char *batch_unsupported(const fans_ctx *ctx, const fans_solve_params *p)
{
    if (!ctx->materials_ready || !ctx->ms_ready || !ctx->gamma_ready) return "microstructure, materials and reference stiffness must be set first";
    if (!ctx->all_linear) return "batched solves need linear material models (the reference's perturbation loop remains for the others)";
    if (!stencil_supported(ctx)) return "batched solves use the stencil form of the linear operator (even n_z, one stiffness per phase)";
    if (ctx->P > 1) return "batched solves run on one GPU (slab-decomposed problems solve the load cases one after the other)";
    if (ctx->any_fft) return "batched solves need power-of-two grid dimensions";
    if (((size_t)ctx->h * ctx->nloc) % 2) return "batched solves need an even number of field values";
    if (p->method != FANS_METHOD_CG) return "batched solves use the CG method";
    if (ctx->mixed) return "batched solves are strain-driven (disable mixed boundary conditions)";
    return nullptr;
}

extern "C" int fans_solve_batch(fans_ctx *ctx, int32_t nb, const double *macro, const fans_solve_params *p, fans_solve_result *res,
                                double *stress_out, double *err_hist)
{
    if (!ctx || !macro || !p || !res || nb < 1 || nb > FANS_MAX_BATCH) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (const char *why = batch_unsupported(ctx, p)) {
        fans_set_error(ctx, FANS_ERR_STATE, std::string("fans_solve_batch: ") + why);
        return FANS_ERR_STATE;
    }
    if (p->measure < FANS_MEASURE_L1 || p->measure > FANS_MEASURE_LINF) {
        fans_set_error(ctx, FANS_ERR_ARG, "Unknown measure type");
        return FANS_ERR_ARG;
    }
    if (p->err_type != FANS_ERR_ABSOLUTE && p->err_type != FANS_ERR_RELATIVE) {
        fans_set_error(ctx, FANS_ERR_ARG, "Unknown error type");
        return FANS_ERR_ARG;
    }
    FANS_CHECK(batch_arena_ensure(ctx, nb));
    auto &A = ctx->arena;
    const size_t fN = (size_t)ctx->h * ctx->nloc;
    const int nstr = ctx->nstr;
    double *U = A.field[0], *R = A.field[1], *S = A.field[2], *D = A.field[3], *Dalt = A.field[4], *KD = A.field[5];
    memset(res, 0, sizeof(*res) * nb);
    if (err_hist)
        for (size_t i = 0; i < (size_t)nb * (p->n_it + 1); ++i) err_hist[i] = 0.0;

    // the context's own scalars, spectrum and gradient are swapped for the arena's while the lanes run
    struct Swap {
        fans_ctx *c;
        double2 *spec;
        double *red, *h_red, *part;
        unsigned int *ticket;
        double g0[9];
        ~Swap()
        {
            c->spec = spec, c->d_red = red, c->h_red = h_red, c->d_part = part, c->d_ticket = ticket, c->nb = 1;
            memcpy(c->g0, g0, sizeof(g0));
        }
    } swap{ctx, ctx->spec, ctx->d_red, ctx->h_red, ctx->d_part, ctx->d_ticket, {}};
    memcpy(swap.g0, ctx->g0, sizeof(swap.g0));
    ctx->spec = A.spec, ctx->d_red = A.red, ctx->h_red = A.h_red, ctx->d_part = A.part, ctx->d_ticket = A.ticket;

    const int evals0 = ctx->n_residual_evals;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    conv_time_resolve(ctx);
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->st));
    for (double *f : {U, S, D}) CUDA_TRY(ctx, cudaMemsetAsync(f, 0, sizeof(double) * fN * nb, ctx->st));
    CUDA_TRY(ctx, cudaMemsetAsync(A.red, 0, sizeof(double) * S_COUNT * nb, ctx->st));

    // initial residuals r_l = residual(u = 0; g0 = macro[l]) and their error norms (solverCG.h:76-80): one element sweep and one
    // reduction per lane, each into the lane's own scalar block, ONE read-back for all of them
    std::vector<ErrState> es(nb);
    std::vector<double> err(nb, 0.0);
    std::vector<char> active(nb, 1);
    for (int l = 0; l < nb; ++l) {
        es[l] = ErrState{p->measure, p->err_type, 0.0, err_hist ? err_hist + (size_t)l * (p->n_it + 1) : nullptr, 0};
        for (int i = 0; i < nstr; ++i) ctx->g0[i] = macro[(size_t)l * nstr + i];
        ctx->n_residual_evals++;
        FANS_CHECK(sweep_run(ctx, SWEEP_RESIDUAL, U + l * fN, R + l * fN, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
        FANS_CHECK(vec_reduce4(ctx, R + l * fN, nullptr, ctx->d_red + (size_t)l * S_COUNT + S_GEN));
    }
    ctx->nb = nb;   // read_scalars: all lanes' blocks
    FANS_CHECK(read_scalars(ctx));
    ctx->nb = 1;
    FANS_CHECK(check_fault_cached(ctx));
    for (int l = 0; l < nb; ++l) {
        err[l] = error_from_scalars(ctx, es[l], l * S_COUNT + S_GEN);
        if (p->verbose) printf("lane %d it %3d .... err %16.8e\n", l, 0, es[l].hist ? es[l].hist[0] : err[l]);
    }
    // delta = 1, deltamid = <r, s> = 0 (s = 0), nothing frozen
    for (int l = 0; l < nb; ++l) {
        for (int i = 0; i < S_COUNT; ++i) A.h_red[(size_t)l * S_COUNT + i] = 0.0;
        A.h_red[(size_t)l * S_COUNT + S_DELTA] = 1.0;
        A.h_flag[l] = 1.0;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(A.red, A.h_red, sizeof(double) * S_COUNT * nb, cudaMemcpyHostToDevice, ctx->st));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    auto freeze = [&](int l) -> int {
        active[l] = 0;
        CUDA_TRY(ctx, cudaMemcpyAsync(A.red + (size_t)l * S_COUNT + S_FREEZE, A.h_flag + l, sizeof(double), cudaMemcpyHostToDevice, ctx->st));
        return FANS_OK;
    };
    int n_active = 0;
    for (int l = 0; l < nb; ++l) {
        if (p->n_it > 0 && err[l] > p->tol) n_active++;
        else FANS_CHECK(freeze(l));
    }
    ctx->nb = nb;
    const bool graph_ok = !ctx->prof && iter_graph_wanted(ctx, nb);
    bool graph_tried = false;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_loop0, ctx->st));
    int sweeps = 0;
    while (n_active > 0) {
        // d = s + beta d, K d, <d, K d>, r / u update, norms: one pass each over ALL lanes; replayed as a graph on small grids
        if (graph_ok && iter_graph_ready(ctx, A.graph, R, S, U, KD, D, Dalt)) {
            CUDA_TRY(ctx, cudaGraphLaunch(A.graph.exec[D == A.graph.dA ? 0 : 1], ctx->st));
            ctx->launches += A.graph.launches;
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
        } else {
            FANS_CHECK(linear_iteration(ctx, R, S, U, KD, D, Dalt, true));
            if (graph_ok && !A.graph.failed && !graph_tried) {
                graph_tried = true;
                FANS_CHECK(iter_graph_build(ctx, A.graph, R, S, U, KD, Dalt, D));
            }
        }
        std::swap(D, Dalt);
        sweeps++;
        for (int l = 0; l < nb; ++l) {
            if (!active[l]) continue;
            es[l].iter++;
            err[l] = error_from_scalars(ctx, es[l], l * S_COUNT + S_L1);
            if (p->verbose) printf("lane %d it %3d .... err %16.8e\n", l, es[l].iter, es[l].hist ? es[l].hist[es[l].iter] : err[l]);
            if (!(es[l].iter < p->n_it && err[l] > p->tol)) {
                FANS_CHECK(freeze(l));
                n_active--;
            }
        }
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_loop1, ctx->st));
    ctx->nb = 1;
    A.field[3] = D, A.field[4] = Dalt;
    ctx->n_residual_evals += sweeps * nb;
    // homogenized stress of every lane (solver.h:707-737), lane by lane through the element sweep
    if (stress_out) {
        const double N = (double)ctx->nx * ctx->ny * ctx->nz;
        for (int l = 0; l < nb; ++l) {   // every lane's sums into its own scalar block, one read-back
            for (int i = 0; i < nstr; ++i) ctx->g0[i] = macro[(size_t)l * nstr + i];
            FANS_CHECK(sweep_run(ctx, SWEEP_STRAINSTRESS, U + l * fN, nullptr, nullptr, nullptr, nullptr, ctx->d_red + (size_t)l * S_COUNT + S_STRESS,
                                 nullptr, nullptr));
        }
        ctx->nb = nb;
        FANS_CHECK(read_scalars(ctx));
        ctx->nb = 1;
        for (int l = 0; l < nb; ++l)
            for (int i = 0; i < nstr; ++i) stress_out[(size_t)l * nstr + i] = ctx->h_red[(size_t)l * S_COUNT + S_STRESS + i] / N;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->st));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0.f, lms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    CUDA_TRY(ctx, cudaEventElapsedTime(&lms, ctx->ev_loop0, ctx->ev_loop1));
    prof_resolve(ctx);
    const double fft_ms = conv_time_resolve(ctx);
    for (int l = 0; l < nb; ++l) {   // the device times are those of the whole batch
        res[l].iters = es[l].iter;
        res[l].err_last = err[l];
        res[l].elapsed_ms = ms, res[l].loop_ms = lms, res[l].fft_ms = fft_ms;
        res[l].n_residual_evals = (ctx->n_residual_evals - evals0) / nb;
    }
    return check_fault(ctx);
}

// displacement of lane `lane` of the last batched solve -> a field of the context (then fans_field_download, fans_strain_stress, ...)
extern "C" int fans_batch_load_displacement(fans_ctx *ctx, int32_t lane, int32_t dst_field)
{
    if (!ctx) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (lane < 0 || lane >= ctx->arena.lanes || !ctx->arena.field[0]) {
        fans_set_error(ctx, FANS_ERR_STATE, "fans_batch_load_displacement: no such lane (run fans_solve_batch first)");
        return FANS_ERR_STATE;
    }
    FANS_CHECK(ensure_fields(ctx, {dst_field}));
    const size_t fN = (size_t)ctx->h * ctx->nloc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->field[dst_field], ctx->arena.field[0] + (size_t)lane * fN, sizeof(double) * fN, cudaMemcpyDeviceToDevice, ctx->st));
    return FANS_OK;
}

// releases the lane buffers of batched solves (they are kept between calls; fans_destroy frees them as well)
extern "C" int fans_batch_release(fans_ctx *ctx)
{
    if (!ctx) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    batch_arena_free(ctx);
    return FANS_OK;
}
