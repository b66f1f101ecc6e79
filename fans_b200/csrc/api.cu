// api.cu — C ABI of libfans_gpu (include/fans_gpu.h): context lifetime, problem data, fields, operators.
#include "internal.h"
#include "stencil.h"
#include <cstdlib>
#include <cmath>
#include <algorithm>

static std::string g_create_error;

void fans_set_error(fans_ctx *ctx, int code, const std::string &msg)
{
    (void)code;
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
}

static cudaEvent_t prof_get_event(fans_ctx *ctx)
{
    if (!ctx->prof_pool.empty()) {
        cudaEvent_t e = ctx->prof_pool.back();
        ctx->prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void prof_begin(fans_ctx *ctx, int cls)
{
    if (!ctx->prof) return;
    fans_ctx::ProfRec r{prof_get_event(ctx), prof_get_event(ctx), cls};
    cudaEventRecord(r.a, ctx->st);
    ctx->prof_pending.push_back(r);
}
void prof_end(fans_ctx *ctx)
{
    if (!ctx->prof || ctx->prof_pending.empty()) return;
    cudaEventRecord(ctx->prof_pending.back().b, ctx->st);
}
void prof_resolve(fans_ctx *ctx)
{
    for (auto &r : ctx->prof_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            ctx->prof_ms[r.cls] += ms;
            ctx->prof_n[r.cls] += 1;
        }
        ctx->prof_pool.push_back(r.a);
        ctx->prof_pool.push_back(r.b);
    }
    ctx->prof_pending.clear();
}

static const char *PROF_NAMES[FANS_PROF_CLASSES] = {"fft_z_fwd", "fft_y_fwd", "fft_x_gamma", "fft_y_inv", "fft_z_inv", "sweep_linear",
                                                    "sweep_residual", "sweep_strainstress", "cg_update", "reduce", "axpy", "other",
                                                    "comm_alltoall", "comm_halo", "comm_scalars", ""};

extern "C" int fans_set_profiling(fans_ctx *ctx, int32_t on)
{
    if (!ctx) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    prof_resolve(ctx);
    ctx->prof = on != 0;
    for (int i = 0; i < FANS_PROF_CLASSES; ++i) ctx->prof_ms[i] = 0.0, ctx->prof_n[i] = 0;
    return FANS_OK;
}

extern "C" int fans_get_profile(fans_ctx *ctx, int32_t cls, const char **name, double *ms, int64_t *count)
{
    if (!ctx || cls < 0 || cls >= FANS_PROF_CLASSES) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    prof_resolve(ctx);
    if (name) *name = PROF_NAMES[cls];
    if (ms) *ms = ctx->prof_ms[cls];
    if (count) *count = ctx->prof_n[cls];
    return FANS_OK;
}

extern "C" const char *fans_last_error(const fans_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
extern "C" int fans_version(void) { return 100; }
extern "C" int64_t fans_launch_count(const fans_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------------
// element matrices on the host (tiny): include/matmodel.h:104-188, 237-253, 284-304; LargeStrainMechModel.h:209-225
// ------------------------------------------------------------------------------------------------
static void basic_B(double x, double y, double z, const double le[3], double b[3][8])
{
    const double v0[8] = {-(1 - y) * (1 - z), (1 - y) * (1 - z), -y * (1 - z), y * (1 - z), -(1 - y) * z, (1 - y) * z, -y * z, y * z};
    const double v1[8] = {-(1 - x) * (1 - z), -x * (1 - z), (1 - x) * (1 - z), x * (1 - z), -(1 - x) * z, -x * z, (1 - x) * z, x * z};
    const double v2[8] = {-(1 - x) * (1 - y), -x * (1 - y), -(1 - x) * y, -x * y, (1 - x) * (1 - y), x * (1 - y), (1 - x) * y, x * y};
    for (int a = 0; a < 8; ++a) {
        b[0][a] = v0[a] / le[0];
        b[1][a] = v1[a] / le[1];
        b[2][a] = v2[a] / le[2];
    }
}

// full strain-displacement matrix (n_str x 8h, row-major) from the basic gradient
static void full_B(int nstr, int h, const double b[3][8], std::vector<double> &B)
{
    const int nd = 8 * h;
    B.assign((size_t)nstr * nd, 0.0);
    const double rs = 7.071067811865476e-01;
    for (int q = 0; q < 8; ++q) {
        if (nstr == 3) {
            for (int j = 0; j < 3; ++j) B[j * nd + q] = b[j][q];
        } else if (nstr == 6) {
            B[0 * nd + 3 * q + 0] = b[0][q];
            B[1 * nd + 3 * q + 1] = b[1][q];
            B[2 * nd + 3 * q + 2] = b[2][q];
            B[3 * nd + 3 * q + 0] = rs * b[1][q];
            B[4 * nd + 3 * q + 0] = rs * b[2][q];
            B[5 * nd + 3 * q + 1] = rs * b[2][q];
            B[3 * nd + 3 * q + 1] = rs * b[0][q];
            B[4 * nd + 3 * q + 2] = rs * b[0][q];
            B[5 * nd + 3 * q + 2] = rs * b[1][q];
        } else {
            for (int i = 0; i < 3; ++i)
                for (int J = 0; J < 3; ++J) B[(3 * i + J) * nd + 3 * q + i] = b[J][q];
        }
    }
}

struct ElemOps {
    int ngp, nstr, h, nd;
    std::vector<std::vector<double>> Bint;  // per GP: n_str x 8h
    double vw;
};

static void build_elem(fans_ctx *ctx, ElemOps &E)
{
    E.nstr = ctx->nstr;
    E.h = ctx->h;
    E.nd = 8 * ctx->h;
    E.ngp = ctx->ngp;
    E.vw = ctx->ve / ctx->ngp;
    const double xp = 0.5 + std::sqrt(3.0) / 6.0, xm = 0.5 - std::sqrt(3.0) / 6.0;
    const double xi[8][3] = {{xm, xm, xm}, {xp, xm, xm}, {xm, xp, xm}, {xp, xp, xm}, {xm, xm, xp}, {xp, xm, xp}, {xm, xp, xp}, {xp, xp, xp}};
    ctx->Bgp.assign(9 * 24, 0.0);
    double bc[3][8];
    basic_B(0.5, 0.5, 0.5, ctx->le, bc);
    for (int j = 0; j < 3; ++j)
        for (int a = 0; a < 8; ++a) ctx->Bgp[(8 * 3 + j) * 8 + a] = bc[j][a];
    std::vector<double> Bvol;
    full_B(E.nstr, E.h, bc, Bvol);
    E.Bint.clear();
    for (int g = 0; g < E.ngp; ++g) {
        double b[3][8];
        if (ctx->fe == FANS_FE_HEX8R) basic_B(0.5, 0.5, 0.5, ctx->le, b);
        else basic_B(xi[g][0], xi[g][1], xi[g][2], ctx->le, b);
        for (int j = 0; j < 3; ++j)
            for (int a = 0; a < 8; ++a) ctx->Bgp[(g * 3 + j) * 8 + a] = b[j][a];
        std::vector<double> B;
        full_B(E.nstr, E.h, b, B);
        if (ctx->fe == FANS_FE_BBAR && E.nstr > 3) {  // matmodel.h:113-140
            for (int col = 0; col < E.nd; ++col) {
                const double vf = (B[0 * E.nd + col] + B[1 * E.nd + col] + B[2 * E.nd + col]) / 3.0;
                const double vb = (Bvol[0 * E.nd + col] + Bvol[1 * E.nd + col] + Bvol[2 * E.nd + col]) / 3.0;
                for (int r = 0; r < 3; ++r) B[r * E.nd + col] = B[r * E.nd + col] - vf + vb;
            }
        }
        E.Bint.push_back(B);
    }
}

// K = sum_gp B^T C B v_e/n_gp   (LinearModel::phase_stiffness, e.g. LinearElastic.h:27-40; matmodel.h:237-246)
static void elem_stiffness(const ElemOps &E, const double *C, std::vector<double> &K)
{
    const int nd = E.nd, ns = E.nstr;
    K.assign((size_t)nd * nd, 0.0);
    std::vector<double> CB((size_t)ns * nd);
    for (int g = 0; g < E.ngp; ++g) {
        const std::vector<double> &B = E.Bint[g];
        for (int i = 0; i < ns; ++i)
            for (int c = 0; c < nd; ++c) {
                double s = 0.0;
                for (int j = 0; j < ns; ++j) s += C[i * ns + j] * B[j * nd + c];
                CB[i * nd + c] = s;
            }
        for (int r = 0; r < nd; ++r)
            for (int c = 0; c < nd; ++c) {
                double s = 0.0;
                for (int i = 0; i < ns; ++i) s += B[i * nd + r] * CB[i * nd + c];
                K[r * nd + c] += s * E.vw;
            }
    }
}

// ------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------
static int create_rest(fans_ctx *ctx);
int check_fault(fans_ctx *ctx);
static int ensure_field(fans_ctx *ctx, int f)
{
    if (f < 0 || f >= FANS_N_FIELDS) {
        fans_set_error(ctx, FANS_ERR_ARG, "invalid field id");
        return FANS_ERR_ARG;
    }
    if (!ctx->field[f]) {
        // one extra (always zero) value: the double2 vector passes of vecops.cu round an odd h*nloc up (grids with odd dimensions)
        const size_t bytes = sizeof(double) * ((size_t)ctx->h * ctx->nloc + 1);
        CUDA_TRY(ctx, cudaMalloc(&ctx->field[f], bytes));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->field[f], 0, bytes, ctx->st));
    }
    return FANS_OK;
}

int ensure_dalt(fans_ctx *ctx)
{
    if (!ctx->d_alt) {
        const size_t bytes = sizeof(double) * ((size_t)ctx->h * ctx->nloc + 1);
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_alt, bytes));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_alt, 0, bytes, ctx->st));
    }
    return FANS_OK;
}

int ensure_fields(fans_ctx *ctx, std::initializer_list<int> ids)
{
    for (int f : ids) FANS_CHECK(ensure_field(ctx, f));
    return FANS_OK;
}

extern "C" int fans_create(fans_ctx **out, const fans_config *cfg)
{
    if (!out || !cfg) {
        fans_set_error(nullptr, FANS_ERR_ARG, "null argument");
        return FANS_ERR_ARG;
    }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        fans_set_error(nullptr, FANS_ERR_CUDA, "no CUDA device available: libfans_gpu has no CPU fallback");
        return FANS_ERR_CUDA;
    }
    fans_ctx *ctx = new fans_ctx();
    ctx->cfg = *cfg;
    ctx->nx = cfg->dims[0];
    ctx->ny = cfg->dims[1];
    ctx->nz = cfg->dims[2];
    ctx->h = cfg->howmany;
    ctx->nstr = cfg->n_str;
    ctx->fe = cfg->fe_type;
    ctx->P = cfg->world_size < 1 ? 1 : cfg->world_size;
    ctx->rank = cfg->world_rank;
    ctx->n0 = cfg->local_n0;
    ctx->x0 = cfg->local_0_start;
    ctx->n1 = cfg->local_n1;
    ctx->y1 = cfg->local_1_start;
    auto fail = [&](int code, const std::string &m) {
        fans_set_error(nullptr, code, m);
        delete ctx;
        return code;
    };
    if (!((ctx->h == 1 && ctx->nstr == 3) || (ctx->h == 3 && (ctx->nstr == 6 || ctx->nstr == 9))))
        return fail(FANS_ERR_ARG, "(howmany, n_str) must be (1,3), (3,6) or (3,9)");
    if (ctx->fe < FANS_FE_HEX8 || ctx->fe > FANS_FE_BBAR)
        return fail(FANS_ERR_ARG, "Unknown FE_type. Supported types: HEX8, HEX8R, BBAR");
    bool pow2 = true;
    for (int d = 0; d < 3; ++d) {
        const int n = cfg->dims[d];
        if (n < 4 || n > 2048) return fail(FANS_ERR_ARG, "grid dimensions must lie in [4, 2048] (got " + std::to_string(n) + ")");
        if ((n & (n - 1)) != 0) pow2 = false;
    }
    // any size is accepted like the reference's FFTW plans (src/reader.cpp:300-305): sizes that are not powers of two take the
    // Bluestein passes of fft_any.cu (single GPU); the slab decomposition with its fused transposes exists for 2^k grids only
    // solver.h:391,400: the reference multiplies the frequencies in pairs, (n_y n_x (n_z/2+1)) / 2 of them — with an odd count the last
    // one is left untouched and the result is garbage ("it is important that at least one of the dimensions n_x and n_z is divisible
    // by two").  Such a grid is refused instead of reproducing that.
    if ((((long long)cfg->dims[0] * cfg->dims[1] * (cfg->dims[2] / 2 + 1)) & 1LL) != 0)
        return fail(FANS_ERR_ARG, "n_x * n_y * (n_z/2 + 1) is odd: the reference's convolution skips the last frequency on such a grid "
                                  "(include/solver.h:391-400); choose a grid with an even n_x, n_y or (n_z/2 + 1)");
    ctx->any_fft = !pow2;
    if (ctx->any_fft && ctx->P > 1) return fail(FANS_ERR_ARG, "world_size > 1 needs power-of-two grid dimensions (the fused NVLink transposes are radix-2^k)");
    // slab sizes as fftw_mpi_local_size_many_transposed hands them out for these grids (src/reader.cpp:311-331)
    if (ctx->P > 1 && ((ctx->P & (ctx->P - 1)) != 0 || ctx->nx % ctx->P != 0 || ctx->ny % ctx->P != 0))
        return fail(FANS_ERR_ARG, "world_size must be a power of two dividing n_x and n_y");
    if (ctx->rank < 0 || ctx->rank >= ctx->P) return fail(FANS_ERR_ARG, "world_rank out of range");
    if (ctx->n0 != ctx->nx / ctx->P || ctx->x0 != ctx->rank * ctx->n0 || ctx->n1 != ctx->ny / ctx->P || ctx->y1 != ctx->rank * ctx->n1)
        return fail(FANS_ERR_ARG, "slab sizes must be local_n0 = n_x/P at rank*local_n0 and local_n1 = n_y/P at rank*local_n1");
    if (ctx->P > 1 && ctx->nx / 4 < ctx->P) return fail(FANS_ERR_ARG, "[ FANS3D_Grid ] ERROR: Number of processes too large");  // reader.cpp:306
    ctx->ngp = (ctx->fe == FANS_FE_HEX8R) ? 1 : 8;
    for (int d = 0; d < 3; ++d) {
        ctx->L[d] = cfg->L[d];
        ctx->le[d] = cfg->L[d] / cfg->dims[d];
    }
    ctx->ve = ctx->le[0] * ctx->le[1] * ctx->le[2];
    ctx->nloc = (size_t)ctx->n0 * ctx->ny * ctx->nz;
    for (int i = 0; i < 9; ++i) ctx->g0[i] = 0.0;

    ctx->device = cfg->device;
    if (ctx->device < 0) cudaGetDevice(&ctx->device);
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(FANS_ERR_CUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess) return fail(FANS_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major < 10) return fail(FANS_ERR_CUDA, std::string("device ") + prop.name + " is not sm_100-class; libfans_gpu only carries sm_100a code");
    if (cfg->stream) {
        ctx->st = (cudaStream_t)cfg->stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess) return fail(FANS_ERR_CUDA, "cudaStreamCreate failed");
        ctx->own_stream = true;
    }
    int rc = comm_check(ctx);
    if (rc == FANS_OK) rc = create_rest(ctx);
    if (rc != FANS_OK) {
        g_create_error = ctx->err;
        fans_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return FANS_OK;
}

static int create_rest(fans_ctx *ctx)
{
    CUDA_TRY(ctx, cudaEventCreate(&ctx->ev0));
    CUDA_TRY(ctx, cudaEventCreate(&ctx->ev1));
    CUDA_TRY(ctx, cudaEventCreate(&ctx->ev_loop0));
    CUDA_TRY(ctx, cudaEventCreate(&ctx->ev_loop1));

    ctx->kzc = ctx->nz / 2 + 1;
    ctx->kzp = (ctx->kzc + 7) / 8 * 8;
    ctx->gT = ctx->any_fft ? 4 : fft_x_tile_width(ctx->nx, ctx->h);
    ctx->gE = ctx->any_fft ? 1 : (ctx->nx < 8 ? ctx->nx : 8);
    ctx->yT = 8;  // 128-byte rows: 512^3: 1.04 ms vs 1.39 ms with T=4; n_y = 1024 over NVLink (2 GPUs): 2.35 / 2.51 ms vs 3.69 / 3.00 ms
    if (const char *e = getenv("FANS_YT")) ctx->yT = (atoi(e) == 4) ? 4 : 8;
    if (ctx->any_fft) {
        FANS_CHECK(any_plan_init(ctx, ctx->anyx, ctx->nx));
        FANS_CHECK(any_plan_init(ctx, ctx->anyy, ctx->ny));
        FANS_CHECK(any_plan_init(ctx, ctx->anyz, ctx->nz));
        ctx->planx.pos_host.resize(ctx->nx);   // natural frequency order along x and y
        ctx->plany.pos_host.resize(ctx->ny);
        for (int f = 0; f < ctx->nx; ++f) ctx->planx.pos_host[f] = f;
        for (int f = 0; f < ctx->ny; ++f) ctx->plany.pos_host[f] = f;
    } else {
        FANS_CHECK(fft_plan_init(ctx, ctx->planx, ctx->nx, ctx->nx));
        FANS_CHECK(fft_plan_init(ctx, ctx->plany, ctx->ny, ctx->ny));
        FANS_CHECK(fft_plan_init(ctx, ctx->planz, ctx->nz / 2, ctx->nz));
    }
    if (const char *env = getenv("FANS_XPAD")) ctx->xpad = atoi(env);
    const size_t spec_elems = (size_t)ctx->P * ctx->h * ctx->n0 * ((size_t)ctx->n1 * ctx->kzp + ctx->xpad);
    CUDA_TRY(ctx, cudaMalloc(&ctx->spec, sizeof(double2) * spec_elems));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->spec, 0, sizeof(double2) * spec_elems, ctx->st));
    if (ctx->P > 1) {  // the transposed spectrum: this rank's y rows for all x
        CUDA_TRY(ctx, cudaMalloc(&ctx->specB, sizeof(double2) * spec_elems));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->specB, 0, sizeof(double2) * spec_elems, ctx->st));
        CUDA_TRY(ctx, cudaMalloc(&ctx->ms_lo, sizeof(uint16_t) * ctx->ny * ctx->nz));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    FANS_CHECK(comm_map_peers(ctx));
    if (ctx->P > 1 && ctx->p2p && ctx->h > 1) {  // component pipeline of the slab convolution (solve.cu, conv_run_pipelined)
        int lo = 0, hi = 0;
        CUDA_TRY(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(ctx, cudaStreamCreateWithPriority(&ctx->st2, cudaStreamNonBlocking, hi));
        for (int i = 0; i < 8; ++i) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_pipe[i], cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_gate, sizeof(int)));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_gate, 0, sizeof(int), ctx->st));
        ctx->pipe = 1;
        ctx->y_grid = 96;  // SMs given to an NVLink-bound y pass while a z pass runs beside it (4 GPUs, 512x1024x1024: 64 -> 21.16, 96 -> 20.54, 120 -> 20.98, no pipeline 21.83 ms/iteration)
        if (const char *e = getenv("FANS_PIPE")) ctx->pipe = atoi(e) ? 1 : 0;
        if (const char *e = getenv("FANS_Y_GRID")) ctx->y_grid = atoi(e);
        // kz-chunked pipeline (the x pass of chunk q under the NVLink transposes of the neighbouring chunks): opt-in, measured slower
        // than the component pipeline at 2 and 8 GPUs and equal at 4 (profiles/r2r_chunked_pipeline_8gpu.txt) — the x pass starves on
        // the SMs the persistent y pass leaves, and the z passes lose their overlap
        ctx->chunks = 0;
        if (const char *e = getenv("FANS_CHUNKS")) ctx->chunks = (ctx->nx >= 64) ? std::min(4, std::max(0, atoi(e))) : 0;
    }

    CUDA_TRY(ctx, cudaMalloc(&ctx->d_part, sizeof(double) * (1 << 20)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_red, sizeof(double) * S_COUNT));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_red, 0, sizeof(double) * S_COUNT, ctx->st));
    CUDA_TRY(ctx, cudaMallocHost(&ctx->h_red, sizeof(double) * S_COUNT));
    CUDA_TRY(ctx, cudaMallocHost(&ctx->h_stage, sizeof(double) * 4));
    CUDA_TRY(ctx, cudaMallocHost(&ctx->h_fault, sizeof(int)));
    *ctx->h_fault = 0;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_ticket, 2 * sizeof(unsigned int)));   // [0] grid-reduce ticket, [1] scratch word (phase-id maximum)
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_ticket, 0, 2 * sizeof(unsigned int), ctx->st));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_flag, sizeof(int)));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->st));
    FANS_CHECK(ensure_fields(ctx, {FANS_FIELD_U, FANS_FIELD_R, FANS_FIELD_U_PREV}));
    ElemOps E;
    build_elem(ctx, E);  // fills ctx->Bgp
    ctx->const_stamp = sweep_new_stamp();
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    return FANS_OK;
}

extern "C" void fans_destroy(fans_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->st) cudaStreamSynchronize(ctx->st);
    if (ctx->P > 1 && ctx->p2p) {
        comm_barrier(ctx);  // nobody may still be storing into this rank's spectrum when it is freed
        cudaStreamSynchronize(ctx->st);
        comm_unmap_peers(ctx);
    }
    for (int f = 0; f < FANS_N_FIELDS; ++f)
        if (ctx->field[f]) cudaFree(ctx->field[f]);
    void *ptrs[] = {ctx->specB, ctx->ms_lo, ctx->halo_send_lo, ctx->halo_send_hi, ctx->halo_lo, ctx->halo_hi, ctx->d_alt, ctx->stage_io, ctx->ms, ctx->phidx, ctx->spec, ctx->gamma, ctx->d_phase, ctx->d_K, ctx->phase_lut,
                    ctx->hist, ctx->hist_t, ctx->hidx, ctx->pflag, ctx->d_Stab, ctx->d_part, ctx->d_red, ctx->d_ticket, ctx->d_flag, ctx->d_C};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    batch_arena_free(ctx);
    iter_graph_free(ctx);
    if (ctx->h_red) cudaFreeHost(ctx->h_red);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_fault) cudaFreeHost(ctx->h_fault);
    fft_plan_free(ctx->planx);
    fft_plan_free(ctx->plany);
    fft_plan_free(ctx->planz);
    any_plan_free(ctx->anyx);
    any_plan_free(ctx->anyy);
    any_plan_free(ctx->anyz);
    prof_resolve(ctx);
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    for (auto &pr : ctx->conv_pending) cudaEventDestroy(pr.first), cudaEventDestroy(pr.second);
    for (cudaEvent_t e : ctx->conv_pool) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_loop0) cudaEventDestroy(ctx->ev_loop0);
    if (ctx->ev_loop1) cudaEventDestroy(ctx->ev_loop1);
    for (int i = 0; i < 8; ++i)
        if (ctx->ev_pipe[i]) cudaEventDestroy(ctx->ev_pipe[i]);
    if (ctx->st2) cudaStreamDestroy(ctx->st2);
    if (ctx->d_gate) cudaFree(ctx->d_gate);
    if (ctx->own_stream && ctx->st) cudaStreamDestroy(ctx->st);
    delete ctx;
}

// ------------------------------------------------------------------------------------------------
// problem data
// ------------------------------------------------------------------------------------------------
// largest phase id of the image (validated against the material table, MaterialManager.h:232-235), found on the device: the host
// never walks the 134 M voxels of a 512^3 slab
__global__ void k_u16_max(const uint16_t *__restrict__ a, size_t n, unsigned int *out)
{
    unsigned int m = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = max(m, (unsigned int)a[i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

extern "C" int fans_set_microstructure(fans_ctx *ctx, const uint16_t *ms)
{
    if (!ctx || !ms) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (!ctx->phidx) CUDA_TRY(ctx, cudaMalloc(&ctx->phidx, sizeof(uint16_t) * ctx->nloc));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->phidx, ms, sizeof(uint16_t) * ctx->nloc, cudaMemcpyHostToDevice, ctx->st));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_ticket + 1, 0, sizeof(unsigned int), ctx->st));
    k_u16_max<<<FANS_SMS * 8, 256, 0, ctx->st>>>(ctx->phidx, ctx->nloc, ctx->d_ticket + 1);
    ctx->launches++;
    unsigned int mx32 = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&mx32, ctx->d_ticket + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    const uint16_t mx = (uint16_t)mx32;
    ctx->ms_max = mx;
    if (ctx->materials_ready && (int)mx >= ctx->n_phases) {
        ctx->ms_ready = false;
        fans_set_error(ctx, FANS_ERR_MATERIAL, "MaterialManager: Phase " + std::to_string(mx) + " not assigned");
        return FANS_ERR_MATERIAL;
    }
    if (ctx->P > 1) {  // phases of element plane -1 (the previous rank's last plane) for the gather-form stencil
        const size_t plane = (size_t)ctx->ny * ctx->nz;
        FANS_CHECK(comm_halo(ctx, nullptr, nullptr, ctx->phidx + (size_t)(ctx->n0 - 1) * plane, ctx->ms_lo, sizeof(uint16_t) * plane));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    ctx->ms_ready = true;
    ctx->hist_ready = false;  // the compact history index follows the phase image
    return FANS_OK;
}

extern "C" int fans_set_materials(fans_ctx *ctx, int32_t n_phases, const fans_phase_desc *ph)
{
    if (!ctx || !ph || n_phases < 1 || n_phases > 65536) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (ctx->ms_ready && (int)ctx->ms_max >= n_phases) {
        fans_set_error(ctx, FANS_ERR_MATERIAL, "MaterialManager: Phase " + std::to_string(ctx->ms_max) + " not assigned");
        return FANS_ERR_MATERIAL;
    }
    ElemOps E;
    build_elem(ctx, E);
    const int nd = E.nd, ns = ctx->nstr;
    std::vector<PhaseDev> dev(n_phases);
    std::vector<double> Ktab, Ctab;
    int nk = 0, nhist = 0;
    bool all_lin = true, any_flag = false;
    for (int i = 0; i < n_phases; ++i) {
        PhaseDev &d = dev[i];
        memset(&d, 0, sizeof(d));
        d.model = ph[i].model;
        d.local_mat = ph[i].local_mat;
        d.group_n_mat = ph[i].group_n_mat;
        d.k_index = -1;
        int npar = 0;
        bool ok = true;
        switch (ph[i].model) {
        case FANS_MAT_LINEAR: {
            d.k_index = nk++;
            std::vector<double> K;
            elem_stiffness(E, ph[i].params, K);
            Ktab.insert(Ktab.end(), K.begin(), K.end());
            Ctab.insert(Ctab.end(), ph[i].params, ph[i].params + ns * ns);
            // exactly isotropic tangent (what LinearElasticIsotropic / LinearThermalIsotropic build, LinearElastic.h:33-34,
            // LinearThermal.h:24-31): the Gauss-point law then is the reference's own closed form instead of a dense C.eps
            const double *Cm = ph[i].params;
            bool iso = (ns == 6 || ns == 3);
            if (ns == 6) {
                const double lam = Cm[1], mu2 = Cm[3 * 6 + 3];
                for (int r = 0; r < 6 && iso; ++r)
                    for (int c = 0; c < 6; ++c) {
                        const double want = (r < 3 && c < 3 ? lam : 0.0) + (r == c ? mu2 : 0.0);
                        if (Cm[r * 6 + c] != want) iso = false;
                    }
                d.params[0] = lam, d.params[1] = mu2;
            } else if (ns == 3) {
                for (int r = 0; r < 3 && iso; ++r)
                    for (int c = 0; c < 3; ++c)
                        if (Cm[r * 3 + c] != (r == c ? Cm[0] : 0.0)) iso = false;
                d.params[0] = Cm[0];
            }
            d.lin_iso = iso ? 1 : 0;
            break;
        }
        case FANS_MAT_PSEUDOPLASTIC_LINEAR:
        case FANS_MAT_PSEUDOPLASTIC_NONLIN:
            npar = 6, any_flag = true, ok = (ns == 6);
            break;
        case FANS_MAT_J2_LINEAR_ISO:
            npar = 7, nhist = std::max(nhist, 13), ok = (ns == 6);
            break;
        case FANS_MAT_J2_NONLIN_ISO:
            npar = 9, nhist = std::max(nhist, 13), ok = (ns == 6);
            break;
        case FANS_MAT_J2NEW_LINEAR_ISO:
            npar = 4, nhist = std::max(nhist, 7), ok = (ns == 6);
            break;
        case FANS_MAT_SVK:
        case FANS_MAT_NEOHOOKE:
            npar = 2, ok = (ns == 9);
            break;
        default:
            ok = false;
        }
        if (!ok) {
            fans_set_error(ctx, FANS_ERR_MATERIAL, "material model id " + std::to_string(ph[i].model) + " is not valid for n_str = " + std::to_string(ns));
            return FANS_ERR_MATERIAL;
        }
        for (int k = 0; k < npar; ++k) d.params[k] = ph[i].params[k];
        d.has_hist = (ph[i].model == FANS_MAT_J2_LINEAR_ISO || ph[i].model == FANS_MAT_J2_NONLIN_ISO || ph[i].model == FANS_MAT_J2NEW_LINEAR_ISO);
        if (ph[i].model != FANS_MAT_LINEAR) all_lin = false;
    }
    ctx->phases.assign(ph, ph + n_phases);
    ctx->K_host = Ktab;
    ctx->n_phases = n_phases;
    ctx->n_k = nk;
    ctx->all_linear = all_lin;
    ctx->any_history = nhist > 0;
    ctx->any_flag = any_flag;
    if (ctx->d_K) cudaFree(ctx->d_K), ctx->d_K = nullptr;
    if (ctx->d_C) cudaFree(ctx->d_C), ctx->d_C = nullptr;
    if (nk > 0) {
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_K, sizeof(double) * Ktab.size()));
        CUDA_TRY(ctx, cudaMemcpy(ctx->d_K, Ktab.data(), sizeof(double) * Ktab.size(), cudaMemcpyHostToDevice));
        CUDA_TRY(ctx, cudaMalloc(&ctx->d_C, sizeof(double) * Ctab.size()));
        CUDA_TRY(ctx, cudaMemcpy(ctx->d_C, Ctab.data(), sizeof(double) * Ctab.size(), cudaMemcpyHostToDevice));
        for (int i = 0; i < n_phases; ++i)
            if (dev[i].k_index >= 0) dev[i].tangent = ctx->d_C + (size_t)dev[i].k_index * ns * ns;
    }
    if (ctx->d_phase) cudaFree(ctx->d_phase), ctx->d_phase = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_phase, sizeof(PhaseDev) * n_phases));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_phase, dev.data(), sizeof(PhaseDev) * n_phases, cudaMemcpyHostToDevice));
    // internal variables: MaterialManager::initialize_internal_variables (solver.h:139).  The reference allocates them for EVERY
    // element of every model (J2Plasticity.h:47-56); here only the elements whose phase reads them get storage (history_prepare)
    if (ctx->hist) cudaFree(ctx->hist), ctx->hist = nullptr;
    if (ctx->hist_t) cudaFree(ctx->hist_t), ctx->hist_t = nullptr;
    if (ctx->pflag) cudaFree(ctx->pflag), ctx->pflag = nullptr;
    ctx->n_hist = nhist;
    ctx->hist_ready = false;
    ctx->has_hist_host.assign(n_phases, 0);
    for (int i = 0; i < n_phases; ++i) ctx->has_hist_host[i] = (uint8_t)dev[i].has_hist;
    if (any_flag) {
        CUDA_TRY(ctx, cudaMalloc(&ctx->pflag, sizeof(int) * ctx->ngp * ctx->nloc));
        CUDA_TRY(ctx, cudaMemset(ctx->pflag, 0, sizeof(int) * ctx->ngp * ctx->nloc));
    }
    ctx->const_stamp = sweep_new_stamp();
    ctx->materials_ready = true;
    return FANS_OK;
}

extern "C" int fans_set_reference_stiffness(fans_ctx *ctx, const double *kapparef)
{
    if (!ctx || !kapparef) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    ElemOps E;
    build_elem(ctx, E);
    const int nd = E.nd, h = ctx->h;
    for (int i = 0; i < ctx->nstr * ctx->nstr; ++i) ctx->kapparef[i] = kapparef[i];
    std::vector<double> K, Ker0((size_t)nd * nd);
    elem_stiffness(E, kapparef, K);
    // component-major reordering, matmodel.h:247-251
    for (int i = 0; i < nd; ++i)
        for (int j = 0; j < nd; ++j) Ker0[((i % h) * 8 + i / h) * nd + (j % h) * 8 + j / h] = K[i * nd + j];
    double *dK = nullptr;
    int *frqx = nullptr, *frqy = nullptr;
    std::vector<int> fx(ctx->nx), fy(ctx->ny);
    for (int f = 0; f < ctx->nx; ++f) fx[ctx->planx.pos_host[f]] = f;
    for (int f = 0; f < ctx->ny; ++f) fy[ctx->plany.pos_host[f]] = f;
    CUDA_TRY(ctx, cudaMalloc(&dK, sizeof(double) * nd * nd));
    CUDA_TRY(ctx, cudaMalloc(&frqx, sizeof(int) * ctx->nx));
    CUDA_TRY(ctx, cudaMalloc(&frqy, sizeof(int) * ctx->ny));
    CUDA_TRY(ctx, cudaMemcpy(dK, Ker0.data(), sizeof(double) * nd * nd, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(frqx, fx.data(), sizeof(int) * ctx->nx, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(frqy, fy.data(), sizeof(int) * ctx->ny, cudaMemcpyHostToDevice));
    const int T = ctx->gT, nTiles = (ctx->kzc + T - 1) / T, NG = h * (h + 1) / 2;
    if (!ctx->gamma) CUDA_TRY(ctx, cudaMalloc(&ctx->gamma, sizeof(double) * (size_t)ctx->n1 * nTiles * NG * ctx->nx * T));
    int rc = gamma_build(ctx, dK, frqx, frqy);
    cudaStreamSynchronize(ctx->st);
    cudaFree(dK);
    cudaFree(frqx);
    cudaFree(frqy);
    if (rc == FANS_OK) ctx->gamma_ready = true;
    return rc;
}

extern "C" int fans_set_gradient(fans_ctx *ctx, const double *g0)
{
    if (!ctx || !g0) return FANS_ERR_ARG;
    for (int i = 0; i < ctx->nstr; ++i) ctx->g0[i] = g0[i];
    return FANS_OK;
}

extern "C" int fans_get_gradient(fans_ctx *ctx, double *g0)
{
    if (!ctx || !g0) return FANS_ERR_ARG;
    for (int i = 0; i < ctx->nstr; ++i) g0[i] = ctx->g0[i];
    return FANS_OK;
}

// ------------------------------------------------------------------------------------------------
// compact history storage: hidx[element] = running index over the elements whose phase carries history (memory order), so the
// history arrays of BASELINE config 3 (J2 matrix, elastic fibres, 512^3) shrink with the J2 volume fraction and stay coalesced
// ------------------------------------------------------------------------------------------------
__global__ void k_hist_rowcount(const uint16_t *__restrict__ ph, const uint8_t *__restrict__ has, int nz, size_t nrows, unsigned *cnt)
{
    const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int lane = threadIdx.x & 31;
    unsigned n = 0;
    for (int z = lane; z < nz; z += 32) n += has[ph[row * nz + z]];
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if (lane == 0) cnt[row] = n;
}
__global__ void k_hist_fill(const uint16_t *__restrict__ ph, const uint8_t *__restrict__ has, int nz, size_t nrows,
                            const unsigned *__restrict__ rowoff, unsigned *__restrict__ hidx)
{
    const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int lane = threadIdx.x & 31;
    unsigned off = rowoff[row];
    for (int z0 = 0; z0 < nz; z0 += 32) {
        const int z = z0 + lane;
        const bool f = (z < nz) && has[ph[row * nz + z]];
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (z < nz) hidx[row * nz + z] = f ? off + __popc(m & ((1u << lane) - 1u)) : 0xffffffffu;
        off += __popc(m);
    }
}

int history_prepare(fans_ctx *ctx)
{
    if (ctx->hist_ready || ctx->n_hist == 0) return FANS_OK;
    if (!ctx->ms_ready || !ctx->materials_ready) {
        fans_set_error(ctx, FANS_ERR_STATE, "microstructure and materials must be set before the history variables exist");
        return FANS_ERR_STATE;
    }
    if (ctx->hist) cudaFree(ctx->hist), ctx->hist = nullptr;
    if (ctx->hist_t) cudaFree(ctx->hist_t), ctx->hist_t = nullptr;
    const size_t nrows = (size_t)ctx->n0 * ctx->ny;
    uint8_t *d_has = nullptr;
    unsigned *d_cnt = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&d_has, ctx->n_phases));
    CUDA_TRY(ctx, cudaMalloc(&d_cnt, sizeof(unsigned) * nrows));
    if (!ctx->hidx) CUDA_TRY(ctx, cudaMalloc(&ctx->hidx, sizeof(unsigned) * ctx->nloc));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_has, ctx->has_hist_host.data(), ctx->n_phases, cudaMemcpyHostToDevice, ctx->st));
    const unsigned nb = (unsigned)((nrows + 7) / 8);
    k_hist_rowcount<<<nb, 256, 0, ctx->st>>>(ctx->phidx, d_has, ctx->nz, nrows, d_cnt);
    std::vector<unsigned> cnt(nrows);
    CUDA_TRY(ctx, cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(unsigned) * nrows, cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    size_t run = 0;
    for (size_t r = 0; r < nrows; ++r) {
        const unsigned c = cnt[r];
        cnt[r] = (unsigned)run;
        run += c;
    }
    ctx->nh = run;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_cnt, cnt.data(), sizeof(unsigned) * nrows, cudaMemcpyHostToDevice, ctx->st));
    k_hist_fill<<<nb, 256, 0, ctx->st>>>(ctx->phidx, d_has, ctx->nz, nrows, d_cnt, ctx->hidx);
    ctx->launches += 2;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    cudaFree(d_has);
    cudaFree(d_cnt);
    CUDA_TRY(ctx, cudaGetLastError());
    const size_t bytes = sizeof(double) * ctx->n_hist * ctx->ngp * (ctx->nh ? ctx->nh : 1);
    size_t fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    if (2 * bytes > fr) {
        fans_set_error(ctx, FANS_ERR_CUDA, "history variables need " + std::to_string(2 * bytes >> 20) + " MiB, only " + std::to_string(fr >> 20) + " MiB free");
        return FANS_ERR_CUDA;
    }
    CUDA_TRY(ctx, cudaMalloc(&ctx->hist, bytes));
    CUDA_TRY(ctx, cudaMalloc(&ctx->hist_t, bytes));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->hist, 0, bytes, ctx->st));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->hist_t, 0, bytes, ctx->st));
    ctx->hist_ready = true;
    return FANS_OK;
}

// history variables [first, first + cnt) of the committed state, expanded to every element (zero where the phase has none):
// out[(v * ngp + g) * cnt + k]
static int history_expand(fans_ctx *ctx, int first, int cnt, std::vector<double> &out)
{
    FANS_CHECK(history_prepare(ctx));
    const size_t N = ctx->nloc, nh = ctx->nh;
    const int ngp = ctx->ngp;
    std::vector<double> tmp((size_t)cnt * ngp * (nh ? nh : 1));
    std::vector<unsigned> idx(N);
    if (nh) CUDA_TRY(ctx, cudaMemcpyAsync(tmp.data(), ctx->hist_t + (size_t)first * ngp * nh, sizeof(double) * cnt * ngp * nh, cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(ctx, cudaMemcpyAsync(idx.data(), ctx->hidx, sizeof(unsigned) * N, cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    out.assign((size_t)cnt * ngp * N, 0.0);
    for (size_t v = 0; v < N; ++v) {
        const unsigned he = idx[v];
        if (he == 0xffffffffu) continue;
        for (int g = 0; g < ngp; ++g)
            for (int k = 0; k < cnt; ++k) out[(v * ngp + g) * cnt + k] = tmp[((size_t)k * ngp + g) * nh + he];
    }
    return FANS_OK;
}

// ------------------------------------------------------------------------------------------------
// fields
// ------------------------------------------------------------------------------------------------
static int ensure_stage(fans_ctx *ctx)
{
    if (!ctx->stage_io) CUDA_TRY(ctx, cudaMalloc(&ctx->stage_io, sizeof(double) * ctx->h * ctx->nloc));
    return FANS_OK;
}

extern "C" int fans_field_upload(fans_ctx *ctx, int32_t f, const double *host)
{
    if (!ctx || !host) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    FANS_CHECK(ensure_field(ctx, f));
    const size_t bytes = sizeof(double) * ctx->h * ctx->nloc;
    if (ctx->h == 1) {
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->field[f], host, bytes, cudaMemcpyHostToDevice, ctx->st));
    } else {
        FANS_CHECK(ensure_stage(ctx));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->stage_io, host, bytes, cudaMemcpyHostToDevice, ctx->st));
        FANS_CHECK(vec_aos_to_soa(ctx, ctx->stage_io, ctx->field[f]));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    return FANS_OK;
}

extern "C" int fans_field_download(fans_ctx *ctx, int32_t f, double *host)
{
    if (!ctx || !host) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    FANS_CHECK(ensure_field(ctx, f));
    const size_t bytes = sizeof(double) * ctx->h * ctx->nloc;
    if (ctx->h == 1) {
        CUDA_TRY(ctx, cudaMemcpyAsync(host, ctx->field[f], bytes, cudaMemcpyDeviceToHost, ctx->st));
    } else {
        FANS_CHECK(ensure_stage(ctx));
        FANS_CHECK(vec_soa_to_aos(ctx, ctx->field[f], ctx->stage_io));
        CUDA_TRY(ctx, cudaMemcpyAsync(host, ctx->stage_io, bytes, cudaMemcpyDeviceToHost, ctx->st));
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    return FANS_OK;
}

extern "C" int fans_field_zero(fans_ctx *ctx, int32_t f)
{
    if (!ctx) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    FANS_CHECK(ensure_field(ctx, f));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->field[f], 0, sizeof(double) * ctx->h * ctx->nloc, ctx->st));
    return FANS_OK;
}

extern "C" int fans_field_copy(fans_ctx *ctx, int32_t dst, int32_t src)
{
    if (!ctx) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    FANS_CHECK(ensure_fields(ctx, {dst, src}));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->field[dst], ctx->field[src], sizeof(double) * ctx->h * ctx->nloc, cudaMemcpyDeviceToDevice, ctx->st));
    return FANS_OK;
}

// ------------------------------------------------------------------------------------------------
// operators
// ------------------------------------------------------------------------------------------------
// the fault flag as of the last read_scalars() (same stream position as the scalars it returned); slabs: all-reduce first
int check_fault_cached(fans_ctx *ctx)
{
    if (ctx->P > 1) return check_fault(ctx);
    if (*ctx->h_fault == FANS_ERR_NEG_JACOBIAN) {
        fans_set_error(ctx, FANS_ERR_NEG_JACOBIAN, "Negative Jacobian determinant in CompressibleNeoHookean!");
        cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->st);
        *ctx->h_fault = 0;
        return FANS_ERR_NEG_JACOBIAN;
    }
    return FANS_OK;
}

int check_fault(fans_ctx *ctx)
{
    int f = 0;
    FANS_CHECK(comm_allreduce_int_max(ctx, ctx->d_flag));   // all slabs see the fault of any slab (every rank calls check_fault at the same points)
    CUDA_TRY(ctx, cudaMemcpyAsync(&f, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->st));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    if (f == FANS_ERR_NEG_JACOBIAN) {
        fans_set_error(ctx, f, "Negative Jacobian determinant in CompressibleNeoHookean!");
        cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->st);
        return f;
    }
    return FANS_OK;
}

extern "C" int fans_residual(fans_ctx *ctx, int32_t fo, int32_t fu)
{
    if (!ctx) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    FANS_CHECK(ensure_fields(ctx, {fo, fu}));
    if (fo == fu) {
        fans_set_error(ctx, FANS_ERR_ARG, "residual: output must differ from input");
        return FANS_ERR_ARG;
    }
    ctx->n_residual_evals++;
    FANS_CHECK(sweep_run(ctx, SWEEP_RESIDUAL, ctx->field[fu], ctx->field[fo], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
    return check_fault(ctx);
}

extern "C" int fans_apply_linear(fans_ctx *ctx, int32_t fo, int32_t fd)
{
    if (!ctx) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    FANS_CHECK(ensure_fields(ctx, {fo, fd}));
    if (fo == fd) {
        fans_set_error(ctx, FANS_ERR_ARG, "apply_linear: output must differ from input");
        return FANS_ERR_ARG;
    }
    if (!ctx->all_linear) {
        fans_set_error(ctx, FANS_ERR_STATE, "apply_linear requires all phases to be LinearModel (MaterialManager::all_linear)");
        return FANS_ERR_STATE;
    }
    ctx->n_residual_evals++;
    if (stencil_supported(ctx) && !getenv("FANS_LINEAR_SWEEP"))
        FANS_CHECK(stencil_run(ctx, ctx->field[fd], ctx->field[fo], nullptr, nullptr, nullptr, nullptr));
    else
        FANS_CHECK(sweep_run(ctx, SWEEP_LINEAR, ctx->field[fd], ctx->field[fo], nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    return FANS_OK;
}

extern "C" int fans_convolution(fans_ctx *ctx, int32_t fi, int32_t fo)
{
    if (!ctx) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    FANS_CHECK(ensure_fields(ctx, {fi, fo}));
    FANS_CHECK(conv_run(ctx, ctx->field[fi], ctx->field[fo], 1.0, nullptr, nullptr));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
    return FANS_OK;
}

extern "C" int fans_dot(fans_ctx *ctx, int32_t a, int32_t b, double *out)
{
    if (!ctx || !out) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    FANS_CHECK(ensure_fields(ctx, {a, b}));
    FANS_CHECK(vec_reduce4(ctx, ctx->field[a], ctx->field[b], ctx->d_red + S_GEN));
    FANS_CHECK(read_scalars(ctx));
    *out = ctx->h_red[S_GEN + 2];
    return FANS_OK;
}

extern "C" int fans_axpy(fans_ctx *ctx, int32_t y, double alpha, int32_t x)
{
    if (!ctx) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    FANS_CHECK(ensure_fields(ctx, {y, x}));
    return vec_axpy(ctx, ctx->field[y], alpha, ctx->field[x]);
}

extern "C" int fans_norm(fans_ctx *ctx, int32_t f, int32_t measure, double *out)
{
    if (!ctx || !out) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    FANS_CHECK(ensure_field(ctx, f));
    FANS_CHECK(vec_reduce4(ctx, ctx->field[f], nullptr, ctx->d_red + S_GEN));
    FANS_CHECK(read_scalars(ctx));
    if (measure == FANS_MEASURE_L1) *out = ctx->h_red[S_GENMAX];
    else if (measure == FANS_MEASURE_L2) *out = std::sqrt(ctx->h_red[S_GENMAX + 1]);
    else if (measure == FANS_MEASURE_LINF) *out = ctx->h_red[S_GENMAX + 3];
    else {
        fans_set_error(ctx, FANS_ERR_ARG, "Unknown measure type");
        return FANS_ERR_ARG;
    }
    return FANS_OK;
}

extern "C" int fans_commit_history(fans_ctx *ctx)
{
    if (!ctx) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    if (ctx->n_hist > 0) {
        FANS_CHECK(history_prepare(ctx));
        if (ctx->nh) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hist_t, ctx->hist, sizeof(double) * ctx->n_hist * ctx->ngp * ctx->nh, cudaMemcpyDeviceToDevice, ctx->st));
    }
    return FANS_OK;
}

extern "C" int fans_extrapolate_displacement(fans_ctx *ctx)
{
    if (!ctx) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    FANS_CHECK(ensure_fields(ctx, {FANS_FIELD_U, FANS_FIELD_U_PREV}));
    return vec_extrapolate(ctx, ctx->field[FANS_FIELD_U], ctx->field[FANS_FIELD_U_PREV]);
}

// ------------------------------------------------------------------------------------------------
// postprocess data sources
// ------------------------------------------------------------------------------------------------
// ONE getStrainStress sweep (Solver::postprocess, solver.h:497-545): element averages [x][y][z][n_str] and / or all Gauss-point values
// [x][y][z][n_gp][n_str]; any pointer may be NULL.  The material law runs once per Gauss point, history side effects included.
extern "C" int fans_strain_stress_gp(fans_ctx *ctx, double *strain_host, double *stress_host, double *strain_gp_host, double *stress_gp_host)
{
    if (!ctx) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    const size_t N = ctx->nloc;
    const int ns = ctx->nstr, ngp = ctx->ngp;
    FANS_CHECK(ensure_field(ctx, FANS_FIELD_U));
    double *host[4] = {strain_host, stress_host, strain_gp_host, stress_gp_host};
    double *dev[4] = {nullptr, nullptr, nullptr, nullptr};
    const size_t cnt[4] = {(size_t)ns * N, (size_t)ns * N, (size_t)ns * ngp * N, (size_t)ns * ngp * N};
    auto release = [&]() {
        for (double *d : dev)
            if (d) cudaFree(d);
    };
    for (int k = 0; k < 4; ++k)
        if (host[k] && cudaMalloc(&dev[k], sizeof(double) * cnt[k]) != cudaSuccess) {
            release();
            fans_set_error(ctx, FANS_ERR_CUDA, "fans_strain_stress: out of device memory");
            return FANS_ERR_CUDA;
        }
    int rc = sweep_run(ctx, SWEEP_STRAINSTRESS, ctx->field[FANS_FIELD_U], nullptr, nullptr, nullptr, nullptr, nullptr, dev[0], dev[1], dev[2], dev[3]);
    if (rc == FANS_OK) {
        std::vector<double> tmp;
        for (int k = 0; k < 4; ++k) {
            if (!host[k]) continue;
            tmp.resize(cnt[k]);
            cudaMemcpyAsync(tmp.data(), dev[k], sizeof(double) * cnt[k], cudaMemcpyDeviceToHost, ctx->st);
            cudaStreamSynchronize(ctx->st);
            const size_t per = cnt[k] / N;   // device order [n_str][ngp][element] -> host [element][ngp][n_str]
            const int g_n = (int)(per / ns);
            for (int i = 0; i < ns; ++i)
                for (int g = 0; g < g_n; ++g) {
                    const double *src = tmp.data() + ((size_t)i * g_n + g) * N;
                    for (size_t v = 0; v < N; ++v) host[k][(v * g_n + g) * ns + i] = src[v];
                }
        }
        rc = check_fault(ctx);
    }
    release();
    return rc;
}

extern "C" int fans_strain_stress(fans_ctx *ctx, double *strain_host, double *stress_host)
{
    return fans_strain_stress_gp(ctx, strain_host, stress_host, nullptr, nullptr);
}

extern "C" int fans_get_field(fans_ctx *ctx, const char *name, void *dst, size_t bytes)
{
    if (!ctx || !name || !dst) return FANS_ERR_ARG;
    cudaSetDevice(ctx->device);
    const std::string n(name);
    const size_t N = ctx->nloc;
    auto need = [&](size_t b) {
        if (bytes < b) {
            fans_set_error(ctx, FANS_ERR_ARG, "fans_get_field(" + n + "): buffer too small, need " + std::to_string(b) + " bytes");
            return false;
        }
        return true;
    };
    if (n == "strain" || n == "stress") {
        if (!need(sizeof(double) * ctx->nstr * N)) return FANS_ERR_ARG;
        return n == "strain" ? fans_strain_stress(ctx, (double *)dst, nullptr) : fans_strain_stress(ctx, nullptr, (double *)dst);
    }
    if (n == "strain_gp" || n == "stress_gp") {
        if (!need(sizeof(double) * ctx->nstr * ctx->ngp * N)) return FANS_ERR_ARG;
        return n == "strain_gp" ? fans_strain_stress_gp(ctx, nullptr, nullptr, (double *)dst, nullptr)
                                : fans_strain_stress_gp(ctx, nullptr, nullptr, nullptr, (double *)dst);
    }
    if (n == "plastic_flag") {
        if (!ctx->pflag) {
            fans_set_error(ctx, FANS_ERR_STATE, "no PseudoPlastic model present");
            return FANS_ERR_STATE;
        }
        if (!need(sizeof(float) * N)) return FANS_ERR_ARG;
        std::vector<int> tmp((size_t)ctx->ngp * N);
        CUDA_TRY(ctx, cudaMemcpyAsync(tmp.data(), ctx->pflag, sizeof(int) * ctx->ngp * N, cudaMemcpyDeviceToHost, ctx->st));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
        float *o = (float *)dst;
        for (size_t v = 0; v < N; ++v) {  // PseudoPlastic.h:55-63: mean over the Gauss points, float
            float s = 0.f;
            for (int g = 0; g < ctx->ngp; ++g) s += (float)tmp[g * N + v];
            o[v] = s / (float)ctx->ngp;
        }
        return FANS_OK;
    }
    if (n == "plastic_strain" || n == "kinematic_hardening_variable" || n == "isotropic_hardening_variable" ||
        n == "plastic_strain_gp" || n == "kinematic_hardening_variable_gp" || n == "isotropic_hardening_variable_gp") {
        // J2Plasticity.h:245-322: element outputs are the Gauss-point means of the committed (_t) values; the *_gp outputs are all
        // Gauss-point values [element][gp][component] (J2Plasticity.h:279-307: writeSlab(..., {n_gp, n_str}))
        if (ctx->n_hist == 0) {
            fans_set_error(ctx, FANS_ERR_STATE, "no history-dependent model present");
            return FANS_ERR_STATE;
        }
        const bool gp = n.size() > 3 && n.compare(n.size() - 3, 3, "_gp") == 0;
        const std::string base = gp ? n.substr(0, n.size() - 3) : n;
        const int first = (base == "plastic_strain") ? 0 : (base == "isotropic_hardening_variable" ? 6 : 7);
        const int cnt = (base == "isotropic_hardening_variable") ? 1 : 6;
        if (first + cnt > ctx->n_hist) {
            fans_set_error(ctx, FANS_ERR_STATE, n + " is not a variable of the active model");
            return FANS_ERR_STATE;
        }
        if (!need(sizeof(double) * cnt * (gp ? ctx->ngp : 1) * N)) return FANS_ERR_ARG;
        std::vector<double> all;
        FANS_CHECK(history_expand(ctx, first, cnt, all));
        double *o = (double *)dst;
        if (gp) {
            memcpy(o, all.data(), sizeof(double) * all.size());
        } else {
            for (size_t v = 0; v < N; ++v)
                for (int k = 0; k < cnt; ++k) {
                    double sum = 0.0;
                    for (int g = 0; g < ctx->ngp; ++g) sum += all[(v * ctx->ngp + g) * cnt + k];
                    o[v * cnt + k] = sum / ctx->ngp;
                }
        }
        return FANS_OK;
    }
    if (n == "fundamental_solution") {
        if (!ctx->gamma_ready) {
            fans_set_error(ctx, FANS_ERR_STATE, "reference stiffness not set");
            return FANS_ERR_STATE;
        }
        const int h = ctx->h, NG = h * (h + 1) / 2, T = ctx->gT, nTiles = (ctx->kzc + T - 1) / T;
        const size_t cnt = (size_t)ctx->n1 * ctx->nx * ctx->kzc * NG;
        if (!need(sizeof(double) * cnt)) return FANS_ERR_ARG;
        const size_t gsz = (size_t)ctx->n1 * nTiles * NG * ctx->nx * T;
        std::vector<double> tmp(gsz);
        CUDA_TRY(ctx, cudaMemcpyAsync(tmp.data(), ctx->gamma, sizeof(double) * gsz, cudaMemcpyDeviceToHost, ctx->st));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->st));
        double *o = (double *)dst;  // natural order [ky][kx][kz][NG]
        const size_t NT = (size_t)ctx->nx * T;
        const int E = ctx->gE, TPC = ctx->nx / E;
        for (int ky = 0; ky < ctx->ny; ++ky) {
            const int py = ctx->plany.pos_host[ky];
            if (py < ctx->y1 || py >= ctx->y1 + ctx->n1) continue;
            for (int kx = 0; kx < ctx->nx; ++kx) {
                const int px = ctx->planx.pos_host[kx];
                for (int kz = 0; kz < ctx->kzc; ++kz) {
                    const int tile = kz / T, t = kz % T;
                    for (int k = 0; k < NG; ++k)
                        o[(((size_t)ky * ctx->nx + kx) * ctx->kzc + kz) * NG + k] =
                            tmp[((((size_t)(py - ctx->y1)) * nTiles + tile) * NG + k) * NT + ((size_t)(px % E) * TPC + px / E) * T + t];
                }
            }
        }
        return FANS_OK;
    }
    fans_set_error(ctx, FANS_ERR_ARG, "fans_get_field: unknown field '" + n + "'");
    return FANS_ERR_ARG;
}
