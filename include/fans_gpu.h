/*
 * fans_gpu.h — flat C ABI of libfans_gpu.so, the B200 (sm_100a, FP64) implementation of FANS's
 * per-iteration solve loop.  One opaque fans_ctx per Solver instance per GPU (= per slab rank).
 *
 * The reference (DataAnalyticsEngineering/FANS v0.6.2) has no FFI for this path; its seams are C++
 * templates (Solver<howmany,n_str>, Matmodel<howmany,n_str>, MaterialManager<howmany,n_str>).  Each entry
 * point below names the reference member it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; fans_last_error() gives the text.
 *   - host field buffers use the reference's memory order  [x_local][y][z][howmany]  (z fastest among
 *     nodes, components interleaved, UNPADDED — i.e. the layout of Solver::v_u, include/solver.h:124-127).
 *     The device keeps component-major planes; conversion happens in upload/download.
 *   - microstructure: uint16 [x_local][y][z]  (Reader::ms after the zyx->xyz transpose, src/reader.cpp:385-394).
 *   - strain-like vectors use the reference's ordering: thermal (3), Mandel small strain (6),
 *     row-major deformation gradient (9).
 *   - no CPU fallback: every compute entry fails with FANS_ERR_CUDA if no sm_100-class device is present.
 */
#ifndef FANS_GPU_H
#define FANS_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fans_ctx fans_ctx;

enum { FANS_OK = 0, FANS_ERR_ARG = 1, FANS_ERR_CUDA = 2, FANS_ERR_STATE = 3, FANS_ERR_MATERIAL = 4, FANS_ERR_NCCL = 5,
       FANS_ERR_NEG_JACOBIAN = 6 };

/* FE_type, include/matmodel.h:104-155 */
enum { FANS_FE_HEX8 = 0, FANS_FE_HEX8R = 1, FANS_FE_BBAR = 2 };

/* device fields (Solver::v_u, v_r, v_u_prev include/solver.h:35-38; SolverCG::s,d,rnew include/solverCG.h:20-22) */
enum { FANS_FIELD_U = 0, FANS_FIELD_R = 1, FANS_FIELD_S = 2, FANS_FIELD_D = 3, FANS_FIELD_RNEW = 4, FANS_FIELD_U_PREV = 5,
       FANS_N_FIELDS = 6 };

/* error_parameters.measure / .type, include/solver.h:414-452 */
enum { FANS_MEASURE_L1 = 0, FANS_MEASURE_L2 = 1, FANS_MEASURE_LINF = 2 };
enum { FANS_ERR_ABSOLUTE = 0, FANS_ERR_RELATIVE = 1 };
enum { FANS_METHOD_CG = 0, FANS_METHOD_FP = 1 };

/* material model ids — one per class registered in include/setup.h:21-73 */
enum {
    FANS_MAT_LINEAR = 0,               /* any LinearModel: params = tangent C (n_str*n_str, row-major); covers
                                          LinearThermalIsotropic/Triclinic (LinearThermal.h), LinearElasticIsotropic/
                                          Triclinic (LinearElastic.h) and GBDiffusion (a LinearModel) */
    FANS_MAT_PSEUDOPLASTIC_LINEAR = 1, /* PseudoPlastic.h:78-124  params = K, G, sigma_y, H, eps_crit, E_s */
    FANS_MAT_PSEUDOPLASTIC_NONLIN = 2, /* PseudoPlastic.h:126-173 params = K, G, sigma_y, n, eps_0, eps_crit */
    FANS_MAT_J2_LINEAR_ISO = 3,        /* J2Plasticity.h:165-178  params = K, G, sigma_y, K_iso, H, eta, dt */
    FANS_MAT_J2_NONLIN_ISO = 4,        /* J2Plasticity.h:180-243  params = K, G, sigma_y, K_iso, H, eta, dt, sigma_inf, delta */
    FANS_MAT_J2NEW_LINEAR_ISO = 5,     /* J2PlasticityNew.h:7-150 params = K, G, sigma_y, K_iso */
    FANS_MAT_SVK = 6,                  /* SaintVenantKirchhoff.h:33-38   params = lambda, mu */
    FANS_MAT_NEOHOOKE = 7              /* CompressibleNeoHookean.h:35-48 params = lambda, mu */
};
#define FANS_MAX_PARAMS 84

/* One entry per PHASE id found in the microstructure: MaterialManager::phase_to_info (MaterialManager.h:13-19,232-235)
 * flattened with the owning model's parsed parameters for that local_mat_id. */
typedef struct {
    int32_t model;     /* FANS_MAT_* */
    int32_t local_mat; /* local_mat_id inside its material group (only used for plastic_flag output) */
    int32_t group_n_mat; /* number of phases in the group (PseudoPlastic plastic_flag = n_mat + mat_index) */
    int32_t reserved;
    double  params[FANS_MAX_PARAMS];
} fans_phase_desc;

/* Solver ctor arguments, include/solver.h:105-142 + Reader slab sizes src/reader.cpp:311-331 */
typedef struct {
    int32_t dims[3];       /* n_x, n_y, n_z (global) */
    double  L[3];          /* microstructure.L */
    int32_t howmany;       /* 1 thermal, 3 mechanics */
    int32_t n_str;         /* 3, 6, 9 */
    int32_t fe_type;       /* FANS_FE_* */
    int32_t world_size;    /* number of x-slabs (ranks) */
    int32_t world_rank;
    int32_t local_n0;      /* x-planes owned, starting at local_0_start: n_x/world_size at world_rank*local_n0 */
    int32_t local_0_start;
    int32_t local_n1;      /* Fourier-space y-planes owned (transposed layout), starting at local_1_start */
    int32_t local_1_start;
    int32_t device;        /* CUDA device ordinal, -1 = current */
    void   *nccl_comm;     /* ncclComm_t or NULL when world_size == 1 */
    void   *stream;        /* cudaStream_t or NULL (library creates its own) */
} fans_config;

/* MixedBC after finalize(), include/mixedBCs.h:15-46 */
typedef struct {
    int32_t n_F;          /* number of stress-controlled components */
    int32_t idx_F[9];
    double  M[81];        /* (Q_F^T C0 Q_F)^+  row-major n_F x n_F */
    double  P_target[9];  /* P_F_path.row(step) */
} fans_mixed_bc;

/* Reader fields consumed by internalSolve: n_it, TOL, errorParameters, ls_* (src/reader.cpp:95-105) */
typedef struct {
    int32_t method;      /* FANS_METHOD_* */
    int32_t n_it;
    double  tol;
    int32_t measure;     /* FANS_MEASURE_* */
    int32_t err_type;    /* FANS_ERR_* */
    int32_t ls_max_iter; /* linesearch_parameters.max_iter (default 5) */
    double  ls_tol;      /* linesearch_parameters.tol (default 1e-2) */
    int32_t verbose;
    int32_t force_nonlinear; /* debug: take the line-search branch even if all phases are linear */
} fans_solve_params;

typedef struct {
    int32_t iters;            /* Solver::iter at exit */
    int32_t n_residual_evals; /* number of compute_residual / K.d sweeps */
    double  err_last;         /* last value returned by compute_error */
    double  elapsed_ms;       /* device time of internalSolve (CUDA events) */
    double  fft_ms;           /* device time inside convolution() (solver.h:293 "Total FFT Time"), CUDA events around every call */
    double  loop_ms;          /* device time of the iteration loop alone (after the initial residual), CUDA events */
} fans_solve_result;

/* ---- lifetime: Solver ctor/dtor include/solver.h:105-142, 780-805 ---- */
int  fans_create(fans_ctx **ctx, const fans_config *cfg);
void fans_destroy(fans_ctx *ctx);
const char *fans_last_error(const fans_ctx *ctx); /* ctx may be NULL: error of the last failed fans_create */
int  fans_version(void);

/* ---- slab communicator: replaces MPI_COMM_WORLD of the reference (src/main.cpp:60-61).  One rank per GPU; rank order = slab
 * order.  fans_comm_unique_id on rank 0, ship the 128 bytes to every rank (MPI_Bcast / torch.distributed / a file), then every
 * rank calls fans_comm_create and passes the handle in fans_config.nccl_comm.  (An ncclComm_t the host created itself with the
 * same libnccl is accepted as well.) ---- */
int fans_comm_unique_id(void *id128 /* out: 128 bytes */);
int fans_comm_create(void **comm, int32_t n_ranks, int32_t rank, const void *id128, int32_t device /* -1: current */);
int fans_comm_destroy(void *comm);
/* MPI_Allreduce(MPI_IN_PLACE, host_inout, n, MPI_DOUBLE, MPI_SUM) over the slabs of ctx's communicator: what Solver::postprocess does
 * with its averages (include/solver.h:556-571).  No-op for world_size == 1. */
int fans_allreduce_sum(fans_ctx *ctx, double *host_inout, int32_t n);

/* ---- problem data ---- */
int fans_set_microstructure(fans_ctx *ctx, const uint16_t *ms);                               /* Solver::ms solver.h:34,121 */
int fans_set_materials(fans_ctx *ctx, int32_t n_phases, const fans_phase_desc *phases);      /* MaterialManager ctor MaterialManager.h:38-159 */
int fans_set_reference_stiffness(fans_ctx *ctx, const double *kapparef /* n_str*n_str */);    /* computeFundamentalSolution solver.h:144-204 */
int fans_set_gradient(fans_ctx *ctx, const double *g0 /* n_str */);                           /* MaterialManager::set_gradient :216-221 */
int fans_get_gradient(fans_ctx *ctx, double *g0);
int fans_set_mixed_bc(fans_ctx *ctx, const fans_mixed_bc *mbc /* NULL disables */);           /* enableMixedBC/disableMixedBC solver.h:70-81 */
int fans_update_mixed_bc(fans_ctx *ctx);                                                      /* MixedBCController::update mixedBCs.h:160-178 */

/* ---- fields ---- */
int fans_field_upload(fans_ctx *ctx, int32_t field, const double *host);
int fans_field_download(fans_ctx *ctx, int32_t field, double *host);
int fans_field_zero(fans_ctx *ctx, int32_t field);
int fans_field_copy(fans_ctx *ctx, int32_t dst, int32_t src);

/* ---- operators ---- */
int fans_residual(fans_ctx *ctx, int32_t field_out, int32_t field_u);       /* compute_residual<pad> solver.h:229-280 */
int fans_apply_linear(fans_ctx *ctx, int32_t field_out, int32_t field_d);   /* linear-operator lambda solverCG.h:98-103 */
int fans_convolution(fans_ctx *ctx, int32_t field_in, int32_t field_out);   /* convolution solver.h:387-412 */
int fans_dot(fans_ctx *ctx, int32_t a, int32_t b, double *out);             /* SolverCG::dotProduct solverCG.h:52-59 */
int fans_axpy(fans_ctx *ctx, int32_t y, double alpha, int32_t x);           /* y += alpha*x (Eigen updates solverCG.h:106-107,130,146) */
int fans_norm(fans_ctx *ctx, int32_t field, int32_t measure, double *out);  /* compute_error solver.h:414-431 (MAX over ranks) */

/* ---- drivers ---- */
int fans_solve(fans_ctx *ctx, const fans_solve_params *p, fans_solve_result *res, double *err_hist /* n_it+1 or NULL */);
                                                                            /* Solver::solve solver.h:282-300 -> internalSolve */
int fans_homogenized_stress(fans_ctx *ctx, double *out /* n_str */);        /* get_homogenized_stress solver.h:707-737 */
/* Batched linear solves: the loop of Solver::get_homogenized_tangent (solver.h:762-775: set_gradient, solve, get_homogenized_stress
 * per column) as ONE CG loop over n_b macroscopic strains ("lanes").  Every pass of the iteration is one launch over all lanes;
 * Gamma_hat and the phase image are read from HBM once per pass.  Lane l starts from u = 0 and reproduces fans_solve with
 * g0 = macro[l]; a converged lane is frozen.  The context's own displacement, gradient and history are left untouched.
 *   macro      [n_b][n_str] macroscopic strains           res        [n_b]
 *   stress_out [n_b][n_str] homogenized stresses or NULL  err_hist   [n_b][n_it+1] or NULL
 * FANS_ERR_STATE (with the reason in fans_last_error) when the problem cannot be batched: nonlinear models, mixed BCs, slabs,
 * non-power-of-two grids, method != CG — the caller then runs the load cases one after the other like the reference. */
#define FANS_MAX_BATCH 16
int fans_solve_batch(fans_ctx *ctx, int32_t n_b, const double *macro, const fans_solve_params *p, fans_solve_result *res,
                     double *stress_out, double *err_hist);
int fans_batch_load_displacement(fans_ctx *ctx, int32_t lane, int32_t dst_field); /* lane's u -> a field of the context */
int fans_batch_release(fans_ctx *ctx);                                            /* frees the lane buffers (kept between calls) */
int fans_commit_history(fans_ctx *ctx);                                     /* update_internal_variables MaterialManager.h:208-213 */
int fans_extrapolate_displacement(fans_ctx *ctx);                           /* extrapolateDisplacement solver.h:302-311 */

/* postprocess data sources, solver.h:454-705 and the models' postprocess().  name is one of
 *   "strain","stress"            double [x][y][z][n_str]          (GP average, absolute ue)
 *   "strain_gp","stress_gp"      double [x][y][z][n_gp][n_str]
 *   "plastic_flag"               float  [x][y][z]                 (PseudoPlastic.h:55-63)
 *   "plastic_strain","kinematic_hardening_variable"   double [x][y][z][6], "isotropic_hardening_variable" double [x][y][z]
 *   "fundamental_solution"       double [y_local][x][kz][h*(h+1)/2]  natural frequency order (debug / tests) */
int fans_get_field(fans_ctx *ctx, const char *name, void *host_dst, size_t bytes);
/* ONE getStrainStress sweep (the loop of Solver::postprocess, solver.h:497-530): element-averaged strain and stress,
 * each double [x][y][z][n_str]; either pointer may be NULL.  Like the reference's sweep it calls the material law once per
 * Gauss point, with the same side effects on the history variables (J2Plasticity.h:103-104). */
int fans_strain_stress(fans_ctx *ctx, double *strain_host, double *stress_host);
/* The same single sweep with the Gauss-point outputs of Solver::postprocess (solver.h:534-542, 677-680: "strain_gp" / "stress_gp",
 * double [x][y][z][n_gp][n_str]) next to the element averages; any pointer may be NULL. */
int fans_strain_stress_gp(fans_ctx *ctx, double *strain_host, double *stress_host, double *strain_gp_host, double *stress_gp_host);

/* per-kernel-class device timing (CUDA events on the library's stream). cls = 0.. until *name is "" */
int fans_set_profiling(fans_ctx *ctx, int32_t on);  /* resets the accumulators */
int fans_get_profile(fans_ctx *ctx, int32_t cls, const char **name, double *ms, int64_t *count);

/* number of CUDA kernel launches issued by this ctx since creation (bench.py gpu_launches) */
int64_t fans_launch_count(const fans_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
