// internal.h — cross-file host entry points of libfans_gpu (not part of the C ABI)
#pragma once
#include "common.cuh"
#include "scalars.h"

// fft.cu
int  fft_plan_init(fans_ctx *ctx, FftPlan &p, int N, int ntab);
void fft_plan_free(FftPlan &p);
int  fft_pass_z_fwd(fans_ctx *ctx, const double *in);
int  fft_pass_y(fans_ctx *ctx, bool inverse);
struct YLaunch {  // one launch of the y pass: stream, component range, CTA cap, optional start signal (see k_fft_y)
    cudaStream_t st;
    int c0, nc, grid;
    int *gate;
    int gate_val;
    int tile0 = 0, ntile = 0;  // kz-tile range [tile0, tile0 + ntile) of this launch (ntile = 0: all tiles) — chunked slab pipeline
};
int  fft_pass_y_part(fans_ctx *ctx, bool inverse, const YLaunch &yl);
int  fft_pass_z_fwd_part(fans_ctx *ctx, const double *in, int c0, int nc, cudaStream_t st);
int  fft_pass_z_inv_part(fans_ctx *ctx, double *out, double scale, const double *dotw, double *red_out, int c0, int nc, bool accumulate,
                         cudaStream_t st);
int  conv_gate(fans_ctx *ctx, int gate_val, cudaStream_t st);
int  fft_pass_x_gamma(fans_ctx *ctx);
int  fft_pass_x_gamma_part(fans_ctx *ctx, cudaStream_t st, int tile0, int ntile, int grid_cap);  // kz tiles [tile0, tile0 + ntile)
int  fft_pass_z_inv(fans_ctx *ctx, double *out, double scale, const double *dotw, double *red_out);
int  fft_x_tile_width(int nx, int h);
struct SpecGeom;
SpecGeom spec_geom_A(const fans_ctx *ctx);

// fft_any.cu: grids that are not powers of two
int  any_plan_init(fans_ctx *ctx, AnyPlan &p, int n);
void any_plan_free(AnyPlan &p);
int  conv_run_any(fans_ctx *ctx, const double *in, double *out, double scale, const double *dotw, double *red_out);

// gamma.cu
int gamma_build(fans_ctx *ctx, const double *Ker0_dev, const int *frqx, const int *frqy);

// sweep.cu
enum { SWEEP_LINEAR = 0, SWEEP_RESIDUAL = 1, SWEEP_STRAINSTRESS = 2 };
uint64_t sweep_new_stamp();
int sweep_run(fans_ctx *ctx, int mode, const double *in, double *out, const double *s_in, double *d_new,
              const double *beta_dev, double *red_out, double *eps_out, double *sig_out, double *eps_gp = nullptr, double *sig_gp = nullptr);

// api.cu: (re)builds the compact history index + arrays when microstructure or materials changed since the last build
int history_prepare(fans_ctx *ctx);

// vecops.cu
int vec_cg_update(fans_ctx *ctx, double *r, const double *kd, double *u, const double *d, const double *s);
int vec_reduce4(fans_ctx *ctx, const double *a, const double *b, double *out_dev);
int vec_axpy(fans_ctx *ctx, double *y, double alpha, const double *x);
int vec_xpby(fans_ctx *ctx, double *y, double beta, const double *x);
int vec_xpby_dev(fans_ctx *ctx, double *y, const double *beta_dev, const double *x);
int vec_extrapolate(fans_ctx *ctx, double *u, double *up);
int vec_aos_to_soa(fans_ctx *ctx, const double *aos, double *soa);
int vec_soa_to_aos(fans_ctx *ctx, const double *soa, double *aos);
int vec_scalars_after_conv(fans_ctx *ctx);
int halo_exchange_both(fans_ctx *ctx, const double *in, const double *s, const double *beta_dev);
int halo_exchange_up(fans_ctx *ctx, const double *in, const double *s, const double *beta_dev);
int halo_add_down(fans_ctx *ctx, double *r);

// comm.cu (slab exchanges over NCCL)
int comm_check(fans_ctx *ctx);
int comm_allreduce(fans_ctx *ctx, const double *in, double *out, int n, bool is_max);
int comm_allreduce_int_max(fans_ctx *ctx, int *d_val);
int comm_halo(fans_ctx *ctx, const void *to_prev, void *from_next, const void *to_next, void *from_prev, size_t bytes);
int comm_alltoall(fans_ctx *ctx, const double2 *src, double2 *dst);
int comm_map_peers(fans_ctx *ctx);
void comm_unmap_peers(fans_ctx *ctx);
int comm_barrier(fans_ctx *ctx);
int comm_barrier_on(fans_ctx *ctx, cudaStream_t st);

// solve.cu
int conv_run(fans_ctx *ctx, const double *in, double *out, double scale, const double *dotw, double *red_out);
int read_scalars(fans_ctx *ctx);  // d_red -> h_red, synchronises the stream
void batch_arena_free(fans_ctx *ctx);  // lane buffers of fans_solve_batch
void iter_graph_free(fans_ctx *ctx);   // CUDA graphs of the linear CG iteration
