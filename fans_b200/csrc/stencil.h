// stencil.h — node-gather linear operator (stencil.cu)
#pragma once
#include "common.cuh"
#define STENCIL_MAXQ 4
bool stencil_supported(const fans_ctx *ctx);
int stencil_run(fans_ctx *ctx, const double *d_old, double *out, const double *s_in, double *d_new, const double *beta_dev,
                double *red_out);
void stencil_from_element_matrix(int h, const double *K, double *S);
