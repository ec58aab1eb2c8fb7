// Internal launch wrappers shared between the translation units of libproxmin_b200.so.
#pragma once
#include "prox.cuh"

struct AdaArgs;

int launch_zero(pmx_ctx* ctx, cudaStream_t st, float* p, size_t n, const int* done);
int launch_extrapolate(pmx_ctx* ctx, const float* X, const float* Xold, float* Xe, size_t n, float omega,
                       const int* done);
int launch_split_bf16(pmx_ctx* ctx, const float* X, int rows, int cols, void* hi, void* lo, int rows_pad, int ld_dst,
                      const int* done);
int launch_gram(pmx_ctx* ctx, cudaStream_t st, const float* X, int rows, int cols, bool tall, double* gram,
                const int* done);
int launch_lambda_max(pmx_ctx* ctx, cudaStream_t st, const double* gram, int C, pmx_ctl* ctl, int which);
int launch_grad_simt(pmx_ctx* ctx, const float* Y, int ldY, const float* A, const float* S, int M, int N, int K, float* GA,
                     float* GS, double* loss, const int* done);
