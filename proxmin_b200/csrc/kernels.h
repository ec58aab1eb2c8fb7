// Internal launch wrappers shared between the translation units of libproxmin_b200.so.
#pragma once
#include "prox.cuh"

int launch_zero(pmx_ctx* ctx, cudaStream_t st, float* p, size_t n, const int* done);
int launch_zero3(pmx_ctx* ctx, cudaStream_t st, float* p0, size_t n0, float* p1, size_t n1, float* p2, size_t n2,
                 const int* done);
int launch_extrapolate(pmx_ctx* ctx, const float* X, const float* Xold, float* Xe, size_t n, float omega,
                       const int* done);
int launch_split_bf16(pmx_ctx* ctx, const float* X, int rows, int cols, void* hi, void* lo, int rows_pad, int ld_dst,
                      const int* done);
int launch_gram(pmx_ctx* ctx, cudaStream_t st, const float* X, int rows, int cols, bool tall, double* gram,
                const int* done);
int launch_gram_reduce(pmx_ctx* ctx, cudaStream_t st, const float* part, int nblocks, int C, double* gram, const int* done);
int launch_lambda_max(pmx_ctx* ctx, cudaStream_t st, const double* gram, int C, pmx_ctl* ctl, int which);
int launch_lambda_max2(pmx_ctx* ctx, cudaStream_t st, const double* gram0, int which0, const double* gram1, int which1,
                       int C, pmx_ctl* ctl);
// rows [m0, m0 + nrows) x columns [col0, col0 + ncols) of Y from a row-major staging buffer (pitch floats per row) into the
// tiled device copy (grad_umma.h); m0 is a multiple of 4
int launch_y_interleave(pmx_ctx* ctx, cudaStream_t st, const float* stage, int pitch, int nrows, int ncols, float* Yb, int ldY,
                        int m0, int col0);
int launch_grad_simt(pmx_ctx* ctx, const float* Y, const float* W, int ldY, int y_blocked, const float* A, const float* S, int M, int N, int K, float* GA,
                     float* GS, double* loss, const int* done);

// solver_kernels.cu
int launch_diff_norms(pmx_ctx* ctx, const float* X, const float* Xold, size_t n, double* norms, const int* done);
int upd_cols_blocks(pmx_ctx* ctx, int cols);   // grid of the column-owner update kernel = fused Gram partials
int launch_axis_sum(pmx_ctx* ctx, const float* X, int rows, int cols, int axis, double* out, const int* done);
int apply_chain_general(pmx_ctx* ctx, const ProxChain& ch, float* X, int rows, int cols, const StepSpec& step,
                        double* sums_scratch, const int* done);
int launch_alpha_means(pmx_ctx* ctx, const float* X, int rows, int cols, int axis, double* sums, float* alpha,
                       const int* done);
int launch_sub_begin(pmx_ctx* ctx, pmx_ctl* ctl, int block);
int launch_sub_finalize(pmx_ctx* ctx, pmx_ctl* ctl, float e2, int max_tau);
int launch_sub_commit(pmx_ctx* ctx, float* X, const float* Z0, const float* Z1, size_t n, pmx_ctl* ctl, int block,
                      unsigned short* hi = nullptr, unsigned short* lo = nullptr, int cols = 0, int ld_split = 0);
int launch_clear_pause(pmx_ctx* ctx, pmx_ctl* ctl);
int launch_adaprox_finalize(pmx_ctx* ctx, pmx_ctl* ctl, float e2A, float e2S, int check);
int launch_bsdmm_block(pmx_ctx* ctx, pmx_ctl* ctl, int block, float* X, const float* G, float* const* Z, float* const* U,
                       float* T, double* sums_scratch, int rows, int cols, int n_g, const ProxChain& direct,
                       const ProxChain* g, const float* step_f, double* norms, float e_rel, float e_abs, bool sharded,
                       double n_elems_global, long long xchg_off = -1);
int launch_alpha_from_sums(pmx_ctx* ctx, const double* sums, int n, double count, float* alpha, const int* done);
int launch_bsdmm_iter_finalize(pmx_ctx* ctx, pmx_ctl* ctl);

// elementwise.cu: adaprox moment update
struct AdaArgs {
  const float* G;
  float* M;
  float* V;
  float* Vhat;   // may be null (quirk: then the running max is never applied, algorithms.py:176-177)
  float* X;
  float* Psi;    // out
  float* Z;      // out: copy of the stepped X
  float* Xold;   // optional out: X before the step (convergence test, algorithms.py:371-372)
  float* psimax; // out (atomicMax on the bit pattern; Psi >= 0)
  const int* done;
  size_t n;
  int rows, cols;
  StepSpec alpha;
  int scheme;
  double b1, b1_prev;  // b1 is a float64 array in the reference (algorithms.py:327-328): M is formed in double
  double b2, eps, p;   // Python floats in the reference: scalar recipes in double, one rounding to fp32 where they meet an array
  int t;  // it + 1
};
int launch_adaprox_moments(pmx_ctx* ctx, const AdaArgs& a);
