/*
 * proxmin_b200 -- C ABI of the B200-native proximal-update hot path.
 *
 * The reference (pmelchior/proxmin v0.6.12) is pure Python and has no FFI; its
 * boundary is the Python call/callback contract of proxmin/algorithms.py,
 * proxmin/nmf.py and proxmin/operators.py.  This header is the native surface a
 * maintainer of the reference would bind with ctypes (INTEGRATION.md shows the
 * stub): every entry point below names the reference lines it replaces.
 *
 * Conventions
 *   - every function returns 0 on success and a negative pmx_status otherwise;
 *     pmx_last_error() returns a thread-local, human readable message.
 *   - all matrices are fp32, row-major (C-contiguous), exactly as NumPy hands
 *     them over; "device" pointers come from pmx_malloc.
 *   - no exceptions, no C++ types, no torch types cross this boundary.
 *   - one pmx_ctx per (process, GPU); not thread-safe (same as the reference).
 */
#ifndef PROXMIN_B200_H
#define PROXMIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PMX_OK = 0,
  PMX_ERR_CUDA = -1,     /* a CUDA runtime/driver call failed          */
  PMX_ERR_ARG = -2,      /* invalid argument (shape, NULL, enum range) */
  PMX_ERR_NCCL = -3,     /* NCCL missing or a collective failed        */
  PMX_ERR_UNSUPPORTED = -4,
  PMX_ERR_NONFINITE = -5 /* non-finite Gram matrix (reference: numpy.linalg.LinAlgError, utils.py:34) */
} pmx_status;

typedef struct pmx_ctx pmx_ctx;
typedef struct pmx_nmf pmx_nmf;
typedef struct pmx_admm pmx_admm;

/* ------------------------------------------------------------------ context */
const char* pmx_last_error(void);
int pmx_version(void);
int pmx_device_count(int* count);
int pmx_ctx_create(int device, pmx_ctx** out);
int pmx_ctx_destroy(pmx_ctx* ctx);
int pmx_ctx_sync(pmx_ctx* ctx);
/* Solver handles take their device buffers from a per-context cache of freed blocks (a solve on host arrays then
 * costs no cudaMalloc/cudaFree after the first one of a shape); this returns the cached blocks to the driver. */
int pmx_ctx_trim(pmx_ctx* ctx);
/* number of kernels this library has launched on the context since creation */
int pmx_ctx_launch_count(pmx_ctx* ctx, long long* count);
/* name/SM count/memory of the device behind the context */
int pmx_ctx_device_info(pmx_ctx* ctx, char* name, int name_len, int* sm_count, size_t* total_mem);

/* per-launch CUDA-event timing of the dominant kernel (the fused gradient kernel), recorded on the
 * stream it is launched on; read returns the summed duration and the number of launches timed */
int pmx_ctx_profile(pmx_ctx* ctx, int enable);
int pmx_ctx_profile_read(pmx_ctx* ctx, float* total_ms, int* launches);

/* multi-GPU: one process per GPU; `unique_id` (128 bytes) is produced by
 * pmx_comm_unique_id on rank 0 and distributed by the host (bench.py uses
 * torch.distributed/gloo for that plumbing).  Collectives are NCCL all-reduces
 * over NVLink on the context's stream. */
int pmx_comm_unique_id(void* unique_id_128);
int pmx_comm_init(pmx_ctx* ctx, const void* unique_id_128, int world, int rank);
int pmx_comm_allreduce_sum(pmx_ctx* ctx, float* dev_buf, size_t count);
/* *enabled = 1 when the ranks of the communicator exchange over CUDA-IPC peer memory (one box), 0 = NCCL only */
int pmx_comm_peer_enabled(pmx_ctx* ctx, int* enabled);

/* ------------------------------------------------------------------- memory */
int pmx_malloc(pmx_ctx* ctx, size_t bytes, void** dev_ptr);
int pmx_free(pmx_ctx* ctx, void* dev_ptr);
int pmx_memset(pmx_ctx* ctx, void* dev_ptr, int byte, size_t bytes);
int pmx_h2d(pmx_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes);
int pmx_d2h(pmx_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes);
int pmx_d2d(pmx_ctx* ctx, void* dev_dst, const void* dev_src, size_t bytes);
/* pinned host staging buffers (bench.py end-to-end leg) */
int pmx_host_alloc(size_t bytes, void** host_ptr);
int pmx_host_free(void* host_ptr);
/* CUDA-event timing on the context's stream: returns milliseconds between two marks */
int pmx_timer_start(pmx_ctx* ctx);
int pmx_timer_stop(pmx_ctx* ctx, float* ms);

/* ------------------------------------------------------ proximal operators
 * Replaces proxmin/operators.py:20-160 and AlternatingProjections (:187-211).
 * A pmx_prox is a chain of primitive ops in APPLICATION order (the Python side
 * reverses AlternatingProjections lists and expands `repeat`, operators.py:207-211;
 * prox_unity_plus = PLUS,UNITY (:48-52); prox_soft_plus = SOFT,PLUS (:153-160);
 * prox_hard_plus = HARD,PLUS (:128-135)). */
typedef enum {
  PMX_OP_ID = 0,    /* operators.py:20-23   */
  PMX_OP_ZERO = 1,  /* operators.py:26-30   */
  PMX_OP_PLUS = 2,  /* operators.py:33-38   X[X<0] = 0                    */
  PMX_OP_UNITY = 3, /* operators.py:41-45   X /= sum(X, axis, keepdims)   */
  PMX_OP_MIN = 4,   /* operators.py:55-69   X[X-t<0] = t                  */
  PMX_OP_MAX = 5,   /* operators.py:72-84   X[X-t>0] = t                  */
  PMX_OP_HARD = 6,  /* operators.py:109-125 X[|X|<t] = 0                  */
  PMX_OP_SOFT = 7,  /* operators.py:138-150 sign(X)*max(|X|-t,0)          */
  PMX_OP_MAXENT = 8,   /* operators.py:163-184 X[X>0] = t W(exp(X/t - 1)/t), W = Lambert W, t = gamma (*step);
                          fp32 semantics of the reference: exp overflows to inf for X/t > 89 */
  PMX_OP_MAXENT64 = 9  /* same for a caller array of dtype float64: the reference then evaluates exp in fp64 */
} pmx_op_code;

typedef struct {
  int32_t op;       /* pmx_op_code */
  int32_t relative; /* 1: t = thresh*step (type="relative", operators.py:4-14); 0: t = thresh */
  int32_t axis;     /* PMX_OP_UNITY only: 0 = sum over rows, 1 = sum over columns */
  float thresh;
} pmx_prox_op;

#define PMX_MAX_OPS 8
typedef struct {
  int32_t n_ops;
  pmx_prox_op ops[PMX_MAX_OPS];
} pmx_prox;

/* X[rows x cols] (device, in place) <- prox(X, step).  operators.py:20-160. */
int pmx_prox_apply(pmx_ctx* ctx, const pmx_prox* prox, float* dev_X, int rows, int cols, float step);

/* ------------------------------------------------ NMF building blocks
 * nmf.py:28-41 grad_likelihood (W == 1): G_A = (A S - Y) S^T, G_S = A^T (A S - Y),
 * one pass over Y; *loss (optional, device double) = sum((A S - Y)^2)/2 (nmf.py:13-25).
 * Y is M x N, A is M x K, S is K x N (all device). `kernel`: 0 = auto,
 * 1 = SIMT fp32, 2 = tcgen05 (3xBF16 split, fp32 accumulate). */
int pmx_nmf_grad(pmx_ctx* ctx, const float* Y, const float* A, const float* S, int M, int N, int K,
                 float* G_A, float* G_S, double* loss_or_null, int kernel);
/* nmf.py:44-65 + utils.py:14-35: Lipschitz constants lambda_max(S S^T), lambda_max(A^T A);
 * host outputs; steps are their reciprocals. Returns PMX_ERR_NONFINITE like eigvals would raise. */
int pmx_nmf_lipschitz(pmx_ctx* ctx, const float* A, const float* S, int M, int N, int K,
                      float* lip_A_host, float* lip_S_host);

/* ------------------------------------------------ NMF solver object
 * Replaces the iteration loops of algorithms.py:87-135 (pgm), :365-410 (adaprox),
 * :800-844 (bsdmm as driven by nmf.py:178-203) for the NMF objective.  The object
 * owns device copies of Y (local column stripe), A, S and all solver state. */
typedef enum { PMX_A = 0, PMX_S = 1, PMX_GA = 2, PMX_GS = 3, PMX_MA = 4, PMX_MS = 5, PMX_VA = 6, PMX_VS = 7,
               PMX_VHA = 8, PMX_VHS = 9 } pmx_which;

int pmx_nmf_create(pmx_ctx* ctx, int M, int N_local, int K, pmx_nmf** out);
int pmx_nmf_destroy(pmx_nmf* h);
/* upload rows x cols host block (leading dimension ld, in elements) of Y starting at column col0 */
int pmx_nmf_set_Y(pmx_nmf* h, const float* host_Y, size_t ld, int col0, int ncols);
/* weights of the weighted likelihood (nmf.py:25, 40: D = W (A S - Y), sum W (Y - A S)^2 / 2): an M x N host matrix
 * uploaded like Y; once set, every gradient / loss of this handle is weighted */
int pmx_nmf_set_W(pmx_nmf* h, const float* host_W, size_t ld, int col0, int ncols);
/* nmf.py:28-41 at the handle's current (A, S): gradients into the GA / GS buffers (pmx_nmf_get), optionally the
 * log-likelihood (nmf.py:13-25) */
int pmx_nmf_gradient(pmx_nmf* h, double* loss_host_or_null);
int pmx_nmf_set(pmx_nmf* h, int which, const float* host_src);
int pmx_nmf_get(pmx_nmf* h, int which, float* host_dst);
int pmx_nmf_device_ptr(pmx_nmf* h, int which, float** dev_ptr);
int pmx_nmf_loss(pmx_nmf* h, double* loss_host); /* nmf.py:13-25 at the current (A,S) */

typedef struct {
  pmx_prox prox_A, prox_S; /* algorithms.py:57-64 (None -> ID) */
  int32_t accelerated;     /* algorithms.py:83,93-95 Nesterov */
  float e_rel_A, e_rel_S;  /* algorithms.py:66-68,130-133 */
  int32_t kernel;          /* gradient kernel selector, see pmx_nmf_grad */
  int32_t check_every;     /* host polls the device-side stop flag every this many iterations */
} pmx_pgm_opts;

/* start a PGM run: resets the Nesterov sequence and iteration counter */
int pmx_nmf_pgm_begin(pmx_nmf* h, const pmx_pgm_opts* opts);
/* run up to n_iter more iterations (stops early when both blocks converged);
 * out: iterations executed in this call, convergence flags of the last executed iteration,
 * the steps used in it (algorithms.py:144 returns converged, G, S). */
int pmx_nmf_pgm_run(pmx_nmf* h, int n_iter, int* iters_done, int* conv_A, int* conv_S,
                    float* step_A, float* step_S);

typedef enum { PMX_ADAM = 0, PMX_NADAM = 1, PMX_AMSGRAD = 2, PMX_PADAM = 3, PMX_ADAMX = 4, PMX_RADAM = 5 } pmx_scheme;
typedef struct {
  pmx_prox prox_A, prox_S;
  int32_t has_prox_A, has_prox_S; /* algorithms.py:380: prox None skips the sub-iterations */
  int32_t scheme;                 /* algorithms.py:338-345 */
  double b2, eps, p;              /* algorithms.py:255-258 (Python floats in the reference: kept in double) */
  float e_rel_A, e_rel_S;
  int32_t check_convergence;      /* algorithms.py:257,371,403 */
  int32_t prox_max_iter;          /* algorithms.py:261,386 */
  int32_t has_vhat;               /* caller supplied Vhat (quirk: otherwise the running max is never applied) */
  int32_t kernel;
  int32_t step_mode;              /* 0: nmf.py:91-93 step_adaprox (means/10); 1: fixed alpha_A, alpha_S */
  float alpha_A, alpha_S;
} pmx_adaprox_opts;

int pmx_nmf_adaprox_begin(pmx_nmf* h, const pmx_adaprox_opts* opts);
/* b1: per-iteration first-moment decay for iterations [it0, it0+n_iter) (algorithms.py:327-330) */
int pmx_nmf_adaprox_run(pmx_nmf* h, int n_iter, const double* b1, const double* b1_prev, int* iters_done,
                        int* conv_A, int* conv_S, long long* sub_A, long long* sub_S);

typedef struct {
  pmx_prox prox_A, prox_S;       /* direct constraints inside prox_f (nmf.py:181-185) */
  int32_t n_g_A, n_g_S;          /* number of ADMM constraints per block (0 = proxs_g[j] is None) */
  pmx_prox proxs_g_A[4], proxs_g_S[4];
  float e_rel_A, e_rel_S, e_abs_A, e_abs_S;
  int32_t kernel;
} pmx_bsdmm_opts;

int pmx_nmf_bsdmm_begin(pmx_nmf* h, const pmx_bsdmm_opts* opts);
int pmx_nmf_bsdmm_run(pmx_nmf* h, int n_iter, int* iters_done, int* conv_A, int* conv_S);

/* ------------------------------------------------ ADMM / SDMM on one vector block, L = identity
 * Replaces utils.py:295-391 (update_variables, do_the_mm, check_constraint_convergence)
 * for the built-in f = 0.5*||X - b||^2 gradient-step prox_f (README.md:82-84 pattern,
 * prox_f(X, s) = X - s (X - b)) and built-in prox_g chains; algorithms.py:478-514, :603-644. */
typedef struct {
  int32_t n_g;            /* number of constraints: 1 = admm, >1 = sdmm */
  pmx_prox proxs_g[4];
  float e_rel, e_abs;
  int32_t dual_uses_step_g; /* 0 reproduces the admm quirk (algorithms.py:494-496), 1 for sdmm */
} pmx_admm_opts;

int pmx_admm_create(pmx_ctx* ctx, size_t n, const pmx_admm_opts* opts, pmx_admm** out);
int pmx_admm_destroy(pmx_admm* h);
int pmx_admm_set(pmx_admm* h, const float* host_X, const float* host_b);
int pmx_admm_get(pmx_admm* h, float* host_X);
/* (re)initialise Z = X, U = 0 (utils.py:244-254) */
int pmx_admm_init_zu(pmx_admm* h);
/* one pass of utils.py:307-346 + :366-391 with step_f (already multiplied by slack) and
 * step_g_i = step_f * n_g (utils.py:279); fills errors[4*i..] = e_pri, e_dual, |R|, |S| per constraint,
 * *converged, and *stalled = (X and every R_i bit-identical to the previous pass, algorithms.py:506,635) */
int pmx_admm_step(pmx_admm* h, double step_f, int* converged, int* stalled, double* errors);
/* fused loop of admm/sdmm iterations with constant step_f: device-side stop flag, no host round trips */
int pmx_admm_run(pmx_admm* h, double step_f, int max_iter, int* iters_logged, int* converged, double* errors);
/* passes executed by the last pmx_admm_run (the halved-slack restarts of algorithms.py:503-512 reset the iteration
 * counter, so this can exceed max_iter) and the number of restarts */
int pmx_admm_stats(pmx_admm* h, long long* passes, int* restarts);

/* ------------------------------------------------ elementwise solver primitives (generic callback path)
 * Used by the Python solvers when grad/step/prox are arbitrary user callables
 * (algorithms.py:107-108, 130-133): X_new = prox(Xe - step*G); returns the two norms. */
int pmx_pgm_update(pmx_ctx* ctx, const pmx_prox* prox, const float* dev_Xe, const float* dev_G, float* dev_X,
                   int rows, int cols, float step, double* norm_diff_sq, double* norm_new_sq);

/* single elementwise expression on device arrays (callback loops; all operands length n).
 * red_host (optional, 5 doubles) receives the reductions of the opcode. */
typedef enum {
  PMX_EW_EXTRAP = 0,  /* o0 = a + s0*(a - b)                     algorithms.py:95                 */
  PMX_EW_ADD = 1,     /* o0 = a + b                               utils.py:297                     */
  PMX_EW_SUB = 2,     /* o0 = a - b                               utils.py:317,338                 */
  PMX_EW_DX_ACC = 3,  /* o0 = d + s0*(a - b + c)   (d may be NULL) utils.py:316,333                */
  PMX_EW_ZU = 4,      /* a=X b=Z' c=Z d=U s0=-1/step_g s1=step_g|0: o0=R o1=S o2=U+R;
                         red = |X|^2 |Z'|^2 |U'(/s1)|^2 |R|^2 |S|^2   utils.py:299-303,349-363 */
  PMX_EW_DOT_DIFF = 5,/* red = sum((a-b)*c), sum((a-b)^2)         algorithms.py:118                */
  PMX_EW_MAXABS = 6,  /* red[0] = max|s0*a|                       algorithms.py:121                */
  PMX_EW_SUMSQ = 7,   /* red[0] = sum(a^2)                        utils.py:257-260                 */
  PMX_EW_AXPY = 9,    /* o0 = s0*a + b   (b may be NULL)              utils.py:316,333 with a dense L  */
  PMX_EW_BB = 8       /* a=X b=X_prev c=G d=G_prev: red = sum(S^2) sum(S*Y) sum(Y^2) sum(G^2), S = a-b, Y = c-d
                         (Barzilai-Borwein step sizes)                utils.py:225-239                 */
} pmx_ew_op;
int pmx_ew(pmx_ctx* ctx, int op, size_t n, const float* a, const float* b, const float* c, const float* d, float s0,
           float s1, float* o0, float* o1, float* o2, double* red_host);

/* dense linear operator of admm / sdmm / bsdmm (utils.py:38-101 MatrixAdapter.dot, :316, :333, :299-303):
 * O[p x m] = op(L)[p x n] X[n x m], all device, row-major, fp32; trans != 0: op(L) = L^T with L stored n x p. */
int pmx_matmul(pmx_ctx* ctx, const float* L, const float* X, float* O, int p, int n, int m, int trans);

/* sums along an axis of a device matrix: out_host has cols (axis 0) or rows (axis 1) doubles (nmf.py:91-93 means) */
int pmx_axis_sum(pmx_ctx* ctx, const float* X, int rows, int cols, int axis, double* out_host);

/* adaprox building blocks on caller-owned device arrays (callback loop of algorithms.py:365-410).
 * alpha: mode 0 = scalar alpha_value, 2 = per-column vector alpha_dev[cols], 3 = per-row vector alpha_dev[rows]. */
int pmx_adaprox_moments(pmx_ctx* ctx, int scheme, const float* G, float* M, float* V, float* Vhat_or_null, float* X,
                        float* Psi, int rows, int cols, const float* alpha_dev, int alpha_mode, float alpha_value,
                        double b1, double b1_prev, double b2, double eps, double p, int t, float* psimax_host);
int pmx_adaprox_sub(pmx_ctx* ctx, const pmx_prox* prox, const float* Z, const float* X, const float* Psi, float* Zout,
                    int rows, int cols, const float* alpha_dev, int alpha_mode, float alpha_value, float psimax,
                    double* norms_host);

#ifdef __cplusplus
}
#endif
#endif /* PROXMIN_B200_H */
