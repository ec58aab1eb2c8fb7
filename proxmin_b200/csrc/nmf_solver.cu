// NMF solver object: device-resident state and the iteration loops of
//   algorithms.py:87-135 (pgm), :365-410 (adaprox), :800-844 (bsdmm via nmf.py:178-203)
// specialised to the Gaussian-likelihood NMF objective of nmf.py:13-41.
//
// One iteration is a short, fixed sequence of kernels on the context's stream; the host never
// waits for the device inside an iteration.  Convergence is decided on the device (pmx_ctl): once
// `done` is set every later kernel returns immediately, so (A, S, G) stay frozen at exactly the
// iteration where the reference would `break`, while the host polls the flag only every
// `check_every` iterations.
#include <math.h>
#include <stdlib.h>

#include "kernels.h"
#include "grad_umma.h"
#include "pgm_tail.h"

int pmx_comm_allreduce_internal(pmx_ctx* ctx, void* buf, size_t count, int kind, cudaStream_t st);

bool pmx_comm_has_aux(pmx_ctx* ctx);
bool pmx_peer_available(pmx_ctx* ctx);
int pmx_peer_arena(pmx_ctx* ctx, size_t bytes, pmx_peer_region** out);
int pmx_peer_reset(pmx_ctx* ctx);
int pmx_peer_signal(pmx_ctx* ctx, int set, cudaStream_t st, const int* done, const double* copy_src,
                    size_t copy_offset_bytes, size_t n);
int pmx_peer_sum(pmx_ctx* ctx, int set, size_t offset_bytes, size_t n, void* dst, int kind, cudaStream_t st, const int* done,
                 int* fault = nullptr);
size_t pmx_peer_small_bytes();
int pmx_peer_small_allreduce(pmx_ctx* ctx, int set, size_t off_bytes, void* buf, int n, int kind, cudaStream_t st,
                             const int* done, int* fault);

struct pmx_nmf {
  pmx_ctx* ctx;
  int M, N, K;   // N = local number of columns (this rank's stripe of Y and S)
  double N_global; // total number of columns over all ranks (row means of S in step_adaprox)
  float* W;      // weights of the weighted likelihood (nmf.py:25, 40) in the layout of Y, or nullptr (W = 1)
  int ldY;       // the device copy of Y is tiled (grad_umma.h): ldY = 128-column tiles per row block = ceil(N / 128)
  float *Y, *A, *S, *A_old, *S_old, *Ae, *Se, *GA, *GS;
  double *gramA, *gramS;
  pmx_ctl* ctl;     // device
  pmx_ctl* h_ctl;   // pinned host mirror
  UmmaPlan* plan;   // tcgen05 gradient kernel state (tensor maps, bf16 operand buffers); lazily built
  float* gram_part; // per-block partial Gram matrices of S written by the fused S update
  bool gram_pending; // the fused update left per-block Gram partials of S that still have to be summed into gramS
  int gram_pending_blocks;
  bool gramS_valid; // gramS already holds S S^T of the current S (fused update): skip the standalone Gram pass
  bool split_valid; // plan's bf16 operand buffers hold the split of the current (A, S) (written by the update kernels)
  bool used_umma;   // the last gradient evaluation went through the tcgen05 kernel
  // ---- pgm
  pmx_pgm_opts pgm;
  ProxChain chA, chS;
  double nest_t;
  int it_enqueued;
  bool peer_mode;              // sharded PGM: exchanges through peer memory instead of NCCL
  size_t peer_off_gram;        // arena offset of the (Gram(S), norms) pair
  cudaGraphExec_t pgm_graph;   // steady-state iteration captured once, replayed per iteration (no launch gaps)
  long long pgm_graph_launches; // kernels inside the graph (for the launch counter)
  // ---- fused PGM tail (pgm_tail.cu): iteration = gradient kernel -> [peer signal] -> one tail kernel
  bool tail_mode;              // this PGM solve runs the fused tail
  bool tail_ready;             // steps / bf16 operands / arena copies match the current (A, S)
  float *GA2, *GS2;            // gradient buffer pairs indexed by the iteration parity (GA2: single-GPU runs only;
                               // sharded runs keep the G_A pair in the peer arena)
  size_t ga_stride, gs_stride; // elements between the two buffers of a pair
  float* tail_gram_rep;        // [2][PMX_TAIL_NREP][K*K] replicated Gram accumulators
  double* tail_acc;            // [2][4] norm partials
  unsigned* tail_tickets;      // [4]
  size_t off_GA, off_A, off_Ahi, off_Alo, off_inbox;   // regions of the peer arena (bytes)
  cudaGraphExec_t tail_graph;
  long long tail_graph_launches;
  // ---- sharded adaprox / bsdmm: exchanges over peer memory (arena: [G_A pair | small all-reduce inboxes])
  bool xchg_peer;
  size_t xchg_ga_stride, xchg_off_small;
  bool tail_G_pending;         // the last gradients still sit in the parity buffers (tail_publish_G)
  size_t tail_G_par;
  // ---- adaprox
  pmx_adaprox_opts ada;
  float *MA, *MS, *VA, *VS, *VhA, *VhS, *Psi, *Z0, *Z1, *alphaA, *alphaS;
  int ada_it;
  int ada_spec[2];   // sub-iterations enqueued speculatively per block (no host round trip), follows the last count
  int ada_enq[2];    // sub-iterations enqueued for the block update in flight
  bool ada_fuse_split; // both blocks have a prox: their commits write the bf16 operands of the next gradient kernel
  // ---- bsdmm
  pmx_bsdmm_opts bs;
  float* Zg[2][4];
  float* Ug[2][4];
  double* bs_norms;  // device: [2 blocks][4 constraints][5 norms]
  int bs_it;
};

namespace {

__global__ void k_pgm_finalize(pmx_ctl* ctl, double* normsS, float e2A, float e2S) {
  if (ctl->done) return;
  if (normsS) {   // sharded run: the S-block norms arrived through the packed all-reduce buffer
    for (int i = 0; i < 3; ++i) {
      ctl->norms[3 + i] = normsS[i];
      normsS[i] = 0.0;
    }
  }
  // algorithms.py:130-133: l2sq(X - X_) <= e_rel**2 * l2sq(X), evaluated in fp32 like the reference
  const bool cA = (float)ctl->norms[0] <= e2A * (float)ctl->norms[1];
  const bool cS = (float)ctl->norms[3] <= e2S * (float)ctl->norms[4];
  ctl->conv[0] = cA;
  ctl->conv[1] = cS;
  ctl->it += 1;
  if (cA && cS) ctl->done = 1;
  for (int i = 0; i < 8; ++i) ctl->norms[i] = 0.0;   // ready for the next iteration's update kernels
}

__global__ void k_ctl_clear_norms(pmx_ctl* ctl) {
  if (ctl->done) return;
  for (int i = threadIdx.x; i < 8; i += blockDim.x) ctl->norms[i] = 0.0;
}

int alloc_f(pmx_ctx* ctx, float** p, size_t n) {
  return pmx_dev_alloc(ctx, (void**)p, sizeof(float) * (n ? n : 1));
}

int pull_ctl(pmx_nmf* h) {
  PMX_CUDA(cudaMemcpyAsync(h->h_ctl, h->ctl, sizeof(pmx_ctl), cudaMemcpyDeviceToHost, h->ctx->stream));
  PMX_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return PMX_OK;
}

}  // namespace

// gradient at (A, S) into (GA, GS) [+ loss], kernel selection, multi-GPU sum of the G_A partials
// defer_reduce: the caller sums the G_A partials over the ranks itself (PGM does it on the side stream)
// ga_epoch: peer-memory mode of the tcgen05 kernel (GA = base of the buffer pair in the arena, no reduction here)
// want: bit 0 = G_A, bit 1 = G_S (tcgen05 kernel only: an unwanted gradient costs no MMAs and no flush)
int nmf_gradient(pmx_nmf* h, const float* A, const float* S, float* GA, float* GS, double* loss, int kernel,
                 const int* done, bool defer_reduce = false, const unsigned* ga_epoch = nullptr, size_t ga_stride = 0,
                 int want = 3) {
  pmx_ctx* ctx = h->ctx;
  bool use_umma = false;
  if (kernel == 2) use_umma = true;
  if (kernel == 0) use_umma = umma_supported(h->M, h->N, h->K) && (long long)h->M * h->N >= 128LL * 128;
  if (use_umma) {
    if (!umma_supported(h->M, h->N, h->K)) {
      pmx_set_error("tcgen05 gradient kernel does not support M=%d N=%d K=%d", h->M, h->N, h->K);
      return PMX_ERR_UNSUPPORTED;
    }
    if (!h->plan) PMX_CHECK(umma_plan_create(ctx, h->Y, h->ldY, h->M, h->N, h->K, &h->plan, 1));
    PMX_CHECK(umma_plan_set_W(h->plan, h->W));
    const int skip = (h->split_valid && A == h->A && S == h->S) ? 1 : 0;
    if (h->xchg_peer && ctx->world > 1 && !defer_reduce && !ga_epoch && (want & 1) && GA) {
      // sharded adaprox / bsdmm: the G_A partials land in the arena pair, every rank sums them out of peer memory
      // (one-shot sum in rank order: bit-identical on every rank) instead of an NCCL all-reduce of M x K floats
      float* ga_pair = reinterpret_cast<float*>(ctx->peer_arena.local);
      PMX_CHECK(launch_grad_umma(ctx, h->plan, A, S, ga_pair, GS, loss, done, skip, &ctx->peer_epoch[0], h->xchg_ga_stride, want));
      PMX_CHECK(pmx_peer_signal(ctx, 0, ctx->stream, done, nullptr, 0, 0));
      PMX_CHECK(pmx_peer_sum(ctx, 0, 0, h->xchg_ga_stride, GA, 0, ctx->stream, done, &h->ctl->fault));
      h->used_umma = true;
      if (loss) PMX_CHECK(pmx_comm_allreduce_internal(ctx, loss, 1, 1, ctx->stream));
      return PMX_OK;
    }
    PMX_CHECK(launch_grad_umma(ctx, h->plan, A, S, GA, GS, loss, done, skip, ga_epoch, ga_stride, want));
    h->used_umma = true;
  } else {
    h->used_umma = false;
    PMX_CHECK(launch_grad_simt(ctx, h->Y, h->W, h->ldY, 1, A, S, h->M, h->N, h->K, GA, GS, loss, done));
  }
  if (ctx->world > 1) {
    if (!defer_reduce && (want & 1) && GA) PMX_CHECK(pmx_comm_allreduce_internal(ctx, GA, (size_t)h->M * h->K, 0, ctx->stream));
    if (loss) PMX_CHECK(pmx_comm_allreduce_internal(ctx, loss, 1, 1, ctx->stream));
  }
  return PMX_OK;
}

// prox_unity(axis=1) on the column-sharded S block divides by a row sum over ALL ranks' columns: not implemented,
// so a sharded solve must refuse it instead of normalising by the local partial sum
static int reject_sharded_row_unity(pmx_nmf* h, const ProxChain& chS, const char* what) {
  const int ax = chain_unity_axis(chS);
  if (h->ctx->world > 1 && (ax == 1 || ax == 2)) {
    pmx_set_error("%s: prox_unity(axis=1) on the column-sharded S block needs a cross-rank row sum (not implemented)", what);
    return PMX_ERR_UNSUPPORTED;
  }
  return PMX_OK;
}

static bool nmf_uses_umma(pmx_nmf* h, int kernel) {
  return kernel == 2 || (kernel == 0 && umma_supported(h->M, h->N, h->K) && (long long)h->M * h->N >= 128LL * 128);
}

// sharded adaprox / bsdmm (no fused tail): lay the peer arena out as [G_A pair | 8 small all-reduce inboxes].
// Collective (every rank calls it from the solver's begin function); without symmetric memory NCCL stays.
static inline size_t xchg_align(size_t x) { return (x + 255) & ~(size_t)255; }
static int xchg_setup(pmx_nmf* h, int kernel) {
  pmx_ctx* ctx = h->ctx;
  h->xchg_peer = false;
  if (ctx->world <= 1 || !pmx_peer_available(ctx) || getenv("PMX_NO_PEER_XCHG")) return PMX_OK;
  const size_t mk = (size_t)h->M * h->K;
  h->xchg_ga_stride = mk;    // (pmx_peer_sum clears the other-parity buffer: the pair is 2 x mk floats)
  h->xchg_off_small = xchg_align(2 * mk * sizeof(float));
  pmx_peer_region* ar = nullptr;
  if (pmx_peer_arena(ctx, h->xchg_off_small + 8 * pmx_peer_small_bytes(), &ar) != PMX_OK) return PMX_OK;
  PMX_CHECK(pmx_peer_reset(ctx));
  h->xchg_peer = true;
  (void)kernel;
  return PMX_OK;
}
// in-place all-reduce of a few scalars (kind 1: doubles, sum; kind 2: int32 max): peer inbox `slot` or NCCL
static int nmf_allreduce(pmx_nmf* h, void* buf, size_t count, int kind, cudaStream_t st, int slot) {
  pmx_ctx* ctx = h->ctx;
  if (ctx->world <= 1) return PMX_OK;
  if (h->xchg_peer && kind != 0 && count <= PMX_SMALL_MAX && st == ctx->stream)
    return pmx_peer_small_allreduce(ctx, 1, h->xchg_off_small + (size_t)slot * pmx_peer_small_bytes(), buf, (int)count, kind, st,
                                    &h->ctl->done, &h->ctl->fault);
  return pmx_comm_allreduce_internal(ctx, buf, count, kind, st);
}

// lip/step of both blocks at (A, S): step[0] = 1/lambda_max(S S^T), step[1] = 1/lambda_max(A^T A)
// (nmf.py:44-49).  Runs on the aux stream so that it overlaps the gradient kernel.
int nmf_steps(pmx_nmf* h, const float* A, const float* S, bool need_A, bool need_S) {
  pmx_ctx* ctx = h->ctx;
  PMX_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
  PMX_CUDA(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
  if (need_A && h->gramS_valid && S == h->S && h->gram_pending) {
    // partial Gram matrices written by the previous S update: summed here, off the main stream's critical path
    PMX_CHECK(launch_gram_reduce(ctx, ctx->aux, h->gram_part, h->gram_pending_blocks, h->K, h->gramS, &h->ctl->done));
    if (ctx->world > 1) PMX_CHECK(pmx_comm_allreduce_internal(ctx, h->gramS, (size_t)h->K * h->K, 1, ctx->aux));
    h->gram_pending = false;
  } else if (need_A && !(h->gramS_valid && S == h->S)) {  // Gram of S: sum over this rank's columns, then over ranks
    PMX_CHECK(launch_gram(ctx, ctx->aux, S, h->K, h->N, false, h->gramS, &h->ctl->done));
    if (ctx->world > 1) PMX_CHECK(pmx_comm_allreduce_internal(ctx, h->gramS, (size_t)h->K * h->K, 1, ctx->aux));
  }
  if (need_S) PMX_CHECK(launch_gram(ctx, ctx->aux, A, h->M, h->K, true, h->gramA, &h->ctl->done));
  if (need_A && need_S)
    PMX_CHECK(launch_lambda_max2(ctx, ctx->aux, h->gramS, 0, h->gramA, 1, h->K, h->ctl));
  else if (need_A)
    PMX_CHECK(launch_lambda_max(ctx, ctx->aux, h->gramS, h->K, h->ctl, 0));
  else if (need_S)
    PMX_CHECK(launch_lambda_max(ctx, ctx->aux, h->gramA, h->K, h->ctl, 1));
  PMX_CUDA(cudaEventRecord(ctx->ev_join, ctx->aux));
  return PMX_OK;
}
int nmf_global_cols(pmx_nmf* h);
int nmf_steps_join(pmx_nmf* h) {
  PMX_CUDA(cudaStreamWaitEvent(h->ctx->stream, h->ctx->ev_join, 0));
  return PMX_OK;
}

// total number of columns over all ranks (one tiny all-reduce; cached)
int nmf_global_cols(pmx_nmf* h) {
  h->N_global = (double)h->N;
  if (h->ctx->world <= 1) return PMX_OK;
  double* d = nullptr;
  PMX_CUDA(cudaMalloc((void**)&d, sizeof(double)));
  PMX_CUDA(cudaMemcpyAsync(d, &h->N_global, sizeof(double), cudaMemcpyHostToDevice, h->ctx->stream));
  int st = pmx_comm_allreduce_internal(h->ctx, d, 1, 1, h->ctx->stream);
  if (st == PMX_OK) {
    cudaError_t e = cudaMemcpyAsync(&h->N_global, d, sizeof(double), cudaMemcpyDeviceToHost, h->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->ctx->stream);
    if (e != cudaSuccess) {
      pmx_set_error("nmf_global_cols: %s", cudaGetErrorString(e));
      st = PMX_ERR_CUDA;
    }
  }
  cudaFree(d);
  return st;
}

extern "C" {

static int tail_publish_G(pmx_nmf* h);
static int tail_detach(pmx_nmf* h);

int pmx_nmf_create(pmx_ctx* ctx, int M, int N_local, int K, pmx_nmf** out) {
  PMX_REQUIRE(ctx && out, "NULL argument");
  PMX_REQUIRE(M > 0 && N_local > 0 && K > 0, "shape must be positive");
  PMX_REQUIRE(K <= 128, "K <= 128 is supported");
  pmx_nmf* h = new pmx_nmf();
  memset(h, 0, sizeof(*h));
  h->ctx = ctx;
  h->M = M;
  h->N = N_local;
  h->K = K;
  h->ldY = pmx_div_up(N_local, 128);
  h->N_global = (double)N_local;
  PMX_CUDA(cudaSetDevice(ctx->device));
  const size_t mk = (size_t)M * K, kn = (size_t)K * N_local;
  PMX_CHECK(alloc_f(h->ctx, &h->Y, (size_t)pmx_div_up(M, 128) * h->ldY * 16384));   // tiled, see grad_umma.h
  PMX_CHECK(alloc_f(h->ctx, &h->A, mk));
  PMX_CHECK(alloc_f(h->ctx, &h->S, kn));
  PMX_CHECK(alloc_f(h->ctx, &h->A_old, mk));
  PMX_CHECK(alloc_f(h->ctx, &h->S_old, kn));
  PMX_CHECK(alloc_f(h->ctx, &h->GA, mk));
  PMX_CHECK(alloc_f(h->ctx, &h->GS, kn));
  PMX_CUDA(cudaMalloc((void**)&h->gramA, sizeof(double) * K * K));
  // gramS carries 4 extra doubles: the S-block norms of a sharded PGM iteration travel in the same all-reduce
  PMX_CUDA(cudaMalloc((void**)&h->gramS, sizeof(double) * (K * K + 4)));
  PMX_CUDA(cudaMemset(h->gramS, 0, sizeof(double) * (K * K + 4)));
  PMX_CUDA(cudaMalloc((void**)&h->ctl, sizeof(pmx_ctl)));
  PMX_CUDA(cudaMemsetAsync(h->ctl, 0, sizeof(pmx_ctl), ctx->stream));
  PMX_CUDA(cudaMemsetAsync(h->Y, 0, sizeof(float) * (size_t)pmx_div_up(M, 128) * h->ldY * 16384, ctx->stream));
  PMX_CUDA(cudaMemsetAsync(h->GA, 0, sizeof(float) * mk, ctx->stream));
  PMX_CUDA(cudaMemsetAsync(h->GS, 0, sizeof(float) * kn, ctx->stream));
  PMX_CUDA(cudaMallocHost((void**)&h->h_ctl, sizeof(pmx_ctl)));
  *out = h;
  return PMX_OK;
}

int pmx_nmf_destroy(pmx_nmf* h) {
  if (!h) return PMX_OK;
  cudaSetDevice(h->ctx->device);
  cudaStreamSynchronize(h->ctx->stream);
  cudaStreamSynchronize(h->ctx->aux);
  if (h->gram_part) pmx_dev_free(h->ctx, h->gram_part);
  float* bufs[] = {h->Y, h->W, h->A, h->S, h->A_old, h->S_old, h->Ae, h->Se, h->GA, h->GS, h->MA, h->MS, h->VA, h->VS,
                   h->VhA, h->VhS, h->Psi, h->Z0, h->Z1, h->alphaA, h->alphaS};
  for (float* b : bufs)
    if (b) pmx_dev_free(h->ctx, b);
  for (int j = 0; j < 2; ++j)
    for (int i = 0; i < 4; ++i) {
      if (h->Zg[j][i]) pmx_dev_free(h->ctx, h->Zg[j][i]);
      if (h->Ug[j][i]) pmx_dev_free(h->ctx, h->Ug[j][i]);
    }
  if (h->bs_norms) cudaFree(h->bs_norms);
  cudaFree(h->gramA);
  cudaFree(h->gramS);
  cudaFree(h->ctl);
  cudaFreeHost(h->h_ctl);
  if (h->pgm_graph) cudaGraphExecDestroy(h->pgm_graph);
  if (h->tail_graph) cudaGraphExecDestroy(h->tail_graph);
  if (h->GA2) pmx_dev_free(h->ctx, h->GA2);
  if (h->GS2) pmx_dev_free(h->ctx, h->GS2);
  if (h->tail_gram_rep) cudaFree(h->tail_gram_rep);
  if (h->tail_acc) cudaFree(h->tail_acc);
  if (h->tail_tickets) cudaFree(h->tail_tickets);
  if (h->plan) umma_plan_destroy(h->ctx, h->plan);
  delete h;
  return PMX_OK;
}

static int upload_tiled(pmx_nmf* h, float* dst, const float* host_Y, size_t ld, int col0, int ncols) {
  PMX_REQUIRE(col0 >= 0 && ncols >= 0 && col0 + ncols <= h->N, "column range outside the local stripe");
  if (ncols == 0) return PMX_OK;
  // The device copy is tiled (grad_umma.h).  Row chunks travel through two row-major staging buffers:
  // the copy of chunk i + 1 (main stream) overlaps the interleave kernel of chunk i (side stream).
  pmx_ctx* ctx = h->ctx;
  const size_t pitch = ((size_t)ncols + 3) & ~(size_t)3;
  size_t rows = ((size_t)48 << 20) / (pitch * sizeof(float));    // ~48 MB per staging buffer
  rows = rows < 4 ? 4 : (rows & ~(size_t)3);
  if (rows > (size_t)((h->M + 3) & ~3)) rows = (size_t)((h->M + 3) & ~3);
  if (rows > 4 * 32768) rows = 4 * 32768;                          // grid.y of the interleave kernel
  float* stage[2] = {nullptr, nullptr};
  PMX_CHECK(alloc_f(ctx, &stage[0], rows * pitch));
  PMX_CHECK(alloc_f(ctx, &stage[1], rows * pitch));
  cudaEvent_t copied[2], used[2];
  for (int i = 0; i < 2; ++i) {
    PMX_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
    PMX_CUDA(cudaEventCreateWithFlags(&used[i], cudaEventDisableTiming));
  }
  int st = PMX_OK, chunk = 0;
  for (size_t m0 = 0; m0 < (size_t)h->M && st == PMX_OK; m0 += rows, ++chunk) {
    const int b = chunk & 1;
    const size_t nr = (size_t)h->M - m0 < rows ? (size_t)h->M - m0 : rows;
    cudaError_t e = cudaSuccess;
    if (chunk >= 2) e = cudaStreamWaitEvent(ctx->stream, used[b], 0);
    if (e == cudaSuccess)
      e = cudaMemcpy2DAsync(stage[b], sizeof(float) * pitch, host_Y + m0 * ld, sizeof(float) * ld, sizeof(float) * ncols, nr,
                            cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(copied[b], ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->aux, copied[b], 0);
    if (e != cudaSuccess) {
      pmx_set_error("pmx_nmf_set_Y / set_W: %s", cudaGetErrorString(e));
      st = PMX_ERR_CUDA;
      break;
    }
    st = launch_y_interleave(ctx, ctx->aux, stage[b], (int)pitch, (int)nr, ncols, dst, h->ldY, (int)m0, col0);
    if (st == PMX_OK && cudaEventRecord(used[b], ctx->aux) != cudaSuccess) st = PMX_ERR_CUDA;
  }
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->aux);
  for (int i = 0; i < 2; ++i) {
    cudaEventDestroy(copied[i]);
    cudaEventDestroy(used[i]);
  }
  pmx_dev_free(ctx, stage[0]);
  pmx_dev_free(ctx, stage[1]);
  return st;
}

int pmx_nmf_set_Y(pmx_nmf* h, const float* host_Y, size_t ld, int col0, int ncols) {
  PMX_REQUIRE(h && host_Y, "NULL argument");
  return upload_tiled(h, h->Y, host_Y, ld, col0, ncols);
}

int pmx_nmf_set_W(pmx_nmf* h, const float* host_W, size_t ld, int col0, int ncols) {
  PMX_REQUIRE(h && host_W, "NULL argument");
  if (!h->W) {
    const size_t n = (size_t)pmx_div_up(h->M, 128) * h->ldY * 16384;
    PMX_CHECK(alloc_f(h->ctx, &h->W, n));
    PMX_CUDA(cudaMemsetAsync(h->W, 0, sizeof(float) * n, h->ctx->stream));
    if (h->plan) PMX_CHECK(umma_plan_set_W(h->plan, h->W));
  }
  return upload_tiled(h, h->W, host_W, ld, col0, ncols);
}

int pmx_nmf_gradient(pmx_nmf* h, double* loss_host_or_null) {
  PMX_REQUIRE(h != nullptr, "NULL argument");
  PMX_CHECK(tail_detach(h));
  h->tail_G_pending = false;
  double* dl = loss_host_or_null ? &h->ctl->norms[6] : nullptr;
  PMX_CHECK(nmf_gradient(h, h->A, h->S, h->GA, h->GS, dl, 0, nullptr));
  if (loss_host_or_null) {
    PMX_CHECK(pull_ctl(h));
    *loss_host_or_null = h->h_ctl->norms[6];
  } else {
    PMX_CUDA(cudaStreamSynchronize(h->ctx->stream));
  }
  return PMX_OK;
}

static int which_ptr(pmx_nmf* h, int which, float** p, size_t* n) {
  const size_t mk = (size_t)h->M * h->K, kn = (size_t)h->K * h->N;
  switch (which) {
    case PMX_A: *p = h->A; *n = mk; break;
    case PMX_S: *p = h->S; *n = kn; break;
    case PMX_GA: *p = h->GA; *n = mk; break;
    case PMX_GS: *p = h->GS; *n = kn; break;
    case PMX_MA: *p = h->MA; *n = mk; break;
    case PMX_MS: *p = h->MS; *n = kn; break;
    case PMX_VA: *p = h->VA; *n = mk; break;
    case PMX_VS: *p = h->VS; *n = kn; break;
    case PMX_VHA: *p = h->VhA; *n = mk; break;
    case PMX_VHS: *p = h->VhS; *n = kn; break;
    default: pmx_set_error("unknown buffer id %d", which); return PMX_ERR_ARG;
  }
  if (!*p) {
    pmx_set_error("buffer %d is not allocated in this solver state", which);
    return PMX_ERR_ARG;
  }
  return PMX_OK;
}

int pmx_nmf_set(pmx_nmf* h, int which, const float* host_src) {
  PMX_REQUIRE(h && host_src, "NULL argument");
  float* p; size_t n;
  PMX_CHECK(which_ptr(h, which, &p, &n));
  if (which == PMX_A || which == PMX_S) {
    h->split_valid = false;
    h->gramS_valid = false;
    h->gram_pending = false;
    h->tail_ready = false;
  }
  return pmx_h2d(h->ctx, p, host_src, n * sizeof(float));
}

int pmx_nmf_get(pmx_nmf* h, int which, float* host_dst) {
  PMX_REQUIRE(h && host_dst, "NULL argument");
  float* p; size_t n;
  PMX_CHECK(which_ptr(h, which, &p, &n));
  if (which == PMX_GA || which == PMX_GS) PMX_CHECK(tail_publish_G(h));
  return pmx_d2h(h->ctx, host_dst, p, n * sizeof(float));
}

int pmx_nmf_device_ptr(pmx_nmf* h, int which, float** dev_ptr) {
  PMX_REQUIRE(h && dev_ptr, "NULL argument");
  size_t n;
  if (which == PMX_GA || which == PMX_GS) PMX_CHECK(tail_publish_G(h));
  return which_ptr(h, which, dev_ptr, &n);
}

// A sharded fused-tail solve leaves the plan's A operands pointing into the peer arena (every rank's tail pushes its
// rows there).  Anything that evaluates the gradient kernel outside that loop must not depend on the arena -- another
// solve may re-lay it out: back to the plan's own buffers, operands re-split from h->A on the next launch.
static int tail_detach(pmx_nmf* h) {
  if (h->tail_mode && h->ctx->world > 1 && h->plan) {
    PMX_CHECK(umma_plan_use_A(h->plan, nullptr, nullptr));
    h->split_valid = false;
    h->tail_ready = false;
  }
  return PMX_OK;
}

int pmx_nmf_loss(pmx_nmf* h, double* loss_host) {
  PMX_REQUIRE(h && loss_host, "NULL argument");
  PMX_CHECK(tail_detach(h));
  // the loss is a by-product of the residual pass: the tcgen05 kernel runs it alone (no gradient GEMMs, no flush,
  // no scratch); the SIMT kernel (tiny problems) still produces both gradients into scratch buffers
  int st;
  if (nmf_uses_umma(h, 0)) {
    st = nmf_gradient(h, h->A, h->S, nullptr, nullptr, &h->ctl->norms[6], 0, nullptr, false, nullptr, 0, 0);
    if (st == PMX_OK) st = pull_ctl(h);
  } else {
    float *ga, *gs;
    PMX_CHECK(alloc_f(h->ctx, &ga, (size_t)h->M * h->K));
    PMX_CHECK(alloc_f(h->ctx, &gs, (size_t)h->K * h->N));
    st = nmf_gradient(h, h->A, h->S, ga, gs, &h->ctl->norms[6], 0, nullptr);
    if (st == PMX_OK) st = pull_ctl(h);
    pmx_dev_free(h->ctx, ga);
    pmx_dev_free(h->ctx, gs);
  }
  *loss_host = h->h_ctl->norms[6];
  return st;
}

// debug (env PMX_STAGE_TIMES): CUDA-event timeline of one PGM iteration, averaged and printed every 50 iterations
struct StageTimer {
  bool on;
  cudaEvent_t ev[10];
  double acc[10];
  int n;
  StageTimer() : on(getenv("PMX_STAGE_TIMES") != nullptr), n(0) {
    for (int i = 0; i < 10; ++i) { ev[i] = nullptr; acc[i] = 0; }
  }
  void mark(int i, cudaStream_t st) {
    if (!on) return;
    if (!ev[i]) cudaEventCreate(&ev[i]);
    cudaEventRecord(ev[i], st);
  }
  void finish(pmx_ctx* ctx) {
    if (!on) return;
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->aux);
    static const char* names[10] = {"start", "aux:steps done", "zero+grad done", "join steps", "aux:AR(G_A)+A update done",
                                    "S update done", "gram reduce + AR2 done", "finalize done", "", ""};
    for (int i = 1; i < 8; ++i) {
      float ms = 0;
      if (ev[i] && cudaEventElapsedTime(&ms, ev[0], ev[i]) == cudaSuccess) acc[i] += ms;
    }
    if (++n % 50 == 0) {
      for (int i = 1; i < 8; ++i) printf("STAGE %-28s at %8.1f us\n", names[i], 1e3 * acc[i] / 50), acc[i] = 0;
      fflush(stdout);
    }
  }
};
static StageTimer g_stage;

// ------------------------------------------------------------------ fused PGM tail: host side
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// decides whether this PGM solve takes the fused tail (pgm_tail.cu) and sets up its buffers
static int tail_setup(pmx_nmf* h) {
  pmx_ctx* ctx = h->ctx;
  h->tail_mode = false;
  h->tail_ready = false;
  h->tail_G_pending = false;
  if (h->tail_graph) {
    cudaGraphExecDestroy(h->tail_graph);
    h->tail_graph = nullptr;
  }
  if (h->plan) PMX_CHECK(umma_plan_use_A(h->plan, nullptr, nullptr));
  const int axA = chain_unity_axis(h->chA), axS = chain_unity_axis(h->chS);
  const bool ok = !getenv("PMX_NO_FUSED_TAIL") && !h->pgm.accelerated && nmf_uses_umma(h, h->pgm.kernel) && h->K <= 64 &&
                  axA == -1 && (axS == -1 || axS == 0) && (ctx->world == 1 || pmx_peer_available(ctx)) &&
                  h->M >= ctx->world;
  if (!ok) return PMX_OK;
  const size_t mk = (size_t)h->M * h->K, kn = (size_t)h->K * h->N, kk = (size_t)h->K * h->K;
  h->ga_stride = (mk + 63) & ~(size_t)63;
  h->gs_stride = (kn + 63) & ~(size_t)63;
  if (ctx->world > 1) {
    const size_t Mp = (size_t)pmx_div_up(h->M, 128) * 128;
    h->off_GA = 0;
    h->off_A = align256(h->off_GA + 2 * h->ga_stride * sizeof(float));
    h->off_Ahi = align256(h->off_A + mk * sizeof(float));
    h->off_Alo = align256(h->off_Ahi + Mp * 64 * 2);
    h->off_inbox = align256(h->off_Alo + Mp * 64 * 2);
    const size_t total = h->off_inbox + sizeof(double) * (kk + 4) * PMX_MAX_WORLD * 4;
    pmx_peer_region* ar = nullptr;
    if (pmx_peer_arena(ctx, total, &ar) != PMX_OK) return PMX_OK;   // no symmetric memory: the NCCL path takes over
    PMX_CHECK(pmx_peer_reset(ctx));
  } else {
    if (!h->GA2) PMX_CHECK(alloc_f(ctx, &h->GA2, 2 * h->ga_stride));
    PMX_CUDA(cudaMemsetAsync(h->GA2, 0, sizeof(float) * 2 * h->ga_stride, ctx->stream));
  }
  if (!h->GS2) PMX_CHECK(alloc_f(ctx, &h->GS2, 2 * h->gs_stride));
  PMX_CUDA(cudaMemsetAsync(h->GS2, 0, sizeof(float) * 2 * h->gs_stride, ctx->stream));
  if (!h->tail_gram_rep) PMX_CUDA(cudaMalloc((void**)&h->tail_gram_rep, sizeof(float) * 2 * PMX_TAIL_NREP * kk));
  if (!h->tail_acc) PMX_CUDA(cudaMalloc((void**)&h->tail_acc, sizeof(double) * 8));
  if (!h->tail_tickets) PMX_CUDA(cudaMalloc((void**)&h->tail_tickets, sizeof(unsigned) * 4));
  PMX_CUDA(cudaMemsetAsync(h->tail_gram_rep, 0, sizeof(float) * 2 * PMX_TAIL_NREP * kk, ctx->stream));
  PMX_CUDA(cudaMemsetAsync(h->tail_acc, 0, sizeof(double) * 8, ctx->stream));
  PMX_CUDA(cudaMemsetAsync(h->tail_tickets, 0, sizeof(unsigned) * 4, ctx->stream));
  h->tail_mode = true;
  return PMX_OK;
}

static float* tail_ga_base(pmx_nmf* h) {
  return h->ctx->world > 1 ? reinterpret_cast<float*>(static_cast<char*>(h->ctx->peer_arena.local) + h->off_GA) : h->GA2;
}

// before the first fused iteration (and after the host replaced A or S): bf16 operands, arena copies of A, and the
// steps of the coming iteration from the stand-alone Gram / lambda_max kernels (nmf.py:44-49)
static int tail_prologue(pmx_nmf* h) {
  pmx_ctx* ctx = h->ctx;
  const int* done = &h->ctl->done;
  if (!h->plan) PMX_CHECK(umma_plan_create(ctx, h->Y, h->ldY, h->M, h->N, h->K, &h->plan, 1));
  PMX_CHECK(umma_plan_set_W(h->plan, h->W));
  if (ctx->world > 1) {
    char* loc = static_cast<char*>(ctx->peer_arena.local);
    PMX_CHECK(umma_plan_use_A(h->plan, loc + h->off_Ahi, loc + h->off_Alo));
    PMX_CUDA(cudaMemcpyAsync(loc + h->off_A, h->A, sizeof(float) * (size_t)h->M * h->K, cudaMemcpyDeviceToDevice, ctx->stream));
  } else {
    PMX_CHECK(umma_plan_use_A(h->plan, nullptr, nullptr));
  }
  void *Ahi, *Alo, *Shi, *Slo;
  int ldA, ldS;
  umma_plan_buffers(h->plan, &Ahi, &Alo, &Shi, &Slo, &ldA, &ldS);
  const int Mp = pmx_div_up(h->M, 128) * 128;
  PMX_CHECK(launch_split_bf16(ctx, h->A, h->M, h->K, Ahi, Alo, Mp, ldA, done));
  PMX_CHECK(launch_split_bf16(ctx, h->S, h->K, h->N, Shi, Slo, ldA, ldS, done));
  h->gramS_valid = false;
  h->gram_pending = false;
  PMX_CHECK(nmf_steps(h, h->A, h->S, true, true));
  PMX_CHECK(nmf_steps_join(h));
  PMX_CHECK(launch_tail_seed_steps(ctx, h->ctl));
  h->split_valid = true;
  h->used_umma = true;
  h->tail_ready = true;
  return PMX_OK;
}

// argument block of the two tail kernels (fixed for a solve: the kernels replay inside the iteration's CUDA graph)
static int tail_args(pmx_nmf* h, PgmTailArgs* out) {
  pmx_ctx* ctx = h->ctx;
  PgmTailArgs& a = *out;
  memset(&a, 0, sizeof(a));
  a.M = h->M; a.N = h->N; a.K = h->K;
  a.world = ctx->world; a.rank = ctx->rank;
  a.ctl = h->ctl;
  void *Ahi, *Alo, *Shi, *Slo;
  int ldA, ldS;
  umma_plan_buffers(h->plan, &Ahi, &Alo, &Shi, &Slo, &ldA, &ldS);
  a.S = h->S; a.GS2 = h->GS2; a.gs_stride = (long long)h->gs_stride;
  a.Shi = (unsigned short*)Shi; a.Slo = (unsigned short*)Slo; a.ldS = ldS;
  a.chS = h->chS;
  a.nS = pgm_tail_s_blocks(ctx, h->N, &a.n_tiles_S);
  // rows of A this rank updates (reduce-scatter slice): an even split of the M rows, a.ra rows per block
  a.m_lo = (int)((long long)h->M * ctx->rank / ctx->world);
  a.m_hi = (int)((long long)h->M * (ctx->rank + 1) / ctx->world);
  // (a sharded run has few rows per rank and every A block starts with remote loads: smaller blocks keep at least
  // ~64 of them in flight so that the exchange is one NVLink round trip deep instead of several)
  a.ra = PMX_TAIL_RA;
  while (a.ra > 8 && pmx_div_up(a.m_hi - a.m_lo, a.ra) < 64) a.ra >>= 1;
  a.nA = pmx_div_up(a.m_hi - a.m_lo, a.ra);
  a.chA = h->chA;
  a.ldA = ldA;
  a.ga_stride = (long long)h->ga_stride;
  a.A_loc = h->A; a.GA2_loc = h->GA2; a.Ahi_loc = (unsigned short*)Ahi; a.Alo_loc = (unsigned short*)Alo;
  if (ctx->world > 1) {
    for (int r = 0; r < PMX_MAX_WORLD; ++r) {
      a.arena.p[r] = ctx->peer_arena.peer[r];
      a.flags.p[r] = ctx->peer_flags.peer[r];
    }
    a.off_GA = h->off_GA; a.off_A = h->off_A; a.off_Ahi = h->off_Ahi; a.off_Alo = h->off_Alo; a.off_inbox = h->off_inbox;
    a.epoch = ctx->peer_epoch;
    a.my_flags = reinterpret_cast<const unsigned*>(ctx->peer_flags.local);
  }
  a.gram_rep = h->tail_gram_rep; a.acc = h->tail_acc; a.tickets = h->tail_tickets;
  const float eA = h->pgm.e_rel_A, eS = h->pgm.e_rel_S;
  a.e2A = (float)((double)eA * (double)eA);
  a.e2S = (float)((double)eS * (double)eS);
  if (a.nA < 1) {   // more ranks than rows of A: not a shape the fused tail is meant for
    pmx_set_error("fused PGM tail: M = %d is too small for %d ranks", h->M, ctx->world);
    return PMX_ERR_UNSUPPORTED;
  }
  return PMX_OK;
}

// One iteration:   side stream: k_tail_final of the PREVIOUS iteration (Gram totals, lambda_max -> the steps this
//                               iteration's tail uses, convergence test) -- next to the gradient kernel, which leaves
//                               it one SM
//                  main stream: gradient kernel -> [peer signal] -> (join) -> k_pgm_tail
static int tail_enqueue_iteration(pmx_nmf* h) {
  pmx_ctx* ctx = h->ctx;
  const int* done = &h->ctl->done;
  PgmTailArgs a;
  PMX_CHECK(tail_args(h, &a));
  PMX_CUDA(cudaEventRecord(ctx->ev_fork, ctx->stream));
  PMX_CUDA(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork, 0));
  PMX_CHECK(launch_tail_final(ctx, ctx->aux, a));
  PMX_CUDA(cudaEventRecord(ctx->ev_join, ctx->aux));
  PMX_CHECK(launch_grad_umma(ctx, h->plan, h->A, h->S, tail_ga_base(h), h->GS2, nullptr, done, 1, &h->ctl->par_ctr,
                             h->ga_stride, 3, h->gs_stride, /*reserve_sms=*/1));
  // the signal is unconditional (no `done` test): the stop flag may be raised by the concurrent k_tail_final while
  // this stream passes here, and the epoch counters of the ranks must stay in lockstep
  if (ctx->world > 1) PMX_CHECK(pmx_peer_signal(ctx, 0, ctx->stream, nullptr, nullptr, 0, 0));
  PMX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  PMX_CHECK(launch_pgm_tail(ctx, a));
  h->it_enqueued += 1;
  return PMX_OK;
}

// closes the last enqueued iteration (no-op when nothing is pending): before the host looks at the control block
static int tail_close(pmx_nmf* h) {
  PgmTailArgs a;
  PMX_CHECK(tail_args(h, &a));
  return launch_tail_final(h->ctx, h->ctx->stream, a);
}

// after a run: sharded, the replicated A leaves the arena for the plain buffer that pmx_nmf_get and the other solvers
// read.  The last gradients (algorithms.py:144 hands them back) stay in their parity buffers until somebody asks for
// them (tail_publish_G, from pmx_nmf_get / pmx_nmf_device_ptr): two 16 MB copies and, sharded, an all-reduce of the
// G_A partials that a loop which only polls the stop flag never needs.
static int tail_publish(pmx_nmf* h) {
  pmx_ctx* ctx = h->ctx;
  const size_t mk = (size_t)h->M * h->K;
  h->tail_G_pending = h->h_ctl->it > 0;
  h->tail_G_par = (size_t)(h->h_ctl->it & 1);
  if (ctx->world > 1 && h->tail_ready)
    PMX_CUDA(cudaMemcpyAsync(h->A, static_cast<char*>(ctx->peer_arena.local) + h->off_A, sizeof(float) * mk,
                             cudaMemcpyDeviceToDevice, ctx->stream));
  PMX_CUDA(cudaStreamSynchronize(ctx->stream));
  return PMX_OK;
}

static int tail_publish_G(pmx_nmf* h) {
  if (!h->tail_G_pending) return PMX_OK;
  pmx_ctx* ctx = h->ctx;
  const size_t mk = (size_t)h->M * h->K, kn = (size_t)h->K * h->N;
  const size_t par = h->tail_G_par;
  PMX_CUDA(cudaMemcpyAsync(h->GS, h->GS2 + par * h->gs_stride, sizeof(float) * kn, cudaMemcpyDeviceToDevice, ctx->stream));
  PMX_CUDA(cudaMemcpyAsync(h->GA, tail_ga_base(h) + par * h->ga_stride, sizeof(float) * mk, cudaMemcpyDeviceToDevice, ctx->stream));
  if (ctx->world > 1) PMX_CHECK(pmx_comm_allreduce_internal(ctx, h->GA, mk, 0, ctx->stream));
  PMX_CUDA(cudaStreamSynchronize(ctx->stream));
  h->tail_G_pending = false;
  return PMX_OK;
}

// ------------------------------------------------------------------ PGM
int pmx_nmf_pgm_begin(pmx_nmf* h, const pmx_pgm_opts* opts) {
  PMX_REQUIRE(h && opts, "NULL argument");
  h->pgm = *opts;
  h->split_valid = false;
  h->gramS_valid = false;
  h->gram_pending = false;
  if (h->pgm_graph) {
    cudaGraphExecDestroy(h->pgm_graph);
    h->pgm_graph = nullptr;
  }
  if (h->pgm.check_every <= 0) h->pgm.check_every = 8;
  h->chA = make_chain(&opts->prox_A);
  h->chS = make_chain(&opts->prox_S);
  PMX_CHECK(reject_sharded_row_unity(h, h->chS, "pgm"));
  h->nest_t = 1.0;  // utils.py:195
  h->it_enqueued = 0;
  if (opts->accelerated) {
    if (!h->Ae) PMX_CHECK(alloc_f(h->ctx, &h->Ae, (size_t)h->M * h->K));
    if (!h->Se) PMX_CHECK(alloc_f(h->ctx, &h->Se, (size_t)h->K * h->N));
  }
  PMX_CUDA(cudaMemsetAsync(h->ctl, 0, sizeof(pmx_ctl), h->ctx->stream));
  h->xchg_peer = false;
  PMX_CHECK(tail_setup(h));
  // sharded run over peer memory (comm.cu): [G_A partials x 2 | (Gram(S) partial, 3 norms, pad) x 2] in the arena
  h->peer_mode = false;
  if (!h->tail_mode && pmx_peer_available(h->ctx) && nmf_uses_umma(h, h->pgm.kernel) && umma_supported(h->M, h->N, h->K) &&
      !getenv("PMX_NO_PEER_PGM")) {
    const size_t mk = (size_t)h->M * h->K, kk4 = (size_t)h->K * h->K + 4;
    h->peer_off_gram = (2 * mk * sizeof(float) + 255) & ~(size_t)255;
    pmx_peer_region* ar = nullptr;
    if (pmx_peer_arena(h->ctx, h->peer_off_gram + 2 * kk4 * sizeof(double), &ar) == PMX_OK) {
      PMX_CHECK(pmx_peer_reset(h->ctx));
      h->peer_mode = true;
    }
  }
  return PMX_OK;
}

static int pgm_enqueue_iteration(pmx_nmf* h) {
  pmx_ctx* ctx = h->ctx;
  const int* done = &h->ctl->done;
  const size_t mk = (size_t)h->M * h->K, kn = (size_t)h->K * h->N;
  // Nesterov sequence (utils.py:198-206), in double like the reference's Python floats
  double omega = 0.0;
  if (h->pgm.accelerated) {
    const double t_next = 0.5 * (1.0 + sqrt(4.0 * h->nest_t * h->nest_t + 1.0));
    omega = (h->nest_t - 1.0) / t_next;
    h->nest_t = t_next;
  }
  const float* Ae = h->A;
  const float* Se = h->S;
  if (omega > 0.0) {  // algorithms.py:94-95
    PMX_CHECK(launch_extrapolate(ctx, h->A, h->A_old, h->Ae, mk, (float)omega, done));
    PMX_CHECK(launch_extrapolate(ctx, h->S, h->S_old, h->Se, kn, (float)omega, done));
    Ae = h->Ae;
    Se = h->Se;
  }
  g_stage.mark(0, ctx->stream);
  // (the norms were zeroed by pgm_begin / the previous iteration's finalize)
  // steps on the side stream, gradient on the main stream (both read the same point, algorithms.py:105-106)
  PMX_CHECK(nmf_steps(h, Ae, Se, true, true));
  // sharded: the sum of the G_A partials runs on the side stream next to the S update -- a one-shot sum over peer
  // memory (comm.cu) or, as the fallback, an NCCL all-reduce on its own communicator
  const bool peer = h->peer_mode && ctx->world > 1;
  const bool ga_on_aux = peer || pmx_comm_has_aux(ctx);
  if (peer) {
    PMX_CHECK(nmf_gradient(h, Ae, Se, static_cast<float*>(ctx->peer_arena.local), h->GS, nullptr, h->pgm.kernel, done,
                           true, ctx->peer_epoch + 0, mk));
    PMX_CHECK(pmx_peer_signal(ctx, 0, ctx->stream, done, nullptr, 0, 0));
  } else {
    PMX_CHECK(nmf_gradient(h, Ae, Se, h->GA, h->GS, nullptr, h->pgm.kernel, done, ga_on_aux));
  }
  g_stage.mark(2, ctx->stream);
  PMX_CHECK(nmf_steps_join(h));
  g_stage.mark(3, ctx->stream);
  // X[j][:] = prox[j](_X[j] - S[j]*G[j], S[j])   (algorithms.py:107-108) + norms (:130-133) + X_ copy (:102)
  UpdIO io;
  memset(&io, 0, sizeof(io));
  io.done = done;
  io.step.mode = 1;
  io.step.scale = 1.f;
  // when the tcgen05 kernel is in use (and the next gradient is taken at (A, S) itself, i.e. no extrapolation) the
  // update kernels also write the bf16 (hi, lo) operands of the next iteration's GEMMs
  void *Ahi = nullptr, *Alo = nullptr, *Shi = nullptr, *Slo = nullptr;
  int ldA = 64, ldS = 0;
  const bool fuse_split = h->used_umma && h->plan && !h->pgm.accelerated;
  if (fuse_split) umma_plan_buffers(h->plan, &Ahi, &Alo, &Shi, &Slo, &ldA, &ldS);
  io.Xin = Ae; io.G = h->GA; io.Xprev = h->A; io.Xout = h->A; io.Xold_out = h->A_old;
  io.norms = &h->ctl->norms[0]; io.rows = h->M; io.cols = h->K; io.step.ptr = &h->ctl->step[0];
  io.hi = (unsigned short*)Ahi; io.lo = (unsigned short*)Alo; io.ld_split = ldA;
  {  // the two block updates are independent (Jacobi, algorithms.py:105-108): A on the side stream, S on the main one
    PMX_CUDA(cudaEventRecord(ctx->ev_fork2, ctx->stream));
    PMX_CUDA(cudaStreamWaitEvent(ctx->aux, ctx->ev_fork2, 0));
    if (peer) PMX_CHECK(pmx_peer_sum(ctx, 0, 0, mk, h->GA, 0, ctx->aux, done, &h->ctl->fault));
    else if (ga_on_aux) PMX_CHECK(pmx_comm_allreduce_internal(ctx, h->GA, mk, 0, ctx->aux));
    cudaStream_t main_stream = ctx->stream;
    ctx->stream = ctx->aux;
    const int st = launch_update(ctx, IN_PGM, h->chA, io);
    ctx->stream = main_stream;
    PMX_CHECK(st);
    g_stage.mark(4, ctx->aux);
  }
  io.Xin = Se; io.G = h->GS; io.Xprev = h->S; io.Xout = h->S; io.Xold_out = h->S_old;
  io.norms = &h->ctl->norms[3]; io.rows = h->K; io.cols = h->N; io.step.ptr = &h->ctl->step[1];
  const bool fuse_gram_pre = !h->pgm.accelerated && h->K <= 64 && chain_unity_axis(h->chS) == 0;
  if (ctx->world > 1 && fuse_gram_pre) io.norms = h->gramS + (size_t)h->K * h->K;   // packed with the Gram all-reduce
  io.hi = (unsigned short*)Shi; io.lo = (unsigned short*)Slo; io.ld_split = ldS;
  // S S^T of the new S as a by-product of the update (column-owner kernel only): the next iteration's step_A
  const bool fuse_gram = !h->pgm.accelerated && h->K <= 64 && chain_unity_axis(h->chS) == 0;
  const int nblk = upd_cols_blocks(ctx, h->N);
  if (fuse_gram) {
    if (!h->gram_part) PMX_CHECK(alloc_f(h->ctx, &h->gram_part, (size_t)nblk * h->K * h->K));
    io.gram_part = h->gram_part;
  }
  PMX_CHECK(launch_update(ctx, IN_PGM, h->chS, io));
  g_stage.mark(5, ctx->stream);
  h->split_valid = fuse_split;
  h->gramS_valid = fuse_gram;
  h->gram_pending = fuse_gram;     // single GPU: the partials are summed on the side stream at the start of the next iteration
  h->gram_pending_blocks = nblk;
  double* normsS = nullptr;
  if (ctx->world > 1) {
    // sharded: one all-reduce carries [S S^T partial sums | the three S-block norms] (an NCCL kernel on the side
    // stream would have to wait for an SM next to the persistent gradient kernel anyway)
    const size_t kk = (size_t)h->K * h->K;
    if (fuse_gram) {
      PMX_CHECK(launch_gram_reduce(ctx, ctx->stream, h->gram_part, nblk, h->K, h->gramS, done));
      h->gram_pending = false;
      normsS = h->gramS + kk;
      if (peer) {
        PMX_CHECK(pmx_peer_signal(ctx, 1, ctx->stream, done, h->gramS, h->peer_off_gram, kk + 4));
        PMX_CHECK(pmx_peer_sum(ctx, 1, h->peer_off_gram, kk + 4, h->gramS, 1, ctx->stream, done, &h->ctl->fault));
      } else {
        PMX_CHECK(pmx_comm_allreduce_internal(ctx, h->gramS, kk + 3, 1, ctx->stream));
      }
    } else {
      PMX_CHECK(pmx_comm_allreduce_internal(ctx, &h->ctl->norms[3], 3, 1, ctx->stream));
    }
  }
  g_stage.mark(6, ctx->stream);
  // join the A update (side stream)
  PMX_CUDA(cudaEventRecord(ctx->ev_join2, ctx->aux));
  PMX_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_join2, 0));
  const float eA = h->pgm.e_rel_A, eS = h->pgm.e_rel_S;
  k_pgm_finalize<<<1, 1, 0, ctx->stream>>>(h->ctl, normsS, (float)((double)eA * (double)eA),
                                           (float)((double)eS * (double)eS));
  PMX_LAUNCHED(ctx);
  g_stage.mark(7, ctx->stream);
  g_stage.finish(ctx);
  h->it_enqueued += 1;
  return pmx_check_launch(ctx, "pgm iteration");
}

int pmx_nmf_pgm_run(pmx_nmf* h, int n_iter, int* iters_done, int* conv_A, int* conv_S, float* step_A, float* step_S) {
  PMX_REQUIRE(h != nullptr, "NULL handle");
  PMX_REQUIRE(n_iter >= 0, "n_iter must be >= 0");
  PMX_CHECK(pull_ctl(h));
  const int it0 = h->h_ctl->it;
  bool stopped = h->h_ctl->done != 0;
  static const bool no_graph = getenv("PMX_NO_GRAPH") != nullptr;
  if (h->tail_mode) {
    // fused tail: gradient kernel -> [peer signal] -> tail kernel on one stream, replayed as a CUDA graph
    if (!h->tail_ready && n_iter > 0 && !stopped) PMX_CHECK(tail_prologue(h));
    for (int i = 0; i < n_iter && !stopped; ++i) {
      static const bool tail_trace = getenv("PMX_TAIL_TRACE") != nullptr;
      if (!no_graph && !h->ctx->profile && !tail_trace) {
        if (!h->tail_graph) {
          cudaGraph_t graph = nullptr;
          const long long l0 = h->ctx->launches;
          PMX_CUDA(cudaStreamBeginCapture(h->ctx->stream, cudaStreamCaptureModeThreadLocal));
          int st = tail_enqueue_iteration(h);
          cudaError_t ce = cudaStreamEndCapture(h->ctx->stream, &graph);
          h->it_enqueued -= 1;               // the capture did not execute anything
          h->tail_graph_launches = h->ctx->launches - l0;
          h->ctx->launches = l0;
          if (st != PMX_OK) return st;
          if (ce != cudaSuccess || !graph) {
            pmx_set_error("CUDA graph capture of the PGM iteration failed: %s", cudaGetErrorString(ce));
            return PMX_ERR_CUDA;
          }
          PMX_CUDA(cudaGraphInstantiate(&h->tail_graph, graph, nullptr, nullptr, 0));
          cudaGraphDestroy(graph);
        }
        PMX_CUDA(cudaGraphLaunch(h->tail_graph, h->ctx->stream));
        h->ctx->launches += h->tail_graph_launches;
        h->it_enqueued += 1;
      } else {
        PMX_CHECK(tail_enqueue_iteration(h));
      }
      if ((i + 1) % h->pgm.check_every == 0 && i + 1 < n_iter) {
        PMX_CHECK(tail_close(h));
        PMX_CHECK(pull_ctl(h));
        stopped = h->h_ctl->done != 0;
      }
    }
    if (h->tail_ready) PMX_CHECK(tail_close(h));
    PMX_CHECK(pull_ctl(h));
    PMX_CHECK(tail_publish(h));
    n_iter = 0;   // (skips the unfused loop below; the common epilogue reports the results)
  }
  for (int i = 0; i < n_iter && !stopped; ++i) {
    // Steady state (tcgen05 kernel, operands split by the update kernels, no extrapolation, no per-launch
    // profiling): the iteration is a fixed kernel sequence on two streams (+ NCCL) -> replay it as a CUDA graph.
    const bool steady = !no_graph && !g_stage.on && !h->ctx->profile && !h->pgm.accelerated && h->split_valid && h->used_umma &&
                        h->it_enqueued >= 2;
    if (steady) {
      if (!h->pgm_graph) {
        cudaGraph_t graph = nullptr;
        const long long l0 = h->ctx->launches;
        PMX_CUDA(cudaStreamBeginCapture(h->ctx->stream, cudaStreamCaptureModeThreadLocal));
        int st = pgm_enqueue_iteration(h);
        cudaError_t ce = cudaStreamEndCapture(h->ctx->stream, &graph);
        h->it_enqueued -= 1;               // the capture did not execute anything
        h->pgm_graph_launches = h->ctx->launches - l0;
        h->ctx->launches = l0;
        if (st != PMX_OK) return st;
        if (ce != cudaSuccess || !graph) {
          pmx_set_error("CUDA graph capture of the PGM iteration failed: %s", cudaGetErrorString(ce));
          return PMX_ERR_CUDA;
        }
        PMX_CUDA(cudaGraphInstantiate(&h->pgm_graph, graph, nullptr, nullptr, 0));
        cudaGraphDestroy(graph);
      }
      PMX_CUDA(cudaGraphLaunch(h->pgm_graph, h->ctx->stream));
      h->ctx->launches += h->pgm_graph_launches;
      h->it_enqueued += 1;
    } else {
      PMX_CHECK(pgm_enqueue_iteration(h));
    }
    if ((i + 1) % h->pgm.check_every == 0 && i + 1 < n_iter) {
      PMX_CHECK(pull_ctl(h));
      stopped = h->h_ctl->done != 0;
    }
  }
  PMX_CHECK(pull_ctl(h));
  if (iters_done) *iters_done = h->h_ctl->it - it0;
  if (conv_A) *conv_A = h->h_ctl->conv[0];
  if (conv_S) *conv_S = h->h_ctl->conv[1];
  if (step_A) *step_A = h->h_ctl->step[0];
  if (step_S) *step_S = h->h_ctl->step[1];
  if (h->h_ctl->fault) {
    pmx_set_error("multi-GPU exchange timed out: a peer did not reach iteration %d", h->h_ctl->it);
    return PMX_ERR_NCCL;
  }
  if (h->h_ctl->nonfinite) {
    pmx_set_error("Gram matrix contains infs or NaNs (iteration %d)", h->h_ctl->it);
    return PMX_ERR_NONFINITE;
  }
  return PMX_OK;
}

// ------------------------------------------------------------------ adaprox (algorithms.py:248-423)
int pmx_nmf_adaprox_begin(pmx_nmf* h, const pmx_adaprox_opts* opts) {
  PMX_REQUIRE(h && opts, "NULL argument");
  PMX_REQUIRE(opts->scheme >= PMX_ADAM && opts->scheme <= PMX_RADAM, "unknown adaprox scheme");
  h->ada = *opts;
  h->tail_mode = false;
  if (h->plan) PMX_CHECK(umma_plan_use_A(h->plan, nullptr, nullptr));
  h->split_valid = false;
  h->gramS_valid = false;
  h->gram_pending = false;
  h->chA = make_chain(&opts->prox_A);
  h->chS = make_chain(&opts->prox_S);
  PMX_CHECK(reject_sharded_row_unity(h, h->chS, "adaprox"));
  const size_t mk = (size_t)h->M * h->K, kn = (size_t)h->K * h->N;
  const size_t big = mk > kn ? mk : kn;
  float** bufs[] = {&h->MA, &h->VA, &h->MS, &h->VS};
  const size_t sizes[] = {mk, mk, kn, kn};
  for (int i = 0; i < 4; ++i) {
    if (!*bufs[i]) PMX_CHECK(alloc_f(h->ctx, bufs[i], sizes[i]));
    PMX_CUDA(cudaMemsetAsync(*bufs[i], 0, sizeof(float) * sizes[i], h->ctx->stream));   // algorithms.py:348-355
  }
  if (opts->has_vhat) {
    if (!h->VhA) PMX_CHECK(alloc_f(h->ctx, &h->VhA, mk));
    if (!h->VhS) PMX_CHECK(alloc_f(h->ctx, &h->VhS, kn));
  }
  if (!h->Psi) PMX_CHECK(alloc_f(h->ctx, &h->Psi, big));
  if (!h->Z0) PMX_CHECK(alloc_f(h->ctx, &h->Z0, big));
  if (!h->Z1) PMX_CHECK(alloc_f(h->ctx, &h->Z1, big));
  if (!h->alphaA) PMX_CHECK(alloc_f(h->ctx, &h->alphaA, h->K));
  if (!h->alphaS) PMX_CHECK(alloc_f(h->ctx, &h->alphaS, h->K));
  if (!h->bs_norms) PMX_CUDA(cudaMalloc((void**)&h->bs_norms, sizeof(double) * 256));
  h->ada_it = 0;
  h->ada_spec[0] = h->ada_spec[1] = 3;
  PMX_CUDA(cudaMemsetAsync(h->ctl, 0, sizeof(pmx_ctl), h->ctx->stream));
  PMX_CHECK(nmf_global_cols(h));
  PMX_CHECK(xchg_setup(h, opts->kernel));
  return PMX_OK;
}

// row means of S over ALL columns (nmf.py:91-93): local row sums, all-reduce, divide by the global column count
static int launch_alpha_means_global(pmx_nmf* h) {
  pmx_ctx* ctx = h->ctx;
  double* sums = h->bs_norms + 128;
  PMX_CHECK(launch_axis_sum(ctx, h->S, h->K, h->N, 1, sums, &h->ctl->done));
  PMX_CHECK(nmf_allreduce(h, sums, (size_t)h->K, 1, ctx->stream, 0));
  return launch_alpha_from_sums(ctx, sums, h->K, h->N_global, h->alphaS, &h->ctl->done);
}

// commit of block j (algorithms.py:398-400) with the bf16 split of the new block for the next gradient kernel
static int adaprox_commit(pmx_nmf* h, int j) {
  const int rows = j == 0 ? h->M : h->K, cols = j == 0 ? h->K : h->N;
  float* X = j == 0 ? h->A : h->S;
  unsigned short *hi = nullptr, *lo = nullptr;
  int ld = 0;
  if (h->plan && h->ada_fuse_split) {
    void *Ahi, *Alo, *Shi, *Slo;
    int ldA, ldS;
    umma_plan_buffers(h->plan, &Ahi, &Alo, &Shi, &Slo, &ldA, &ldS);
    hi = (unsigned short*)(j == 0 ? Ahi : Shi);
    lo = (unsigned short*)(j == 0 ? Alo : Slo);
    ld = j == 0 ? ldA : ldS;
  }
  return launch_sub_commit(h->ctx, X, h->Z0, h->Z1, (size_t)rows * cols, h->ctl, j, hi, lo, cols, ld);
}

// one proximal sub-iteration of block j (algorithms.py:387-389); `enq` = sub-iterations enqueued before this one
static int adaprox_sub_enqueue(pmx_nmf* h, int j, int enq) {
  pmx_ctx* ctx = h->ctx;
  pmx_ctl* ctl = h->ctl;
  const pmx_adaprox_opts& o = h->ada;
  const int rows = j == 0 ? h->M : h->K, cols = j == 0 ? h->K : h->N;
  float* X = j == 0 ? h->A : h->S;
  StepSpec alpha;
  memset(&alpha, 0, sizeof(alpha));
  alpha.scale = 1.f;
  if (o.step_mode == 0) {
    alpha.ptr = j == 0 ? h->alphaA : h->alphaS;
    alpha.mode = j == 0 ? 2 : 3;
  } else {
    alpha.mode = 0;
    alpha.value = j == 0 ? o.alpha_A : o.alpha_S;
  }
  const ProxChain& ch = j == 0 ? h->chA : h->chS;
  const float e = j == 0 ? o.e_rel_A : o.e_rel_S;
  const float e2 = (float)((double)e * (double)e);
  UpdIO io;
  memset(&io, 0, sizeof(io));
  const bool odd = enq & 1;
  io.Xin = odd ? h->Z1 : h->Z0;
  io.Xprev = io.Xin;
  io.Xout = odd ? h->Z0 : h->Z1;
  io.X0 = X;
  io.G = h->Psi;
  io.psimax = &ctl->psi_max[j];
  io.norms = &ctl->norms[8];
  io.done = &ctl->done;
  io.done2 = &ctl->sub_done;
  io.rows = rows; io.cols = cols;
  io.step = alpha;
  PMX_CHECK(launch_update(ctx, IN_ADASUB, ch, io));
  if (j == 1)   // the sub-iteration stop test is a global norm (algorithms.py:389)
    PMX_CHECK(nmf_allreduce(h, &ctl->norms[8], 3, 1, ctx->stream, 2));
  return launch_sub_finalize(ctx, ctl, e2, o.prox_max_iter);
}

// Block j of an adaprox iteration (algorithms.py:375-400), enqueued WITHOUT a host round trip: moments, then `spec`
// speculative sub-iterations (kernels that return once the stopping rule of :389 fired), then the commit -- which
// freezes the solve (done = 2) if the rule has not fired yet; pmx_nmf_adaprox_run then finishes the block with
// adaprox_block_resume.  spec follows the count the block needed last time.
static int adaprox_block(pmx_nmf* h, int j, int it, double b1, double b1_prev, int spec) {
  pmx_ctx* ctx = h->ctx;
  pmx_ctl* ctl = h->ctl;
  const pmx_adaprox_opts& o = h->ada;
  const int rows = j == 0 ? h->M : h->K, cols = j == 0 ? h->K : h->N;
  const size_t n = (size_t)rows * cols;
  float* X = j == 0 ? h->A : h->S;
  StepSpec alpha;
  memset(&alpha, 0, sizeof(alpha));
  alpha.scale = 1.f;
  if (o.step_mode == 0) {   // nmf.py:91-93: A gets a K-vector (per column), S a K x 1 column (per row)
    alpha.ptr = j == 0 ? h->alphaA : h->alphaS;
    alpha.mode = j == 0 ? 2 : 3;
  } else {
    alpha.mode = 0;
    alpha.value = j == 0 ? o.alpha_A : o.alpha_S;
  }
  PMX_CHECK(launch_sub_begin(ctx, ctl, j));
  AdaArgs a;
  memset(&a, 0, sizeof(a));
  a.G = j == 0 ? h->GA : h->GS;
  a.M = j == 0 ? h->MA : h->MS;
  a.V = j == 0 ? h->VA : h->VS;
  a.Vhat = o.has_vhat ? (j == 0 ? h->VhA : h->VhS) : nullptr;
  a.X = X;
  a.Psi = h->Psi;
  a.Z = h->Z0;
  a.Xold = o.check_convergence ? (j == 0 ? h->A_old : h->S_old) : nullptr;
  a.psimax = &ctl->psi_max[j];
  a.done = &ctl->done;
  a.n = n; a.rows = rows; a.cols = cols;
  a.alpha = alpha;
  a.scheme = o.scheme;
  a.b1 = b1; a.b1_prev = b1_prev; a.b2 = o.b2; a.eps = o.eps; a.p = o.p;
  a.t = it + 1;
  PMX_CHECK(launch_adaprox_moments(ctx, a));
  // column-sharded S block: max(Psi) is a maximum over all ranks (algorithms.py:384)
  if (j == 1) PMX_CHECK(nmf_allreduce(h, &ctl->psi_max[1], 1, 2, ctx->stream, 1));
  const bool has_prox = j == 0 ? o.has_prox_A : o.has_prox_S;
  if (!has_prox) return PMX_OK;   // algorithms.py:380
  if (spec > o.prox_max_iter) spec = o.prox_max_iter;
  for (int enq = 0; enq < spec; ++enq) PMX_CHECK(adaprox_sub_enqueue(h, j, enq));
  h->ada_enq[j] = spec;
  return adaprox_commit(h, j);
}

// the solve is frozen inside block j (done == 2): continue its sub-iterations with a look at the flag every two, commit
static int adaprox_block_resume(pmx_nmf* h, int j) {
  pmx_ctx* ctx = h->ctx;
  const pmx_adaprox_opts& o = h->ada;
  const int rows = j == 0 ? h->M : h->K, cols = j == 0 ? h->K : h->N;
  float* X = j == 0 ? h->A : h->S;
  PMX_CHECK(launch_clear_pause(ctx, h->ctl));
  int enq = h->ada_enq[j];
  while (enq < o.prox_max_iter) {
    for (int rep = 0; rep < 2 && enq < o.prox_max_iter; ++rep, ++enq) PMX_CHECK(adaprox_sub_enqueue(h, j, enq));
    PMX_CHECK(pull_ctl(h));
    if (h->h_ctl->sub_done || h->h_ctl->done) break;
  }
  (void)rows; (void)cols; (void)X; (void)ctx;
  return adaprox_commit(h, j);
}

// rest of an iteration after block j_done (exclusive): the remaining block, the convergence norms, the finalize step
static int adaprox_iteration_tail(pmx_nmf* h, int from_block, int it, double b1, double b1_prev) {
  pmx_ctx* ctx = h->ctx;
  const pmx_adaprox_opts& o = h->ada;
  for (int j = from_block; j < 2; ++j) PMX_CHECK(adaprox_block(h, j, it, b1, b1_prev, h->ada_spec[j]));
  if (o.check_convergence) {   // algorithms.py:403-410
    PMX_CHECK(launch_diff_norms(ctx, h->A, h->A_old, (size_t)h->M * h->K, &h->ctl->norms[0], &h->ctl->done));
    PMX_CHECK(launch_diff_norms(ctx, h->S, h->S_old, (size_t)h->K * h->N, &h->ctl->norms[3], &h->ctl->done));
    PMX_CHECK(nmf_allreduce(h, &h->ctl->norms[3], 3, 1, ctx->stream, 3));
  }
  const float eA = o.e_rel_A, eS = o.e_rel_S;
  return launch_adaprox_finalize(ctx, h->ctl, (float)((double)eA * eA), (float)((double)eS * eS), o.check_convergence);
}

int pmx_nmf_adaprox_run(pmx_nmf* h, int n_iter, const double* b1, const double* b1_prev, int* iters_done, int* conv_A,
                        int* conv_S, long long* sub_A, long long* sub_S) {
  PMX_REQUIRE(h && b1 && b1_prev, "NULL argument");
  pmx_ctx* ctx = h->ctx;
  const pmx_adaprox_opts& o = h->ada;
  PMX_CHECK(pull_ctl(h));
  const int it0 = h->h_ctl->it;
  const int ada_it0 = h->ada_it;
  // No host round trip inside an iteration: the control block is read every `poll` iterations.  A block whose
  // speculative sub-iterations did not suffice freezes the device side (done = 2, see k_sub_commit); the host then
  // finishes that block, re-enqueues the rest of its iteration and continues behind it (the frozen kernels of the
  // iterations that were already enqueued did nothing).
  const int poll = 4;
  h->ada_fuse_split = o.has_prox_A && o.has_prox_S && nmf_uses_umma(h, o.kernel);
  int i = 0;
  while (i < n_iter) {
    const int it = ada_it0 + i;
    PMX_CHECK(nmf_gradient(h, h->A, h->S, h->GA, h->GS, nullptr, o.kernel, &h->ctl->done));   // algorithms.py:369
    if (o.step_mode == 0) {                                                                     // algorithms.py:370
      PMX_CHECK(launch_alpha_means(ctx, h->A, h->M, h->K, 0, h->bs_norms, h->alphaA, &h->ctl->done));
      PMX_CHECK(launch_alpha_means_global(h));
    }
    PMX_CHECK(adaprox_iteration_tail(h, 0, it, b1[i], b1_prev[i]));
    h->split_valid = h->ada_fuse_split && h->used_umma && h->plan != nullptr;   // the commits wrote the bf16 operands
    ++i;
    if (i % poll != 0 && i < n_iter) continue;
    PMX_CHECK(pull_ctl(h));
    while (h->h_ctl->done == 2) {
      const int k = h->h_ctl->it - it0;          // iteration (index inside this call) the device froze in
      const int j = h->h_ctl->paused_block - 1;
      if (k < 0 || k >= n_iter || j < 0 || j > 1) {
        pmx_set_error("adaprox: inconsistent pause state (iteration %d, block %d)", k, j);
        return PMX_ERR_CUDA;
      }
      PMX_CHECK(adaprox_block_resume(h, j));
      PMX_CHECK(adaprox_iteration_tail(h, j + 1, ada_it0 + k, b1[k], b1_prev[k]));
      PMX_CHECK(pull_ctl(h));
      if (h->h_ctl->done != 2) i = k + 1;        // continue behind the repaired iteration
    }
    for (int j = 0; j < 2; ++j) {                // next time: what the block needed last time, plus a margin
      const int want = h->h_ctl->sub_last[j] + 1;
      h->ada_spec[j] = want < 2 ? 2 : (want > 64 ? 64 : want);
    }
    if (h->h_ctl->done) break;
  }
  h->ada_it = ada_it0 + (h->h_ctl->it - it0);
  PMX_CHECK(pull_ctl(h));
  if (iters_done) *iters_done = h->h_ctl->it - it0;
  if (conv_A) *conv_A = h->h_ctl->conv[0];
  if (conv_S) *conv_S = h->h_ctl->conv[1];
  if (sub_A) *sub_A = h->h_ctl->sub_total[0];
  if (sub_S) *sub_S = h->h_ctl->sub_total[1];
  if (h->h_ctl->fault) {
    pmx_set_error("multi-GPU exchange timed out: a peer did not reach iteration %d", h->h_ctl->it);
    return PMX_ERR_NCCL;
  }
  return PMX_OK;
}

// ------------------------------------------------------------------ bsdmm (algorithms.py:653-850 via nmf.py:178-203)
int pmx_nmf_bsdmm_begin(pmx_nmf* h, const pmx_bsdmm_opts* opts) {
  PMX_REQUIRE(h && opts, "NULL argument");
  PMX_REQUIRE(opts->n_g_A >= 0 && opts->n_g_A <= 4 && opts->n_g_S >= 0 && opts->n_g_S <= 4, "0..4 constraints per block");
  if (h->ctx->world > 1)
    for (int i = 0; i < opts->n_g_S; ++i)
      for (int k = 0; k < opts->proxs_g_S[i].n_ops; ++k)
        if (opts->proxs_g_S[i].ops[k].op == PMX_OP_UNITY && opts->proxs_g_S[i].ops[k].axis == 1) {
          pmx_set_error("prox_unity(axis=1) on the column-sharded S block needs a cross-rank row sum (not implemented)");
          return PMX_ERR_UNSUPPORTED;
        }
  PMX_CHECK(reject_sharded_row_unity(h, make_chain(&opts->prox_S), "bsdmm (direct prox_S)"));
  h->bs = *opts;
  h->tail_mode = false;
  if (h->plan) PMX_CHECK(umma_plan_use_A(h->plan, nullptr, nullptr));
  h->split_valid = false;
  h->gramS_valid = false;
  h->gram_pending = false;
  const size_t mk = (size_t)h->M * h->K, kn = (size_t)h->K * h->N;
  const size_t big = mk > kn ? mk : kn;
  for (int j = 0; j < 2; ++j) {
    const int ng = j == 0 ? opts->n_g_A : opts->n_g_S;
    const size_t n = j == 0 ? mk : kn;
    const float* X = j == 0 ? h->A : h->S;
    for (int i = 0; i < ng; ++i) {   // Z_ji = X_j.copy(), U_ji = 0   (algorithms.py:787-790, utils.py:244-254)
      if (!h->Zg[j][i]) PMX_CHECK(alloc_f(h->ctx, &h->Zg[j][i], n));
      if (!h->Ug[j][i]) PMX_CHECK(alloc_f(h->ctx, &h->Ug[j][i], n));
      PMX_CUDA(cudaMemcpyAsync(h->Zg[j][i], X, sizeof(float) * n, cudaMemcpyDeviceToDevice, h->ctx->stream));
      PMX_CUDA(cudaMemsetAsync(h->Ug[j][i], 0, sizeof(float) * n, h->ctx->stream));
    }
  }
  if (!h->Z0) PMX_CHECK(alloc_f(h->ctx, &h->Z0, big));
  if (!h->bs_norms) PMX_CUDA(cudaMalloc((void**)&h->bs_norms, sizeof(double) * 256));
  PMX_CUDA(cudaMemsetAsync(h->bs_norms, 0, sizeof(double) * 256, h->ctx->stream));
  PMX_CUDA(cudaMemsetAsync(h->ctl, 0, sizeof(pmx_ctl), h->ctx->stream));
  h->bs_it = 0;
  PMX_CHECK(nmf_global_cols(h));
  PMX_CHECK(xchg_setup(h, opts->kernel));
  return PMX_OK;
}

int pmx_nmf_bsdmm_run(pmx_nmf* h, int n_iter, int* iters_done, int* conv_A, int* conv_S) {
  PMX_REQUIRE(h != nullptr, "NULL handle");
  pmx_ctx* ctx = h->ctx;
  const pmx_bsdmm_opts& o = h->bs;
  PMX_CHECK(pull_ctl(h));
  const int it0 = h->h_ctl->it;
  ProxChain dA = make_chain(&o.prox_A), dS = make_chain(&o.prox_S);
  ProxChain gA[4], gS[4];
  for (int i = 0; i < 4; ++i) {
    gA[i] = make_chain(i < o.n_g_A ? &o.proxs_g_A[i] : nullptr);
    gS[i] = make_chain(i < o.n_g_S ? &o.proxs_g_S[i] : nullptr);
  }
  double* sums = h->bs_norms + 64;   // scratch for UNITY sums (up to 128 doubles)
  for (int i = 0; i < n_iter; ++i) {
    if (h->h_ctl->done) break;
    // block A (Gauss-Seidel: uses the current S), then block S with the updated A  (algorithms.py:805-839)
    // (the reference's closure evaluates BOTH gradients per block and keeps one, nmf.py:181-185: 12 MNK flop per outer
    // iteration; the tcgen05 kernel skips the unused GEMM and its flush: 8 MNK)
    PMX_CHECK(nmf_steps(h, h->A, h->S, true, false));
    PMX_CHECK(nmf_gradient(h, h->A, h->S, h->GA, h->GS, nullptr, o.kernel, &h->ctl->done, false, nullptr, 0, 1));
    PMX_CHECK(nmf_steps_join(h));
    PMX_CHECK(launch_bsdmm_block(ctx, h->ctl, 0, h->A, h->GA, h->Zg[0], h->Ug[0], h->Z0, sums, h->M, h->K, o.n_g_A, dA,
                                 gA, &h->ctl->step[0], h->bs_norms, o.e_rel_A, o.e_abs_A, false,
                                 (double)h->M * h->K));
    PMX_CHECK(nmf_steps(h, h->A, h->S, false, true));
    PMX_CHECK(nmf_gradient(h, h->A, h->S, h->GA, h->GS, nullptr, o.kernel, &h->ctl->done, false, nullptr, 0, 2));
    PMX_CHECK(nmf_steps_join(h));
    PMX_CHECK(launch_bsdmm_block(ctx, h->ctl, 1, h->S, h->GS, h->Zg[1], h->Ug[1], h->Z0, sums, h->K, h->N, o.n_g_S, dS,
                                 gS, &h->ctl->step[1], h->bs_norms + 32, o.e_rel_S, o.e_abs_S, true,
                                 (double)h->K * h->N_global,
                                 h->xchg_peer ? (long long)(h->xchg_off_small + 4 * pmx_peer_small_bytes()) : -1));
    PMX_CHECK(launch_bsdmm_iter_finalize(ctx, h->ctl));
    if ((i + 1) % 8 == 0) PMX_CHECK(pull_ctl(h));
  }
  PMX_CHECK(pull_ctl(h));
  if (iters_done) *iters_done = h->h_ctl->it - it0;
  if (conv_A) *conv_A = h->h_ctl->conv[0];
  if (conv_S) *conv_S = h->h_ctl->conv[1];
  if (h->h_ctl->fault) {
    pmx_set_error("multi-GPU exchange timed out: a peer did not reach iteration %d", h->h_ctl->it);
    return PMX_ERR_NCCL;
  }
  if (h->h_ctl->nonfinite) {
    pmx_set_error("Gram matrix contains infs or NaNs (iteration %d)", h->h_ctl->it);
    return PMX_ERR_NONFINITE;
  }
  return PMX_OK;
}

int pmx_nmf_grad(pmx_ctx* ctx, const float* Y, const float* A, const float* S, int M, int N, int K, float* G_A,
                 float* G_S, double* loss_or_null, int kernel) {
  PMX_REQUIRE(ctx && Y && A && S && G_A && G_S, "NULL argument");
  PMX_REQUIRE(M > 0 && N > 0 && K > 0, "shape must be positive");
  bool use_umma = (kernel == 2) || (kernel == 0 && umma_supported(M, N, K) && (long long)M * N >= 128LL * 128);
  if (use_umma) {
    if (!umma_supported(M, N, K)) {
      pmx_set_error("tcgen05 gradient kernel needs K <= 128 (M=%d N=%d K=%d)", M, N, K);
      return PMX_ERR_UNSUPPORTED;
    }
    UmmaPlan* plan = nullptr;
    PMX_CHECK(umma_plan_create(ctx, Y, N, M, N, K, &plan));
    int st = launch_grad_umma(ctx, plan, A, S, G_A, G_S, loss_or_null, nullptr);
    cudaStreamSynchronize(ctx->stream);
    umma_plan_destroy(ctx, plan);
    return st;
  }
  return launch_grad_simt(ctx, Y, nullptr, N, 0, A, S, M, N, K, G_A, G_S, loss_or_null, nullptr);
}

int pmx_nmf_lipschitz(pmx_ctx* ctx, const float* A, const float* S, int M, int N, int K, float* lip_A_host,
                      float* lip_S_host) {
  PMX_REQUIRE(ctx && A && S, "NULL argument");
  double* gram;
  pmx_ctl* ctl;
  const size_t kk2 = ((size_t)K * K + 1) & ~(size_t)1;   // the second Gram starts 16-byte aligned (vector zero-fill)
  PMX_CUDA(cudaMalloc((void**)&gram, sizeof(double) * kk2 * 2));
  PMX_CUDA(cudaMalloc((void**)&ctl, sizeof(pmx_ctl)));
  PMX_CUDA(cudaMemsetAsync(ctl, 0, sizeof(pmx_ctl), ctx->stream));
  int st = launch_gram(ctx, ctx->stream, S, K, N, false, gram, nullptr);
  if (st == PMX_OK) st = launch_lambda_max(ctx, ctx->stream, gram, K, ctl, 0);
  if (st == PMX_OK) st = launch_gram(ctx, ctx->stream, A, M, K, true, gram + kk2, nullptr);
  if (st == PMX_OK) st = launch_lambda_max(ctx, ctx->stream, gram + kk2, K, ctl, 1);
  pmx_ctl hc;
  memset(&hc, 0, sizeof(hc));
  if (st == PMX_OK) {
    cudaError_t e = cudaMemcpyAsync(&hc, ctl, sizeof(hc), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      pmx_set_error("pmx_nmf_lipschitz: %s", cudaGetErrorString(e));
      st = PMX_ERR_CUDA;
    }
  }
  cudaFree(gram);
  cudaFree(ctl);
  if (st != PMX_OK) return st;
  if (lip_A_host) *lip_A_host = hc.lip[0];  // Lipschitz constant of grad_A = lambda_max(S S^T)
  if (lip_S_host) *lip_S_host = hc.lip[1];
  if (hc.nonfinite) {
    pmx_set_error("Array must not contain infs or NaNs");
    return PMX_ERR_NONFINITE;
  }
  return PMX_OK;
}

}  // extern "C"
