// Linearised ADMM / SDMM on one block with L = identity, fully fused:
//   utils.py:307-346  update_variables (X step through prox_f, then do_the_mm per constraint)
//   utils.py:295-304  do_the_mm        (Z' = prox_g(X + U), R = X - Z', S = -(Z' - Z)/step_g, U += R)
//   utils.py:349-391  get_variable_errors / check_constraint_convergence (5 norms per constraint)
//   algorithms.py:478-514, 603-644  iteration counter, stall detection and the halved-slack restart
// One elementwise kernel per iteration reads X, b, Z_i, U_i and writes X, Z_i, U_i (7 fp32 streams for
// one constraint = the 28 bytes/element of SURVEY 8-d); all norms and the "nothing changed" test are
// fused into it, and a one-thread kernel keeps the iteration / restart state machine on the device.
// The arithmetic uses explicit round-to-nearest intrinsics in the reference's operation order (no FMA
// contraction), so X, Z, U are bit-identical to NumPy's fp32 results.
#include <math.h>
#include <stdlib.h>

#include "kernels.h"

namespace {

constexpr int kT = 256;
constexpr int MAXG = 4;

__device__ __forceinline__ uint64_t l2_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_evict_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld4_hint(const float4* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st4_hint(float4* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
               "l"(pol)
               : "memory");
}

struct admm_ctl {
  int done, it, converged, reinit, restarts, max_iter;
  int changed_x, changed_r;
  float slack;
  float pad;
  long long passes;        // passes executed since pmx_admm_run started (restarts reset `it`, not this)
  double norms[MAXG][5];   // |LX|^2, |Z|^2, |U(/step_g)|^2, |R|^2, |S|^2   (after the pass)
  double errors[MAXG][4];  // e_pri, e_dual, |R|, |S|
};

struct PassArgs {
  float* X;
  const float* b;
  float* Z[MAXG];
  float* U[MAXG];
  size_t n;
  int n_g;
  ProxChain chain[MAXG];
  double step_base;   // step_f(X, it) of the caller; multiplied by ctl->slack when use_slack
  int use_slack;
  int dual_uses_step_g;
  admm_ctl* ctl;
  int l2_mode;        // L2 residency policy of the streams (see k_admm_pass)
  float* part;        // [gridDim.x][MAXG * 5] per-block partial norms (summed by k_admm_finalize: no same-address atomics)
};

// NG = number of constraints (compile time: the accumulators and the Z/U registers of unused slots would cost
// half of the occupancy)
template <int NG>
__global__ void __launch_bounds__(kT) k_admm_pass(PassArgs a) {
  admm_ctl* ctl = a.ctl;
  if (ctl->done) return;
  const bool reinit = ctl->reinit != 0;   // utils.py:244-254 folded into the pass: Z = X, U = 0
  // scalar recipe in double, then one rounding to fp32 (NumPy weak-scalar promotion of Python floats)
  const double sf_d = a.use_slack ? (double)ctl->slack * a.step_base : a.step_base;   // algorithms.py:482
  const double sg_d = sf_d * 1 * 1 * (NG > 1 ? NG : 1);                         // utils.py:279
  const float sf = (float)sf_d;
  const float sg = (float)sg_d;
  const float ratio = (float)(sf_d / sg_d);     // step_f / step_g   (utils.py:316,333)
  const float cS = (float)(-1.0 / sg_d);        // -1 / step_g       (utils.py:300)
  float acc[NG][5];
#pragma unroll
  for (int i = 0; i < NG; ++i)
#pragma unroll
    for (int q = 0; q < 5; ++q) acc[i][q] = 0.f;
  int ch_x = 0, ch_r = 0;
  // one element of the pass: (x, b, z_i, u_i) -> (x', z_i', u_i'), norms and change flags accumulated
  auto element = [&](float x, float bv, float (&z)[NG], float (&u)[NG], float& xn_out) {
    // dX = sum_i step_f/step_g_i (X - Z_i + U_i)    (left-to-right like np.sum over the list, utils.py:331-337)
    float dX = __fmul_rn(ratio, __fadd_rn(__fsub_rn(x, z[0]), u[0]));
#pragma unroll
    for (int i = 1; i < NG; ++i)
      dX = __fadd_rn(dX, __fmul_rn(ratio, __fadd_rn(__fsub_rn(x, z[i]), u[i])));
    const float xa = __fsub_rn(x, dX);
    // prox_f(Xa, step_f) = Xa - step_f (Xa - b)       (utils.LeastSquaresProx, README.md:82-84)
    const float xn = __fsub_rn(xa, __fmul_rn(sf, __fsub_rn(xa, bv)));
    xn_out = xn;
    ch_x |= (xn != x) && !(xn != xn && x != x);
#pragma unroll
    for (int i = 0; i < NG; ++i)
      {
        const float zn = chain_segment(a.chain[i], 0, a.chain[i].n, __fadd_rn(xn, u[i]), sg);  // utils.py:297
        const float r = __fsub_rn(xn, zn);                                                      // utils.py:299
        const float sv = __fmul_rn(cS, __fsub_rn(zn, z[i]));                                    // utils.py:300
        const float un = __fadd_rn(u[i], r);                                                    // utils.py:303
        const float r_prev = __fsub_rn(x, z[i]);  // the R of the previous pass (same fp32 subtraction)
        ch_r |= (r != r_prev);
        const float uq = a.dual_uses_step_g ? __fdiv_rn(un, sg) : un;                           // utils.py:359-362
        acc[i][0] = fmaf(xn, xn, acc[i][0]);
        acc[i][1] = fmaf(zn, zn, acc[i][1]);
        acc[i][2] = fmaf(uq, uq, acc[i][2]);
        acc[i][3] = fmaf(r, r, acc[i][3]);
        acc[i][4] = fmaf(sv, sv, acc[i][4]);
        z[i] = zn;
        u[i] = un;
      }
  };
  // main part: 16-byte accesses, the loads of TWO groups per thread issued before the arithmetic (the pass is a pure
  // memory stream: 7 arrays for one constraint); the buffers come from cudaMalloc, i.e. are 16-byte aligned.
  // L2 residency: X and U_0 are read AND rewritten by every pass -- they carry an evict_last policy on both sides, b
  // and Z an evict_first one.  At n = 1e7 (40 MB per array, 126 MB of L2) the two kept arrays are served from L2 by
  // the next pass: DRAM sees b + Z (12 of the 28 algorithmic bytes per element) instead of all seven streams.
  // (a.l2_mode, env PMX_ADMM_L2 for experiments: 0 = no distinction (evict_normal everywhere), 1 = X and U_0 kept
  //  [default], 2 = only X kept, 3 = everything evict_first)
  const uint64_t pol_n = l2_evict_normal();
  const uint64_t pol_stream = a.l2_mode == 0 ? pol_n : l2_evict_first();
  const uint64_t pol_keep = a.l2_mode == 0 ? pol_n : (a.l2_mode == 3 ? pol_stream : l2_evict_last());
  const uint64_t pol_keep_u = a.l2_mode == 2 ? pol_stream : pol_keep;
  const size_t n4 = a.n >> 2;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t g0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g0 < n4; g0 += 2 * stride) {
    float4 x4[2], b4[2], z4[2][NG], u4[2][NG];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const size_t g = g0 + h * stride;
      if (g < n4) {
        x4[h] = ld4_hint(reinterpret_cast<const float4*>(a.X) + g, pol_keep);
        b4[h] = ld4_hint(reinterpret_cast<const float4*>(a.b) + g, pol_stream);
#pragma unroll
        for (int i = 0; i < NG; ++i) {
          z4[h][i] = reinit ? x4[h] : ld4_hint(reinterpret_cast<const float4*>(a.Z[i]) + g, pol_stream);
          u4[h][i] = reinit ? make_float4(0.f, 0.f, 0.f, 0.f)
                            : ld4_hint(reinterpret_cast<const float4*>(a.U[i]) + g, i == 0 ? pol_keep_u : pol_stream);
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const size_t g = g0 + h * stride;
      if (g >= n4) break;
      float xo[4];
      const float xi[4] = {x4[h].x, x4[h].y, x4[h].z, x4[h].w}, bi[4] = {b4[h].x, b4[h].y, b4[h].z, b4[h].w};
      float zo[NG][4], uo[NG][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float z[NG], u[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) {
          z[i] = reinterpret_cast<const float*>(&z4[h][i])[c];
          u[i] = reinterpret_cast<const float*>(&u4[h][i])[c];
        }
        element(xi[c], bi[c], z, u, xo[c]);
#pragma unroll
        for (int i = 0; i < NG; ++i) {
          zo[i][c] = z[i];
          uo[i][c] = u[i];
        }
      }
      st4_hint(reinterpret_cast<float4*>(a.X) + g, make_float4(xo[0], xo[1], xo[2], xo[3]), pol_keep);
#pragma unroll
      for (int i = 0; i < NG; ++i) {
        st4_hint(reinterpret_cast<float4*>(a.Z[i]) + g, make_float4(zo[i][0], zo[i][1], zo[i][2], zo[i][3]), pol_stream);
        st4_hint(reinterpret_cast<float4*>(a.U[i]) + g, make_float4(uo[i][0], uo[i][1], uo[i][2], uo[i][3]),
                 i == 0 ? pol_keep_u : pol_stream);
      }
    }
  }
  // tail (n not a multiple of 4): scalar, the first threads of block 0
  if (blockIdx.x == 0) {
    for (size_t idx = (n4 << 2) + threadIdx.x; idx < a.n; idx += blockDim.x) {
      const float x = a.X[idx];
      float z[NG], u[NG];
#pragma unroll
      for (int i = 0; i < NG; ++i)
        {
          z[i] = reinit ? x : a.Z[i][idx];
          u[i] = reinit ? 0.f : a.U[i][idx];
        }
      float xn;
      element(x, a.b[idx], z, u, xn);
      a.X[idx] = xn;
#pragma unroll
      for (int i = 0; i < NG; ++i)
        {
          a.Z[i][idx] = z[i];
          a.U[i][idx] = u[i];
        }
    }
  }
  // block reduction: norms and the two change flags
  __shared__ float red[kT / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NG; ++i)
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float v = acc[i][q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      __syncthreads();
      if (lane == 0) red[w] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < kT / 32; ++k) t += red[k];
        a.part[(size_t)blockIdx.x * (MAXG * 5) + i * 5 + q] = t;
      }
    }
  const int any_x = __syncthreads_or(ch_x);
  const int any_r = __syncthreads_or(ch_r);
  if (threadIdx.x == 0) {
    if (any_x) ctl->changed_x = 1;
    if (any_r) ctl->changed_r = 1;
  }
}

// convergence test + iteration / restart state machine (algorithms.py:494-514)
__global__ void __launch_bounds__(256) k_admm_finalize(admm_ctl* ctl, int n_g, double n_elems, float e_rel, float e_abs,
                                                       int manage, const float* part, int nblocks) {
  if (ctl->done) return;
  // sum the per-block partial norms (fp64), 256 threads
  __shared__ double s_acc[MAXG * 5][8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int v = 0; v < n_g * 5; ++v) {
    double acc = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) acc += (double)part[(size_t)b * (MAXG * 5) + v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_acc[v][w] = acc;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int v = 0; v < n_g * 5; ++v) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += s_acc[v][k];
    ctl->norms[v / 5][v % 5] = t;
  }
  bool all = true;
  for (int i = 0; i < n_g; ++i) {
    const float lLX = sqrtf((float)ctl->norms[i][0]);
    const float lZ = sqrtf((float)ctl->norms[i][1]);
    const float lU = sqrtf((float)ctl->norms[i][2]);
    const float lR = sqrtf((float)ctl->norms[i][3]);
    const float lS = sqrtf((float)ctl->norms[i][4]);
    // np.sqrt(p) * e_abs is float64 in the reference, the e_rel product fp32 (utils.py:357-362)
    const double e_pri = sqrt(n_elems) * (double)e_abs + (double)(e_rel * fmaxf(lLX, lZ));
    const double e_dual = sqrt(n_elems) * (double)e_abs + (double)(e_rel * lU);
    ctl->errors[i][0] = e_pri;
    ctl->errors[i][1] = e_dual;
    ctl->errors[i][2] = lR;
    ctl->errors[i][3] = lS;
    all = all && ((double)lR <= e_pri) && ((double)lS <= e_dual);                                            // utils.py:390
    for (int q = 0; q < 5; ++q) ctl->norms[i][q] = 0.0;
  }
  ctl->passes += 1;
  const bool stalled = !ctl->changed_x && !ctl->changed_r;
  ctl->changed_x = 0;
  ctl->changed_r = 0;
  ctl->converged = all ? 1 : 0;
  ctl->reinit = 0;
  if (!manage) {
    ctl->restarts = stalled ? 1 : 0;   // step mode: report "nothing changed" to the host
    return;
  }
  if (all) {
    ctl->done = 1;
    return;
  }
  ctl->it += 1;
  if (ctl->it > 1 && stalled) {        // algorithms.py:503-512: halve the slack, restart
    ctl->slack *= 0.5f;
    ctl->it = 0;
    ctl->reinit = 1;
    ctl->restarts += 1;
    if (ctl->restarts > 64) ctl->done = 1;
  }
  if (ctl->it >= ctl->max_iter) ctl->done = 1;   // `while it < max_iter`
}

}  // namespace

struct pmx_admm {
  pmx_ctx* ctx;
  size_t n;
  float* part;
  int nblocks;
  pmx_admm_opts opts;
  float *X, *b;
  float* Z[MAXG];
  float* U[MAXG];
  admm_ctl* ctl;
  admm_ctl* h_ctl;
  cudaGraphExec_t graph;   // one batch of iterations (pass + finalize) captured for pmx_admm_run
  double graph_step;       // step_f the graph was captured with
  int graph_len;
};

static int admm_pull(pmx_admm* h) {
  PMX_CUDA(cudaMemcpyAsync(h->h_ctl, h->ctl, sizeof(admm_ctl), cudaMemcpyDeviceToHost, h->ctx->stream));
  PMX_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return PMX_OK;
}

static int admm_enqueue(pmx_admm* h, double step_base, int use_slack, int manage) {
  pmx_ctx* ctx = h->ctx;
  PassArgs a;
  memset(&a, 0, sizeof(a));
  a.X = h->X;
  a.b = h->b;
  a.n = h->n;
  a.n_g = h->opts.n_g;
  for (int i = 0; i < h->opts.n_g; ++i) {
    a.Z[i] = h->Z[i];
    a.U[i] = h->U[i];
    a.chain[i] = make_chain(&h->opts.proxs_g[i]);
  }
  a.step_base = step_base;
  a.use_slack = use_slack;
  a.dual_uses_step_g = h->opts.dual_uses_step_g;
  a.ctl = h->ctl;
  a.part = h->part;
  {
    static const char* m = getenv("PMX_ADMM_L2");
    a.l2_mode = m ? atoi(m) : 1;
  }
  const bool prof = ctx->profile && ctx->prof_n < PMX_PROF_MAX;   // per-launch events around the pass (bench.py roofline)
  if (prof) PMX_CUDA(cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n], ctx->stream));
  switch (h->opts.n_g) {
    case 1: k_admm_pass<1><<<h->nblocks, kT, 0, ctx->stream>>>(a); break;
    case 2: k_admm_pass<2><<<h->nblocks, kT, 0, ctx->stream>>>(a); break;
    case 3: k_admm_pass<3><<<h->nblocks, kT, 0, ctx->stream>>>(a); break;
    default: k_admm_pass<4><<<h->nblocks, kT, 0, ctx->stream>>>(a); break;
  }
  if (prof) {
    PMX_CUDA(cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n + 1], ctx->stream));
    ctx->prof_n++;
  }
  PMX_LAUNCHED(ctx);
  k_admm_finalize<<<1, 256, 0, ctx->stream>>>(h->ctl, h->opts.n_g, (double)h->n, h->opts.e_rel, h->opts.e_abs, manage,
                                              h->part, h->nblocks);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "admm pass");
}

extern "C" {

int pmx_admm_create(pmx_ctx* ctx, size_t n, const pmx_admm_opts* opts, pmx_admm** out) {
  PMX_REQUIRE(ctx && opts && out, "NULL argument");
  PMX_REQUIRE(opts->n_g >= 1 && opts->n_g <= MAXG, "1..4 constraints are supported by the fused ADMM loop");
  for (int i = 0; i < opts->n_g; ++i)
    for (int k = 0; k < opts->proxs_g[i].n_ops; ++k)
      if (opts->proxs_g[i].ops[k].op == PMX_OP_UNITY) {
        pmx_set_error("prox_unity inside a vector ADMM constraint needs a global sum; use the callback loop");
        return PMX_ERR_UNSUPPORTED;
      }
  pmx_admm* h = new pmx_admm();
  memset(h, 0, sizeof(*h));
  h->ctx = ctx;
  h->n = n;
  h->opts = *opts;
  PMX_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = sizeof(float) * (n ? n : 1);
  PMX_CUDA(cudaMalloc((void**)&h->X, bytes));
  PMX_CUDA(cudaMalloc((void**)&h->b, bytes));
  for (int i = 0; i < opts->n_g; ++i) {
    PMX_CUDA(cudaMalloc((void**)&h->Z[i], bytes));
    PMX_CUDA(cudaMalloc((void**)&h->U[i], bytes));
  }
  {
    long long blocks = (long long)((n / 4 + kT - 1) / kT);
    // one full wave of resident blocks (grid-stride loop): a partial second wave leaves SMs half empty for the
    // second half of an HBM-bound pass
    int occ = 0;
    cudaError_t oe = cudaSuccess;
    switch (opts->n_g) {
      case 1: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_admm_pass<1>, kT, 0); break;
      case 2: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_admm_pass<2>, kT, 0); break;
      case 3: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_admm_pass<3>, kT, 0); break;
      default: oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_admm_pass<4>, kT, 0); break;
    }
    if (oe != cudaSuccess || occ < 1) {
      cudaGetLastError();
      occ = 4;
    }
    const long long cap = (long long)ctx->sm_count * occ;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    h->nblocks = (int)blocks;
  }
  PMX_CUDA(cudaMalloc((void**)&h->part, sizeof(float) * (size_t)h->nblocks * MAXG * 5));
  PMX_CUDA(cudaMalloc((void**)&h->ctl, sizeof(admm_ctl)));
  PMX_CUDA(cudaMallocHost((void**)&h->h_ctl, sizeof(admm_ctl)));
  *out = h;
  return pmx_admm_init_zu(h);
}

int pmx_admm_destroy(pmx_admm* h) {
  if (!h) return PMX_OK;
  cudaStreamSynchronize(h->ctx->stream);
  cudaFree(h->X);
  cudaFree(h->b);
  for (int i = 0; i < MAXG; ++i) {
    if (h->Z[i]) cudaFree(h->Z[i]);
    if (h->U[i]) cudaFree(h->U[i]);
  }
  if (h->graph) cudaGraphExecDestroy(h->graph);
  cudaFree(h->part);
  cudaFree(h->ctl);
  cudaFreeHost(h->h_ctl);
  delete h;
  return PMX_OK;
}

int pmx_admm_set(pmx_admm* h, const float* host_X, const float* host_b) {
  PMX_REQUIRE(h != nullptr, "NULL handle");
  if (host_X) PMX_CHECK(pmx_h2d(h->ctx, h->X, host_X, sizeof(float) * h->n));
  if (host_b) PMX_CHECK(pmx_h2d(h->ctx, h->b, host_b, sizeof(float) * h->n));
  return PMX_OK;
}

int pmx_admm_get(pmx_admm* h, float* host_X) {
  PMX_REQUIRE(h && host_X, "NULL argument");
  return pmx_d2h(h->ctx, host_X, h->X, sizeof(float) * h->n);
}

int pmx_admm_init_zu(pmx_admm* h) {
  PMX_REQUIRE(h != nullptr, "NULL handle");
  admm_ctl c;
  memset(&c, 0, sizeof(c));
  c.slack = 1.0f;
  c.reinit = 1;  // the next pass treats Z = X, U = 0 (utils.py:244-254)
  c.max_iter = 1 << 30;
  memcpy(h->h_ctl, &c, sizeof(c));
  PMX_CUDA(cudaMemcpyAsync(h->ctl, h->h_ctl, sizeof(admm_ctl), cudaMemcpyHostToDevice, h->ctx->stream));
  PMX_CUDA(cudaStreamSynchronize(h->ctx->stream));
  return PMX_OK;
}

int pmx_admm_step(pmx_admm* h, double step_f, int* converged, int* stalled, double* errors) {
  PMX_REQUIRE(h != nullptr, "NULL handle");
  PMX_CHECK(admm_enqueue(h, step_f, 0, 0));
  PMX_CHECK(admm_pull(h));
  if (converged) *converged = h->h_ctl->converged;
  if (stalled) *stalled = h->h_ctl->restarts;
  if (errors) memcpy(errors, h->h_ctl->errors, sizeof(double) * 4 * h->opts.n_g);
  return PMX_OK;
}

int pmx_admm_run(pmx_admm* h, double step_f, int max_iter, int* iters_logged, int* converged, double* errors) {
  PMX_REQUIRE(h != nullptr, "NULL handle");
  PMX_REQUIRE(max_iter >= 0, "max_iter must be >= 0");
  PMX_CHECK(pmx_admm_init_zu(h));
  h->h_ctl->max_iter = max_iter;
  if (max_iter == 0) h->h_ctl->done = 1;
  PMX_CUDA(cudaMemcpyAsync(h->ctl, h->h_ctl, sizeof(admm_ctl), cudaMemcpyHostToDevice, h->ctx->stream));
  PMX_CUDA(cudaStreamSynchronize(h->ctx->stream));
  // A batch of iterations is replayed as one CUDA graph (every kernel honours ctl->done, so overshooting the stopping
  // iteration is harmless): the host then launches once per batch instead of twice per iteration.
  const int batch = 16;
  const bool no_graph = getenv("PMX_NO_GRAPH") != nullptr || h->ctx->profile;
  if (!no_graph && (!h->graph || h->graph_step != step_f || h->graph_len != batch)) {
    if (h->graph) {
      cudaGraphExecDestroy(h->graph);
      h->graph = nullptr;
    }
    cudaGraph_t g = nullptr;
    const long long l0 = h->ctx->launches;
    PMX_CUDA(cudaStreamBeginCapture(h->ctx->stream, cudaStreamCaptureModeThreadLocal));
    int st = PMX_OK;
    for (int i = 0; i < batch && st == PMX_OK; ++i) st = admm_enqueue(h, step_f, 1, 1);
    cudaError_t ce = cudaStreamEndCapture(h->ctx->stream, &g);
    h->ctx->launches = l0;   // the capture did not execute anything
    if (st != PMX_OK) return st;
    if (ce != cudaSuccess || !g) {
      pmx_set_error("CUDA graph capture of the ADMM batch failed: %s", cudaGetErrorString(ce));
      return PMX_ERR_CUDA;
    }
    PMX_CUDA(cudaGraphInstantiate(&h->graph, g, nullptr, nullptr, 0));
    cudaGraphDestroy(g);
    h->graph_step = step_f;
    h->graph_len = batch;
  }
  long long guard = 0;
  while (!h->h_ctl->done) {
    if (h->graph && !no_graph) {
      PMX_CUDA(cudaGraphLaunch(h->graph, h->ctx->stream));
      h->ctx->launches += 2LL * batch;
    } else {
      for (int i = 0; i < batch; ++i) PMX_CHECK(admm_enqueue(h, step_f, 1, 1));
    }
    PMX_CHECK(admm_pull(h));
    if (++guard > (1LL << 26)) break;
  }
  if (iters_logged) *iters_logged = h->h_ctl->it + 1;   // "Completed it + 1 iterations" (algorithms.py:516)
  if (converged) *converged = h->h_ctl->converged;
  if (errors) memcpy(errors, h->h_ctl->errors, sizeof(double) * 4 * h->opts.n_g);
  if (h->h_ctl->restarts > 64) {
    pmx_set_error("ADMM restarted more than 64 times without progress");
    return PMX_ERR_UNSUPPORTED;
  }
  return PMX_OK;
}

int pmx_admm_stats(pmx_admm* h, long long* passes, int* restarts) {
  PMX_REQUIRE(h != nullptr, "NULL handle");
  if (passes) *passes = h->h_ctl->passes;
  if (restarts) *restarts = h->h_ctl->restarts;
  return PMX_OK;
}

}  // extern "C"
