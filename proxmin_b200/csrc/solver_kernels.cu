// Small elementwise / reduction kernels used by the adaprox and bsdmm device loops
// (algorithms.py:365-410, :800-844; utils.py:295-391).  All are single coalesced passes.
#include <cuda_bf16.h>

#include "kernels.h"

int pmx_comm_allreduce_internal(pmx_ctx* ctx, void* buf, size_t count, int kind, cudaStream_t st);
int pmx_peer_small_allreduce(pmx_ctx* ctx, int set, size_t off_bytes, void* buf, int n, int kind, cudaStream_t st,
                             const int* done, int* fault);

namespace {

constexpr int kT = 256;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[k] (double, pre-zeroed) += partial sums of `vals[k]` over the block (k < nvals <= 8)
template <int NV>
__device__ __forceinline__ void block_add(const float (&vals)[NV], double* out) {
  __shared__ float red[NV][kT / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const float v = wsum(vals[k]);
    if (lane == 0) red[k][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float t = 0.f;
    for (int i = 0; i < kT / 32; ++i) t += red[threadIdx.x][i];
    atomicAdd(out + threadIdx.x, (double)t);
  }
}

inline int grid_for(pmx_ctx* ctx, size_t n) {
  long long blocks = (long long)((n + kT - 1) / kT);
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

// ---- |X - Xold|^2, |X|^2, |Xold|^2 --------------------------------------------------------
__global__ void __launch_bounds__(kT) k_diff_norms(const float* __restrict__ X, const float* __restrict__ Xold,
                                                   size_t n, double* norms, const int* done) {
  if (done && *done) return;
  float v[3] = {0.f, 0.f, 0.f};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float x = X[i], o = Xold[i], d = x - o;
    v[0] = fmaf(d, d, v[0]);
    v[1] = fmaf(x, x, v[1]);
    v[2] = fmaf(o, o, v[2]);
  }
  block_add<3>(v, norms);
}

// ---- sums along an axis (double accumulators, pre-zeroed) ---------------------------------
// axis 0: out[c] = sum_r X[r, c];  thread = (row lane, column), coalesced along c
__global__ void __launch_bounds__(kT) k_colsum(const float* __restrict__ X, int rows, int cols, double* out,
                                               const int* done) {
  if (done && *done) return;
  const int cpb = cols < kT ? cols : kT;         // columns handled per block pass
  const int rpb = kT / cpb;                      // row lanes per block
  const int c_in = threadIdx.x % cpb, r_in = threadIdx.x / cpb;
  if (r_in >= rpb) return;
  for (int c0 = 0; c0 < cols; c0 += cpb) {
    const int c = c0 + c_in;
    if (c >= cols) continue;
    float acc = 0.f;
    for (int r = blockIdx.x * rpb + r_in; r < rows; r += gridDim.x * rpb) acc += X[(size_t)r * cols + c];
    atomicAdd(out + c, (double)acc);
  }
}
// axis 1: out[r] = sum_c X[r, c];  grid (chunks, rows)
__global__ void __launch_bounds__(kT) k_rowsum(const float* __restrict__ X, int rows, int cols, double* out,
                                               const int* done) {
  if (done && *done) return;
  const int r = blockIdx.y;
  float v[1] = {0.f};
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x) v[0] += X[(size_t)r * cols + c];
  block_add<1>(v, out + r);
}
// X[r, c] /= sums[c] (axis 0) or sums[r] (axis 1)            operators.py:44
__global__ void __launch_bounds__(kT) k_div_axis(float* __restrict__ X, int rows, int cols, const double* sums, int axis,
                                                 const int* done) {
  if (done && *done) return;
  const size_t n = (size_t)rows * cols;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (size_t)r * cols);
    X[i] = X[i] / (float)sums[axis == 0 ? c : r];
  }
}
// alpha[k] = mean / 10                                         nmf.py:91-93
__global__ void k_alpha_from_sums(const double* sums, int n, double count, float* alpha, const int* done) {
  if (done && *done) return;
  for (int k = threadIdx.x; k < n; k += blockDim.x) alpha[k] = ((float)(sums[k] / count)) / 10.0f;
}

// ---- adaprox sub-iteration control (algorithms.py:386-400) ---------------------------------
__global__ void k_sub_begin(pmx_ctl* ctl, int block) {
  if (ctl->done) return;
  ctl->sub_done = 0;
  ctl->sub_tau = 0;
  ctl->sub_parity = 0;
  ctl->psi_max[block] = 0.f;
  for (int i = 8; i < 11; ++i) ctl->norms[i] = 0.0;
}
__global__ void k_sub_finalize(pmx_ctl* ctl, float e2, int max_tau) {
  if (ctl->done || ctl->sub_done) return;
  ctl->sub_tau += 1;
  ctl->sub_parity ^= 1;
  // l2sq(z_ - z) <= e_rel**2 * l2sq(z)   with the OLD z on the right (algorithms.py:389)
  const bool conv = (float)ctl->norms[8] <= e2 * (float)ctl->norms[10];
  for (int i = 8; i < 11; ++i) ctl->norms[i] = 0.0;
  if (conv || ctl->sub_tau >= max_tau) ctl->sub_done = 1;
}
// X <- z (the buffer selected by the parity), Sub_iter[j] += tau   (algorithms.py:398-400)
// hi / lo (optional): bf16 split of the committed block, element (r, c) at [r * ld_split + c] -- the operands of the
// next gradient kernel (saves the two separate split passes of an adaprox iteration)
__global__ void __launch_bounds__(kT) k_sub_commit(float* __restrict__ X, const float* __restrict__ Z0,
                                                   const float* __restrict__ Z1, size_t n, pmx_ctl* ctl, int block,
                                                   unsigned short* __restrict__ hi, unsigned short* __restrict__ lo,
                                                   int cols, int ld_split) {
  if (ctl->done) return;
  if (!ctl->sub_done) {
    // the sub-iterations enqueued speculatively (no host round trip) did not reach the stopping rule: freeze the solve
    // (done = 2: every following kernel returns) until the host has finished this block the slow way
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      ctl->paused_block = block + 1;
      __threadfence();
      ctl->done = 2;
    }
    return;
  }
  const float* src = ctl->sub_parity ? Z1 : Z0;
  if (hi && n < 0xffffffffull) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < (unsigned)n; i += gridDim.x * blockDim.x) {
      const float v = src[i];
      X[i] = v;
      const unsigned r = i / (unsigned)cols, c = i - r * (unsigned)cols;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      hi[(size_t)r * ld_split + c] = __bfloat16_as_ushort(h);
      lo[(size_t)r * ld_split + c] = __bfloat16_as_ushort(l);
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) X[i] = src[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->sub_total[block] += ctl->sub_tau;
    ctl->sub_last[block] = ctl->sub_tau;
  }
}
__global__ void k_clear_pause(pmx_ctl* ctl) {
  if (ctl->done == 2) {
    ctl->done = 0;
    ctl->paused_block = 0;
  }
}
// sub-iteration with ping-pong buffers needs the buffer roles resolved on the device
__global__ void k_adaprox_finalize(pmx_ctl* ctl, float e2A, float e2S, int check) {
  if (ctl->done) return;
  if (check) {
    const bool cA = (float)ctl->norms[0] <= e2A * (float)ctl->norms[1];
    const bool cS = (float)ctl->norms[3] <= e2S * (float)ctl->norms[4];
    ctl->conv[0] = cA;
    ctl->conv[1] = cS;
    if (cA && cS) ctl->done = 1;
  }
  ctl->it += 1;
  for (int i = 0; i < 6; ++i) ctl->norms[i] = 0.0;
}

// ---- bsdmm (utils.py:307-346 with prox_f = prox_j(X - step*grad_j), nmf.py:181-185) -------
struct BsArgs {
  float* X;
  const float* G;
  float* Z[4];
  float* U[4];
  float* T;            // scratch (X + U_i) for the unfused path
  size_t n;
  int rows, cols;
  int n_g;
  int N_blocks;        // number of variable blocks (2)
  ProxChain direct;    // prox_j of nmf.py:185 (elementwise here)
  ProxChain g[4];
  const float* step_f; // device scalar (1 / Lipschitz)
  double* norms;       // [n_g][5]
  float e_dummy;
  pmx_ctl* ctl;
};

// X <- prox_j((X - dX) - step_f * G_j),  dX = sum_i step_f/step_g_i (X - Z_i + U_i)     utils.py:331-338
__global__ void __launch_bounds__(kT) k_bsdmm_x(BsArgs a) {
  if (a.ctl->done) return;
  const float sf = a.step_f[0];
  const float sg = __fmul_rn(__fmul_rn(__fmul_rn(sf, 1.0f), (float)a.N_blocks), (float)(a.n_g > 0 ? a.n_g : 1));  // utils.py:279
  const float ratio = __fdiv_rn(sf, sg);
  float v[3] = {0.f, 0.f, 0.f};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x * blockDim.x) {
    const float x = a.X[i];
    float xa = x;
    if (a.n_g > 0) {
      float dX = __fmul_rn(ratio, __fadd_rn(__fsub_rn(x, a.Z[0][i]), a.U[0][i]));
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (k < a.n_g) dX = __fadd_rn(dX, __fmul_rn(ratio, __fadd_rn(__fsub_rn(x, a.Z[k][i]), a.U[k][i])));
      xa = __fsub_rn(x, dX);
    }
    float xn = __fsub_rn(xa, __fmul_rn(sf, a.G[i]));                    // nmf.py:185
    xn = chain_segment(a.direct, 0, a.direct.n, xn, sf);
    a.X[i] = xn;
    const float d = xn - x;
    v[0] = fmaf(d, d, v[0]);   // only used when the block has no constraint (utils.py:319-327: S = X_new - X_old)
    v[1] = fmaf(xn, xn, v[1]);
  }
  if (a.n_g == 0) block_add<3>(v, a.norms);
}

// fused do_the_mm for elementwise prox_g (utils.py:295-304) + the five norms of utils.py:349-391
__global__ void __launch_bounds__(kT) k_bsdmm_zu(BsArgs a, int i_g, int fused) {
  if (a.ctl->done) return;
  const float sf = a.step_f[0];
  const float sg = __fmul_rn(__fmul_rn(__fmul_rn(sf, 1.0f), (float)a.N_blocks), (float)a.n_g);
  const float cS = __fdiv_rn(-1.0f, sg);
  float* Z = a.Z[i_g];
  float* U = a.U[i_g];
  float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x * blockDim.x) {
    const float x = a.X[i], u = U[i], z = Z[i];
    float zn;
    if (fused) zn = chain_segment(a.g[i_g], 0, a.g[i_g].n, __fadd_rn(x, u), sg);
    else zn = a.T[i];                                   // prox_g already applied to T = X + U by the general path
    const float r = __fsub_rn(x, zn);
    const float s = __fmul_rn(cS, __fsub_rn(zn, z));
    const float un = __fadd_rn(u, r);
    Z[i] = zn;
    U[i] = un;
    const float uq = __fdiv_rn(un, sg);
    v[0] = fmaf(x, x, v[0]);
    v[1] = fmaf(zn, zn, v[1]);
    v[2] = fmaf(uq, uq, v[2]);
    v[3] = fmaf(r, r, v[3]);
    v[4] = fmaf(s, s, v[4]);
  }
  block_add<5>(v, a.norms + 5 * i_g);
}

__global__ void __launch_bounds__(kT) k_add(const float* __restrict__ X, const float* __restrict__ U, float* __restrict__ T,
                                            size_t n, const int* done) {
  if (done && *done) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    T[i] = __fadd_rn(X[i], U[i]);
}

// per-block convergence of bsdmm (utils.py:373-391); norms: [n_g][5] doubles, cleared afterwards
__global__ void k_bsdmm_block_finalize(pmx_ctl* ctl, double* norms, int n_g, double n_elems, float e_rel, float e_abs,
                                       int block) {
  if (ctl->done) return;
  bool all = true;
  if (n_g == 0) {
    // no constraint: R = 0, S = X_new - X_old, e_dual = sqrt(n) e_abs + e_rel * |U| with U = 0
    const float lS = sqrtf((float)norms[0]);
    const float lX = sqrtf((float)norms[1]);
    const double e_pri = sqrt(n_elems) * (double)e_abs + (double)(e_rel * lX);
    const double e_dual = sqrt(n_elems) * (double)e_abs + 0.0;
    all = (0.0 <= e_pri) && ((double)lS <= e_dual);
    for (int q = 0; q < 5; ++q) norms[q] = 0.0;
  }
  for (int i = 0; i < n_g; ++i) {
    double* nr = norms + 5 * i;
    const float lLX = sqrtf((float)nr[0]), lZ = sqrtf((float)nr[1]), lU = sqrtf((float)nr[2]);
    const float lR = sqrtf((float)nr[3]), lS = sqrtf((float)nr[4]);
    const double e_pri = sqrt(n_elems) * (double)e_abs + (double)(e_rel * fmaxf(lLX, lZ));
    const double e_dual = sqrt(n_elems) * (double)e_abs + (double)(e_rel * lU);
    all = all && ((double)lR <= e_pri) && ((double)lS <= e_dual);
    for (int q = 0; q < 5; ++q) nr[q] = 0.0;
  }
  ctl->conv[block] = all ? 1 : 0;
}
__global__ void k_bsdmm_iter_finalize(pmx_ctl* ctl) {
  if (ctl->done) return;
  ctl->it += 1;
  if (ctl->conv[0] && ctl->conv[1]) ctl->done = 1;   // algorithms.py:841-844
}

}  // namespace

// ------------------------------------------------------------------------------------------
int launch_diff_norms(pmx_ctx* ctx, const float* X, const float* Xold, size_t n, double* norms, const int* done) {
  k_diff_norms<<<grid_for(ctx, n), kT, 0, ctx->stream>>>(X, Xold, n, norms, done);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_diff_norms");
}

int launch_axis_sum(pmx_ctx* ctx, const float* X, int rows, int cols, int axis, double* out, const int* done) {
  PMX_CHECK(launch_zero(ctx, ctx->stream, reinterpret_cast<float*>(out), 2 * (size_t)(axis == 0 ? cols : rows), done));
  if (axis == 0) {
    const int cpb = cols < kT ? cols : kT;
    const int rpb = kT / cpb;
    int blocks = pmx_div_up(rows, rpb * 8);
    if (blocks > ctx->sm_count * 4) blocks = ctx->sm_count * 4;
    if (blocks < 1) blocks = 1;
    k_colsum<<<blocks, kT, 0, ctx->stream>>>(X, rows, cols, out, done);
  } else {
    int bx = pmx_div_up(cols, kT * 8);
    if (bx > ctx->sm_count) bx = ctx->sm_count;
    if (bx < 1) bx = 1;
    k_rowsum<<<dim3(bx, rows), kT, 0, ctx->stream>>>(X, rows, cols, out, done);
  }
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "axis sum");
}

// General (multi-kernel) application of a prox chain in place: elementwise segments through the fused
// update kernel, every UNITY as "sum along axis" + "divide" passes.  Works for any chain and shape;
// the fused single-kernel paths are preferred whenever they apply.
int apply_chain_general(pmx_ctx* ctx, const ProxChain& ch, float* X, int rows, int cols, const StepSpec& step,
                        double* sums_scratch, const int* done) {
  int a = 0;
  while (a < ch.n || a == 0) {
    const int b = chain_next_unity(ch, a);
    if (b > a) {
      ProxChain seg;
      memset(&seg, 0, sizeof(seg));
      seg.n = b - a;
      for (int i = a; i < b; ++i) {
        seg.op[i - a] = ch.op[i];
        seg.rel[i - a] = ch.rel[i];
        seg.axis[i - a] = ch.axis[i];
        seg.thr[i - a] = ch.thr[i];
      }
      UpdIO io;
      memset(&io, 0, sizeof(io));
      io.Xin = X;
      io.Xout = X;
      io.rows = rows;
      io.cols = cols;
      io.step = step;
      io.done = done;
      PMX_CHECK(launch_update(ctx, IN_PLAIN, seg, io));
    }
    if (b >= ch.n) break;
    const int axis = ch.axis[b];
    PMX_CHECK(launch_axis_sum(ctx, X, rows, cols, axis, sums_scratch, done));
    k_div_axis<<<grid_for(ctx, (size_t)rows * cols), kT, 0, ctx->stream>>>(X, rows, cols, sums_scratch, axis, done);
    PMX_LAUNCHED(ctx);
    a = b + 1;
    if (a >= ch.n) break;
  }
  return pmx_check_launch(ctx, "apply_chain_general");
}

int launch_alpha_means(pmx_ctx* ctx, const float* X, int rows, int cols, int axis, double* sums, float* alpha,
                       const int* done) {
  PMX_CHECK(launch_axis_sum(ctx, X, rows, cols, axis, sums, done));
  const int n = axis == 0 ? cols : rows;
  const double count = axis == 0 ? rows : cols;
  k_alpha_from_sums<<<1, 128, 0, ctx->stream>>>(sums, n, count, alpha, done);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_alpha_from_sums");
}

int launch_alpha_from_sums(pmx_ctx* ctx, const double* sums, int n, double count, float* alpha, const int* done) {
  k_alpha_from_sums<<<1, 128, 0, ctx->stream>>>(sums, n, count, alpha, done);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_alpha_from_sums");
}

int launch_sub_begin(pmx_ctx* ctx, pmx_ctl* ctl, int block) {
  k_sub_begin<<<1, 1, 0, ctx->stream>>>(ctl, block);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_sub_begin");
}
int launch_sub_finalize(pmx_ctx* ctx, pmx_ctl* ctl, float e2, int max_tau) {
  k_sub_finalize<<<1, 1, 0, ctx->stream>>>(ctl, e2, max_tau);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_sub_finalize");
}
int launch_sub_commit(pmx_ctx* ctx, float* X, const float* Z0, const float* Z1, size_t n, pmx_ctl* ctl, int block,
                      unsigned short* hi, unsigned short* lo, int cols, int ld_split) {
  if (n >= 0xffffffffull) hi = lo = nullptr;
  k_sub_commit<<<grid_for(ctx, n), kT, 0, ctx->stream>>>(X, Z0, Z1, n, ctl, block, hi, lo, cols, ld_split);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_sub_commit");
}
int launch_clear_pause(pmx_ctx* ctx, pmx_ctl* ctl) {
  k_clear_pause<<<1, 1, 0, ctx->stream>>>(ctl);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_clear_pause");
}
int launch_adaprox_finalize(pmx_ctx* ctx, pmx_ctl* ctl, float e2A, float e2S, int check) {
  k_adaprox_finalize<<<1, 1, 0, ctx->stream>>>(ctl, e2A, e2S, check);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_adaprox_finalize");
}

// one block (j) of a bsdmm outer iteration; `g_unfused[i]` marks constraints whose chain contains UNITY
int launch_bsdmm_block(pmx_ctx* ctx, pmx_ctl* ctl, int block, float* X, const float* G, float* const* Z, float* const* U,
                       float* T, double* sums_scratch, int rows, int cols, int n_g, const ProxChain& direct,
                       const ProxChain* g, const float* step_f, double* norms, float e_rel, float e_abs, bool sharded,
                       double n_elems_global, long long xchg_off) {
  BsArgs a;
  memset(&a, 0, sizeof(a));
  a.X = X; a.G = G; a.T = T;
  a.n = (size_t)rows * cols; a.rows = rows; a.cols = cols;
  a.n_g = n_g; a.N_blocks = 2;
  a.direct = direct;
  for (int i = 0; i < n_g; ++i) { a.Z[i] = Z[i]; a.U[i] = U[i]; a.g[i] = g[i]; }
  a.step_f = step_f; a.norms = norms; a.ctl = ctl;
  const int grid = grid_for(ctx, a.n);
  if (chain_unity_axis(direct) != -1) {
    pmx_set_error("bsdmm: prox_unity as a *direct* constraint (prox_A/prox_S) is not supported; pass it in proxs_g");
    return PMX_ERR_UNSUPPORTED;
  }
  k_bsdmm_x<<<grid, kT, 0, ctx->stream>>>(a);
  PMX_LAUNCHED(ctx);
  for (int i = 0; i < n_g; ++i) {
    const bool fused = chain_unity_axis(g[i]) == -1;
    if (!fused) {
      k_add<<<grid, kT, 0, ctx->stream>>>(X, U[i], T, a.n, &ctl->done);
      PMX_LAUNCHED(ctx);
      StepSpec st;  // prox_g sees step_g = step_f * N * M_j; relative thresholds scale with it
      st.ptr = step_f; st.mode = 1; st.scale = (float)(2 * n_g); st.value = 0.f;
      PMX_CHECK(apply_chain_general(ctx, g[i], T, rows, cols, st, sums_scratch, &ctl->done));
    }
    k_bsdmm_zu<<<grid, kT, 0, ctx->stream>>>(a, i, fused ? 1 : 0);
    PMX_LAUNCHED(ctx);
  }
  if (sharded && ctx->world > 1) {   // the S block is column-sharded: its norms are sums over the ranks
    const int cnt = n_g > 0 ? n_g * 5 : 5;
    if (xchg_off >= 0)   // inbox in the peer arena (comm.cu): one small kernel instead of an NCCL all-reduce
      PMX_CHECK(pmx_peer_small_allreduce(ctx, 1, (size_t)xchg_off, norms, cnt, 1, ctx->stream, &ctl->done, &ctl->fault));
    else
      PMX_CHECK(pmx_comm_allreduce_internal(ctx, norms, (size_t)cnt, 1, ctx->stream));
  }
  k_bsdmm_block_finalize<<<1, 1, 0, ctx->stream>>>(ctl, norms, n_g, n_elems_global, e_rel, e_abs, block);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "bsdmm block");
}

int launch_bsdmm_iter_finalize(pmx_ctx* ctx, pmx_ctl* ctl) {
  k_bsdmm_iter_finalize<<<1, 1, 0, ctx->stream>>>(ctl);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_bsdmm_iter_finalize");
}
