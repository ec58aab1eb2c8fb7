// Elementwise expression kernels for the *callback loops* of the Python solvers: when grad / step /
// prox are arbitrary user callables the iterates live in host arrays, but every arithmetic expression
// of the library itself (algorithms.py:94-95,118,121,378,387; utils.py:295-363) still runs here.
// One kernel, selected by opcode; operation order and roundings follow the NumPy expressions.
#include "kernels.h"

namespace {

constexpr int kT = 256;

struct EwArgs {
  int op;
  size_t n;
  const float *a, *b, *c, *d;
  float s0, s1;
  float *o0, *o1, *o2;
  double* red;  // up to 5 reductions (sum) or 1 (max, as float bits)
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = (v != v || t != t) ? (v + t) : fmaxf(v, t);
  }
  return v;
}

__global__ void __launch_bounds__(kT) k_ew(EwArgs e) {
  float r[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  float mx = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += (size_t)gridDim.x * blockDim.x) {
    switch (e.op) {
      case PMX_EW_EXTRAP: {  // X + omega * (X - X_)                      algorithms.py:95
        const float x = e.a[i];
        e.o0[i] = __fadd_rn(x, __fmul_rn(e.s0, __fsub_rn(x, e.b[i])));
        break;
      }
      case PMX_EW_ADD:       // LX + U                                     utils.py:297
        e.o0[i] = __fadd_rn(e.a[i], e.b[i]);
        break;
      case PMX_EW_SUB:       // X - dX                                     utils.py:317,338
        e.o0[i] = __fsub_rn(e.a[i], e.b[i]);
        break;
      case PMX_EW_DX_ACC: {  // d + step_f/step_g * (X - Z + U)            utils.py:316,333 (d may be NULL = 0)
        const float t = __fmul_rn(e.s0, __fadd_rn(__fsub_rn(e.a[i], e.b[i]), e.c[i]));
        e.o0[i] = e.d ? __fadd_rn(e.d[i], t) : t;
        break;
      }
      case PMX_EW_ZU: {      // R, S, U += R and the five norms            utils.py:299-303, 349-363
        const float x = e.a[i], zn = e.b[i], z = e.c[i], u = e.d[i];
        const float rr = __fsub_rn(x, zn);
        const float ss = __fmul_rn(e.s0, __fsub_rn(zn, z));
        const float un = __fadd_rn(u, rr);
        e.o0[i] = rr;
        e.o1[i] = ss;
        e.o2[i] = un;
        const float uq = e.s1 != 0.f ? __fdiv_rn(un, e.s1) : un;
        r[0] = fmaf(x, x, r[0]);
        r[1] = fmaf(zn, zn, r[1]);
        r[2] = fmaf(uq, uq, r[2]);
        r[3] = fmaf(rr, rr, r[3]);
        r[4] = fmaf(ss, ss, r[4]);
        break;
      }
      case PMX_EW_DOT_DIFF: {  // sum((X - X_) * G), sum((X - X_)**2)       algorithms.py:118
        const float d = __fsub_rn(e.a[i], e.b[i]);
        r[0] = fmaf(d, e.c[i], r[0]);
        r[1] = fmaf(d, d, r[1]);
        break;
      }
      case PMX_EW_MAXABS: {    // max(abs(s0 * a))                          algorithms.py:121
        const float v = fabsf(__fmul_rn(e.s0, e.a[i]));
        mx = (v != v) ? v : fmaxf(mx, v);
        break;
      }
      case PMX_EW_SUMSQ:       // l2sq(a)                                   utils.py:257-260
        r[0] = fmaf(e.a[i], e.a[i], r[0]);
        break;
      case PMX_EW_AXPY:        // s0 * a + b (two roundings, like NumPy)       utils.py:316, 333
        e.o0[i] = e.b ? __fadd_rn(__fmul_rn(e.s0, e.a[i]), e.b[i]) : __fmul_rn(e.s0, e.a[i]);
        break;
      case PMX_EW_BB: {        // S = X - X_, Y = G - G_                    utils.py:225-239
        const float sd = e.a[i] - e.b[i], yd = e.c[i] - e.d[i], g = e.c[i];
        r[0] = fmaf(sd, sd, r[0]);
        r[1] = fmaf(sd, yd, r[1]);
        r[2] = fmaf(yd, yd, r[2]);
        r[3] = fmaf(g, g, r[3]);
        break;
      }
    }
  }
  if (!e.red) return;
  __shared__ float red[5][kT / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (e.op == PMX_EW_MAXABS) {
    mx = wmax(mx);
    if (lane == 0) red[0][w] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = red[0][0];
      for (int k = 1; k < kT / 32; ++k) m = (m != m || red[0][k] != red[0][k]) ? (m + red[0][k]) : fmaxf(m, red[0][k]);
      atomicMax(reinterpret_cast<int*>(e.red), __float_as_int(m));  // non-negative floats order like ints; NaN on top
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float v = wsum(r[k]);
    if (lane == 0) red[k][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    float t = 0.f;
    for (int k = 0; k < kT / 32; ++k) t += red[threadIdx.x][k];
    atomicAdd(e.red + threadIdx.x, (double)t);
  }
}

}  // namespace

extern "C" {

int pmx_ew(pmx_ctx* ctx, int op, size_t n, const float* a, const float* b, const float* c, const float* d, float s0,
           float s1, float* o0, float* o1, float* o2, double* red_host) {
  PMX_REQUIRE(ctx != nullptr, "ctx is NULL");
  PMX_REQUIRE(op >= PMX_EW_EXTRAP && op <= PMX_EW_AXPY, "unknown elementwise opcode");
  double* d_red = nullptr;
  if (red_host) {
    PMX_CUDA(cudaMalloc((void**)&d_red, 5 * sizeof(double)));
    PMX_CUDA(cudaMemsetAsync(d_red, 0, 5 * sizeof(double), ctx->stream));
  }
  EwArgs e;
  e.op = op; e.n = n; e.a = a; e.b = b; e.c = c; e.d = d; e.s0 = s0; e.s1 = s1; e.o0 = o0; e.o1 = o1; e.o2 = o2;
  e.red = d_red;
  long long blocks = (long long)((n + kT - 1) / kT);
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_ew<<<(int)blocks, kT, 0, ctx->stream>>>(e);
  PMX_LAUNCHED(ctx);
  int st = pmx_check_launch(ctx, "k_ew");
  if (red_host) {
    double h[5] = {0, 0, 0, 0, 0};
    if (st == PMX_OK) {
      cudaError_t err = cudaMemcpyAsync(h, d_red, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
      if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
      if (err != cudaSuccess) {
        pmx_set_error("pmx_ew: %s", cudaGetErrorString(err));
        st = PMX_ERR_CUDA;
      }
    }
    cudaFree(d_red);
    if (op == PMX_EW_MAXABS) {
      float f;
      memcpy(&f, h, sizeof(float));
      red_host[0] = (double)f;
    } else {
      memcpy(red_host, h, sizeof(h));
    }
  }
  return st;
}

// ---------------------------------------------------------------- dense linear operators  (utils.py:76-77)
// O[p x m] = L[p x n] X[n x m], fp32 with fma accumulation: the L.dot(X) / L.T.dot(..) of admm / sdmm / bsdmm with a
// non-identity L (utils.py:316, 333, 299-303).  L is small on this path (a K x K coupling or a p x n design matrix of a
// few thousand columns): a 32 x 32 shared-memory tiled kernel is all it needs.
}  // extern "C"
namespace {
constexpr int MT = 32;
__global__ void __launch_bounds__(MT * 8) k_matmul(const float* __restrict__ L, const float* __restrict__ X,
                                                   float* __restrict__ O, int p, int n, int m, int trans) {
  __shared__ float sL[MT][MT + 1];
  __shared__ float sX[MT][MT + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads, 4 output rows per thread
  const int col = blockIdx.x * MT + tx;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < n; k0 += MT) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int lr = ty + 8 * r;
      const int gi = blockIdx.y * MT + lr, gk = k0 + tx;
      sL[lr][tx] = (gi < p && gk < n) ? (trans ? L[(size_t)gk * p + gi] : L[(size_t)gi * n + gk]) : 0.f;
      const int xk = k0 + lr;
      sX[lr][tx] = (xk < n && col < m) ? X[(size_t)xk * m + col] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < MT; ++k) {
      const float xv = sX[k][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(sL[ty + 8 * r][k], xv, acc[r]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int gi = blockIdx.y * MT + ty + 8 * r;
    if (gi < p && col < m) O[(size_t)gi * m + col] = acc[r];
  }
}
}  // namespace
extern "C" {

int pmx_matmul(pmx_ctx* ctx, const float* L, const float* X, float* O, int p, int n, int m, int trans) {
  PMX_REQUIRE(ctx && L && X && O, "NULL argument");
  PMX_REQUIRE(p > 0 && n > 0 && m > 0, "shape must be positive");
  dim3 grid((unsigned)pmx_div_up(m, MT), (unsigned)pmx_div_up(p, MT));
  PMX_REQUIRE(grid.y <= 65535, "too many rows for pmx_matmul");
  k_matmul<<<grid, MT * 8, 0, ctx->stream>>>(L, X, O, p, n, m, trans);
  PMX_LAUNCHED(ctx);
  PMX_CHECK(pmx_check_launch(ctx, "k_matmul"));
  PMX_CUDA(cudaStreamSynchronize(ctx->stream));
  return PMX_OK;
}

// sums along an axis of a device matrix (np.mean / np.sum building block, nmf.py:91-93); out_host: cols (axis 0) or rows (axis 1) doubles
int pmx_axis_sum(pmx_ctx* ctx, const float* X, int rows, int cols, int axis, double* out_host) {
  PMX_REQUIRE(ctx && X && out_host, "NULL argument");
  PMX_REQUIRE(axis == 0 || axis == 1, "axis must be 0 or 1");
  const int n = axis == 0 ? cols : rows;
  double* d = nullptr;
  PMX_CUDA(cudaMalloc((void**)&d, sizeof(double) * (n ? n : 1)));
  int st = launch_axis_sum(ctx, X, rows, cols, axis, d, nullptr);
  if (st == PMX_OK) {
    cudaError_t err = cudaMemcpyAsync(out_host, d, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
    if (err != cudaSuccess) {
      pmx_set_error("pmx_axis_sum: %s", cudaGetErrorString(err));
      st = PMX_ERR_CUDA;
    }
  }
  cudaFree(d);
  return st;
}

// adaprox moment update + step on caller-owned device arrays (algorithms.py:147-245, :378); *psimax_host = max(Psi)
int pmx_adaprox_moments(pmx_ctx* ctx, int scheme, const float* G, float* M, float* V, float* Vhat_or_null, float* X,
                        float* Psi, int rows, int cols, const float* alpha_dev, int alpha_mode, float alpha_value,
                        double b1, double b1_prev, double b2, double eps, double p, int t, float* psimax_host) {
  PMX_REQUIRE(ctx && G && M && V && X && Psi && psimax_host, "NULL argument");
  PMX_REQUIRE(scheme >= PMX_ADAM && scheme <= PMX_RADAM, "unknown adaprox scheme");
  float* d_pm = nullptr;
  float* d_z = nullptr;
  const size_t n = (size_t)rows * cols;
  PMX_CUDA(cudaMalloc((void**)&d_pm, sizeof(float)));
  PMX_CUDA(cudaMemsetAsync(d_pm, 0, sizeof(float), ctx->stream));
  PMX_CUDA(cudaMalloc((void**)&d_z, sizeof(float) * (n ? n : 1)));
  AdaArgs a;
  memset(&a, 0, sizeof(a));
  a.G = G; a.M = M; a.V = V; a.Vhat = Vhat_or_null; a.X = X; a.Psi = Psi; a.Z = d_z; a.psimax = d_pm;
  a.n = n; a.rows = rows; a.cols = cols;
  a.alpha.ptr = alpha_dev; a.alpha.mode = alpha_mode; a.alpha.value = alpha_value; a.alpha.scale = 1.f;
  a.scheme = scheme; a.b1 = b1; a.b1_prev = b1_prev; a.b2 = b2; a.eps = eps; a.p = p; a.t = t;
  int st = launch_adaprox_moments(ctx, a);
  if (st == PMX_OK) {
    cudaError_t err = cudaMemcpyAsync(psimax_host, d_pm, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
    if (err != cudaSuccess) {
      pmx_set_error("pmx_adaprox_moments: %s", cudaGetErrorString(err));
      st = PMX_ERR_CUDA;
    }
  }
  cudaFree(d_pm);
  cudaFree(d_z);
  return st;
}

// one proximal sub-iteration z' = prox(z - gamma/Alpha * Psi * (z - X), gamma), gamma = Alpha / psimax
// (algorithms.py:384-389); norms_host = { |z'-z|^2, |z'|^2, |z|^2 }
int pmx_adaprox_sub(pmx_ctx* ctx, const pmx_prox* prox, const float* Z, const float* X, const float* Psi, float* Zout,
                    int rows, int cols, const float* alpha_dev, int alpha_mode, float alpha_value, float psimax,
                    double* norms_host) {
  PMX_REQUIRE(ctx && prox && Z && X && Psi && Zout && norms_host, "NULL argument");
  double* d_norms = nullptr;
  float* d_pm = nullptr;
  PMX_CUDA(cudaMalloc((void**)&d_norms, 3 * sizeof(double)));
  PMX_CUDA(cudaMemsetAsync(d_norms, 0, 3 * sizeof(double), ctx->stream));
  PMX_CUDA(cudaMalloc((void**)&d_pm, sizeof(float)));
  PMX_CUDA(cudaMemcpyAsync(d_pm, &psimax, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  ProxChain ch = make_chain(prox);
  UpdIO io;
  memset(&io, 0, sizeof(io));
  io.Xin = Z; io.Xprev = Z; io.Xout = Zout; io.X0 = X; io.G = Psi; io.psimax = d_pm; io.norms = d_norms;
  io.rows = rows; io.cols = cols;
  io.step.ptr = alpha_dev; io.step.mode = alpha_mode; io.step.value = alpha_value; io.step.scale = 1.f;
  int st = launch_update(ctx, IN_ADASUB, ch, io);
  if (st == PMX_OK) {
    cudaError_t err = cudaMemcpyAsync(norms_host, d_norms, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->stream);
    if (err != cudaSuccess) {
      pmx_set_error("pmx_adaprox_sub: %s", cudaGetErrorString(err));
      st = PMX_ERR_CUDA;
    }
  }
  cudaFree(d_norms);
  cudaFree(d_pm);
  return st;
}

}  // extern "C"
