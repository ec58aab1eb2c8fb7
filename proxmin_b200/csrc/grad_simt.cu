// SIMT fp32 gradient kernel: G_A = (A S - Y) S^T, G_S = A^T (A S - Y), loss = |A S - Y|^2 / 2
// (nmf.py:13-41).  It handles every shape (any M, N, K <= 128) and is the path for shapes the
// tcgen05 kernel does not take (K > 64) as well as its on-device cross-check.  One CTA owns a
// 64 x 64 tile of Y: the residual tile never leaves shared memory, so Y is read exactly once and
// no M x N temporary exists (the reference materialises three of them, nmf.py:40).
#include "common.cuh"
#include "grad_umma.h"

int launch_zero(pmx_ctx* ctx, cudaStream_t st, float* p, size_t n, const int* done);

namespace {

constexpr int BM = 64, BN = 64;
constexpr int LDR = BN + 1;  // residual tile row stride (odd: conflict-free row-strided reads)

__global__ void __launch_bounds__(256) k_grad_simt(const float* __restrict__ Y, const float* __restrict__ W, int ldY, int y_blocked, const float* __restrict__ A,
                                                   const float* __restrict__ S, int M, int N, int K,
                                                   float* __restrict__ GA, float* __restrict__ GS,
                                                   double* __restrict__ loss, const int* done) {
  if (done && *done) return;
  extern __shared__ __align__(16) float smem[];
  const int ldA = K + 1;
  const int ldS = BN + 4;
  float* sA = smem;                 // [BM][ldA]
  float* sS = sA + BM * ldA;        // [K][ldS]   (offset keeps 16B alignment only if BM*ldA % 4 == 0, see below)
  // keep sS 16-byte aligned for the float4 reads
  sS = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sS) + 15) & ~uintptr_t(15));
  float* sR = sS + (size_t)K * ldS; // [BM][LDR]

  const int tid = threadIdx.x;
  const int tiles_n = (N + BN - 1) / BN;
  const int tiles_m = (M + BM - 1) / BM;
  const long long ntiles = (long long)tiles_m * tiles_n;
  float loss_part = 0.f;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tm = (int)(tile / tiles_n), tn = (int)(tile % tiles_n);
    const int m_base = tm * BM, n_base = tn * BN;
    __syncthreads();
    for (int idx = tid; idx < BM * K; idx += blockDim.x) {
      const int m = idx / K, k = idx - m * K;
      sA[m * ldA + k] = (m_base + m < M) ? A[(size_t)(m_base + m) * K + k] : 0.f;
    }
    for (int idx = tid; idx < K * BN; idx += blockDim.x) {
      const int k = idx / BN, n = idx - k * BN;
      sS[k * ldS + n] = (n_base + n < N) ? S[(size_t)k * N + n_base + n] : 0.f;
    }
    __syncthreads();
    // residual micro-tile 4 x 4
    const int m0 = (tid >> 4) * 4, n0 = (tid & 15) * 4;
    float r[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) r[p][q] = 0.f;
    for (int k = 0; k < K; ++k) {
      const float4 s4 = *reinterpret_cast<const float4*>(sS + k * ldS + n0);
      const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float a = sA[(m0 + p) * ldA + k];
#pragma unroll
        for (int q = 0; q < 4; ++q) r[p][q] = fmaf(a, sv[q], r[p][q]);
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = m_base + m0 + p, n = n_base + n0 + q;
        float d = 0.f;
        float rr = 0.f;
        if (m < M && n < N) {
          const size_t yi = pmx_y_index(m, n, ldY, y_blocked);
          rr = r[p][q] - Y[yi];                                                  // nmf.py:40
          d = W ? W[yi] * rr : rr;                                               // D = W (A S - Y)
        }
        sR[(m0 + p) * LDR + n0 + q] = d;
        loss_part = fmaf(d, rr, loss_part);                                      // nmf.py:25: sum W (Y - A S)^2 / 2
      }
    __syncthreads();
    // G_A[m, k] += sum_n R[m, n] S[k, n]         (nmf.py:41, D.dot(S.T))
    {
      const int m = tid >> 2;
      if (m_base + m < M)
        for (int k = tid & 3; k < K; k += 4) {
          float acc = 0.f;
          const float* rr = sR + m * LDR;
          const float* ss = sS + k * ldS;
#pragma unroll 8
          for (int n = 0; n < BN; ++n) acc = fmaf(rr[n], ss[n], acc);
          atomicAdd(&GA[(size_t)(m_base + m) * K + k], acc);
        }
    }
    // G_S[k, n] += sum_m A[m, k] R[m, n]         (nmf.py:41, A.T.dot(D))
    {
      const int n = tid & 63;
      if (n_base + n < N)
        for (int k = tid >> 6; k < K; k += 4) {
          float acc = 0.f;
#pragma unroll 8
          for (int m = 0; m < BM; ++m) acc = fmaf(sA[m * ldA + k], sR[m * LDR + n], acc);
          atomicAdd(&GS[(size_t)k * N + n_base + n], acc);
        }
    }
  }
  if (loss) {
    __shared__ float red[8];
    for (int o = 16; o > 0; o >>= 1) loss_part += __shfl_xor_sync(0xffffffffu, loss_part, o);
    if ((tid & 31) == 0) red[tid >> 5] = loss_part;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w];
      atomicAdd(loss, 0.5 * (double)t);
    }
  }
}

}  // namespace

// G_A, G_S (and *loss) must be zeroed by the caller-facing wrapper: done here on the same stream.
int launch_grad_simt(pmx_ctx* ctx, const float* Y, const float* W, int ldY, int y_blocked, const float* A, const float* S, int M, int N, int K, float* GA,
                     float* GS, double* loss, const int* done) {
  if (K > 128) {
    pmx_set_error("SIMT gradient kernel supports K <= 128 (got %d)", K);
    return PMX_ERR_UNSUPPORTED;
  }
  PMX_CHECK(launch_zero(ctx, ctx->stream, GA, (size_t)M * K, done));
  PMX_CHECK(launch_zero(ctx, ctx->stream, GS, (size_t)K * N, done));
  if (loss) PMX_CHECK(launch_zero(ctx, ctx->stream, reinterpret_cast<float*>(loss), 2, done));
  const size_t smem = sizeof(float) * ((size_t)BM * (K + 1) + (size_t)K * (BN + 4) + (size_t)BM * LDR) + 16;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    PMX_CUDA(cudaFuncSetAttribute(k_grad_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const long long ntiles = (long long)pmx_div_up(M, BM) * pmx_div_up(N, BN);
  long long blocks = ntiles < (long long)ctx->sm_count * 4 ? ntiles : (long long)ctx->sm_count * 4;
  if (blocks < 1) return PMX_OK;
  k_grad_simt<<<(int)blocks, 256, smem, ctx->stream>>>(Y, W, ldY, y_blocked, A, S, M, N, K, GA, GS, loss, done);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_grad_simt");
}

// ---------------------------------------------------------------- row-major staging buffer -> tiled Y (pmx_nmf_set_Y)
namespace {
__global__ void __launch_bounds__(256) k_y_interleave(const float* __restrict__ stage, int pitch, int nrows, int ncols,
                                                      float* __restrict__ Yb, int ldY, int m0, int col0) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int q = blockIdx.y;                     // row quad inside the chunk
  if (n >= ncols) return;
  float v[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) v[r] = 4 * q + r < nrows ? __ldcs(stage + (size_t)(4 * q + r) * pitch + n) : 0.f;
  float4* dst = reinterpret_cast<float4*>(Yb + pmx_y_index(m0 + 4 * q, col0 + n, ldY, 1));
  *dst = make_float4(v[0], v[1], v[2], v[3]);
}
}  // namespace

int launch_y_interleave(pmx_ctx* ctx, cudaStream_t st, const float* stage, int pitch, int nrows, int ncols, float* Yb, int ldY,
                        int m0, int col0) {
  if (nrows <= 0 || ncols <= 0) return PMX_OK;
  dim3 grid((unsigned)pmx_div_up(ncols, 256), (unsigned)pmx_div_up(nrows, 4));
  k_y_interleave<<<grid, 256, 0, st>>>(stage, pitch, nrows, ncols, Yb, ldY, m0, col0);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_y_interleave");
}
