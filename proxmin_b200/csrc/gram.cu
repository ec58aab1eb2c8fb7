// Lipschitz constants of the NMF gradient (nmf.py:44-49 step_A/step_S, utils.py:14-35
// get_spectral_norm dense branch): lambda_max of the K x K Gram matrices S S^T and A^T A.
//
//   k_gram<TALL>   : Gram += X^T X (tall M x K operand) or X X^T (wide K x N operand), fp32 tiles in
//                    shared memory, 4x4 register blocks, persistent blocks, one fp64 atomic flush per block.
//   k_lambda_max   : one 256-thread CTA per Gram (algorithm in lambda_max.cuh: repeated squaring + fp64 power steps
//                    + Rayleigh quotient, ~1e-7 relative).  Writes lip / step = 1/lip into the control block; flags
//                    non-finite input (reference: numpy.linalg.LinAlgError from eigvals, utils.py:34).
#include "common.cuh"
#include "lambda_max.cuh"

int launch_zero(pmx_ctx* ctx, cudaStream_t st, float* p, size_t n, const int* done);
int launch_gram_reduce(pmx_ctx* ctx, cudaStream_t st, const float* part, int nblocks, int C, double* gram, const int* done);

namespace {

constexpr int kMaxK = 128;
constexpr int kTileLen = 32;  // rows (tall) / columns (wide) staged per step

// acc[4][4] += sum_l T[l][i0..i0+3] * T[l][j0..j0+3]
__device__ __forceinline__ void gram_block(const float* T, int len, int ldt, int i0, int j0, float acc[4][4]) {
  for (int l = 0; l < len; ++l) {
    const float* row = T + (size_t)l * ldt;
    const float4 a = *reinterpret_cast<const float4*>(row + i0);
    const float4 b = *reinterpret_cast<const float4*>(row + j0);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
  }
}

// X is rows x cols row-major.  TALL: Gram (cols x cols) over rows.  !TALL: Gram (rows x rows) over cols.
template <bool TALL>
__global__ void __launch_bounds__(256) k_gram(const float* __restrict__ X, int rows, int cols, float* __restrict__ part,
                                              const int* done) {
  if (done && *done) return;
  extern __shared__ __align__(16) float smem[];
  const int C = TALL ? cols : rows;        // Gram dimension
  const long long L = TALL ? rows : cols;  // contraction length
  const int C4 = (C + 3) & ~3;
  const int ldt = C4 + 4;                  // padded row of the staged tile
  float* T = smem;                         // [kTileLen][ldt]
  const int nt = C4 / 4;
  const int ntiles = nt * nt;
  // each thread owns up to 4 register blocks (C <= 128 -> at most 1024 blocks for 256 threads)
  float acc[4][4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[u][p][q] = 0.f;

  const long long nchunks = (L + kTileLen - 1) / kTileLen;
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long l0 = ch * kTileLen;
    const int len = (int)min((long long)kTileLen, L - l0);
    __syncthreads();
    if (TALL) {  // rows l0..l0+len of X, contiguous in memory
      for (int idx = threadIdx.x; idx < kTileLen * C4; idx += blockDim.x) {
        const int l = idx / C4, c = idx - l * C4;
        T[l * ldt + c] = (l < len && c < C) ? X[(size_t)(l0 + l) * cols + c] : 0.f;
      }
    } else {  // columns l0..l0+len of every row: coalesced along the column index
      for (int idx = threadIdx.x; idx < kTileLen * C4; idx += blockDim.x) {
        const int c = idx / kTileLen, l = idx - c * kTileLen;
        T[l * ldt + c] = (l < len && c < C) ? X[(size_t)c * cols + (l0 + l)] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int tile = threadIdx.x + u * 256;
      if (tile < ntiles) gram_block(T, len, ldt, (tile / nt) * 4, (tile % nt) * 4, acc[u]);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int tile = threadIdx.x + u * 256;
    if (tile < ntiles) {
      const int i0 = (tile / nt) * 4, j0 = (tile % nt) * 4;
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (i0 + p < C && j0 + q < C) part[(size_t)blockIdx.x * C * C + (size_t)(i0 + p) * C + (j0 + q)] = acc[u][p][q];
    }
  }
}

// gram[e] += sum over a slice of the blocks of part[b][e], in fp64 (gram pre-zeroed; gridDim.y slices run in
// parallel so that no thread walks more than ~32 partials; the per-block partials are fp32 sums of <= a few
// hundred terms)
__global__ void __launch_bounds__(256) k_gram_reduce(const float* __restrict__ part, int nblocks, int n, double* __restrict__ gram,
                                                     const int* done) {
  if (done && *done) return;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int per = (nblocks + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * per, b1 = min(nblocks, b0 + per);
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int b = b0;
  for (; b + 8 <= b1; b += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = part[(size_t)(b + u) * n + e];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] += (double)v[u];
  }
  for (; b < b1; ++b) acc[0] += (double)part[(size_t)b * n + e];
  if (b1 > b0) atomicAdd(&gram[e], ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7])));
}

// One CTA per Gram matrix (gridDim.x = 1 or 2): 256 threads for C <= 64, 1024 above (the squarings are C^3 FMAs).  gram: C x C fp64 (symmetric PSD).
// which: 0 -> lip[0]/step[0], 1 -> lip[1]/step[1].  The algorithm lives in lambda_max.cuh (shared with the fused PGM
// tail); 36 KB of shared memory for C = 64.
__global__ void __launch_bounds__(1024) k_lambda_max(const double* __restrict__ gram0, int which0,
                                                    const double* __restrict__ gram1, int which1, int C,
                                                    pmx_ctl* ctl, int squarings) {
  if (ctl->done) return;
  const double* __restrict__ gram = blockIdx.x == 0 ? gram0 : gram1;
  const int which = blockIdx.x == 0 ? which0 : which1;
  extern __shared__ __align__(16) unsigned char lm_smem[];
  int status = 0;
  const double lam = lmax::block_lambda_max(gram, C, lm_smem, squarings, &status, lmax::whole_block());
  if (threadIdx.x != 0) return;
  if (status == 1) {        // reference: numpy.linalg.LinAlgError from eigvals (utils.py:34)
    ctl->nonfinite = 1;
    ctl->done = 1;
    ctl->lip[which] = __int_as_float(0x7fc00000);
    ctl->step[which] = __int_as_float(0x7fc00000);
  } else if (status == 2) {  // zero matrix: lambda_max = 0, step = 1/0 = inf (the reference divides by zero too)
    ctl->lip[which] = 0.f;
    ctl->step[which] = __int_as_float(0x7f800000);
  } else {
    const float lf = (float)lam;             // the reference's eigvals runs in fp32 for fp32 inputs
    ctl->lip[which] = lf;
    ctl->step[which] = 1.0f / lf;            // nmf.py:45,49
  }
}

}  // namespace


int launch_gram(pmx_ctx* ctx, cudaStream_t st, const float* X, int rows, int cols, bool tall, double* gram,
                const int* done) {
  const int C = tall ? cols : rows;
  if (C > kMaxK) {
    pmx_set_error("Gram dimension K=%d exceeds the supported maximum %d", C, kMaxK);
    return PMX_ERR_UNSUPPORTED;
  }
  const long long L = tall ? rows : cols;
  const int C4 = (C + 3) & ~3;
  const size_t smem = (size_t)kTileLen * (C4 + 4) * sizeof(float);
  long long nchunks = (L + kTileLen - 1) / kTileLen;
  // enough blocks to keep loads in flight on every SM, few enough that the fp64 reduction stays tiny
  int blocks = (int)(nchunks < (long long)ctx->sm_count * 2 ? nchunks : (long long)ctx->sm_count * 2);
  if (blocks < 1) blocks = 1;
  // scratch for the per-block partials: owned by the context, one buffer per (tall/wide, stream) so that two
  // contexts of one process or the two streams of a context never share it; grown on demand
  const int w = (tall ? 1 : 0) + (st == ctx->aux ? 2 : 0);
  const size_t need = sizeof(float) * (size_t)blocks * C * C;
  if (need > ctx->gram_scratch_bytes[w]) {
    if (ctx->gram_scratch[w]) {
      cudaStreamSynchronize(st);   // a previous launch may still read the old buffer
      cudaFree(ctx->gram_scratch[w]);
      ctx->gram_scratch[w] = nullptr;
      ctx->gram_scratch_bytes[w] = 0;
    }
    if (cudaMalloc((void**)&ctx->gram_scratch[w], need) != cudaSuccess) {
      pmx_set_error("cudaMalloc of the Gram scratch failed");
      return PMX_ERR_CUDA;
    }
    ctx->gram_scratch_bytes[w] = need;
  }
  float* part = ctx->gram_scratch[w];
  if (tall)
    k_gram<true><<<blocks, 256, smem, st>>>(X, rows, cols, part, done);
  else
    k_gram<false><<<blocks, 256, smem, st>>>(X, rows, cols, part, done);
  PMX_LAUNCHED(ctx);
  return launch_gram_reduce(ctx, st, part, blocks, C, gram, done);
}

int launch_lambda_max2(pmx_ctx* ctx, cudaStream_t st, const double* gram0, int which0, const double* gram1, int which1,
                       int C, pmx_ctl* ctl);

int launch_lambda_max(pmx_ctx* ctx, cudaStream_t st, const double* gram, int C, pmx_ctl* ctl, int which) {
  return launch_lambda_max2(ctx, st, gram, which, nullptr, 0, C, ctl);
}

// one launch for one or two Gram matrices (two CTAs run concurrently)
int launch_lambda_max2(pmx_ctx* ctx, cudaStream_t st, const double* gram0, int which0, const double* gram1, int which1,
                       int C, pmx_ctl* ctl) {
  if (C > kMaxK) {
    pmx_set_error("Gram dimension K=%d exceeds the supported maximum %d", C, kMaxK);
    return PMX_ERR_UNSUPPORTED;
  }
  const size_t smem = lmax::smem_bytes(C);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_lambda_max, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) {
      pmx_set_error("cudaFuncSetAttribute(k_lambda_max): %s", cudaGetErrorString(e));
      return PMX_ERR_CUDA;
    }
    attr_set = true;
  }
  k_lambda_max<<<gram1 ? 2 : 1, C <= 64 ? 256 : 1024, smem, st>>>(gram0, which0, gram1, which1, C, ctl, 20);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_lambda_max");
}

// partial Gram matrices written by another kernel (fused S update): sum them into `gram` (fp64)
int launch_gram_reduce(pmx_ctx* ctx, cudaStream_t st, const float* part, int nblocks, int C, double* gram, const int* done) {
  PMX_CHECK(launch_zero(ctx, st, reinterpret_cast<float*>(gram), 2 * (size_t)C * C, done));
  const int slices = nblocks >= 64 ? 16 : (nblocks >= 8 ? 4 : 1);
  k_gram_reduce<<<dim3(pmx_div_up(C * C, 256), slices), 256, 0, st>>>(part, nblocks, C * C, gram, done);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_gram_reduce");
}
