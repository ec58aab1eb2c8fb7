// lambda_max of a small symmetric PSD Gram matrix, computed cooperatively by ONE thread block.
//
// Replaces np.linalg.eigvals(L^T L).max() of utils.py:20,34 (the Lipschitz constants of nmf.py:44-49) on the device:
// repeated squaring  B <- B^2 / trace  drives B to v1 v1^T for ANY spectral gap (adaptive exit once the power is rank
// one to fp32 resolution: 2-3 squarings for the positive Gram matrices of NMF), then fp64 power steps and a Rayleigh
// quotient with the ORIGINAL fp64 Gram give lambda_max to ~1e-7 relative -- the accuracy LAPACK geev gives the
// reference in fp32.  Used by k_lambda_max (gram.cu) and by the final phase of the fused PGM tail (pgm_tail.cu).
// Every reduction is a warp-shuffle tree + one shared-memory hop; no single-thread loops.
#pragma once
#include "common.cuh"

namespace lmax {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// The cooperating threads: a whole block (bar = 0: __syncthreads) or a group of `nthr` consecutive threads of a larger
// block (bar > 0: named barrier `bar`, nthr a multiple of 32) -- the final phase of the fused PGM tail runs two
// eigen-solves side by side in one 512-thread block.
struct Group {
  int tid, nthr, bar;
  __device__ __forceinline__ void sync() const {
    if (bar == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthr) : "memory");
  }
};
__device__ __forceinline__ Group whole_block() {
  Group g;
  g.tid = threadIdx.x;
  g.nthr = blockDim.x;
  g.bar = 0;
  return g;
}

// sum of one double per thread over the group; every thread gets the result.  red: >= 33 doubles of shared memory
__device__ __forceinline__ double block_sum_d(double v, double* red, const Group& grp) {
  const int lane = grp.tid & 31, w = grp.tid >> 5, nw = (grp.nthr + 31) >> 5;
  v = warp_sum_d(v);
  grp.sync();   // red may still be read from a previous call
  if (lane == 0) red[w] = v;
  grp.sync();
  if (w == 0) {
    double t = lane < nw ? red[lane] : 0.0;
    t = warp_sum_d(t);
    if (lane == 0) red[32] = t;
  }
  grp.sync();
  return red[32];
}

// acc[4][4] += sum_l T[l][i0..i0+3] * T[l][j0..j0+3]   (T symmetric: T^T T = T^2)
__device__ __forceinline__ void sq_block(const float* T, int len, int ldt, int i0, int j0, float acc[4][4]) {
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

// shared memory the routine needs for dimension C (bytes); layout: B0, B1 [C4][C4+4] floats, v, w [C4] doubles,
// red [40] doubles
__host__ __device__ inline size_t smem_bytes(int C) {
  const int C4 = (C + 3) & ~3;
  return 2 * (size_t)C4 * (C4 + 4) * sizeof(float) + (2 * (size_t)C4 + 40) * sizeof(double);
}

// gram: C x C fp64, row-major (global or shared memory; read-only here).  Returns lambda_max to every thread.
// status (same for every thread): 0 ok, 1 non-finite input, 2 zero matrix (lambda = 0).
// smem: smem_bytes(C) bytes, 16-byte aligned.  All threads of the group must call this.
__device__ inline double block_lambda_max(const double* gram, int C, unsigned char* smem, int max_squarings, int* status,
                                          const Group& grp) {
  const int C4 = (C + 3) & ~3;
  const int ldt = C4 + 4;
  float* B0 = reinterpret_cast<float*>(smem);
  float* B1 = B0 + (size_t)C4 * ldt;
  double* v = reinterpret_cast<double*>(B1 + (size_t)C4 * ldt);
  double* w = v + C4;
  double* red = w + C4;            // [40]: 0..32 block_sum scratch, 33 flag, 34 best index
  const int tid = grp.tid, nthr = grp.nthr;
  grp.sync();
  if (tid == 0) red[33] = 0.0;
  grp.sync();
  // load (fp64 -> fp32 working copy), detect non-finite entries, trace
  double tr = 0.0;
  bool bad = false;
  for (int idx = tid; idx < C4 * C4; idx += nthr) {
    const int i = idx / C4, j = idx - i * C4;
    const double g = (i < C && j < C) ? gram[(size_t)i * C + j] : 0.0;
    if (!isfinite(g)) bad = true;
    if (i == j) tr += g;
    B0[i * ldt + j] = (float)g;
  }
  if (bad) red[33] = 1.0;
  const double trace0 = block_sum_d(tr, red, grp);
  if (red[33] != 0.0 || !isfinite(trace0)) {
    *status = 1;
    return trace0;
  }
  if (trace0 <= 0.0) {   // zero matrix: lambda_max = 0 (the reference then divides by zero: step = inf)
    *status = 2;
    return 0.0;
  }
  *status = 0;
  {  // normalise by the trace so that every power keeps its entries in (0, 1]
    const float inv = (float)(1.0 / trace0);
    for (int idx = tid; idx < C4 * ldt; idx += nthr) B0[idx] *= inv;
  }
  grp.sync();
  const int nt = C4 / 4, ntiles = nt * nt;
  float* cur = B0;
  float* nxt = B1;
  for (int s = 0; s < max_squarings; ++s) {
    float tr_part = 0.f;
    for (int tile = tid; tile < ntiles; tile += nthr) {
      const int i0 = (tile / nt) * 4, j0 = (tile % nt) * 4;
      float acc[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
      sq_block(cur, C4, ldt, i0, j0, acc);
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          nxt[(i0 + p) * ldt + j0 + q] = acc[p][q];
          if (i0 + p == j0 + q) tr_part += acc[p][q];
        }
    }
    // tr(B^2) with tr(B) = 1 is sum_i w_i^2 of the normalised spectrum: 1 - tr(B^2) ~ 2 delta, delta = weight outside
    // the leading eigenvector.  The squarings only have to deliver a START vector for the fp64 power steps below: a
    // column of B is v1 contaminated by eps <= sqrt(C) delta, and the Rayleigh quotient after two more power steps
    // is off by eps^2 rho^4 (1 - rho) <= 0.08 eps^2 (rho = lambda_2 / lambda_1).  delta <= 5e-5 (eps <= 4e-4 for
    // C = 64) leaves < 2e-8 relative -- and the threshold sits well above the fp32 noise of the trace (~1e-5), so
    // the exit is reliable (a 2e-6 threshold was not: it sat inside the noise and cost up to 20 squarings).
    const double tr2 = block_sum_d((double)tr_part, red, grp);
    const bool rank_one = (1.0 - tr2) < 1e-4;
    const float inv = (float)(1.0 / tr2);
    for (int idx = tid; idx < C4 * ldt; idx += nthr) nxt[idx] *= inv;
    grp.sync();
    float* tmp = cur;
    cur = nxt;
    nxt = tmp;
    if (rank_one) break;   // block-uniform
  }
  // v = column of the (near rank-one) power with the largest diagonal entry: argmax by one warp
  if (tid < 32) {
    float bd = -1.f;
    int best = 0;
    for (int i = tid; i < C; i += 32) {
      const float dgn = cur[i * ldt + i];
      if (dgn > bd) {
        bd = dgn;
        best = i;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best, o);
      if (ob > bd || (ob == bd && oi < best)) {
        bd = ob;
        best = oi;
      }
    }
    if (tid == 0) red[34] = (double)best;
  }
  grp.sync();
  const int col = (int)red[34];
  for (int i = tid; i < C4; i += nthr) v[i] = (i < C) ? (double)cur[i * ldt + col] : 0.0;
  grp.sync();
  // fp64 power steps with the original Gram (4 lanes per row), then the Rayleigh quotient
  double lam = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    for (int i0 = 0; i0 < C; i0 += nthr >> 2) {   // uniform trip count: the shuffles need whole warps
      const int i = i0 + (tid >> 2);
      double acc = 0.0;
      if (i < C)
        for (int j = tid & 3; j < C; j += 4) acc += gram[(size_t)i * C + j] * v[j];
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (i < C && (tid & 3) == 0) w[i] = acc;
    }
    grp.sync();
    if (rep == 2) {
      double num = 0.0, den = 0.0;
      for (int i = tid; i < C; i += nthr) {
        num += v[i] * w[i];
        den += v[i] * v[i];
      }
      num = block_sum_d(num, red, grp);
      den = block_sum_d(den, red, grp);
      lam = den > 0 ? num / den : 0.0;
      break;
    }
    double nn = 0.0;
    for (int i = tid; i < C; i += nthr) nn += w[i] * w[i];
    nn = block_sum_d(nn, red, grp);
    const double sc = nn > 0 ? 1.0 / sqrt(nn) : 0.0;
    for (int i = tid; i < C; i += nthr) v[i] = w[i] * sc;
    grp.sync();
  }
  return lam;
}

}  // namespace lmax
