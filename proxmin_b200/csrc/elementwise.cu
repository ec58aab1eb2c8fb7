// Elementwise / reduction kernels of the hot path (all HBM- or L2-bound):
//   k_upd_flat / k_upd_cols / k_upd_rows : transform -> prox chain -> store -> norms
//   k_extrapolate                         : Nesterov point  (algorithms.py:94-95)
//   k_split_bf16                          : fp32 -> (hi, lo) bf16 operands for the tcgen05 GEMMs
//   k_adaprox_moments                     : moment update + Phi/Psi step (algorithms.py:147-245, :378)
// Grids are sized to a multiple of the SM count; loads are coalesced (thread index runs
// along the contiguous dimension) and 128-bit where the layout allows.
#include <cuda_bf16.h>

#include "kernels.h"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide reduction of three partial sums, then one double atomicAdd per block and value
__device__ __forceinline__ void block_accumulate3(float a, float b, float c, double* out) {
  __shared__ float red[3][32];
  a = warp_sum(a);
  b = warp_sum(b);
  c = warp_sum(c);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    red[0][w] = a;
    red[1][w] = b;
    red[2][w] = c;
  }
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    a = lane < nw ? red[0][lane] : 0.f;
    b = lane < nw ? red[1][lane] : 0.f;
    c = lane < nw ? red[2][lane] : 0.f;
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    if (lane == 0) {
      atomicAdd(out + 0, (double)a);
      atomicAdd(out + 1, (double)b);
      atomicAdd(out + 2, (double)c);
    }
  }
}

__device__ __forceinline__ bool skip(const UpdIO& io) {
  return (io.done && *io.done) || (io.done2 && *io.done2);
}

__device__ __forceinline__ void store_split(const UpdIO& io, int r, int c, float v) {
  if (io.hi) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    const size_t j = (size_t)r * io.ld_split + c;
    io.hi[j] = __bfloat16_as_ushort(h);
    io.lo[j] = __bfloat16_as_ushort(l);
  }
}

// the value the prox chain starts from, and the step handed to the chain
template <int IN>
__device__ __forceinline__ float transform(const UpdIO& io, size_t i, int r, int c, float& pstep) {
  const float s = step_at(io.step, r, c);
  if (IN == IN_PGM) {  // algorithms.py:108  _X[j] - T[j]*S[j]*G[j]
    pstep = s;
    return __fsub_rn(io.Xin[i], __fmul_rn(s, io.G[i]));  // two roundings like NumPy, no FMA contraction
  }
  if (IN == IN_ADASUB) {  // algorithms.py:384,387  z - gamma/Alpha * Psi * (z - X), prox step gamma
    const float gamma = s / io.psimax[0];
    const float z = io.Xin[i];
    pstep = gamma;
    return z - gamma / s * io.G[i] * (z - io.X0[i]);
  }
  pstep = s;
  return io.Xin[i];
}

// the step the prox chain sees (thresholds of type="relative"), for passes after the first
template <int IN>
__device__ __forceinline__ float prox_step(const UpdIO& io, int r, int c) {
  const float s = step_at(io.step, r, c);
  return (IN == IN_ADASUB) ? s / io.psimax[0] : s;
}

// ---- elementwise-only chains: one thread per element, grid-stride -----------------------
template <int IN>
__global__ void __launch_bounds__(kThreads) k_upd_flat(ProxChain ch, UpdIO io) {
  if (skip(io)) return;
  const size_t n = (size_t)io.rows * io.cols;
  float nd = 0.f, nn = 0.f, np = 0.f;
  const bool need_rc = io.step.mode >= 2 || io.hi != nullptr;
  // fast path: 16-byte accesses, one 32-bit division per four elements (the scalar loop below costs a 64-bit division
  // and three dependent 4-byte loads per element: the adaprox sub-iteration ran at 1.6 TB/s with it)
  const bool vec = (io.cols & 3) == 0 && (n >> 2) < 0xffffffffull && !io.hi && !io.Xold_out &&
                   ((reinterpret_cast<uintptr_t>(io.Xin) | reinterpret_cast<uintptr_t>(io.Xout) |
                     reinterpret_cast<uintptr_t>(io.G) | reinterpret_cast<uintptr_t>(io.X0) |
                     reinterpret_cast<uintptr_t>(io.Xprev)) & 15) == 0;
  if (vec) {
    const unsigned n4 = (unsigned)(n >> 2), cols4 = (unsigned)io.cols >> 2;
    const bool prev_is_in = io.Xprev == io.Xin;
    for (unsigned g = blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += gridDim.x * blockDim.x) {
      unsigned r = 0, c = 0;
      if (need_rc) {
        r = g / cols4;
        c = (g - r * cols4) << 2;
      }
      const float4 xin4 = reinterpret_cast<const float4*>(io.Xin)[g];
      const float4 g4 = (IN != IN_PLAIN) ? reinterpret_cast<const float4*>(io.G)[g] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 x04 = (IN == IN_ADASUB) ? reinterpret_cast<const float4*>(io.X0)[g] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 pv4 = prev_is_in ? xin4 : (io.Xprev ? reinterpret_cast<const float4*>(io.Xprev)[g] : make_float4(0.f, 0.f, 0.f, 0.f));
      const float xi[4] = {xin4.x, xin4.y, xin4.z, xin4.w}, gi[4] = {g4.x, g4.y, g4.z, g4.w};
      const float x0[4] = {x04.x, x04.y, x04.z, x04.w}, pv[4] = {pv4.x, pv4.y, pv4.z, pv4.w};
      float out[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float s = step_at(io.step, (int)r, (int)c + q);
        float ps = s, v;
        if (IN == IN_PGM) {
          v = __fsub_rn(xi[q], __fmul_rn(s, gi[q]));
        } else if (IN == IN_ADASUB) {
          const float gamma = s / io.psimax[0];
          ps = gamma;
          v = xi[q] - gamma / s * gi[q] * (xi[q] - x0[q]);
        } else {
          v = xi[q];
        }
        v = chain_segment(ch, 0, ch.n, v, ps);
        out[q] = v;
        const float d = v - pv[q];
        nd += d * d;
        nn += v * v;
        np += pv[q] * pv[q];
      }
      reinterpret_cast<float4*>(io.Xout)[g] = make_float4(out[0], out[1], out[2], out[3]);
    }
    if (io.norms) block_accumulate3(nd, nn, np, io.norms);
    return;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int r = 0, c = 0;
    if (need_rc) {
      r = (int)(i / io.cols);
      c = (int)(i - (size_t)r * io.cols);
    }
    const float prev = io.Xprev ? io.Xprev[i] : 0.f;
    float ps;
    float v = transform<IN>(io, i, r, c, ps);
    v = chain_segment(ch, 0, ch.n, v, ps);
    if (io.Xold_out) io.Xold_out[i] = prev;
    io.Xout[i] = v;
    store_split(io, r, c, v);
    const float d = v - prev;
    nd += d * d;
    nn += v * v;
    np += prev * prev;
  }
  if (io.norms) block_accumulate3(nd, nn, np, io.norms);
}

// ---- chains with UNITY(axis=0): one thread owns a column (coalesced across threads) -----
// Rows are processed in batches of RB with all loads issued before the arithmetic, so that a thread
// keeps RB x (number of input streams) requests in flight: the pass is latency-bound otherwise.
// Fast path for short columns (rows <= 64, i.e. S with K <= 64) and a chain that ENDS with its only UNITY
// (prox_unity, prox_unity_plus): the column stays in registers between the sum and the division, so S is
// read once and written once (4 HBM streams instead of 7).
template <int IN>
__global__ void __launch_bounds__(128) k_upd_cols_reg(ProxChain ch, UpdIO io) {
  if (skip(io)) return;
  constexpr int RMAX = 64;
  float nd = 0.f, nn = 0.f, np = 0.f;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < io.cols) {
    const bool prev_is_in = (io.Xprev == io.Xin);
    const int b = ch.n - 1;   // position of the UNITY op
    float t[RMAX], pv[RMAX];
    float sum = 0.f;
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      if (r < io.rows) {
        const size_t i = (size_t)r * io.cols + c;
        const float xin = io.Xin[i];
        const float g = (IN != IN_PLAIN) ? io.G[i] : 0.f;
        const float x0 = (IN == IN_ADASUB) ? io.X0[i] : 0.f;
        pv[r] = prev_is_in ? xin : (io.Xprev ? io.Xprev[i] : 0.f);
        const float s = step_at(io.step, r, c);
        float ps = s, v;
        if (IN == IN_PGM) {
          v = __fsub_rn(xin, __fmul_rn(s, g));
        } else if (IN == IN_ADASUB) {
          const float gamma = s / io.psimax[0];
          ps = gamma;
          v = xin - gamma / s * g * (xin - x0);
        } else {
          v = xin;
        }
        v = chain_segment(ch, 0, b, v, ps);
        t[r] = v;
        sum += v;   // row order, like NumPy's axis-0 reduction
      }
    }
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      if (r < io.rows) {
        const size_t i = (size_t)r * io.cols + c;
        const float v = t[r] / sum;                      // operators.py:44
        if (io.Xold_out) io.Xold_out[i] = pv[r];
        io.Xout[i] = v;
        store_split(io, r, c, v);
        const float d = v - pv[r];
        nd += d * d; nn += v * v; np += pv[r] * pv[r];
      }
    }
  }
  if (io.norms) block_accumulate3(nd, nn, np, io.norms);
}

template <int IN>
__global__ void __launch_bounds__(128) k_upd_cols(ProxChain ch, UpdIO io) {
  if (skip(io)) return;
  constexpr int RB = 8;
  extern __shared__ __align__(16) float gtile[];   // [rows][GLD] final values of this block's 128 columns (Gram fusion)
  constexpr int GLD = 129;
  const bool want_gram = io.gram_part != nullptr;
  float nd = 0.f, nn = 0.f, np = 0.f;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (want_gram && c >= io.cols)
    for (int r = 0; r < io.rows; ++r) gtile[r * GLD + threadIdx.x] = 0.f;
  if (c < io.cols) {
    const bool alias = (io.Xprev == io.Xout);
    const float* prev_src = alias ? io.Xold_out : io.Xprev;
    const bool prev_is_in = (io.Xprev == io.Xin);
    int a = 0, b = chain_next_unity(ch, 0);
    // pass 0: transform + first elementwise segment; column sum in row order (NumPy's order for axis=0)
    float sum = 0.f;
    for (int r0 = 0; r0 < io.rows; r0 += RB) {
      float xin[RB], g[RB], x0[RB], prev[RB];
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const int r = r0 + j;
        if (r < io.rows) {
          const size_t i = (size_t)r * io.cols + c;
          xin[j] = io.Xin[i];
          g[j] = (IN != IN_PLAIN) ? io.G[i] : 0.f;
          x0[j] = (IN == IN_ADASUB) ? io.X0[i] : 0.f;
          prev[j] = prev_is_in ? xin[j] : (io.Xprev ? io.Xprev[i] : 0.f);
        }
      }
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const int r = r0 + j;
        if (r < io.rows) {
          const size_t i = (size_t)r * io.cols + c;
          const float s = step_at(io.step, r, c);
          float ps = s, v;
          if (IN == IN_PGM) {
            v = __fsub_rn(xin[j], __fmul_rn(s, g[j]));
          } else if (IN == IN_ADASUB) {
            const float gamma = s / io.psimax[0];
            ps = gamma;
            v = xin[j] - gamma / s * g[j] * (xin[j] - x0[j]);
          } else {
            v = xin[j];
          }
          v = chain_segment(ch, a, b, v, ps);
          if (io.Xold_out) io.Xold_out[i] = prev[j];
          io.Xout[i] = v;
          sum += v;
          if (b >= ch.n) {
            store_split(io, r, c, v);
            if (want_gram) gtile[r * GLD + threadIdx.x] = v;
            const float d = v - prev[j];
            nd += d * d; nn += v * v; np += prev[j] * prev[j];
          }
        }
      }
    }
    // one further pass per UNITY op: divide by the column sum, then the next elementwise segment
    while (b < ch.n) {
      a = b + 1;
      b = chain_next_unity(ch, a);
      const float denom = sum;
      sum = 0.f;
      const bool last = b >= ch.n;
      for (int r0 = 0; r0 < io.rows; r0 += RB) {
        float cur[RB], prev[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          const int r = r0 + j;
          if (r < io.rows) {
            const size_t i = (size_t)r * io.cols + c;
            cur[j] = io.Xout[i];
            prev[j] = (last && prev_src) ? prev_src[i] : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          const int r = r0 + j;
          if (r < io.rows) {
            const size_t i = (size_t)r * io.cols + c;
            float v = cur[j] / denom;                     // operators.py:44
            v = chain_segment(ch, a, b, v, prox_step<IN>(io, r, c));
            io.Xout[i] = v;
            sum += v;
            if (last) {
              store_split(io, r, c, v);
              if (want_gram) gtile[r * GLD + threadIdx.x] = v;
              const float d = v - prev[j];
              nd += d * d; nn += v * v; np += prev[j] * prev[j];
            }
          }
        }
      }
    }
  }
  if (io.norms) block_accumulate3(nd, nn, np, io.norms);
  if (want_gram) {
    // block partial of X X^T over this block's 128 columns: thread -> 4 x 8 outputs (rows <= 64)
    __syncthreads();
    const int i0 = (threadIdx.x >> 3) * 4, j0 = (threadIdx.x & 7) * 8;
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[p][q] = 0.f;
    for (int cc = 0; cc < 128; ++cc) {
      float av[4], bv[8];
#pragma unroll
      for (int p = 0; p < 4; ++p) av[p] = (i0 + p < io.rows) ? gtile[(i0 + p) * GLD + cc] : 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) bv[q] = (j0 + q < io.rows) ? gtile[(j0 + q) * GLD + cc] : 0.f;
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
    }
    float* out = io.gram_part + (size_t)blockIdx.x * io.rows * io.rows;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (i0 + p < io.rows && j0 + q < io.rows) out[(i0 + p) * io.rows + j0 + q] = acc[p][q];
  }
}

// ---- UNITY(axis=0) chains on short columns (rows <= 64: S with K <= 64) ---------------------
// Tile = 32 columns x 64 rows, 256 threads = 32 columns x 8 row groups (a warp reads one 128-byte line per row).
// Every thread first issues all its loads (8 rows x up to 3 streams), the column sums are taken by one warp in
// NumPy's row order from a shared-memory copy of the tile, and the column never leaves the registers between the
// sum and the division: S is read once and written once.  Blocks are persistent over a strided tile sequence so
// that the fused Gram partial (X X^T over the block's columns, 4 x 4 outputs per thread, packed fp32x2 FMAs) is
// written once per block.  The kernel is issue-bound, so the prox chain is decoded once per op for the thread's
// eight elements instead of once per element.
constexpr int CT_COLS = 32, CT_RPT = 8, CT_LD = 68;   // tile stored column-major: tileT[col][row], ld 68

// ops [a, b) of the chain applied to N values that share the step (or have per-value steps when `ps` varies)
template <int N>
__device__ __forceinline__ void chain_segment_vec(const ProxChain& c, int a, int b, float (&v)[N], const float (&ps)[N]) {
  for (int i = a; i < b; ++i) {
    const int op = c.op[i];
    const float thr = c.thr[i];
    const bool rel = c.rel[i] != 0;
    if (op == PMX_OP_PLUS) {
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = (v[j] < 0.0f) ? 0.0f : v[j];
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) v[j] = prox_elem(v[j], op, rel ? __fmul_rn(thr, ps[j]) : thr);
    }
  }
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

template <int IN>
__global__ void __launch_bounds__(256, 3) k_upd_cols_tile(ProxChain ch, UpdIO io, int n_tiles) {
  if (skip(io)) return;
  __shared__ __align__(16) float tileT[CT_COLS * CT_LD];
  __shared__ __align__(16) float tileD[CT_COLS * 2 * CT_LD];   // every value twice (a, a): broadcast operand of fp32x2 FMAs
  __shared__ float colsum[CT_COLS];
  const int cx = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const bool want_gram = io.gram_part != nullptr;
  const bool prev_is_in = (io.Xprev == io.Xin);
  const int rows = io.rows, cols = io.cols;
  float nd = 0.f, nn = 0.f, np = 0.f;
  // Gram outputs of this thread: rows gi0..gi0+3 x columns gj0..gj0+3 of the rows x rows matrix
  const int gi0 = (threadIdx.x >> 4) * 4, gj0 = (threadIdx.x & 15) * 4;
  unsigned long long acc[4][2];   // acc[p][h] = (G[gi0+p][gj0+2h], G[gi0+p][gj0+2h+1])
#pragma unroll
  for (int a = 0; a < 4; ++a) acc[a][0] = acc[a][1] = 0ull;
  const int r0 = rg * CT_RPT;
  const float step_scalar = (io.step.mode <= 1) ? step_at(io.step, 0, 0) : 0.f;
  const float inv_psimax = 0.f;
  (void)inv_psimax;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int c = tile * CT_COLS + cx;
    const bool col_ok = c < cols;
    const size_t i0 = (size_t)r0 * cols + (col_ok ? c : 0);
    float xin[CT_RPT], g[CT_RPT], x0[CT_RPT], v[CT_RPT], ps[CT_RPT];
    {
      const float* pin = io.Xin + i0;
      const float* pg = io.G + i0;
      const float* p0 = io.X0 + i0;
#pragma unroll
      for (int j = 0; j < CT_RPT; ++j) {
        const bool ok = col_ok && (r0 + j < rows);
        xin[j] = ok ? pin[(size_t)j * cols] : 0.f;
        g[j] = (IN != IN_PLAIN && ok) ? pg[(size_t)j * cols] : 0.f;
        x0[j] = (IN == IN_ADASUB && ok) ? p0[(size_t)j * cols] : 0.f;
      }
    }
    int a = 0, b = chain_next_unity(ch, 0);
    const float psimax = (IN == IN_ADASUB) ? io.psimax[0] : 1.f;
#pragma unroll
    for (int j = 0; j < CT_RPT; ++j) {
      const int r = r0 + j;
      const float s = (io.step.mode <= 1) ? step_scalar : step_at(io.step, r < rows ? r : 0, col_ok ? c : 0);
      if (IN == IN_PGM) {
        ps[j] = s;
        v[j] = __fsub_rn(xin[j], __fmul_rn(s, g[j]));
      } else if (IN == IN_ADASUB) {
        const float gamma = s / psimax;
        ps[j] = gamma;
        v[j] = xin[j] - gamma / s * g[j] * (xin[j] - x0[j]);
      } else {
        ps[j] = s;
        v[j] = xin[j];
      }
    }
    chain_segment_vec<CT_RPT>(ch, a, b, v, ps);
#pragma unroll
    for (int j = 0; j < CT_RPT; ++j)
      if (!(col_ok && r0 + j < rows)) v[j] = 0.f;
    while (b < ch.n) {   // one round per UNITY op: column sum in row order, divide, next elementwise segment
      __syncthreads();   // previous readers of tileT are done
      *reinterpret_cast<float4*>(&tileT[cx * CT_LD + r0]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&tileT[cx * CT_LD + r0 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
      __syncthreads();
      if (threadIdx.x < CT_COLS) {
        float sum = 0.f;
        const float4* col = reinterpret_cast<const float4*>(&tileT[cx * CT_LD]);
        for (int r4 = 0; r4 < (rows + 3) / 4; ++r4) {   // rows beyond `rows` hold zeros; NumPy's axis-0 order
          const float4 t = col[r4];
          sum += t.x; sum += t.y; sum += t.z; sum += t.w;
        }
        colsum[cx] = sum;
      }
      __syncthreads();
      const float denom = colsum[cx];
      a = b + 1;
      b = chain_next_unity(ch, a);
#pragma unroll
      for (int j = 0; j < CT_RPT; ++j) v[j] = v[j] / denom;                 // operators.py:44
      chain_segment_vec<CT_RPT>(ch, a, b, v, ps);
#pragma unroll
      for (int j = 0; j < CT_RPT; ++j)
        if (!(col_ok && r0 + j < rows)) v[j] = 0.f;
    }
    // stores + norms
    {
      float* pout = io.Xout + i0;
      float* pold = io.Xold_out ? io.Xold_out + i0 : nullptr;
      const float* pprev = (!prev_is_in && io.Xprev) ? io.Xprev + i0 : nullptr;
      unsigned short* phi = io.hi ? io.hi + (size_t)r0 * io.ld_split + c : nullptr;
      unsigned short* plo = io.hi ? io.lo + (size_t)r0 * io.ld_split + c : nullptr;
#pragma unroll
      for (int j = 0; j < CT_RPT; ++j) {
        if (col_ok && r0 + j < rows) {
          const float prev = prev_is_in ? xin[j] : (pprev ? pprev[(size_t)j * cols] : 0.f);
          if (pold) pold[(size_t)j * cols] = prev;
          pout[(size_t)j * cols] = v[j];
          if (phi) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v[j]);
            const __nv_bfloat16 l = __float2bfloat16_rn(v[j] - __bfloat162float(h));
            phi[(size_t)j * io.ld_split] = __bfloat16_as_ushort(h);
            plo[(size_t)j * io.ld_split] = __bfloat16_as_ushort(l);
          }
          const float d = v[j] - prev;
          nd = fmaf(d, d, nd); nn = fmaf(v[j], v[j], nn); np = fmaf(prev, prev, np);
        }
      }
    }
    if (want_gram) {
      __syncthreads();
      *reinterpret_cast<float4*>(&tileT[cx * CT_LD + r0]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&tileT[cx * CT_LD + r0 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
      float4* dd = reinterpret_cast<float4*>(&tileD[cx * 2 * CT_LD + 2 * r0]);
      dd[0] = make_float4(v[0], v[0], v[1], v[1]);
      dd[1] = make_float4(v[2], v[2], v[3], v[3]);
      dd[2] = make_float4(v[4], v[4], v[5], v[5]);
      dd[3] = make_float4(v[6], v[6], v[7], v[7]);
      __syncthreads();
#pragma unroll 4
      for (int cc = 0; cc < CT_COLS; ++cc) {
        const ulonglong2 a01 = *reinterpret_cast<const ulonglong2*>(&tileD[cc * 2 * CT_LD + 2 * gi0]);       // (a0,a0),(a1,a1)
        const ulonglong2 a23 = *reinterpret_cast<const ulonglong2*>(&tileD[cc * 2 * CT_LD + 2 * gi0 + 4]);   // (a2,a2),(a3,a3)
        const ulonglong2 bv = *reinterpret_cast<const ulonglong2*>(&tileT[cc * CT_LD + gj0]);                // (b0,b1),(b2,b3)
        ffma2(acc[0][0], a01.x, bv.x); ffma2(acc[0][1], a01.x, bv.y);
        ffma2(acc[1][0], a01.y, bv.x); ffma2(acc[1][1], a01.y, bv.y);
        ffma2(acc[2][0], a23.x, bv.x); ffma2(acc[2][1], a23.x, bv.y);
        ffma2(acc[3][0], a23.y, bv.x); ffma2(acc[3][1], a23.y, bv.y);
      }
    }
  }
  if (io.norms) block_accumulate3(nd, nn, np, io.norms);
  if (want_gram) {
    float* out = io.gram_part + (size_t)blockIdx.x * rows * rows;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 f = *reinterpret_cast<const float2*>(&acc[p][h]);
        if (gi0 + p < rows && gj0 + 2 * h < rows) out[(gi0 + p) * rows + gj0 + 2 * h] = f.x;
        if (gi0 + p < rows && gj0 + 2 * h + 1 < rows) out[(gi0 + p) * rows + gj0 + 2 * h + 1] = f.y;
      }
  }
}

// ---- chains with UNITY(axis=1): one warp owns a row ------------------------------------
template <int IN>
__global__ void __launch_bounds__(kThreads) k_upd_rows(ProxChain ch, UpdIO io) {
  if (skip(io)) return;
  float nd = 0.f, nn = 0.f, np = 0.f;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const bool alias = (io.Xprev == io.Xout);
  const float* prev_src = alias ? io.Xold_out : io.Xprev;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < io.rows; r += gridDim.x * warps_per_block) {
    int a = 0, b = chain_next_unity(ch, 0);
    float sum = 0.f;
    for (int c = lane; c < io.cols; c += 32) {
      const size_t i = (size_t)r * io.cols + c;
      const float prev = io.Xprev ? io.Xprev[i] : 0.f;
      float ps;
      float v = transform<IN>(io, i, r, c, ps);
      v = chain_segment(ch, a, b, v, ps);
      if (io.Xold_out) io.Xold_out[i] = prev;
      io.Xout[i] = v;
      sum += v;
      if (b >= ch.n) {
        store_split(io, r, c, v);
        const float d = v - prev;
        nd += d * d; nn += v * v; np += prev * prev;
      }
    }
    while (b < ch.n) {
      a = b + 1;
      b = chain_next_unity(ch, a);
      const float denom = warp_sum(sum);
      sum = 0.f;
      for (int c = lane; c < io.cols; c += 32) {
        const size_t i = (size_t)r * io.cols + c;
        float v = io.Xout[i] / denom;
        v = chain_segment(ch, a, b, v, prox_step<IN>(io, r, c));
        io.Xout[i] = v;
        sum += v;
        if (b >= ch.n) {
          store_split(io, r, c, v);
          const float prev = prev_src ? prev_src[i] : 0.f;
          const float d = v - prev;
          nd += d * d; nn += v * v; np += prev * prev;
        }
      }
    }
  }
  if (io.norms) block_accumulate3(nd, nn, np, io.norms);
}

}  // namespace

// blocks the column-owner update kernel runs with (= number of fused Gram partials it writes)
int upd_cols_blocks(pmx_ctx* ctx, int cols) {
  const int n_tiles = pmx_div_up(cols, CT_COLS);
  const int cap = ctx->sm_count * 3;
  return n_tiles < cap ? n_tiles : cap;
}

namespace {

template <int IN>
int launch_update_t(pmx_ctx* ctx, const ProxChain& chain, const UpdIO& io) {
  const int ax = chain_unity_axis(chain);
  const size_t n = (size_t)io.rows * io.cols;
  if (n == 0) return PMX_OK;
  if (ax == 2) {
    pmx_set_error("prox chain mixes UNITY along both axes; apply it as two chains");
    return PMX_ERR_UNSUPPORTED;
  }
  if (ax >= 0 && io.norms && io.Xprev == io.Xout && !io.Xold_out) {
    pmx_set_error("in-place UNITY update with norms needs an Xold buffer");
    return PMX_ERR_ARG;
  }
  if (io.gram_part && !(ax == 0 && io.rows <= 64)) {
    pmx_set_error("fused Gram output needs the column-owner kernel (UNITY axis 0) and rows <= 64");
    return PMX_ERR_ARG;
  }
  if (ax == -1) {
    long long blocks = (long long)((n + kThreads - 1) / kThreads);
    const long long cap = (long long)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_upd_flat<IN><<<(int)blocks, kThreads, 0, ctx->stream>>>(chain, io);
  } else if (ax == 0) {
    int n_unity = 0;
    for (int i = 0; i < chain.n; ++i) n_unity += chain.op[i] == PMX_OP_UNITY;
    // the register-resident variant measured 4x slower than the two-pass kernel on B200 (204 vs 53 us for
    // S = 64 x 65536): kept for reference, disabled
    const bool reg_path = false && io.rows <= 64 && n_unity == 1 && chain.op[chain.n - 1] == PMX_OP_UNITY;
    if (io.rows <= 64) {
      const int n_tiles = pmx_div_up(io.cols, CT_COLS);
      k_upd_cols_tile<IN><<<upd_cols_blocks(ctx, io.cols), 256, 0, ctx->stream>>>(chain, io, n_tiles);
    } else if (reg_path)
      k_upd_cols_reg<IN><<<pmx_div_up(io.cols, 128), 128, 0, ctx->stream>>>(chain, io);
    else
      k_upd_cols<IN><<<pmx_div_up(io.cols, 128), 128, io.gram_part ? sizeof(float) * io.rows * 129 : 0,
                       ctx->stream>>>(chain, io);
  } else {
    const int wpb = kThreads / 32;
    long long blocks = pmx_div_up(io.rows, wpb);
    const long long cap = (long long)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_upd_rows<IN><<<(int)blocks, kThreads, 0, ctx->stream>>>(chain, io);
  }
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "update kernel");
}

}  // namespace

int launch_update(pmx_ctx* ctx, int in_kind, const ProxChain& chain, const UpdIO& io) {
  switch (in_kind) {
    case IN_PLAIN: return launch_update_t<IN_PLAIN>(ctx, chain, io);
    case IN_PGM: return launch_update_t<IN_PGM>(ctx, chain, io);
    case IN_ADASUB: return launch_update_t<IN_ADASUB>(ctx, chain, io);
  }
  pmx_set_error("bad transform kind %d", in_kind);
  return PMX_ERR_ARG;
}

// =========================================================================================
// zero fill that honours the device-side stop flag (a plain memset would wipe the "last
// gradient" the solvers hand back after the loop froze, algorithms.py:144)
// =========================================================================================
__global__ void __launch_bounds__(kThreads) k_zero(float4* __restrict__ p4, size_t n4, float* __restrict__ tail,
                                                   int ntail, const int* done) {
  if (done && *done) return;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) p4[i] = z;
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) tail[threadIdx.x] = 0.f;
}

// up to three buffers in one launch (the gradient outputs and the loss of one gradient evaluation)
struct ZeroArgs {
  float* p[3];
  size_t n[3];
};
__global__ void __launch_bounds__(kThreads) k_zero3(ZeroArgs a, const int* done) {
  if (done && *done) return;
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    float* p = a.p[b];
    const size_t n = a.n[b];
    if (!p || n == 0) continue;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
      const size_t n4 = n / 4;
      float4* p4 = reinterpret_cast<float4*>(p);
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (blockIdx.x == 0 && threadIdx.x < n - n4 * 4) p[n4 * 4 + threadIdx.x] = 0.f;
    } else {
      for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.f;
    }
  }
}

int launch_zero3(pmx_ctx* ctx, cudaStream_t st, float* p0, size_t n0, float* p1, size_t n1, float* p2, size_t n2,
                 const int* done) {
  ZeroArgs a;
  a.p[0] = p0; a.n[0] = n0; a.p[1] = p1; a.n[1] = n1; a.p[2] = p2; a.n[2] = n2;
  size_t nmax = n0 > n1 ? n0 : n1;
  if (n2 > nmax) nmax = n2;
  if (nmax == 0) return PMX_OK;
  long long blocks = (long long)((nmax / 4 + kThreads - 1) / kThreads);
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_zero3<<<(int)blocks, kThreads, 0, st>>>(a, done);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_zero3");
}

int launch_zero(pmx_ctx* ctx, cudaStream_t st, float* p, size_t n, const int* done) {
  if (n == 0) return PMX_OK;
  const size_t n4 = n / 4;
  long long blocks = (long long)((n4 + kThreads - 1) / kThreads);
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_zero<<<(int)blocks, kThreads, 0, st>>>(reinterpret_cast<float4*>(p), n4, p + n4 * 4, (int)(n - n4 * 4), done);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_zero");
}

// =========================================================================================
// Nesterov extrapolation  Xe = X + omega (X - Xold)   (algorithms.py:94-95)
// =========================================================================================
__global__ void __launch_bounds__(kThreads) k_extrapolate(const float* __restrict__ X, const float* __restrict__ Xold,
                                                          float* __restrict__ Xe, size_t n, float omega,
                                                          const int* done) {
  if (done && *done) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float x = X[i];
    Xe[i] = x + omega * (x - Xold[i]);
  }
}

int launch_extrapolate(pmx_ctx* ctx, const float* X, const float* Xold, float* Xe, size_t n, float omega,
                       const int* done) {
  long long blocks = (long long)((n + kThreads - 1) / kThreads);
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return PMX_OK;
  k_extrapolate<<<(int)blocks, kThreads, 0, ctx->stream>>>(X, Xold, Xe, n, omega, done);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_extrapolate");
}

// =========================================================================================
// fp32 -> (hi, lo) bf16 split, x ~= hi + lo with |x - hi - lo| <= 2^-17 |x|.
// The tcgen05 GEMMs compute hi*hi + hi*lo + lo*hi in fp32 (plain TF32/BF16 inputs
// miss the 1e-4 parity target; the 3-term split clears it with margin).
// Source is rows x cols (ld = cols); destination is rows_pad x ld_dst, zero padded.
// =========================================================================================
__global__ void __launch_bounds__(kThreads) k_split_bf16(const float* __restrict__ X, int rows, int cols,
                                                         __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, int rows_pad, int ld_dst,
                                                         const int* done) {
  if (done && *done) return;
  const size_t n = (size_t)rows_pad * ld_dst;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld_dst);
    const int c = (int)(i - (size_t)r * ld_dst);
    float x = 0.f;
    if (r < rows && c < cols) x = X[(size_t)r * cols + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    hi[i] = h;
    lo[i] = l;
  }
}

int launch_split_bf16(pmx_ctx* ctx, const float* X, int rows, int cols, void* hi, void* lo, int rows_pad, int ld_dst,
                      const int* done) {
  const size_t n = (size_t)rows_pad * ld_dst;
  long long blocks = (long long)((n + kThreads - 1) / kThreads);
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return PMX_OK;
  k_split_bf16<<<(int)blocks, kThreads, 0, ctx->stream>>>(X, rows, cols, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                          rows_pad, ld_dst, done);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_split_bf16");
}

// =========================================================================================
// adaprox moment update (algorithms.py:147-245) fused with the step X -= Alpha*Phi/Psi (:378),
// the copy z = X (:383) and the block-wide max(Psi) (:384).
// =========================================================================================
__global__ void __launch_bounds__(kThreads) k_adaprox_moments(AdaArgs a) {
  if (a.done && *a.done) return;
  float pm = 0.f;
  // scalar recipe pieces, evaluated in fp32 like NumPy does for fp32 arrays with Python scalars
  const double c1 = 1.0 - pow(a.b1, (double)a.t);                   // 1 - b1[it]**t  (np.float64 scalar)
  const float c2 = (float)(1.0 - pow(a.b2, (double)a.t));   // 1 - b2**t      (Python float -> weak fp32)
  const float omb2 = (float)(1.0 - a.b2);
  const float b2f = (float)a.b2, epsf = (float)a.eps, pf = (float)a.p;
  float radam_r = 0.f;
  bool radam_rect = false;
  if (a.scheme == PMX_RADAM) {  // algorithms.py:222-239 (scalar part, double like Python floats)
    const double b2 = a.b2, tt = a.t;
    const double rho_inf = 2.0 / (1.0 - b2) - 1.0;
    const double rho = rho_inf - 2.0 * tt * pow(b2, tt) / (1.0 - pow(b2, tt));
    radam_rect = rho > 4.0;
    if (radam_rect) radam_r = (float)sqrt((rho - 4.0) * (rho - 2.0) * rho_inf / (rho_inf - 4.0) / (rho_inf - 2.0) / rho);
  }
  const bool need_rc = a.alpha.mode >= 2;
  const bool small = a.n < 0xffffffffull;   // 32-bit index arithmetic (a 64-bit division per element otherwise)
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x * blockDim.x) {
    int r = 0, c = 0;
    if (need_rc) {
      if (small) {
        const unsigned ii = (unsigned)i, rr = ii / (unsigned)a.cols;
        r = (int)rr;
        c = (int)(ii - rr * (unsigned)a.cols);
      } else {
        r = (int)(i / a.cols);
        c = (int)(i - (size_t)r * a.cols);
      }
    }
    const float g = a.G[i];
    const float m = (float)((1.0 - a.b1) * (double)g + a.b1 * (double)a.M[i]);
    const float v = omb2 * (g * g) + b2f * a.V[i];
    a.M[i] = m;
    a.V[i] = v;
    double phi;  // Phi is float64 in the reference whenever b1[it] enters it
    float psi;
    switch (a.scheme) {
      case PMX_ADAM:
        phi = (double)m / c1;
        psi = sqrtf(v / c2) + epsf;
        break;
      case PMX_NADAM:
        phi = (a.b1 * (double)m + (1.0 - a.b1) * (double)g) / c1;
        psi = sqrtf(v / c2) + epsf;
        break;
      case PMX_RADAM:
        phi = (double)m / c1;
        psi = radam_rect ? sqrtf(v / c2) / radam_r : 1.0f;
        if (a.eps > 0.0) psi = fmaxf(psi, (float)sqrt(a.eps));
        break;
      default: {  // AMSGRAD / PADAM / ADAMX
        float vh = v;
        if (a.Vhat) {
          float old = a.Vhat[i];
          if (a.scheme == PMX_ADAMX) {
            const double f = ((1.0 - a.b1) * (1.0 - a.b1)) / ((1.0 - a.b1_prev) * (1.0 - a.b1_prev));
            old = (float)(f * (double)old);
          }
          vh = fmaxf(old, v);
          a.Vhat[i] = vh;
        }
        if (a.eps > 0.0) vh = fmaxf(vh, epsf);
        phi = m;
        psi = (a.scheme == PMX_PADAM) ? powf(vh, pf) : sqrtf(vh);
      }
    }
    const float al = step_at(a.alpha, r, c);
    const float xo = a.X[i];
    float xn;  // algorithms.py:378
    if (a.scheme == PMX_ADAM || a.scheme == PMX_NADAM || a.scheme == PMX_RADAM)
      xn = (float)((double)xo - (double)al * phi / (double)psi);
    else
      xn = xo - al * (float)phi / psi;
    if (a.Xold) a.Xold[i] = xo;
    a.X[i] = xn;
    a.Z[i] = xn;
    a.Psi[i] = psi;
    pm = fmaxf(pm, psi);
    if (psi != psi) pm = psi;  // np.max propagates NaN
  }
  // block max -> global
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float other = __shfl_xor_sync(0xffffffffu, pm, o);
    pm = (pm != pm || other != other) ? (pm + other) : fmaxf(pm, other);
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = pm;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m2 = red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) {
      const float o2 = red[w];
      m2 = (m2 != m2 || o2 != o2) ? (m2 + o2) : fmaxf(m2, o2);
    }
    // Psi >= 0 (or NaN, whose bit pattern 0x7fc00000 compares above every finite float)
    atomicMax((int*)a.psimax, __float_as_int(m2));
  }
}

int launch_adaprox_moments(pmx_ctx* ctx, const AdaArgs& a) {
  long long blocks = (long long)((a.n + kThreads - 1) / kThreads);
  const long long cap = (long long)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) return PMX_OK;
  k_adaprox_moments<<<(int)blocks, kThreads, 0, ctx->stream>>>(a);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_adaprox_moments");
}
