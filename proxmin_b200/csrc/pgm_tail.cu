// Fused tail of a PGM iteration on the NMF objective: everything that follows the gradient kernel, in ONE launch.
//
//   algorithms.py:107-108  X_j <- prox_j(X_j - step_j G_j)          both blocks (Jacobi: same gradient point)
//   algorithms.py:130-133  |X_j - X_j_old|^2 <= e_rel^2 |X_j|^2       norms fused into the updates, test on the device
//   nmf.py:44-49, utils.py:20,34   step_j = 1 / lambda_max(Gram)      Gram matrices of the NEW factors as a by-product
//                                                                     of the updates, lambda_max in the last block
// plus the bf16 (hi, lo) split of the new factors for the next gradient kernel and the zero-fill of the gradient
// buffers of the next iteration (the gradient buffers are pairs indexed by the iteration parity).
//
// Round 1 ran this as 9 kernels on two streams (Gram reduce, Gram of A, lambda_max, zero-fill, A update, S update,
// Gram reduce, finalize, ...): ~0.09 ms per iteration that does not shrink when the columns are split over GPUs.
//
// Block roles (one grid, 256 threads per block):
//   S blocks  [0, nS)        : persistent over 32-column tiles of S: forward step + prox chain (UNITY along axis 0 in
//                              NumPy's row order) + store S, S_hi, S_lo + norms + 64x64 Gram partial (packed fp32x2
//                              FMAs) + zero-fill of the other-parity G_S tile.
//   A blocks  [nS, nS + nA)  : 64 rows of A each.  Multi-GPU: they first wait for "gradient done" of every rank, then
//                              reduce-scatter -- sum the G_A partials of all ranks for THEIR rows straight out of peer
//                              memory (rank order: bit-identical on every rank) -- update the rows and all-gather by
//                              pushing A, A_hi, A_lo into every rank's symmetric arena (two-shot exchange: each rank
//                              reads and writes 1/world of what the one-shot sum of round 1 moved).
//   last block of each role  : sums the block Gram partials (8 replicated fp32 accumulators fed by red.add), multi-GPU:
//                              exchanges the K x K partial + 3 norms through per-rank inbox slots and sums them in rank
//                              order, then lambda_max (lambda_max.cuh) -> the step of the OTHER block for the next
//                              iteration.
//   second of those two      : convergence test, iteration counter, stop flag (was k_pgm_finalize).
// No host involvement, no second stream: an iteration is  gradient kernel -> [peer signal] -> this kernel.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "kernels.h"
#include "lambda_max.cuh"
#include "pgm_tail.h"

namespace {

constexpr int TT = 256;
constexpr int CT_COLS = 32, CT_RPT = 8, CT_LD = 68;   // S tile: 32 columns x 64 rows, stored column-major (ld 68)
constexpr int RA = PMX_TAIL_RA;                       // rows of A per A block (upper bound; a.ra is the actual count)
constexpr int NREP = PMX_TAIL_NREP;                   // replicated accumulators of the Gram partials
constexpr long long SPIN_LIMIT = 20000000000LL;        // ~10 s of SM clocks: a lost peer must not hang the GPU

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_sys_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void red_add_f32(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

// Gram partials are symmetric: only the 4 x 4 register tiles on or above the diagonal are computed (136 of the 256
// tiles of a 64 x 64 Gram), by the first 136 threads -- whole warps drop out of the FMA loop -- and k_tail_final
// mirrors them.  Thread t -> tile (ti, tj), ti <= tj, row-major over the upper triangle of an nt x nt tile grid.
__device__ __forceinline__ bool gram_tile_of(int t, int nt, int& ti, int& tj) {
  int row = 0, left = t;
  while (row < nt && left >= nt - row) {
    left -= nt - row;
    ++row;
  }
  ti = row;
  tj = row + left;
  return row < nt;
}

// debug timeline (env PMX_TAIL_TRACE): slot 0 = earliest block start (atomicMin), other slots = latest time a phase
// ended over all blocks (atomicMax), nanoseconds of %globaltimer
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TSTAMP(slot)                                                   \
  do {                                                                 \
    if (a.trace && threadIdx.x == 0) {                                 \
      if ((slot) == 0) atomicMin(a.trace, gtime());                    \
      else atomicMax(a.trace + (slot), gtime());                       \
    }                                                                  \
  } while (0)

// ops [a, b) of the chain applied to N values
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

// three partial sums of the block -> one fp64 atomicAdd each
__device__ __forceinline__ void block_accumulate3(float a, float b, float c, double* out, float (*red)[8]) {
  a = warp_sum(a);
  b = warp_sum(b);
  c = warp_sum(c);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) {
    red[0][w] = a;
    red[1][w] = b;
    red[2][w] = c;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int i = 0; i < TT / 32; ++i) t += red[threadIdx.x][i];
    atomicAdd(out + threadIdx.x, (double)t);
  }
}

// wait until flags[set][r] >= e for every rank r (threads 0..world-1 spin on LOCAL memory); false on timeout
__device__ __forceinline__ bool wait_flags(const unsigned* my_flags, int set, int world, unsigned e, int* s_fault) {
  if ((int)threadIdx.x < world) {
    const unsigned* f = my_flags + set * PMX_MAX_WORLD + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(f) - e) < 0) {
      if (clock64() - t0 > SPIN_LIMIT) {
        *s_fault = 1;
        break;
      }
    }
  }
  __syncthreads();
  return *s_fault == 0;
}

// ------------------------------------------------------------------------------------------ S blocks
__device__ __forceinline__ void s_role(const PgmTailArgs& a, unsigned par, float step, unsigned char* smem) {
  float* tileT = reinterpret_cast<float*>(smem);                       // [32][68]
  float* tileD = tileT + CT_COLS * CT_LD;                              // [32][2 * 68]: every value twice (a, a)
  float* colsum = tileD + CT_COLS * 2 * CT_LD;                         // [32]
  float(*red)[8] = reinterpret_cast<float(*)[8]>(colsum + CT_COLS);    // [3][8]
  const ProxChain& ch = a.chS;
  const int cx = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int rows = a.K, cols = a.N;
  const float* __restrict__ G = a.GS2 + (size_t)par * a.gs_stride;
  float* __restrict__ Gz = a.GS2 + (size_t)(par ^ 1u) * a.gs_stride;   // zero-filled for the next iteration
  float nd = 0.f, nn = 0.f, np = 0.f;
  int gti, gtj;
  const bool gram_on = gram_tile_of(threadIdx.x, (rows + 3) >> 2, gti, gtj);
  const int gi0 = gti * 4, gj0 = gtj * 4;
  unsigned long long acc[4][2];
#pragma unroll
  for (int p = 0; p < 4; ++p) acc[p][0] = acc[p][1] = 0ull;
  const int r0 = rg * CT_RPT;
  const int sblk = blockIdx.x - a.nA;
  // the loads of the next tile are issued before the current one is processed: a tile is a chain of dependent
  // phases (column sums in row order, division, Gram) and would otherwise expose the full DRAM latency every time
  float xin[CT_RPT], g[CT_RPT];
  auto load_tile = [&](int tile, float (&xo)[CT_RPT], float (&go)[CT_RPT]) {
    const int c = tile * CT_COLS + cx;
    const bool col_ok = c < cols;
    const size_t i0 = (size_t)r0 * cols + (col_ok ? c : 0);
#pragma unroll
    for (int j = 0; j < CT_RPT; ++j) {
      const bool ok = col_ok && (r0 + j < rows);
      xo[j] = ok ? __ldcs(a.S + i0 + (size_t)j * cols) : 0.f;
      go[j] = ok ? __ldcs(G + i0 + (size_t)j * cols) : 0.f;
    }
  };
  if (sblk < a.n_tiles_S) load_tile(sblk, xin, g);
  for (int tile = sblk; tile < a.n_tiles_S; tile += a.nS) {
    const int c = tile * CT_COLS + cx;
    const bool col_ok = c < cols;
    const size_t i0 = (size_t)r0 * cols + (col_ok ? c : 0);
    float xn[CT_RPT], gn[CT_RPT], v[CT_RPT], ps[CT_RPT];
    if (tile + a.nS < a.n_tiles_S) load_tile(tile + a.nS, xn, gn);
    int sa = 0, sb = chain_next_unity(ch, 0);
#pragma unroll
    for (int j = 0; j < CT_RPT; ++j) {
      ps[j] = step;
      v[j] = __fsub_rn(xin[j], __fmul_rn(step, g[j]));     // algorithms.py:108, two roundings like NumPy
    }
    chain_segment_vec<CT_RPT>(ch, sa, sb, v, ps);
#pragma unroll
    for (int j = 0; j < CT_RPT; ++j)
      if (!(col_ok && r0 + j < rows)) v[j] = 0.f;
    if (a.trace && v[0] != 12345.f) TSTAMP(13);
    while (sb < ch.n) {   // one round per UNITY op: column sum in row order (NumPy's axis-0 order), divide, next segment
      __syncthreads();
      *reinterpret_cast<float4*>(&tileT[cx * CT_LD + r0]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&tileT[cx * CT_LD + r0 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
      __syncthreads();
      if (threadIdx.x < CT_COLS) {
        float sum = 0.f;
        const float4* col = reinterpret_cast<const float4*>(&tileT[cx * CT_LD]);
        for (int r4 = 0; r4 < (rows + 3) / 4; ++r4) {
          const float4 t = col[r4];
          sum += t.x; sum += t.y; sum += t.z; sum += t.w;
        }
        colsum[cx] = sum;
      }
      __syncthreads();
      const float denom = colsum[cx];
      sa = sb + 1;
      sb = chain_next_unity(ch, sa);
#pragma unroll
      for (int j = 0; j < CT_RPT; ++j) v[j] = v[j] / denom;                 // operators.py:44
      chain_segment_vec<CT_RPT>(ch, sa, sb, v, ps);
#pragma unroll
      for (int j = 0; j < CT_RPT; ++j)
        if (!(col_ok && r0 + j < rows)) v[j] = 0.f;
    }
    TSTAMP(14);
    {  // stores + norms + zero-fill of the other-parity gradient tile
      float* pout = a.S + i0;
      float* pz = Gz + i0;
      unsigned short* phi = a.Shi + (size_t)r0 * a.ldS + c;
      unsigned short* plo = a.Slo + (size_t)r0 * a.ldS + c;
#pragma unroll
      for (int j = 0; j < CT_RPT; ++j) {
        if (col_ok && r0 + j < rows) {
          pout[(size_t)j * cols] = v[j];
          pz[(size_t)j * cols] = 0.f;
          const __nv_bfloat16 h = __float2bfloat16_rn(v[j]);
          const __nv_bfloat16 l = __float2bfloat16_rn(v[j] - __bfloat162float(h));
          phi[(size_t)j * a.ldS] = __bfloat16_as_ushort(h);
          plo[(size_t)j * a.ldS] = __bfloat16_as_ushort(l);
          const float d = v[j] - xin[j];
          nd = fmaf(d, d, nd); nn = fmaf(v[j], v[j], nn); np = fmaf(xin[j], xin[j], np);
        }
      }
    }
    TSTAMP(15);
    {  // Gram partial of this tile: S S^T over its 32 columns, 4 x 4 outputs per thread
      __syncthreads();
      *reinterpret_cast<float4*>(&tileT[cx * CT_LD + r0]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&tileT[cx * CT_LD + r0 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
      float4* dd = reinterpret_cast<float4*>(&tileD[cx * 2 * CT_LD + 2 * r0]);
      dd[0] = make_float4(v[0], v[0], v[1], v[1]);
      dd[1] = make_float4(v[2], v[2], v[3], v[3]);
      dd[2] = make_float4(v[4], v[4], v[5], v[5]);
      dd[3] = make_float4(v[6], v[6], v[7], v[7]);
      __syncthreads();
      if (gram_on)
#pragma unroll 4
      for (int cc = 0; cc < CT_COLS; ++cc) {
        const ulonglong2 a01 = *reinterpret_cast<const ulonglong2*>(&tileD[cc * 2 * CT_LD + 2 * gi0]);
        const ulonglong2 a23 = *reinterpret_cast<const ulonglong2*>(&tileD[cc * 2 * CT_LD + 2 * gi0 + 4]);
        const ulonglong2 bv = *reinterpret_cast<const ulonglong2*>(&tileT[cc * CT_LD + gj0]);
        ffma2(acc[0][0], a01.x, bv.x); ffma2(acc[0][1], a01.x, bv.y);
        ffma2(acc[1][0], a01.y, bv.x); ffma2(acc[1][1], a01.y, bv.y);
        ffma2(acc[2][0], a23.x, bv.x); ffma2(acc[2][1], a23.x, bv.y);
        ffma2(acc[3][0], a23.y, bv.x); ffma2(acc[3][1], a23.y, bv.y);
      }
    }
#pragma unroll
    for (int j = 0; j < CT_RPT; ++j) {
      xin[j] = xn[j];
      g[j] = gn[j];
    }
  }
  TSTAMP(1);
  block_accumulate3(nd, nn, np, a.acc + 4, red);
  float* out = a.gram_rep + ((size_t)NREP + (sblk % NREP)) * rows * rows;   // [1][rep][K*K]
  if (gram_on)
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float2 f = *reinterpret_cast<const float2*>(&acc[p][h]);
      if (gi0 + p < rows && gj0 + 2 * h < rows) red_add_f32(&out[(gi0 + p) * rows + gj0 + 2 * h], f.x);
      if (gi0 + p < rows && gj0 + 2 * h + 1 < rows) red_add_f32(&out[(gi0 + p) * rows + gj0 + 2 * h + 1], f.y);
    }
}

// ------------------------------------------------------------------------------------------ A blocks
__device__ __forceinline__ void a_role(const PgmTailArgs& a, unsigned par, float step, unsigned char* smem, int* s_fault) {
  const int K = a.K, world = a.world;
  const int ablk = blockIdx.x;
  const int m0 = a.m_lo + ablk * a.ra;
  const int nrows = min(a.ra, a.m_hi - m0);
  const int ldt = ((K + 3) & ~3) + 4;                                       // 16-byte aligned rows
  float* at = reinterpret_cast<float*>(smem);                               // [RA][ldt] new values (Gram operand)
  float(*red)[8] = reinterpret_cast<float(*)[8]>(at + RA * ldt);
  if ((K & 3) != 0)   // the columns K .. K4-1 the vector loads of the Gram touch
    for (int i = threadIdx.x; i < RA * 4; i += TT) at[(i >> 2) * ldt + (K & ~3) + (i & 3)] = 0.f;
  const ProxChain& ch = a.chA;
  float nd = 0.f, nn = 0.f, np = 0.f;
  if (world > 1) {
    // "gradient done" of every rank (set 0, signalled by k_peer_signal after the gradient kernel)
    const unsigned e0 = a.epoch[0];
    if (!wait_flags(a.my_flags, 0, world, e0, s_fault)) return;
  }
  const size_t ga_off = (size_t)par * a.ga_stride;
  const int nelem = nrows * K;
  // sources of the G_A partials and the current A; destinations of the new rows (every rank's arena, or the local
  // buffers of a single-GPU run)
  const float* A_src = world > 1 ? reinterpret_cast<const float*>(static_cast<char*>(a.arena.p[a.rank]) + a.off_A) : a.A_loc;
  if ((K & 3) == 0) {
    const int ngroups = nelem >> 2;
    for (int grp = threadIdx.x; grp < ngroups; grp += TT) {
      // all remote loads of a 16-byte group first (one NVLink round trip), then the sum in rank order
      const int idx = grp * 4;
      const size_t e = (size_t)m0 * K + idx;
      float4 part[PMX_MAX_WORLD];
      if (world > 1) {
#pragma unroll
        for (int r = 0; r < PMX_MAX_WORLD; ++r)
          if (r < world)
            part[r] = ld_sys_f4(reinterpret_cast<const float4*>(
                reinterpret_cast<const float*>(static_cast<const char*>(a.arena.p[r]) + a.off_GA) + ga_off + e));
      } else {
        part[0] = *reinterpret_cast<const float4*>(a.GA2_loc + ga_off + e);
      }
      const float4 x = *reinterpret_cast<const float4*>(A_src + e);
      float4 gs = part[0];
#pragma unroll
      for (int r = 1; r < PMX_MAX_WORLD; ++r)
        if (r < world) {
          gs.x += part[r].x; gs.y += part[r].y; gs.z += part[r].z; gs.w += part[r].w;
        }
      float v[4] = {__fsub_rn(x.x, __fmul_rn(step, gs.x)), __fsub_rn(x.y, __fmul_rn(step, gs.y)),
                    __fsub_rn(x.z, __fmul_rn(step, gs.z)), __fsub_rn(x.w, __fmul_rn(step, gs.w))};
      const float ps[4] = {step, step, step, step};
      chain_segment_vec<4>(ch, 0, ch.n, v, ps);
      const float xo[4] = {x.x, x.y, x.z, x.w};
      const int rl = idx / K, cl = idx - rl * K;     // row inside the block, first column of the group
      unsigned short hb[4], lb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float d = v[q] - xo[q];
        nd = fmaf(d, d, nd); nn = fmaf(v[q], v[q], nn); np = fmaf(xo[q], xo[q], np);
        at[rl * ldt + cl + q] = v[q];
        const __nv_bfloat16 h = __float2bfloat16_rn(v[q]);
        const __nv_bfloat16 l = __float2bfloat16_rn(v[q] - __bfloat162float(h));
        hb[q] = __bfloat16_as_ushort(h);
        lb[q] = __bfloat16_as_ushort(l);
      }
      const size_t es = (size_t)(m0 + rl) * a.ldA + cl;     // element index in the bf16 operand buffers
      const float4 vo = make_float4(v[0], v[1], v[2], v[3]);
      const uint2 ho = make_uint2((unsigned)hb[0] | ((unsigned)hb[1] << 16), (unsigned)hb[2] | ((unsigned)hb[3] << 16));
      const uint2 lo = make_uint2((unsigned)lb[0] | ((unsigned)lb[1] << 16), (unsigned)lb[2] | ((unsigned)lb[3] << 16));
      if (world > 1) {
#pragma unroll
        for (int r = 0; r < PMX_MAX_WORLD; ++r)
          if (r < world) {
            char* base = static_cast<char*>(a.arena.p[r]);
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(base + a.off_A) + e) = vo;
            *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(base + a.off_Ahi) + es) = ho;
            *reinterpret_cast<uint2*>(reinterpret_cast<unsigned short*>(base + a.off_Alo) + es) = lo;
          }
      } else {
        *reinterpret_cast<float4*>(a.A_loc + e) = vo;
        *reinterpret_cast<uint2*>(a.Ahi_loc + es) = ho;
        *reinterpret_cast<uint2*>(a.Alo_loc + es) = lo;
      }
    }
  } else {
    for (int i = threadIdx.x; i < nelem; i += TT) {
      const size_t e = (size_t)m0 * K + i;
      float gs = 0.f;
      if (world > 1) {
        float part[PMX_MAX_WORLD];
#pragma unroll
        for (int r = 0; r < PMX_MAX_WORLD; ++r)
          if (r < world)
            part[r] = ld_sys_f(reinterpret_cast<const float*>(static_cast<const char*>(a.arena.p[r]) + a.off_GA) + ga_off + e);
        gs = part[0];
#pragma unroll
        for (int r = 1; r < PMX_MAX_WORLD; ++r)
          if (r < world) gs += part[r];
      } else {
        gs = a.GA2_loc[ga_off + e];
      }
      const float xo = A_src[e];
      float v = __fsub_rn(xo, __fmul_rn(step, gs));
      v = chain_segment(ch, 0, ch.n, v, step);
      const float d = v - xo;
      nd = fmaf(d, d, nd); nn = fmaf(v, v, nn); np = fmaf(xo, xo, np);
      const int rl = i / K, cl = i - rl * K;
      at[rl * ldt + cl] = v;
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      const size_t es = (size_t)(m0 + rl) * a.ldA + cl;
      if (world > 1) {
#pragma unroll
        for (int r = 0; r < PMX_MAX_WORLD; ++r)
          if (r < world) {
            char* base = static_cast<char*>(a.arena.p[r]);
            reinterpret_cast<float*>(base + a.off_A)[e] = v;
            reinterpret_cast<unsigned short*>(base + a.off_Ahi)[es] = __bfloat16_as_ushort(h);
            reinterpret_cast<unsigned short*>(base + a.off_Alo)[es] = __bfloat16_as_ushort(l);
          }
      } else {
        a.A_loc[e] = v;
        a.Ahi_loc[es] = __bfloat16_as_ushort(h);
        a.Alo_loc[es] = __bfloat16_as_ushort(l);
      }
    }
  }
  // zero-fill of the other-parity G_A buffer (the whole M x K local buffer is split evenly over the A blocks)
  {
    const size_t total = (size_t)a.M * K;
    const size_t per = (total + a.nA - 1) / a.nA;
    const size_t z0 = (size_t)ablk * per, z1 = min(total, z0 + per);
    float* Gz = (world > 1 ? reinterpret_cast<float*>(static_cast<char*>(a.arena.p[a.rank]) + a.off_GA) : a.GA2_loc) +
                (size_t)(par ^ 1u) * a.ga_stride;
    if ((z0 & 3) == 0 && (per & 3) == 0) {
      for (size_t i = z0 + 4 * (size_t)threadIdx.x; i + 3 < z1; i += 4 * TT)
        *reinterpret_cast<float4*>(Gz + i) = make_float4(0.f, 0.f, 0.f, 0.f);
      for (size_t i = z0 + ((z1 - z0) & ~(size_t)3) + threadIdx.x; i < z1; i += TT) Gz[i] = 0.f;
    } else {
      for (size_t i = z0 + threadIdx.x; i < z1; i += TT) Gz[i] = 0.f;
    }
  }
  __syncthreads();
  // Gram partial of these rows: A^T A, 4 x 4 outputs per thread, upper-triangular tiles only (see gram_tile_of)
  {
    int ti, tj;
    if (gram_tile_of(threadIdx.x, (K + 3) >> 2, ti, tj)) {
      const int gi0 = ti * 4, gj0 = tj * 4;
      float acc[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
      for (int r = 0; r < nrows; ++r) {
        const float4 a4 = *reinterpret_cast<const float4*>(&at[r * ldt + gi0]);
        const float4 b4 = *reinterpret_cast<const float4*>(&at[r * ldt + gj0]);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
      }
      float* out = a.gram_rep + (size_t)(ablk % NREP) * K * K;     // [0][rep][K*K]
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (gi0 + p < K && gj0 + q < K) red_add_f32(&out[(gi0 + p) * K + gj0 + q], acc[p][q]);
    }
  }
  block_accumulate3(nd, nn, np, a.acc, red);
}

// ------------------------------------------------------------------------------------------ roles kernel
__global__ void __launch_bounds__(TT, 3) k_pgm_tail(const PgmTailArgs a) {
  pmx_ctl* ctl = a.ctl;
  if (ctl->done) return;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_last, s_fault;
  const unsigned pc = ctl->par_ctr;             // iterations whose updates are complete = index of this iteration
  const unsigned par = (pc + 1u) & 1u;          // gradient buffers the gradient kernel of this iteration filled
  const int cur = (int)(pc & 1u);               // steps of this iteration
  if (threadIdx.x == 0) s_fault = 0;
  __syncthreads();
  // A blocks come first in the grid (they are short, and in a sharded run they start with a wait for the peers)
  const int which = (int)blockIdx.x < a.nA ? 0 : 1;
  TSTAMP(0);
  if (which == 1) {
    s_role(a, par, ctl->step2[cur][1], smem);
    TSTAMP(2);
  } else {
    a_role(a, par, ctl->step2[cur][0], smem, &s_fault);
    TSTAMP(3);
    if (a.world > 1) {
      // the last A block tells every rank that this rank's rows of A (fp32 + bf16 operands) are in place (set 2)
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&a.tickets[0], 1u);
        s_last = t == (unsigned)(a.nA - 1);
      }
      __syncthreads();
      if (s_last) {
        __shared__ unsigned s_e;
        __threadfence_system();
        if (threadIdx.x == 0) {
          s_e = a.epoch[2] + 1;
          a.epoch[2] = s_e;
        }
        __syncthreads();
        if ((int)threadIdx.x < a.world)
          st_release_sys(reinterpret_cast<unsigned*>(a.flags.p[threadIdx.x]) + 2 * PMX_MAX_WORLD + a.rank, s_e);
      }
    }
  }
  // ---- last block of the grid: every update of this rank is done.  Sharded: the next gradient kernel reads the rows
  // of A that the OTHER ranks push, so it also waits for their "rows in place" flags before the kernel ends.
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_fault) ctl->fault = 1;
    const unsigned t = atomicAdd(&a.tickets[2], 1u);
    s_last = t == (unsigned)(a.nA + a.nS - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (a.world > 1) {
    const unsigned e2 = *reinterpret_cast<volatile unsigned*>(a.epoch + 2);
    wait_flags(a.my_flags, 2, a.world, e2, &s_fault);
  }
  if (threadIdx.x == 0) {
    if (s_fault) {
      ctl->fault = 1;
      ctl->done = 1;
    }
    a.tickets[0] = 0u;
    a.tickets[2] = 0u;
    ctl->final_pending = 1;
    __threadfence();
    ctl->par_ctr = pc + 1u;
    if (a.trace) atomicMax(a.trace + 12, gtime());
  }
}

// ------------------------------------------------------------------------------------------ final kernel
// One block of 512 threads, off the critical path (side stream, next to the following gradient kernel, which leaves one
// SM free): group 0 (threads 0..255) owns the Gram of A, group 1 the Gram of S.  Each group sums the replicated block
// partials, multi-GPU: all-reduces [Gram partial | 3 norms] through per-rank inbox slots (summed in rank order:
// bit-identical on every rank), runs lambda_max -> step of the OTHER block for the next iteration (nmf.py:44-49).
// Thread 0 then finishes the iteration: convergence test, counter, stop flag (algorithms.py:130-135).
__global__ void __launch_bounds__(2 * TT) k_tail_final(const PgmTailArgs a) {
  pmx_ctl* ctl = a.ctl;
  if (ctl->done || !ctl->final_pending) return;
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_fault2[2];
  __shared__ unsigned s_e2[2];
  const int which = threadIdx.x >> 8;           // 0 = Gram of A, 1 = Gram of S
  lmax::Group grp;
  grp.tid = threadIdx.x & (TT - 1);
  grp.nthr = TT;
  grp.bar = 1 + which;
  const int K = a.K, kk = K * K, world = a.world;
  const unsigned pc = ctl->par_ctr - 1u;        // the iteration whose updates just completed
  const unsigned par = (pc + 1u) & 1u;
  const int cur = (int)(pc & 1u), nxt = cur ^ 1;
  const size_t half = (sizeof(double) * (size_t)(kk + 4) + 16 + lmax::smem_bytes(K) + 15) & ~(size_t)15;
  unsigned char* sm = smem + (size_t)which * half;
  double* Gd = reinterpret_cast<double*>(sm);                       // [K*K] (+ 4)
  unsigned char* lm = sm + sizeof(double) * (size_t)(kk + 4);
  lm += (16 - (reinterpret_cast<size_t>(lm) & 15)) & 15;
  if (grp.tid == 0) s_fault2[which] = 0;
  // local totals: sum of the replicated fp32 accumulators (L2 reads: the roles kernel fed them with red.add), cleared
  // for the next iteration
  float* rep = a.gram_rep + (size_t)which * NREP * kk;
  for (int e0 = 0; e0 < kk; e0 += 4 * TT) {   // 4 elements x NREP loads in flight per thread, then the stores
    float v[4][NREP];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * TT + grp.tid;
#pragma unroll
      for (int r = 0; r < NREP; ++r) v[u][r] = e < kk ? __ldcg(rep + (size_t)r * kk + e) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * TT + grp.tid;
      if (e < kk) {
        double sum = 0.0;
#pragma unroll
        for (int r = 0; r < NREP; ++r) {
          sum += (double)v[u][r];
          rep[(size_t)r * kk + e] = 0.f;
        }
        Gd[e] = sum;
      }
    }
  }
  if (grp.tid < 3) {
    Gd[kk + grp.tid] = __ldcg(a.acc + 4 * which + grp.tid);
    a.acc[4 * which + grp.tid] = 0.0;
  }
  grp.sync();
  // the roles kernel only produced the 4 x 4 tiles on or above the diagonal: mirror them
  for (int e = grp.tid; e < kk; e += TT) {
    const int i = e / K, j = e - i * K;
    if ((i >> 2) > (j >> 2)) Gd[e] = Gd[j * K + i];
  }
  grp.sync();
  if (world > 1) {
    // all-reduce of [Gram partial | 3 norms] through inbox slots: push to slot `rank` of every rank, signal, wait, sum
    // the slots in rank order
    const int set = which == 0 ? 1 : 3;
    const size_t slot_doubles = (size_t)kk + 4;
    const size_t inbox_off = a.off_inbox + sizeof(double) * slot_doubles * PMX_MAX_WORLD * (2 * (size_t)par + which);
    for (int r = 0; r < world; ++r) {
      double* dst = reinterpret_cast<double*>(static_cast<char*>(a.arena.p[r]) + inbox_off) + slot_doubles * a.rank;
      for (int e = grp.tid; e < kk + 3; e += TT) dst[e] = Gd[e];
    }
    __threadfence_system();
    grp.sync();
    if (grp.tid == 0) {
      s_e2[which] = a.epoch[set] + 1;
      a.epoch[set] = s_e2[which];
    }
    grp.sync();
    const unsigned e = s_e2[which];
    if (grp.tid < world)
      st_release_sys(reinterpret_cast<unsigned*>(a.flags.p[grp.tid]) + set * PMX_MAX_WORLD + a.rank, e);
    if (grp.tid < world) {
      const unsigned* f = a.my_flags + set * PMX_MAX_WORLD + grp.tid;
      const long long t0 = clock64();
      while ((int)(ld_acquire_sys(f) - e) < 0) {
        if (clock64() - t0 > SPIN_LIMIT) {
          s_fault2[which] = 1;
          break;
        }
      }
    }
    grp.sync();
    if (!s_fault2[which]) {
      const volatile double* in = reinterpret_cast<const volatile double*>(static_cast<char*>(a.arena.p[a.rank]) + inbox_off);
      for (int i = grp.tid; i < kk + 3; i += TT) {
        double sum = 0.0;
        for (int r = 0; r < world; ++r) sum += in[slot_doubles * r + i];
        Gd[i] = sum;
      }
    }
    grp.sync();
  }
  double nrm[3] = {Gd[kk], Gd[kk + 1], Gd[kk + 2]};
  int status = 0;
  const double lam = lmax::block_lambda_max(Gd, K, lm, 20, &status, grp);
  if (grp.tid == 0) {
    const int j = which == 0 ? 1 : 0;   // Gram of A gives the step of S and vice versa (nmf.py:44-49)
    if (status == 1) {
      ctl->nonfinite = 1;
      ctl->lip[j] = __int_as_float(0x7fc00000);
      ctl->step2[nxt][j] = __int_as_float(0x7fc00000);
    } else if (status == 2) {
      ctl->lip[j] = 0.f;
      ctl->step2[nxt][j] = __int_as_float(0x7f800000);
    } else {
      const float lf = (float)lam;
      ctl->lip[j] = lf;
      ctl->step2[nxt][j] = 1.0f / lf;
    }
    for (int i = 0; i < 3; ++i) ctl->norms[3 * which + i] = nrm[i];   // [0..2] block A, [3..5] block S
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  // algorithms.py:130-133: l2sq(X - X_) <= e_rel**2 * l2sq(X), evaluated in fp32 like the reference
  const bool cA = (float)ctl->norms[0] <= a.e2A * (float)ctl->norms[1];
  const bool cS = (float)ctl->norms[3] <= a.e2S * (float)ctl->norms[4];
  ctl->conv[0] = cA;
  ctl->conv[1] = cS;
  ctl->step[0] = ctl->step2[cur][0];   // the steps this iteration used (algorithms.py:144 returns them)
  ctl->step[1] = ctl->step2[cur][1];
  ctl->it = ctl->it + 1;
  if (s_fault2[0] || s_fault2[1]) ctl->fault = 1;
  if ((cA && cS) || ctl->nonfinite || ctl->fault) ctl->done = 1;
  for (int i = 0; i < 8; ++i) ctl->norms[i] = 0.0;
  ctl->final_pending = 0;
}

__global__ void k_tail_seed_steps(pmx_ctl* ctl) {
  if (ctl->done) return;
  const int cur = (int)(ctl->par_ctr & 1u);
  ctl->step2[cur][0] = ctl->step[0];
  ctl->step2[cur][1] = ctl->step[1];
}

}  // namespace

size_t pgm_tail_smem_bytes(int K) {
  const size_t srole = sizeof(float) * (CT_COLS * CT_LD + CT_COLS * 2 * CT_LD + CT_COLS + 24);
  const size_t arole = sizeof(float) * ((size_t)RA * (((K + 3) & ~3) + 4) + 24);
  return srole > arole ? srole : arole;
}

static size_t tail_final_smem_bytes(int K) {
  const size_t half = (sizeof(double) * ((size_t)K * K + 4) + 16 + lmax::smem_bytes(K) + 15) & ~(size_t)15;
  return 2 * half;
}

int launch_tail_final(pmx_ctx* ctx, cudaStream_t st, const PgmTailArgs& a) {
  static bool attr = false;
  if (!attr) {
    PMX_CUDA(cudaFuncSetAttribute(k_tail_final, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  k_tail_final<<<1, 2 * TT, tail_final_smem_bytes(a.K), st>>>(a);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_tail_final");
}

int pgm_tail_s_blocks(pmx_ctx* ctx, int n_cols, int* n_tiles) {
  const int nt = pmx_div_up(n_cols, CT_COLS);
  *n_tiles = nt;
  const int cap = 3 * ctx->sm_count;   // three 256-thread blocks per SM (registers and 69 KB of shared memory each)
  return nt < cap ? nt : cap;
}

int launch_pgm_tail(pmx_ctx* ctx, const PgmTailArgs& a_in) {
  PgmTailArgs a = a_in;
  // debug: phase timeline of the tail kernel, averaged over the launches and printed every 50
  static const bool tracing = getenv("PMX_TAIL_TRACE") != nullptr;
  static unsigned long long* d_trace = nullptr;
  static double t_acc[16];
  static int t_n = 0;
  if (tracing) {
    if (!d_trace) cudaMalloc((void**)&d_trace, 16 * sizeof(unsigned long long));
    unsigned long long init[16];
    for (int i = 0; i < 16; ++i) init[i] = 0ull;
    init[0] = ~0ull;
    cudaMemcpyAsync(d_trace, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream);
    a.trace = d_trace;
  }
  const size_t smem = pgm_tail_smem_bytes(a.K);
  static bool attr = false;
  if (!attr) {
    PMX_CUDA(cudaFuncSetAttribute(k_pgm_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr = true;
  }
  k_pgm_tail<<<a.nS + a.nA, TT, smem, ctx->stream>>>(a);
  PMX_LAUNCHED(ctx);
  if (tracing) {
    unsigned long long hts[16];
    cudaMemcpyAsync(hts, d_trace, sizeof(hts), cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 1; i < 16; ++i) t_acc[i] += hts[i] ? (double)(hts[i] - hts[0]) * 1e-3 : 0.0;
    if (++t_n % 50 == 0) {
      static const char* names[16] = {"", "S loop end", "S role end (reds)", "A role end", "", "",
                                      "", "", "", "", "", "", "roles kernel end",
                                      "S: loads used", "S: unity done", "S: stores issued"};
      for (int i = 1; i < 16; ++i)
        if (names[i][0]) printf("TAIL %-20s at %8.2f us\n", names[i], t_acc[i] / 50), t_acc[i] = 0;
      fflush(stdout);
    }
  }
  return pmx_check_launch(ctx, "k_pgm_tail");
}

int launch_tail_seed_steps(pmx_ctx* ctx, pmx_ctl* ctl) {
  k_tail_seed_steps<<<1, 1, 0, ctx->stream>>>(ctl);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_tail_seed_steps");
}
