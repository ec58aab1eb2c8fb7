// Fused NMF gradient kernel for sm_100a: one pass over Y computes
//     R = A S - Y,   G_A = R S^T,   G_S = A^T R,   loss = |R|^2 / 2          (nmf.py:13-41)
// with tcgen05 tensor-core MMAs (TMEM accumulators), TMA loads, and warp-specialised roles.
//
// Precision: every GEMM is the 3-term BF16 split  hi*hi + hi*lo + lo*hi  with fp32 accumulation
// (x = hi + lo, |x - hi - lo| <= 2^-17 |x|).  SURVEY 7.3 measured that single-pass TF32/BF16
// operands miss the 1e-4 parity target by 10-100x while this split clears it with margin.
// A and S arrive pre-split (k_split_bf16); R is split in the epilogue registers.
//
// Tiling: a tile is 128 rows (m) x 128 columns (n) of Y; K <= 64 (operands zero-padded to 64).
// The tile sequence is m-block major (all 128-column stripes of a 128-row block are consecutive) and
// is cut into gridDim.x contiguous, equally long ranges -- one persistent CTA per SM -- so the load is
// balanced to within one tile.  Inside a row segment G_A accumulates in TMEM across the stripes and is
// flushed once; the 128 x 64 G_S^T partial of every tile is flushed with red.add into G_S, lanes
// running along n so that every warp-level reduction is one fully coalesced 128-byte line.
//
// Shared memory (all operands are "panels": rows of 128 bytes, 128B-swizzled in 8-row atoms, which
// the same bytes can be read as a K-major or an MN-major UMMA operand):
//   S_hi,S_lo  [2 n-panels][64 k-rows][128B] x 2 slots  64 KB   per tile      (TMA)
//   A_hi,A_lo  [128 m-rows][128B]              32 KB   per row segment   (TMA)
//   R_hi,R_lo  [2 n-panels][128 m-rows][128B]  64 KB   per tile          (epilogue writes)
//   Y ring     4 x [128 m-rows][32 fp32]       64 KB   4 sub-tiles/tile  (TMA)
// TMEM (512 columns): residual accumulator 2 x 128, G_S^T accumulator 2 x 64, G_A accumulator 64.
//
// Warp roles (512 threads): warp 0 = TMA producer, warps 1, 2, 3 = MMA issuers (residual / G_A / G_S; warp 2
// also allocates the tensor memory), warps 4..11 = residual warps (TMEM accumulator -> R in TMEM and SMEM; two warps
// share each TMEM lane quarter and split the column chunks), warps 12..15 = gradient flush warps.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "grad_umma.h"
#include "kernels.h"

namespace {

constexpr int TILE_M = 128, TILE_N = 128, KP = 64;
constexpr int Y_SUB = 32;           // columns per Y sub-tile (128 bytes of fp32)
constexpr int Y_STAGES = 4;
constexpr uint32_t PANEL_S = 64 * 128;    // bytes of one S panel (64 k-rows)
constexpr uint32_t PANEL_R = 128 * 128;   // bytes of one R / A panel (128 m-rows)

// shared-memory map (offsets from the 1024-aligned base)
constexpr uint32_t OFF_S = 0;                               // slot s: hi at OFF_S + s*4*PANEL_S, lo 2 panels later
constexpr uint32_t OFF_A = OFF_S + 8 * PANEL_S;             // A_hi, then A_lo (one m-block at a time)
constexpr uint32_t OFF_R_HI = OFF_A + 2 * PANEL_R;
constexpr uint32_t OFF_R_LO = OFF_R_HI + 2 * PANEL_R;
constexpr uint32_t OFF_Y = OFF_R_LO + 2 * PANEL_R;
constexpr uint32_t OFF_BAR = OFF_Y + Y_STAGES * PANEL_R;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 512 + 1024;       // barriers + alignment slack

enum {  // mbarrier indices
  B_A_FULL = 0, B_A_EMPTY, B_S_FULL, B_S_EMPTY = B_S_FULL + 2, B_Y_FULL = B_S_EMPTY + 2,
  B_Y_EMPTY = B_Y_FULL + Y_STAGES, B_ACC_FULL = B_Y_EMPTY + Y_STAGES, B_ACC_EMPTY = B_ACC_FULL + 2,
  B_RT_FULL = B_ACC_EMPTY + 2, B_RS_FULL = B_RT_FULL + 2, B_RS_EMPTY, B_GS_FULL, B_GS_EMPTY = B_GS_FULL + 2,
  B_GA_FULL = B_GS_EMPTY + 2, B_GA_EMPTY, B_COUNT
};

// TMEM column map (512): residual accumulator / R operand 2 x 128, G_S^T 2 x 64 (double-buffered so that the
// red.add flush of tile t overlaps the MMAs of tile t+1), G_A [hh|hl] 128
constexpr uint32_t TM_ACC = 0, TM_GS = 256, TM_GA = 384;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// single non-blocking probe of a barrier phase (test_wait returns immediately; try_wait may suspend the thread
// for a system-dependent time when the phase is not complete)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
// TMA load with an L2 eviction-priority hint: Y is streamed once (evict_first) while the S / A operand tiles are
// re-read by every m-block row / stripe (evict_last), so the 2 GB Y stream does not push them out of L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], "
      "[%4], %5;" ::"r"(dst),
      "l"(map), "r"(x), "r"(y), "r"(bar), "l"(policy)
      : "memory");
}
// TMA prefetch of one box into L2 (no shared memory, no barrier): decouples the HBM latency from the ring depth
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int x, int y) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (bf16 inputs, fp32 accumulate), M x N x 16
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M x 16 bf16 = 128 lanes x 8 columns) is read from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The issuer warps run warp-uniform code and elect one lane inside the asm statement: descriptors and tensor
// memory addresses then stay in uniform registers instead of being broadcast lane by lane before every MMA.
__device__ __forceinline__ void umma_ss_e(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_ts_e(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_e(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// In-register 4x4 transpose across the 4 lanes of a quad: before, lane r holds row r = (v0..v3);
// after, lane r holds column r.  Two xor-butterfly rounds (2x2 blocks across lane^1, then lane^2).
__device__ __forceinline__ void quad_transpose(float& v0, float& v1, float& v2, float& v3, int lane) {
  const bool b0 = lane & 1, b1 = lane & 2;
  float x, y;
  x = b0 ? v0 : v1; y = __shfl_xor_sync(0xffffffffu, x, 1); if (b0) v0 = y; else v1 = y;
  x = b0 ? v2 : v3; y = __shfl_xor_sync(0xffffffffu, x, 1); if (b0) v2 = y; else v3 = y;
  x = b1 ? v0 : v2; y = __shfl_xor_sync(0xffffffffu, x, 2); if (b1) v0 = y; else v2 = y;
  x = b1 ? v1 : v3; y = __shfl_xor_sync(0xffffffffu, x, 2); if (b1) v1 = y; else v3 = y;
}

// UMMA shared-memory descriptor, 128B swizzle (cute::UMMA::SmemDescriptor layout, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// instruction descriptor: bf16 x bf16 -> fp32 (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Params {
  int M, N, K;
  int NS;                 // 128-column stripes per m-block row
  long long total_tiles;  // m-blocks * NS
  float* GA;
  float* GS;
  double* loss;
  const int* done;
  int y_prefetch;         // tiles of Y prefetched into L2 ahead of the shared-memory ring
  long long* trace;       // debug: clock64 timeline of CTA 0 (env PMX_TRACE), [role][tile][event]
  int ablate;             // debug/timing only (env PMX_ABLATE): bit0 no MMA1, 1 no MMA2, 2 no MMA3, 3 no Y read, 4 no R store, 5 no G_S flush, 6 no epilogue math
};

// timeline tracing (debug only): role r, local tile index t, event e
#define TRACE_TILES 24
#define TRACE_EVENTS 8
#define TR(r, t, e)                                                                                   \
  do {                                                                                                \
    if (p.trace && blockIdx.x == 0 && lane == 0 && (t) < TRACE_TILES)                                 \
      p.trace[((r) * TRACE_TILES + (t)) * TRACE_EVENTS + (e)] = clock64();                            \
  } while (0)

constexpr int NUM_EPI_WARPS = 8;     // residual warps: TMEM accumulator -> R (TMEM + SMEM)
constexpr int NUM_FLUSH_WARPS = 4;   // gradient flush warps: TMEM accumulators -> red.add to global memory
constexpr int NUM_THREADS = 128 + 32 * (NUM_EPI_WARPS + NUM_FLUSH_WARPS);

__global__ void __launch_bounds__(NUM_THREADS, 1)
k_grad_umma(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmAhi,
            const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmShi,
            const __grid_constant__ CUtensorMap tmSlo, Params p) {
  if (p.done && *p.done) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + OFF_BAR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + OFF_BAR + 8 * B_COUNT);
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  // warp index through a shuffle: the compiler then treats it (and every branch on it) as warp-uniform, which keeps
  // the MMA issuers' descriptors and tensor-memory addresses in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  // tile sequence: m-block major (all 128-column stripes of one 128-row block are consecutive)
  const long long g_begin = (long long)blockIdx.x * p.total_tiles / gridDim.x;
  const long long g_end = (long long)(blockIdx.x + 1) * p.total_tiles / gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < B_COUNT; ++i) {
      uint32_t count = 1;
      if (i == B_A_EMPTY) count = 2;                                              // both issuer warps release A
      if (i >= B_Y_EMPTY && i < B_Y_EMPTY + Y_STAGES) count = NUM_EPI_WARPS / 2;   // one column-chunk group
      if (i == B_ACC_EMPTY || i == B_ACC_EMPTY + 1) count = NUM_EPI_WARPS;
      if (i == B_RT_FULL || i == B_RT_FULL + 1) count = NUM_EPI_WARPS;
      if (i == B_RS_FULL) count = NUM_EPI_WARPS;
      if (i == B_GS_EMPTY || i == B_GS_EMPTY + 1) count = NUM_FLUSH_WARPS;
      if (i == B_GA_EMPTY) count = NUM_FLUSH_WARPS;
      mbar_init(bar(i), count);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  auto first_in_seg = [&](long long g) { return g == g_begin || (g % p.NS) == 0; };
  auto last_in_seg = [&](long long g) { return g == g_end - 1 || (g % p.NS) == p.NS - 1; };
  // S slot layout: n-panel pn (64 columns) = [S_hi 64 k-rows | S_lo 64 k-rows] -> 16 KB per panel, so that
  // [S_hi; S_lo] is one 128-row K-major operand (the "stacked" N = 128 operand of the G_A GEMM)
  constexpr uint32_t S_SLOT = 4 * PANEL_S, S_PANEL = 2 * PANEL_S;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      uint32_t t = 0, seg = 0;
      const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
      const bool hints = !(p.ablate & 128);
      for (long long g = g_begin; g < g_end; ++g, ++t) {
        const int mb = (int)(g / p.NS), stripe = (int)(g % p.NS);
        const int m0 = mb * TILE_M, n0 = stripe * TILE_N;
        const uint32_t slot = t & 1;
        mbar_wait(bar(B_S_EMPTY + slot), ((t >> 1) & 1) ^ 1);
        mbar_expect_tx(bar(B_S_FULL + slot), 4 * PANEL_S);
        const uint32_t sb = base + OFF_S + slot * S_SLOT;
        if (hints) {
          tma_load_2d_hint(sb, &tmShi, n0, 0, bar(B_S_FULL + slot), pol_keep);
          tma_load_2d_hint(sb + PANEL_S, &tmSlo, n0, 0, bar(B_S_FULL + slot), pol_keep);
          tma_load_2d_hint(sb + S_PANEL, &tmShi, n0 + 64, 0, bar(B_S_FULL + slot), pol_keep);
          tma_load_2d_hint(sb + S_PANEL + PANEL_S, &tmSlo, n0 + 64, 0, bar(B_S_FULL + slot), pol_keep);
        } else {
          tma_load_2d(sb, &tmShi, n0, 0, bar(B_S_FULL + slot));
          tma_load_2d(sb + PANEL_S, &tmSlo, n0, 0, bar(B_S_FULL + slot));
          tma_load_2d(sb + S_PANEL, &tmShi, n0 + 64, 0, bar(B_S_FULL + slot));
          tma_load_2d(sb + S_PANEL + PANEL_S, &tmSlo, n0 + 64, 0, bar(B_S_FULL + slot));
        }
        TR(0, t, 0);
        for (int q = 0; q < 4; ++q) {
          mbar_wait(bar(B_Y_EMPTY + q), (t & 1) ^ 1);
          mbar_expect_tx(bar(B_Y_FULL + q), PANEL_R);
          if (hints)
            tma_load_2d_hint(base + OFF_Y + q * PANEL_R, &tmY, n0 + q * Y_SUB, m0, bar(B_Y_FULL + q), pol_stream);
          else
            tma_load_2d(base + OFF_Y + q * PANEL_R, &tmY, n0 + q * Y_SUB, m0, bar(B_Y_FULL + q));
          TR(0, t, 1 + q);
        }
        if (first_in_seg(g)) {
          mbar_wait(bar(B_A_EMPTY), (seg & 1) ^ 1);
          mbar_expect_tx(bar(B_A_FULL), 2 * PANEL_R);
          tma_load_2d(base + OFF_A, &tmAhi, 0, m0, bar(B_A_FULL));
          tma_load_2d(base + OFF_A + PANEL_R, &tmAlo, 0, m0, bar(B_A_FULL));
          ++seg;
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer 1: residual GEMM ==============================
    // Three issuer warps (residual / G_A / G_S), each blocking only on its own dependencies; all 32 lanes run
    // the loop and one elected lane issues.  MMA1(t+2) overwrites the accumulator whose columns hold R(t), the
    // TMEM operand of MMA2(t): that hazard is ordered through the S slot -- S(t+2) is loaded only after MMA2(t)
    // committed s_empty, and MMA1(t+2) waits for S(t+2).
    constexpr uint32_t ID_RES = make_idesc(128, 128, 0, 1);   // A tile K-major (SMEM), S tile MN-major
    const uint32_t a_hi = base + OFF_A, a_lo = a_hi + PANEL_R;
    uint32_t t = 0, seg_full = 0;
    for (long long g = g_begin; g < g_end; ++g, ++t) {
      const uint32_t slot = t & 1;
      mbar_wait(bar(B_S_FULL + slot), (t >> 1) & 1);
      if (first_in_seg(g)) {
        mbar_wait(bar(B_A_FULL), seg_full & 1);
        ++seg_full;
      }
      mbar_wait(bar(B_ACC_EMPTY + slot), ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      TR(1, t, 0);
      const uint32_t sb = base + OFF_S + slot * S_SLOT;
      const uint32_t d = tmem + TM_ACC + slot * 128;
      // acc = A_hi S_hi + A_hi S_lo + A_lo S_hi      (K = 64: 4 k-steps of 16)
      const uint32_t a_src[3] = {a_hi, a_hi, a_lo};
      const uint32_t s_src[3] = {sb, sb + PANEL_S, sb};
#pragma unroll
      for (int term = 0; term < 3; ++term)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          if (p.ablate & 1) continue;
          const uint64_t ad = make_desc(a_src[term] + ks * 32, 16, 1024);             // K-major
          const uint64_t bd = make_desc(s_src[term] + ks * 2048, S_PANEL, 1024);      // MN-major: LBO = next 64 n
          umma_ss_e(d, ad, bd, ID_RES, (term | ks) ? 1u : 0u);
        }
      tc_commit_e(bar(B_ACC_FULL + slot));
      TR(1, t, 1);
      if (last_in_seg(g)) tc_commit_e(bar(B_A_EMPTY));   // residual GEMMs of this segment are done with the A tile
    }
  } else if (warp == 2) {
    // ============================== MMA issuer 2: G_A GEMM (after the TMEM allocation above) ================
    // G_A[m, (hh|hl)] += R_hi [S_hi;S_lo]^T ; G_A[m, hh] += R_lo S_hi^T   (K = n: 8 k-steps).  R (bf16 hi/lo)
    // sits in the columns of the residual accumulator it was computed from: every 16 accumulator columns become
    // [hi 8 cols | lo 8 cols] of packed bf16 pairs
    constexpr uint32_t ID_GA2 = make_idesc(128, 128, 0, 0);   // R_hi from TMEM x [S_hi;S_lo] K-major, N = 128
    constexpr uint32_t ID_GA1 = make_idesc(128, 64, 0, 0);    // R_lo from TMEM x S_hi, N = 64
    uint32_t t = 0, seg = 0;
    for (long long g = g_begin; g < g_end; ++g, ++t) {
      const uint32_t slot = t & 1;
      const bool first = first_in_seg(g);
      mbar_wait(bar(B_S_FULL + slot), (t >> 1) & 1);    // already complete (MMA1(t) consumed it): visibility only
      mbar_wait(bar(B_RT_FULL + slot), (t >> 1) & 1);
      if (first) mbar_wait(bar(B_GA_EMPTY), (seg & 1) ^ 1);
      tc_fence_after();
      TR(1, t, 2);
      const uint32_t sb = base + OFF_S + slot * S_SLOT;
      const uint32_t d = tmem + TM_GA;
      const uint32_t racc = tmem + TM_ACC + slot * 128;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        if (p.ablate & 2) continue;
        const uint32_t at = racc + ks * 16;        // R_hi pairs of the 16 columns of k-step ks
        const uint64_t bd = make_desc(sb + (ks >> 2) * S_PANEL + (ks & 3) * 32, 16, 1024);   // K-major, 128 rows
        umma_ts_e(d, at, bd, ID_GA2, (!first || ks) ? 1u : 0u);
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        if (p.ablate & 2) continue;
        const uint32_t at = racc + ks * 16 + 8;    // R_lo pairs
        const uint64_t bd = make_desc(sb + (ks >> 2) * S_PANEL + (ks & 3) * 32, 16, 1024);   // K-major, rows 0..63
        umma_ts_e(d, at, bd, ID_GA1, 1u);
      }
      tc_commit_e(bar(B_S_EMPTY + slot));   // S(t) was last used here
      TR(1, t, 3);
      if (last_in_seg(g)) {
        tc_commit_e(bar(B_GA_FULL));
        ++seg;
      }
    }
  } else if (warp == 3) {
    // ============================== MMA issuer 3: G_S GEMM ==============================
    // G_S^T[n, k] = R_hi^T A_hi + R_hi^T A_lo + R_lo^T A_hi   (M = n, N = k = 64, K = m: 8 k-steps), R^T from SMEM
    constexpr uint32_t ID_GS = make_idesc(128, 64, 1, 1);
    const uint32_t a_hi = base + OFF_A, a_lo = a_hi + PANEL_R;
    const uint32_t r_hi = base + OFF_R_HI, r_lo = base + OFF_R_LO;
    uint32_t t = 0, seg_full = 0;
    for (long long g = g_begin; g < g_end; ++g, ++t) {
      const uint32_t slot = t & 1;
      if (first_in_seg(g)) {
        mbar_wait(bar(B_A_FULL), seg_full & 1);
        ++seg_full;
      }
      mbar_wait(bar(B_RS_FULL), t & 1);
      TR(2, t, 2);
      mbar_wait(bar(B_GS_EMPTY + slot), ((t >> 1) & 1) ^ 1);
      tc_fence_after();
      TR(2, t, 0);
      const uint32_t d = tmem + TM_GS + slot * 64;
      const uint32_t r_src[3] = {r_hi, r_hi, r_lo};
      const uint32_t a_src[3] = {a_hi, a_lo, a_hi};
#pragma unroll
      for (int term = 0; term < 3; ++term)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          if (p.ablate & 4) continue;
          const uint64_t ad = make_desc(r_src[term] + ks * 2048, PANEL_R, 1024);  // MN-major: LBO = next 64 n
          const uint64_t bd = make_desc(a_src[term] + ks * 2048, PANEL_R, 1024);  // MN-major, one 64-wide atom
          umma_ss_e(d, ad, bd, ID_GS, (term | ks) ? 1u : 0u);
        }
      tc_commit_e(bar(B_RS_EMPTY));
      tc_commit_e(bar(B_GS_FULL + slot));
      TR(2, t, 1);
      if (last_in_seg(g)) tc_commit_e(bar(B_A_EMPTY));
    }
  } else if (warp >= 4 && warp < 4 + NUM_EPI_WARPS) {
    // ============================== residual warps (8) ==============================
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may access
    const int grp = (warp - 4) >> 2;         // column-chunk group: chunks q with (q & 1) == grp
    const int row = q4 * 32 + lane;          // row m of the tile
    const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
    float loss_part = 0.f;
    uint32_t t = 0;
    for (long long g = g_begin; g < g_end; ++g, ++t) {
      const uint32_t slot = t & 1;
      mbar_wait(bar(B_ACC_FULL + slot), (t >> 1) & 1);
      tc_fence_after();
      if (warp == 4) TR(3, t, 0);
      // ---- phase A: residual -> bf16 (hi, lo), written back over the accumulator columns it came from
#pragma unroll 1
      for (int qq = 0; qq < 2; ++qq) {
        const int q = qq * 2 + grp;
        mbar_wait(bar(B_Y_FULL + q), t & 1);
        if (warp == 4) TR(3, t, 1 + qq);
        uint32_t acc[32];
        tmem_ld32(lane_addr + TM_ACC + slot * 128 + q * 32, acc);
        const uint8_t* yrow = base_ptr + OFF_Y + q * PANEL_R + row * 128;
        float4 yv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
          yv[c] = (p.ablate & 8) ? make_float4(0.f, 0.f, 0.f, 0.f)
                                 : *reinterpret_cast<const float4*>(yrow + ((c ^ (row & 7)) << 4));
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_Y_EMPTY + q));   // the Y values are in registers: refill the slot now
        tmem_ld_wait();
        const float* yf = reinterpret_cast<const float*>(yv);
        uint32_t hl[32];   // per 16 columns: [hi 8 pairs | lo 8 pairs]
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float r0 = __uint_as_float(acc[j]) - yf[j];          // nmf.py:40  (A S - Y)
          const float r1 = __uint_as_float(acc[j + 1]) - yf[j + 1];
          if (p.loss) {
            loss_part = fmaf(r0, r0, loss_part);
            loss_part = fmaf(r1, r1, loss_part);
          }
          const __nv_bfloat162 h = __floats2bfloat162_rn(r0, r1);
          const float2 hf = __bfloat1622float2(h);
          const __nv_bfloat162 l = __floats2bfloat162_rn(r0 - hf.x, r1 - hf.y);
          hl[(j >> 4) * 16 + ((j >> 1) & 7)] = *reinterpret_cast<const uint32_t*>(&h);
          hl[(j >> 4) * 16 + 8 + ((j >> 1) & 7)] = *reinterpret_cast<const uint32_t*>(&l);
        }
        tmem_st16(lane_addr + TM_ACC + slot * 128 + q * 32, hl);
        tmem_st16(lane_addr + TM_ACC + slot * 128 + q * 32 + 16, hl + 16);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_RT_FULL + slot));       // MMA2(t) may start
      if (warp == 4) TR(3, t, 3);
      // ---- phase B: copy R to shared memory (the MN-major operand of the G_S GEMM) once MMA3(t-1) released it
      mbar_wait(bar(B_RS_EMPTY), (t & 1) ^ 1);
      if (warp == 4) TR(3, t, 4);
#pragma unroll 1
      for (int qq = 0; qq < 2; ++qq) {
        const int q = qq * 2 + grp;
        uint32_t hl[32];   // per 16 columns: [hi 8 | lo 8] pairs
        tmem_ld32(lane_addr + TM_ACC + slot * 128 + q * 32, hl);
        tmem_ld_wait();
        uint8_t* rh = base_ptr + OFF_R_HI + (q >> 1) * PANEL_R + row * 128;
        uint8_t* rl = base_ptr + OFF_R_LO + (q >> 1) * PANEL_R + row * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (p.ablate & 16) continue;
          const int chunk = ((q & 1) * 4 + c) ^ (row & 7);
          const int o = (c >> 1) * 16 + (c & 1) * 4;     // hi pairs of elements 8c .. 8c+7
          *reinterpret_cast<uint4*>(rh + (chunk << 4)) = make_uint4(hl[o], hl[o + 1], hl[o + 2], hl[o + 3]);
          *reinterpret_cast<uint4*>(rl + (chunk << 4)) = make_uint4(hl[o + 8], hl[o + 9], hl[o + 10], hl[o + 11]);
        }
      }
      tc_fence_before();
      fence_async_smem();   // generic-proxy writes of R -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(B_RS_FULL));             // MMA3(t) may start
        mbar_arrive(bar(B_ACC_EMPTY + slot));    // our reads of this accumulator/R buffer are done
      }
      if (warp == 4) TR(3, t, 5);
    }
    if (p.loss) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss_part += __shfl_xor_sync(0xffffffffu, loss_part, o);
      if (lane == 0) atomicAdd(p.loss, 0.5 * (double)loss_part);
    }
  } else if (warp >= 4 + NUM_EPI_WARPS) {
    // ============================== gradient flush warps (4) ==============================
    // The red.add traffic (32 KB per tile into the L2-resident G_S) back-pressures the issuing warp; keeping it
    // on dedicated warps takes it off the residual warps' critical loop.
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;          // n for G_S^T, m for G_A
    const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
    uint32_t t = 0, seg = 0;
    for (long long g = g_begin; g < g_end; ++g, ++t) {
      {  // G_S^T[n, k] = hh + hl halves -> G_S[k, n]: for a fixed k the 32 lanes hit one 128-byte line
        const int n = (int)(g % p.NS) * TILE_N + row;
        const uint32_t slot = t & 1;
        mbar_wait(bar(B_GS_FULL + slot), (t >> 1) & 1);
        tc_fence_after();
        if (warp == 12) TR(4, t, 0);
        uint32_t v[32], w[32];
        tmem_ld32(lane_addr + TM_GS + slot * 64, v);
        tmem_ld32(lane_addr + TM_GS + slot * 64 + 32, w);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_GS_EMPTY + slot));   // values are in registers: the accumulator is free
        if (warp == 12) TR(4, t, 1);
        if (!(p.ablate & 32) && !((p.ablate & 64) && (t & 1))) {
          const int n_tile0 = (int)(g % p.NS) * TILE_N;
          if (p.K == KP && n_tile0 + TILE_N <= p.N && (p.N & 3) == 0) {
            // Full tile: 16 4x4 quad transposes done in lock step (all first-round shuffles, then all second-round
            // shuffles) so that the 64 shuffles pipeline instead of forming 16 dependent chains, then 16-byte
            // reductions with no per-instruction predicate: lane 4j+r ends up with G_S[k = 4i+r][n = 4j .. 4j+3],
            // a warp-level red covers 4 full 128-byte lines.
            const bool b0 = lane & 1, b1 = lane & 2;
            float xa[16], xb[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t* s4 = (i < 8) ? &v[4 * i] : &w[4 * (i - 8)];
              xa[i] = __uint_as_float(b0 ? s4[0] : s4[1]);
              xb[i] = __uint_as_float(b0 ? s4[2] : s4[3]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              xa[i] = __shfl_xor_sync(0xffffffffu, xa[i], 1);
              xb[i] = __shfl_xor_sync(0xffffffffu, xb[i], 1);
            }
            float c0[16], c1[16], c2[16], c3[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t* s4 = (i < 8) ? &v[4 * i] : &w[4 * (i - 8)];
              c0[i] = b0 ? xa[i] : __uint_as_float(s4[0]);
              c1[i] = b0 ? __uint_as_float(s4[1]) : xa[i];
              c2[i] = b0 ? xb[i] : __uint_as_float(s4[2]);
              c3[i] = b0 ? __uint_as_float(s4[3]) : xb[i];
              xa[i] = b1 ? c0[i] : c2[i];
              xb[i] = b1 ? c1[i] : c3[i];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              xa[i] = __shfl_xor_sync(0xffffffffu, xa[i], 2);
              xb[i] = __shfl_xor_sync(0xffffffffu, xb[i], 2);
            }
            float* dst = p.GS + (size_t)(lane & 3) * p.N + (n_tile0 + q4 * 32 + (lane & ~3));
            const size_t stride4 = (size_t)4 * p.N;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float o0 = b1 ? xa[i] : c0[i], o2 = b1 ? c2[i] : xa[i];
              const float o1 = b1 ? xb[i] : c1[i], o3 = b1 ? c3[i] : xb[i];
              red_add_v4(dst + i * stride4, o0, o1, o2, o3);
            }
          } else if ((p.N & 3) == 0) {
            // 4x4 quad transposes: lane 4j+r ends up with G_S[k = 4i+r][n = 4j .. 4j+3] -> one 16-byte red per
            // 4 values (16 instead of 64 reductions per thread; a warp-level red covers 4 full 128-byte lines)
            const int nq = (int)(g % p.NS) * TILE_N + q4 * 32 + (lane & ~3);
            const int r = lane & 3;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float a0 = __uint_as_float(v[4 * i]), a1 = __uint_as_float(v[4 * i + 1]);
              float a2 = __uint_as_float(v[4 * i + 2]), a3 = __uint_as_float(v[4 * i + 3]);
              quad_transpose(a0, a1, a2, a3, lane);
              const int k = 4 * i + r;
              if (k < p.K && nq < p.N) red_add_v4(p.GS + (size_t)k * p.N + nq, a0, a1, a2, a3);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float a0 = __uint_as_float(w[4 * i]), a1 = __uint_as_float(w[4 * i + 1]);
              float a2 = __uint_as_float(w[4 * i + 2]), a3 = __uint_as_float(w[4 * i + 3]);
              quad_transpose(a0, a1, a2, a3, lane);
              const int k = 32 + 4 * i + r;
              if (k < p.K && nq < p.N) red_add_v4(p.GS + (size_t)k * p.N + nq, a0, a1, a2, a3);
            }
          } else if (n < p.N) {
            float* dst = p.GS + n;
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (k < p.K) atomicAdd(dst + (size_t)k * p.N, __uint_as_float(v[k]));
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (32 + k < p.K) atomicAdd(dst + (size_t)(32 + k) * p.N, __uint_as_float(w[k]));
          }
        }
        if (warp == 12) TR(4, t, 2);
      }
      if (last_in_seg(g)) {
        // G_A[m, k] = hh + hl halves of the whole segment: once per m-block row and CTA
        mbar_wait(bar(B_GA_FULL), seg & 1);
        tc_fence_after();
        const int m = (int)(g / p.NS) * TILE_M + row;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32], w[32];
          tmem_ld32(lane_addr + TM_GA + half * 32, v);
          tmem_ld32(lane_addr + TM_GA + 64 + half * 32, w);
          tmem_ld_wait();
          if (half == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_GA_EMPTY));
          }
          if (m < p.M) {
            float* dst = p.GA + (size_t)m * p.K + half * 32;
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = __float_as_uint(__uint_as_float(v[k]) + __uint_as_float(w[k]));
            if ((p.K & 3) == 0) {
#pragma unroll
              for (int k = 0; k < 32; k += 4)
                if (half * 32 + k < p.K)
                  red_add_v4(dst + k, __uint_as_float(v[k]), __uint_as_float(v[k + 1]), __uint_as_float(v[k + 2]),
                             __uint_as_float(v[k + 3]));
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (half * 32 + k < p.K) atomicAdd(dst + k, __uint_as_float(v[k]));
            }
          }
        }
        ++seg;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_map(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* ptr, uint64_t cols, uint64_t rows,
             uint64_t pitch_bytes, uint32_t box_cols, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    pmx_set_error("cuTensorMapEncodeTiled is not available from the driver");
    return PMX_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  (void)elem_bytes;
  CUresult r = enc(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    pmx_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PMX_ERR_CUDA;
  }
  return PMX_OK;
}

}  // namespace

struct UmmaPlan {
  int M, N, K, Mp, Np;
  void *Ahi, *Alo, *Shi, *Slo;  // bf16 operand buffers (zero padded)
  CUtensorMap tmY, tmAhi, tmAlo, tmShi, tmSlo;
};

bool umma_supported(int M, int N, int K) { return K >= 1 && K <= KP && M >= 1 && N >= 1; }

int umma_plan_create(pmx_ctx* ctx, const float* Y, int ldY, int M, int N, int K, UmmaPlan** out) {
  PMX_REQUIRE(umma_supported(M, N, K), "unsupported shape for the tcgen05 kernel");
  PMX_REQUIRE((ldY % 4) == 0 && (reinterpret_cast<uintptr_t>(Y) % 16) == 0, "Y must be 16-byte aligned with a pitch multiple of 4");
  UmmaPlan* pl = new UmmaPlan();
  memset(pl, 0, sizeof(*pl));
  pl->M = M; pl->N = N; pl->K = K;
  pl->Mp = pmx_div_up(M, TILE_M) * TILE_M;
  pl->Np = pmx_div_up(N, TILE_N) * TILE_N;
  PMX_CUDA(cudaSetDevice(ctx->device));
  PMX_CHECK(pmx_dev_alloc(ctx, &pl->Ahi, (size_t)pl->Mp * KP * 2));
  PMX_CHECK(pmx_dev_alloc(ctx, &pl->Alo, (size_t)pl->Mp * KP * 2));
  PMX_CHECK(pmx_dev_alloc(ctx, &pl->Shi, (size_t)KP * pl->Np * 2));
  PMX_CHECK(pmx_dev_alloc(ctx, &pl->Slo, (size_t)KP * pl->Np * 2));
  PMX_CUDA(cudaMemsetAsync(pl->Ahi, 0, (size_t)pl->Mp * KP * 2, ctx->stream));
  PMX_CUDA(cudaMemsetAsync(pl->Alo, 0, (size_t)pl->Mp * KP * 2, ctx->stream));
  PMX_CUDA(cudaMemsetAsync(pl->Shi, 0, (size_t)KP * pl->Np * 2, ctx->stream));
  PMX_CUDA(cudaMemsetAsync(pl->Slo, 0, (size_t)KP * pl->Np * 2, ctx->stream));
  // Y: fp32, box 32 columns x 128 rows; out-of-bounds rows/columns are zero-filled by TMA
  PMX_CHECK(make_map(&pl->tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, Y, (uint64_t)N, (uint64_t)M, (uint64_t)ldY * 4, Y_SUB, TILE_M));
  PMX_CHECK(make_map(&pl->tmAhi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl->Ahi, KP, (uint64_t)pl->Mp, KP * 2, KP, TILE_M));
  PMX_CHECK(make_map(&pl->tmAlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl->Alo, KP, (uint64_t)pl->Mp, KP * 2, KP, TILE_M));
  PMX_CHECK(make_map(&pl->tmShi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl->Shi, (uint64_t)pl->Np, KP, (uint64_t)pl->Np * 2, 64, KP));
  PMX_CHECK(make_map(&pl->tmSlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl->Slo, (uint64_t)pl->Np, KP, (uint64_t)pl->Np * 2, 64, KP));
  static bool attr = false;
  if (!attr) {
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr = true;
  }
  *out = pl;
  return PMX_OK;
}

void umma_plan_destroy(pmx_ctx* ctx, UmmaPlan* pl) {
  if (!pl) return;
  pmx_dev_free(ctx, pl->Ahi);
  pmx_dev_free(ctx, pl->Alo);
  pmx_dev_free(ctx, pl->Shi);
  pmx_dev_free(ctx, pl->Slo);
  delete pl;
}

void umma_plan_buffers(UmmaPlan* pl, void** Ahi, void** Alo, void** Shi, void** Slo, int* ldS) {
  *Ahi = pl->Ahi; *Alo = pl->Alo; *Shi = pl->Shi; *Slo = pl->Slo; *ldS = pl->Np;
}

int launch_grad_umma(pmx_ctx* ctx, UmmaPlan* pl, const float* A, const float* S, float* GA, float* GS, double* loss,
                     const int* done, int skip_split) {
  if (!skip_split) {
    PMX_CHECK(launch_split_bf16(ctx, A, pl->M, pl->K, pl->Ahi, pl->Alo, pl->Mp, KP, done));
    PMX_CHECK(launch_split_bf16(ctx, S, pl->K, pl->N, pl->Shi, pl->Slo, KP, pl->Np, done));
  }
  PMX_CHECK(launch_zero3(ctx, ctx->stream, GS, (size_t)pl->K * pl->N, GA, (size_t)pl->M * pl->K,
                         reinterpret_cast<float*>(loss), loss ? 2 : 0, done));
  Params p;
  p.M = pl->M; p.N = pl->N; p.K = pl->K;
  p.NS = pl->Np / TILE_N;
  p.total_tiles = (long long)(pl->Mp / TILE_M) * p.NS;
  p.GA = GA; p.GS = GS; p.loss = loss; p.done = done;
  {
    const char* ab = getenv("PMX_ABLATE");
    p.ablate = ab ? atoi(ab) : 0;
    static long long* d_trace = nullptr;
    p.trace = nullptr;
    if (getenv("PMX_TRACE")) {
      if (!d_trace) cudaMalloc((void**)&d_trace, sizeof(long long) * 5 * TRACE_TILES * TRACE_EVENTS);
      cudaMemsetAsync(d_trace, 0, sizeof(long long) * 5 * TRACE_TILES * TRACE_EVENTS, ctx->stream);
      p.trace = d_trace;
    }
    const char* pf = getenv("PMX_Y_PREFETCH");
    p.y_prefetch = pf ? atoi(pf) : 0;
  }
  int grid = (int)(p.total_tiles < ctx->sm_count ? p.total_tiles : ctx->sm_count);
  const bool prof = ctx->profile && ctx->prof_n < PMX_PROF_MAX;
  if (prof) PMX_CUDA(cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n], ctx->stream));
  k_grad_umma<<<grid, NUM_THREADS, SMEM_BYTES, ctx->stream>>>(pl->tmY, pl->tmAhi, pl->tmAlo, pl->tmShi, pl->tmSlo, p);
  if (prof) {
    PMX_CUDA(cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n + 1], ctx->stream));
    ctx->prof_n++;
  }
  PMX_LAUNCHED(ctx);
  if (p.trace && getenv("PMX_TRACE_DUMP")) {  // debug: print the timeline of this launch
    static int dumped = 0;
    if (dumped++ == 3) {
      long long h[5 * TRACE_TILES * TRACE_EVENTS];
      cudaStreamSynchronize(ctx->stream);
      cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
      long long t0 = h[(0 * TRACE_TILES + 4) * TRACE_EVENTS + 0];
      const char* roles[5] = {"producer", "issuer1", "issuer2", "residual", "flush"};
      for (int t = 4; t < TRACE_TILES; ++t)
        for (int r = 0; r < 5; ++r) {
          printf("TRACE tile %2d %-9s", t, roles[r]);
          for (int e = 0; e < TRACE_EVENTS; ++e) {
            long long v = h[(r * TRACE_TILES + t) * TRACE_EVENTS + e];
            printf(" %7lld", v ? v - t0 : -1);
          }
          printf("\n");
        }
      fflush(stdout);
    }
  }
  return pmx_check_launch(ctx, "k_grad_umma");
}
