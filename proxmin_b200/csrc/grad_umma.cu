// Fused NMF gradient kernel for sm_100a: one pass over Y computes
//     R = A S - Y,   G_A = R S^T,   G_S = A^T R,   loss = |R|^2 / 2          (nmf.py:13-41)
// with tcgen05 tensor-core MMAs (TMEM accumulators), TMA loads, and warp-specialised roles.
//
// Precision: every GEMM is the 3-term BF16 split  hi*hi + hi*lo + lo*hi  with fp32 accumulation
// (x = hi + lo, |x - hi - lo| <= 2^-17 |x|).  SURVEY 7.3 measured that single-pass TF32/BF16
// operands miss the 1e-4 parity target by 10-100x while this split clears it with margin.
// A and S arrive pre-split (k_split_bf16); R is split in the epilogue registers.
//
// Tiling: a tile is 128 rows (m) x 128 columns (n) of Y.  K <= 64: operands zero-padded to 64 (KH = 1).
// 64 < K <= 128 (KH = 2, BASELINE config 5): operands zero-padded to 128 and handled as two 64-deep k-halves -- the
// residual GEMM accumulates both halves into the same accumulator, the gradient GEMMs run per half from the same
// R^T (see "K > 64" below).
// The tile sequence is m-block major (all 128-column stripes of a 128-row block are consecutive) and
// is cut into gridDim.x contiguous, equally long ranges -- one persistent CTA per SM -- so the load is
// balanced to within one tile.
//
// Orientation: the residual is computed TRANSPOSED, acc^T[n, m] = S^T A^T, so that the 128 TMEM lanes run along n,
// the direction in which Y, S and G_S are contiguous in global memory:
//   * Y never touches shared memory: a residual warp reads Y[m, n0 + lane] with plain coalesced loads (one 128-byte
//     line per instruction, a whole tile period in flight) while the producer warp pulls the next tiles into L2;
//   * R^T (bf16 hi/lo) is written back over the accumulator columns it came from and is the TENSOR-MEMORY A operand
//     of the G_S GEMM (G_S^T[n, k] = R^T A), flushed per tile with red.add, lanes along n (coalesced lines);
//   * the shared-memory copy of R^T is the MN-major A operand of the G_A GEMM (G_A[m, k] = R S^T), which accumulates
//     in TMEM across the stripes of a row segment and is flushed once.
// Shared-memory traffic per tile is 352 KB (S tile 32 KB in, R^T 64 KB, MMA operand reads 256 KB) against 512 KB of
// the previous orientation (Y staged through shared memory, both operands of the G_S GEMM read from it): the
// 128 B/clk shared-memory pipe, not the tensor pipe or HBM, was the limiter.
//
// Shared memory (all operands are "panels": rows of 128 bytes, 128B-swizzled in 8-row atoms, which
// the same bytes can be read as a K-major or an MN-major UMMA operand):
//   S_hi,S_lo  [2 n-panels][64 k-rows][128B] x 3 slots  96 KB   per tile (per k-half)   (TMA)
//   A_hi,A_lo  [128 m-rows][128B] x KH            32 KB x KH   per row segment          (TMA)
//   R_hi,R_lo  [2 m-panels][128 n-rows][128B]     64 KB   per tile          (residual warps write)
// TMEM (512 columns): residual accumulator 2 x 128; KH = 1: G_S^T accumulator 2 x 64, G_A accumulator [hh|hl] 128;
// KH = 2: G_S^T accumulator 128 (single buffer), G_A accumulator 2 x 64 (one per k-half, three terms summed in place).
//
// Warp roles (640 threads): warp 0 = TMA producer, warps 1, 2, 3 = MMA issuers (residual / G_S / G_A; warp 2
// also allocates the tensor memory), warps 4..19 = sixteen residual + flush warps (4 TMEM lane quarters x 4 column
// chunks): TMEM accumulator -> R^T in TMEM and SMEM, then the red.add flush of the previous tile's G_S^T.
//
// K > 64 (KH = 2).  Shared memory cannot hold two-deep S tiles of 128 k-rows next to the 64 KB A tile and the 64 KB
// R^T copy, so the S ring holds k-HALF tiles (same 32 KB slot layout) and every half is loaded twice per tile: once
// for the residual GEMM (released by issuer 1) and once, a tile later, for the G_A GEMM (released by issuer 3).  The
// ring is one FIFO over both consumers:  R(0,0) R(0,1) | R(t+1,0) R(t+1,1) G(t,0) G(t,1) | ...   (R = residual use,
// G = G_A use, second index = k-half); every consumer derives slot and phase from the use index.  G_S^T[n, 0..127]
// takes the A tile of both halves as ONE N = 128 MN-major B operand (the halves are 16 KB apart = the LBO).
#include <cuda_bf16.h>
#include <type_traits>
#include <stdlib.h>

#include "grad_umma.h"
#include "kernels.h"

namespace {

#ifndef PMX_REG_DEC
#define PMX_REG_DEC 32     // registers per thread of the producer / issuer warps after setmaxnreg.dec
#endif
#ifndef PMX_REG_INC
#define PMX_REG_INC 112    // registers per thread of the residual warps after setmaxnreg.inc
#endif
#define PMX_STR2(x) #x
#define PMX_STR(x) PMX_STR2(x)
#ifndef PMX_WAIT_HINT_NS
#define PMX_WAIT_HINT_NS 0   // suspend-time hint (ns) of the mbarrier waits; 0 = plain try_wait loop.  Round 2: every
                            // hint (100 ns .. 10 ms) costs 6 % -- the wake-up of a suspended warp sits in the tile chain
#endif
#ifndef PMX_Y_REFILL_EARLY
#define PMX_Y_REFILL_EARLY 0   // 1: refill a Y buffer right after the conversion half that consumed it (round-2 experiment)
#endif
#ifndef PMX_GS_STACK
#define PMX_GS_STACK 0
#endif
constexpr int TILE_M = 128, TILE_N = 128, KP = 64;
constexpr int S_SLOTS = 3;          // S-tile ring
constexpr uint32_t PANEL_S = 64 * 128;    // bytes of one S panel (64 k-rows)
constexpr uint32_t PANEL_R = 128 * 128;   // bytes of one R^T / A panel (128 rows)

// shared-memory map (offsets from the 1024-aligned base)
constexpr uint32_t OFF_S = 0;                               // slot s at OFF_S + s*4*PANEL_S (see S slot layout below)
constexpr uint32_t OFF_A = OFF_S + S_SLOTS * 4 * PANEL_S;   // A_hi[kh] at OFF_A + kh*PANEL_R, A_lo[kh] at OFF_A + (KH + kh)*PANEL_R
template <int KH> struct SmemMap {
  static constexpr uint32_t R_HI = OFF_A + KH * 2 * PANEL_R;
  static constexpr uint32_t R_LO = R_HI + 2 * PANEL_R;
  static constexpr uint32_t BAR = R_LO + 2 * PANEL_R;
  static constexpr uint32_t BYTES = BAR + 512 + 1024;       // barriers + alignment slack
};
static_assert(SmemMap<2>::BYTES <= 232448, "K = 128 configuration exceeds the 227 KB of shared memory per CTA");

enum {  // mbarrier indices
  B_A_FULL = 0, B_A_EMPTY, B_S_FULL, B_S_EMPTY = B_S_FULL + S_SLOTS, B_ACC_FULL = B_S_EMPTY + S_SLOTS,
  B_ACC_EMPTY = B_ACC_FULL + 2, B_RT_FULL = B_ACC_EMPTY + 2, B_RS_FULL = B_RT_FULL + 2, B_RS_EMPTY, B_GS_FULL,
  B_GS_EMPTY = B_GS_FULL + 2, B_GA_FULL = B_GS_EMPTY + 2, B_GA_EMPTY, B_COUNT
};

// TMEM column map (512): residual accumulator / R^T operand 2 x 128, G_S^T 2 x 64 (double-buffered so that the
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
// (round 1 passed a suspend-time hint so that parked warps do not burn issue slots; with the leaner round-2 instruction
// stream every hint value from 100 ns to 10 ms measured 6 % SLOWER than the plain try_wait loop, and non-blocking
// test_wait spin loops for the issuers / the accumulator wait changed nothing: profiles/README.md)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if PMX_WAIT_HINT_NS > 0
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar),
      "r"(parity), "r"((uint32_t)PMX_WAIT_HINT_NS)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
#endif
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
__device__ __forceinline__ uint64_t l2_policy_normal() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// streaming fp32 load (read once): no L1 allocation, L2 eviction priority from `policy`
__device__ __forceinline__ float ld_stream(const float* ptr, uint64_t policy) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ void red_add_f32(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
// two fp32 -> packed bf16 pair (round to nearest even): `lo` in bits 0..15, `hi` in bits 16..31
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// Strided access sequences (the Y loads, the G_S reductions) keep the address as a (lo, hi) register pair and step
// only the low word -- one integer instruction per access instead of a 64-bit multiply-add; the caller checks
// that the low word cannot carry over the whole sequence.
__device__ __forceinline__ float ld_stream_lohi(uint32_t lo, uint32_t hi, uint64_t policy) {
  float v;
  asm volatile(
      "{\n\t.reg .b64 a;\n\t"
      "mov.b64 a, {%1, %2};\n\t"
      "ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [a], %3;\n\t}"
      : "=f"(v)
      : "r"(lo), "r"(hi), "l"(policy));
  return v;
}
__device__ __forceinline__ float4 ld_stream_v4(const float* ptr, uint64_t policy) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(ptr), "l"(policy));
  return v;
}
__device__ __forceinline__ void red_add_f32_lohi(uint32_t lo, uint32_t hi, float v) {
  asm volatile(
      "{\n\t.reg .b64 a;\n\t"
      "mov.b64 a, {%0, %1};\n\t"
      "red.global.add.f32 [a], %2;\n\t}" ::"r"(lo),
      "r"(hi), "f"(v)
      : "memory");
}
template <int J0, int J1, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (J0 < J1) {
    f(std::integral_constant<int, J0>{});
    static_for<J0 + 1, J1>(f);
  }
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
  const float* Y;         // fp32, row pitch ldY floats
  const float* W;         // weights (WGT instantiations only): same layout as Y, tiled
  int ldY;
  float* GA;
  const unsigned* ga_epoch;   // parity counter: the gradients go to buffer (*ga_epoch + 1) & 1 of a pair (sharded runs: the
                              // peer epoch of comm.cu; fused PGM tail: the iteration counter, see pgm_tail.cu)
  long long ga_stride;        // elements between the two G_A buffers
  long long gs_stride;        // elements between the two G_S buffers (0: G_S is a single buffer)
  float* GS;
  double* loss;
  const int* done;
  int y_blocked;          // layout of Y: 0 row-major (pitch ldY), 1 tiled (see grad_umma.h)
  int want_ga, want_gs;   // 0: that gradient GEMM and its flush are skipped (bsdmm needs one gradient per pass, the loss none)
  long long* trace;       // debug: clock64 timeline of CTA 0 (env PMX_TRACE), [role][tile][event]
  int ablate;             // debug/timing only (env PMX_ABLATE): bit0 no MMA1, 1 no MMA2, 2 no MMA3, 3 no Y read, 4 no R store, 5 no G_S flush, 7 no L2 hint on the Y loads
};

// timeline tracing (debug instantiation only): role r, local tile index t, event e
#define TRACE_TILES 24
#define TRACE_EVENTS 8
#define TR(r, t, e)                                                                                   \
  do {                                                                                                \
    if (DBG && p.trace && blockIdx.x == 0 && lane == 0 && (t) < TRACE_TILES)                          \
      p.trace[((r) * TRACE_TILES + (t)) * TRACE_EVENTS + (e)] = clock64();                            \
  } while (0)
#define ABL(bit) (DBG && (p.ablate & (bit)))

constexpr int NUM_EPI_WARPS = 16;    // residual + flush warps: 4 TMEM lane quarters x 4 column chunks
constexpr int NUM_THREADS = 128 + 32 * NUM_EPI_WARPS;

// position in the m-block-major tile sequence, advanced without divisions
struct TilePos {
  int mb, st;
  __device__ __forceinline__ void next(int NS) {
    if (++st == NS) {
      st = 0;
      ++mb;
    }
  }
};

// KH = 2: position of a use in the S-ring FIFO  R(0,0) R(0,1) | R(t+1,0) R(t+1,1) G(t,0) G(t,1) | ...
// (without the G_A GEMM only the R uses exist)
__device__ __forceinline__ uint32_t ring_use_R(int t, int kh, bool with_g) {
  if (!with_g) return 2u * t + kh;
  return t == 0 ? (uint32_t)kh : (uint32_t)(4 * t - 2 + kh);
}
__device__ __forceinline__ uint32_t ring_use_G(int t, int kh, int ntiles) {
  return (uint32_t)((t == ntiles - 1 ? 4 * t + 2 : 4 * t + 4) + kh);
}

// KH: number of 64-deep k-halves (1: K <= 64, 2: K <= 128)
// LOSS: accumulate |R|^2 / 2 (nmf.py:25); DBG: timing ablations and the clock64 trace (never used for results)
// WGT: weighted likelihood, D = W (A S - Y) (nmf.py:25, 40) with a second M x N stream W in the same tiled layout
template <int KH, bool LOSS, bool DBG, bool WGT = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_grad_umma(const __grid_constant__ CUtensorMap tmAhi,
            const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmShi,
            const __grid_constant__ CUtensorMap tmSlo, const Params p) {
  if (p.done && *p.done) return;
  constexpr uint32_t OFF_R_HI = SmemMap<KH>::R_HI, OFF_R_LO = SmemMap<KH>::R_LO, OFF_BAR = SmemMap<KH>::BAR;
  constexpr int KPT = KP * KH;   // padded K
  constexpr bool GS1 = (KH == 2) || PMX_GS_STACK;   // single-buffered 128-column G_S^T accumulator
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

  // tile sequence: m-block major (all 128-column stripes of one 128-row block are consecutive); the only divisions of
  // the kernel are these (the loops advance (mb, st) incrementally)
  const int NS = p.NS;
  const long long g_begin = (long long)blockIdx.x * p.total_tiles / gridDim.x;
  const int ntiles = (int)((long long)(blockIdx.x + 1) * p.total_tiles / gridDim.x - g_begin);
  TilePos pos0;
  pos0.mb = (int)(g_begin / NS);
  pos0.st = (int)(g_begin - (long long)pos0.mb * NS);
  const bool want_ga = p.want_ga != 0, want_gs = p.want_gs != 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < B_COUNT; ++i) {
      uint32_t count = 1;
      if (i == B_A_EMPTY) count = 2;                                              // issuers 1 and 2 release A
      if (i == B_ACC_EMPTY || i == B_ACC_EMPTY + 1) count = NUM_EPI_WARPS + 1;     // residual warps + issuer 2
      if (i == B_RT_FULL || i == B_RT_FULL + 1) count = NUM_EPI_WARPS;
      if (i == B_RS_FULL) count = NUM_EPI_WARPS;
      if (i == B_GS_EMPTY || i == B_GS_EMPTY + 1) count = NUM_EPI_WARPS;
      if (i == B_GA_EMPTY) count = NUM_EPI_WARPS;
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

  // a row segment = the consecutive tiles of one m-block inside this CTA's range
  auto first_in_seg = [&](int t, const TilePos& q) { return t == 0 || q.st == 0; };
  auto last_in_seg = [&](int t, const TilePos& q) { return t == ntiles - 1 || q.st == NS - 1; };
  // S slot layout: n-panel pn (64 columns) = [S_hi 64 k-rows | S_lo 64 k-rows] -> 16 KB per panel, so that
  // [S_hi; S_lo] is one 128-row K-major operand (the "stacked" N = 128 operand of the G_A GEMM, KH = 1)
  constexpr uint32_t S_SLOT = 4 * PANEL_S, S_PANEL = 2 * PANEL_S;
  // A tile: hi half kh at a_hi(kh), lo half at a_lo(kh); the two k-halves of one precision are PANEL_R apart
  auto a_hi = [&](int kh) { return base + OFF_A + (uint32_t)kh * PANEL_R; };
  auto a_lo = [&](int kh) { return base + OFF_A + (uint32_t)(KH + kh) * PANEL_R; };

  // Register split (setmaxnreg, per warpgroup of 4 warps): the producer and the three issuer warps (warpgroup 0) need
  // few registers, the residual warps hold three half-tile Y buffers + R^T + the accumulator chunk.  The kernel is
  // compiled for 96 registers per thread (640 threads); warpgroup 0 drops to 32, warpgroups 1-4 rise to 112.  The
  // increase is served ONLY from what the decrease released (128 x 64 = 8192 = 512 x 16 registers): a larger request
  // blocks forever (launch_grad_umma checks the compiled register count against this budget).
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 " PMX_STR(PMX_REG_DEC) ";");
  if (warp == 0) {
    // ============================== TMA producer ==============================
    // S tiles (per tile) and A tiles (per row segment) into shared memory; Y is NOT staged in shared memory: the
    // residual warps read it from global memory (coalesced along n).
    // (explicit L2 prefetches of Y -- TMA prefetch boxes with and without an evict_first hint, per-lane
    // prefetch.global.L2, 1 to 8 tiles ahead -- were measured in both rounds and only slowed the kernel down,
    // monotonically with the distance: profiles/README.md)
    uint32_t seg = 0;
    const uint64_t pol_keep = l2_policy_evict_last();
    uint32_t slot = 0, use = 0;
    // one S-ring slot: the 64-deep k-half kh of the S tile at column n0
    auto load_s = [&](int n0, int kh) {
      mbar_wait(bar(B_S_EMPTY + slot), (use & 1) ^ 1);
      if (lane == 0) {
        mbar_expect_tx(bar(B_S_FULL + slot), 4 * PANEL_S);
        const uint32_t sb = base + OFF_S + slot * S_SLOT;
        tma_load_2d_hint(sb, &tmShi, n0, kh * KP, bar(B_S_FULL + slot), pol_keep);
        tma_load_2d_hint(sb + PANEL_S, &tmSlo, n0, kh * KP, bar(B_S_FULL + slot), pol_keep);
        tma_load_2d_hint(sb + S_PANEL, &tmShi, n0 + 64, kh * KP, bar(B_S_FULL + slot), pol_keep);
        tma_load_2d_hint(sb + S_PANEL + PANEL_S, &tmSlo, n0 + 64, kh * KP, bar(B_S_FULL + slot), pol_keep);
      }
      __syncwarp();
      if (++slot == S_SLOTS) {
        slot = 0;
        ++use;
      }
    };
    auto load_a = [&](int m0) {
      mbar_wait(bar(B_A_EMPTY), (seg & 1) ^ 1);
      if (lane == 0) {
        mbar_expect_tx(bar(B_A_FULL), KH * 2 * PANEL_R);
#pragma unroll
        for (int kh = 0; kh < KH; ++kh) {
          tma_load_2d(a_hi(kh), &tmAhi, kh * KP, m0, bar(B_A_FULL));
          tma_load_2d(a_lo(kh), &tmAlo, kh * KP, m0, bar(B_A_FULL));
        }
      }
      ++seg;
      __syncwarp();
    };
    if constexpr (KH == 1) {
      TilePos pos = pos0;
      for (int t = 0; t < ntiles; ++t, pos.next(NS)) {
        load_s(pos.st * TILE_N, 0);
        TR(0, t, 0);
        if (first_in_seg(t, pos)) load_a(pos.mb * TILE_M);
      }
    } else {
      // FIFO  R(0,0) R(0,1) | R(t+1,0) R(t+1,1) G(t,0) G(t,1) | ...  (see the file header); the A tile of a row segment
      // is requested right after the first residual half of its first tile
      TilePos pos = pos0, nxt = pos0;
      if (ntiles > 0) {
        load_s(pos.st * TILE_N, 0);
        load_a(pos.mb * TILE_M);
        load_s(pos.st * TILE_N, 1);
      }
      for (int t = 0; t < ntiles; ++t) {
        nxt.next(NS);
        if (t + 1 < ntiles) {
          load_s(nxt.st * TILE_N, 0);
          if (first_in_seg(t + 1, nxt)) load_a(nxt.mb * TILE_M);
          load_s(nxt.st * TILE_N, 1);
        }
        if (want_ga) {
          load_s(pos.st * TILE_N, 0);
          load_s(pos.st * TILE_N, 1);
        }
        pos = nxt;
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer 1: residual GEMM (transposed) ==============================
    // acc^T[n, m] = S^T A^T: the S tile is the (MN-major) A operand, the A tile the (K-major) B operand, so that the
    // accumulator lanes run along n -- the direction in which Y, S and G_S are contiguous in global memory.
    // Three issuer warps (residual / G_S / G_A), each blocking only on its own dependencies; all 32 lanes run the
    // loop and one elected lane issues.
    constexpr uint32_t ID_RES = make_idesc(128, 128, 1, 0);
    uint32_t seg_full = 0, sslot = 0, suse = 0;
    TilePos pos = pos0;
    for (int t = 0; t < ntiles; ++t, pos.next(NS)) {
      const uint32_t slot = t & 1;
      const uint32_t d = tmem + TM_ACC + slot * 128;
#pragma unroll
      for (int kh = 0; kh < KH; ++kh) {
        if constexpr (KH == 2) {
          const uint32_t u = ring_use_R(t, kh, want_ga);
          sslot = u % S_SLOTS;
          suse = u / S_SLOTS;
        }
        mbar_wait(bar(B_S_FULL + sslot), suse & 1);
        if (kh == 0) {
          if (first_in_seg(t, pos)) {
            mbar_wait(bar(B_A_FULL), seg_full & 1);
            ++seg_full;
          }
          mbar_wait(bar(B_ACC_EMPTY + slot), ((t >> 1) & 1) ^ 1);
        }
        tc_fence_after();
        if (kh == 0) TR(1, t, 0);
        const uint32_t sb = base + OFF_S + sslot * S_SLOT;
        // acc += S_hi^T A_hi^T + S_lo^T A_hi^T + S_hi^T A_lo^T      (64 k of this half: 4 k-steps of 16)
        const uint32_t s_src[3] = {sb, sb + PANEL_S, sb};
        const uint32_t a_src[3] = {a_hi(kh), a_hi(kh), a_lo(kh)};
#pragma unroll
        for (int term = 0; term < 3; ++term)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ABL(1)) continue;
            const uint64_t ad = make_desc(s_src[term] + ks * 2048, S_PANEL, 1024);      // MN-major: LBO = next 64 n
            const uint64_t bd = make_desc(a_src[term] + ks * 32, 16, 1024);             // K-major, 128 m-rows
            umma_ss_e(d, ad, bd, ID_RES, (kh | term | ks) ? 1u : 0u);
          }
        if constexpr (KH == 2) tc_commit_e(bar(B_S_EMPTY + sslot));   // this half-tile is reloaded for the G_A GEMM
      }
      tc_commit_e(bar(B_ACC_FULL + slot));
      TR(1, t, 1);
      if (last_in_seg(t, pos)) tc_commit_e(bar(B_A_EMPTY));   // residual GEMMs of this segment are done with the A tile
      if constexpr (KH == 1) {
        if (++sslot == S_SLOTS) {
          sslot = 0;
          ++suse;
        }
      }
    }
  } else if (warp == 2) {
    // ============================== MMA issuer 2: G_S GEMM (after the TMEM allocation above) ================
    // G_S^T[n, k] = R_hi^T A_hi + R_hi^T A_lo + R_lo^T A_hi   (M = n, N = k, K = m: 8 k-steps).  R^T (bf16 hi/lo)
    // sits in the columns of the residual accumulator it was computed from: every 16 accumulator columns become
    // [hi 8 cols | lo 8 cols] of packed bf16 pairs -> the A operand comes from tensor memory, only the A tile
    // (MN-major B operand) is read from shared memory.  KH = 1: N = 64, double-buffered accumulator; KH = 2: the two
    // k-halves of the A tile are one N = 128 operand (LBO = distance of the halves), single accumulator.
    constexpr uint32_t ID_GS = make_idesc(128, 64 * KH, 0, 1);
    constexpr uint32_t ID_GS_STACK = make_idesc(128, 128, 0, 1);
    uint32_t seg_full = 0;
    TilePos pos = pos0;
    for (int t = 0; t < ntiles; ++t, pos.next(NS)) {
      const uint32_t slot = t & 1;
      const uint32_t gslot = GS1 ? 0u : slot, guse = GS1 ? (uint32_t)t : (uint32_t)(t >> 1);
      if (first_in_seg(t, pos)) {
        mbar_wait(bar(B_A_FULL), seg_full & 1);          // already complete (MMA1 consumed it): visibility only
        ++seg_full;
      }
      mbar_wait(bar(B_RT_FULL + slot), (t >> 1) & 1);
      mbar_wait(bar(B_GS_EMPTY + gslot), (guse & 1) ^ 1);
      tc_fence_after();
      TR(1, t, 2);
      const uint32_t d = tmem + TM_GS + gslot * 64;
      const uint32_t racc = tmem + TM_ACC + slot * 128;
      const uint32_t r_off[3] = {0, 0, 8};               // R_hi, R_hi, R_lo pairs of a k-step's 16 columns
      const uint32_t a_src[3] = {a_hi(0), a_lo(0), a_hi(0)};
      if (want_gs) {
        if constexpr (KH == 1 && PMX_GS_STACK) {
          // G_S^T[n, (hh | hl)] = R_hi^T [A_hi | A_lo]  (one N = 128 instruction per k-step: the hi and lo panels of the
          // A tile are PANEL_R apart = the LBO), then G_S^T[n, hh] += R_lo^T A_hi (N = 64); the flush adds the halves
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            if (ABL(2)) continue;
            const uint64_t bd = make_desc(a_hi(0) + ks * 2048, PANEL_R, 1024);
            umma_ts_e(d, racc + ks * 16, bd, ID_GS_STACK, ks ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            if (ABL(2)) continue;
            const uint64_t bd = make_desc(a_hi(0) + ks * 2048, PANEL_R, 1024);
            umma_ts_e(d, racc + ks * 16 + 8, bd, ID_GS, 1u);
          }
        } else {
#pragma unroll
          for (int term = 0; term < 3; ++term)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              if (ABL(2)) continue;
              const uint32_t at = racc + ks * 16 + r_off[term];
              const uint64_t bd = make_desc(a_src[term] + ks * 2048, PANEL_R, 1024);  // MN-major, 64-wide atoms PANEL_R apart
              umma_ts_e(d, at, bd, ID_GS, (term | ks) ? 1u : 0u);
            }
        }
      }
      tc_commit_e(bar(B_GS_FULL + gslot));
      tc_commit_e(bar(B_ACC_EMPTY + slot));  // MMA1(t+2) may overwrite these columns once R^T(t) has been consumed
      TR(1, t, 3);
      if (last_in_seg(t, pos)) tc_commit_e(bar(B_A_EMPTY));
    }
  } else {
    // ============================== MMA issuer 3: G_A GEMM ==============================
    // KH = 1: G_A[m, (hh|hl)] += R_hi [S_hi;S_lo]^T ; G_A[m, hh] += R_lo S_hi^T   (M = m, K = n: 8 k-steps).  R comes
    // from the shared-memory copy of R^T (MN-major A operand), the stacked S slot is the K-major B operand.
    // KH = 2: per k-half, G_A[m, half] += R_hi S_hi^T + R_hi S_lo^T + R_lo S_hi^T (N = 64 each, summed in place).
    constexpr uint32_t ID_GA2 = make_idesc(128, 128, 1, 0);
    constexpr uint32_t ID_GA1 = make_idesc(128, 64, 1, 0);
    const uint32_t r_hi = base + OFF_R_HI, r_lo = base + OFF_R_LO;
    uint32_t seg = 0, sslot = 0, suse = 0;
    TilePos pos = pos0;
    for (int t = 0; t < ntiles; ++t, pos.next(NS)) {
      const bool first = first_in_seg(t, pos);
      if constexpr (KH == 1) {
        mbar_wait(bar(B_S_FULL + sslot), suse & 1);    // already complete (MMA1(t) consumed it): visibility only
        mbar_wait(bar(B_RS_FULL), t & 1);
        if (first) mbar_wait(bar(B_GA_EMPTY), (seg & 1) ^ 1);
        tc_fence_after();
        TR(2, t, 0);
        const uint32_t sb = base + OFF_S + sslot * S_SLOT;
        const uint32_t d = tmem + TM_GA;
        if (want_ga) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            if (ABL(4)) continue;
            const uint64_t ad = make_desc(r_hi + ks * 2048, PANEL_R, 1024);                        // MN-major: LBO = next 64 m
            const uint64_t bd = make_desc(sb + (ks >> 2) * S_PANEL + (ks & 3) * 32, 16, 1024);     // K-major, 128 rows
            umma_ss_e(d, ad, bd, ID_GA2, (!first || ks) ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            if (ABL(4)) continue;
            const uint64_t ad = make_desc(r_lo + ks * 2048, PANEL_R, 1024);
            const uint64_t bd = make_desc(sb + (ks >> 2) * S_PANEL + (ks & 3) * 32, 16, 1024);     // K-major, rows 0..63
            umma_ss_e(d, ad, bd, ID_GA1, 1u);
          }
        }
        tc_commit_e(bar(B_RS_EMPTY));
        tc_commit_e(bar(B_S_EMPTY + sslot));   // S(t) was last used here
        if (++sslot == S_SLOTS) {
          sslot = 0;
          ++suse;
        }
      } else {
        mbar_wait(bar(B_RS_FULL), t & 1);
        if (first) mbar_wait(bar(B_GA_EMPTY), (seg & 1) ^ 1);
        if (want_ga) {
#pragma unroll
          for (int kh = 0; kh < KH; ++kh) {
            const uint32_t u = ring_use_G(t, kh, ntiles);
            sslot = u % S_SLOTS;
            suse = u / S_SLOTS;
            mbar_wait(bar(B_S_FULL + sslot), suse & 1);
            tc_fence_after();
            if (kh == 0) TR(2, t, 0);
            const uint32_t sb = base + OFF_S + sslot * S_SLOT;
            const uint32_t d = tmem + TM_GA + kh * 64;
            const uint32_t r_src[3] = {r_hi, r_hi, r_lo};
            const uint32_t s_off[3] = {0, PANEL_S, 0};         // S_hi rows, S_lo rows, S_hi rows of the n-panel
#pragma unroll
            for (int term = 0; term < 3; ++term)
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                if (ABL(4)) continue;
                const uint64_t ad = make_desc(r_src[term] + ks * 2048, PANEL_R, 1024);
                const uint64_t bd = make_desc(sb + (ks >> 2) * S_PANEL + s_off[term] + (ks & 3) * 32, 16, 1024);   // K-major, 64 k-rows
                umma_ss_e(d, ad, bd, ID_GA1, (!first || term || ks) ? 1u : 0u);
              }
            tc_commit_e(bar(B_S_EMPTY + sslot));
          }
        } else {
          tc_fence_after();
        }
        tc_commit_e(bar(B_RS_EMPTY));
      }
      TR(2, t, 1);
      if (last_in_seg(t, pos)) {
        tc_commit_e(bar(B_GA_FULL));
        ++seg;
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 " PMX_STR(PMX_REG_INC) ";");
    // ============================== residual + flush warps (16) ==============================
    // Warp (q4, grp): TMEM lane quarter q4 (the 32 columns n = 32 q4 + lane of the tile), 32-row chunk grp of the
    // accumulator columns.  Per tile: residual chunk -> R^T (TMEM, registers -> SMEM), then the flush of the previous
    // tile's G_S^T (a quarter of its k-columns) while the tensor pipe works on this one.
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may access
    const int grp = (warp - 4) >> 2;         // m-chunk of the residual / k-quarter of the flushes
    const int row = q4 * 32 + lane;          // column n of the tile = accumulator lane (row m for the G_A flush)
    const uint32_t lane_addr = tmem + ((uint32_t)(q4 * 32) << 16);
    const uint64_t pol = ABL(128) ? l2_policy_normal() : l2_policy_evict_first();
    const int K = p.K, N = p.N, M = p.M, ldY = p.ldY;
    constexpr int KQ = 16 * KH;              // gradient columns (k) this warp flushes: [grp * KQ, grp * KQ + KQ)
    const unsigned parity = p.ga_epoch ? ((*p.ga_epoch + 1u) & 1u) : 0u;
    float* const GA = p.GA + (size_t)parity * p.ga_stride;
    float* const GSb = p.GS + (size_t)parity * p.gs_stride;
    // shared-memory destinations of this thread's R^T chunk (row n, 128B-swizzled 16-byte chunks): loop invariant
    uint8_t* const rh = base_ptr + OFF_R_HI + (grp >> 1) * PANEL_R + row * 128;
    uint8_t* const rl = base_ptr + OFF_R_LO + (grp >> 1) * PANEL_R + row * 128;
    float loss_part = 0.f;
    // Y in registers: three buffers of 16 rows (half of this warp's 32-row chunk of a tile) that rotate so that the
    // loads run one to two tile periods ahead of their use -- tile t reads its first half from buffer (2t) % 3 and its
    // second half from buffer (2t + 1) % 3; once both are consumed they are refilled with the second half of tile
    // t + 1 and the first half of tile t + 2.  With a single 32-row buffer the loads of tile t + 1 could only be issued
    // after the conversion of tile t and were needed right after it: DRAM latency (~700 cycles) and the transfer time
    // of 64 KB through the SM's L2 port (~800 cycles) sat in the residual warps' serial chain of every tile
    // (profiles/README.md, round-2 ablations).  The buffer index must be a compile-time constant (registers), hence
    // the tile loop below is unrolled three times.
    float y[48];
    auto issue_y_half = [&](auto BUF, const TilePos& q, int h, const float* Ybase) {
      constexpr int B0 = 16 * decltype(BUF)::value;
      if (ABL(8)) {
#pragma unroll
        for (int j = 0; j < 16; ++j) y[B0 + j] = 0.f;
        return;
      }
      if (p.y_blocked) {
        // tiled Y (grad_umma.h): the 128 x 128 tile is one contiguous 64 KB block [32 row quads][128 columns][4 rows], so
        // one 16-byte load per lane brings 4 consecutive rows of column n, a warp instruction covers 512 contiguous
        // bytes, a CTA streams its tile range as one contiguous region (the tile sequence is the storage order) and
        // there are no edge cases: the padding of the device copy is zero
        // (timing ablation 64: every tile reads tile (0, 0): same instructions, L2 hits instead of DRAM)
        const float* src4 = Ybase + ((size_t)(ABL(64) ? 0 : q.mb * NS + q.st) * (TILE_M * TILE_N) +
                                   (size_t)(grp * 8 + h * 4) * (TILE_N * 4) + row * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = ld_stream_v4(src4 + i * (TILE_N * 4), pol);
          y[B0 + 4 * i] = v.x; y[B0 + 4 * i + 1] = v.y; y[B0 + 4 * i + 2] = v.z; y[B0 + 4 * i + 3] = v.w;
        }
        return;
      }
      // row-major Y (pmx_nmf_grad on a caller's matrix): for a fixed row the 32 lanes read one 128-byte line
      const int m0 = q.mb * TILE_M + grp * 32 + h * 16, n0 = q.st * TILE_N;
      const float* src = Ybase + (size_t)m0 * ldY + (n0 + row);
      const uint32_t pitch = (uint32_t)ldY * 4u;
      const uint64_t a0 = reinterpret_cast<uint64_t>(src);
      uint32_t lo = (uint32_t)a0;
      const uint32_t hi = (uint32_t)(a0 >> 32);
      const bool interior = (m0 + 16 <= M) && (n0 + TILE_N <= N);
      if (interior && pitch <= (1u << 26) && __all_sync(0xffffffffu, lo <= 0xffffffffu - 16u * pitch)) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          y[B0 + j] = ld_stream_lohi(lo, hi, pol);
          lo += pitch;
        }
      } else {
        const bool col_ok = n0 + row < N;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          y[B0 + j] = 0.f;
          if (col_ok && m0 + j < M) y[B0 + j] = ld_stream(src + (size_t)j * ldY, pol);
        }
      }
    };
    // G_S^T[n, KQ grp .. KQ grp + KQ - 1] of the tile at q (tile index tt) -> red.add into G_S[k, n]: for a fixed k the
    // 32 lanes hit one 128-byte line
    auto flush_gs = [&](const TilePos& q, uint32_t tt) {
      const uint32_t gslot = GS1 ? 0u : (tt & 1), guse = GS1 ? tt : (tt >> 1);
      mbar_wait(bar(B_GS_FULL + gslot), guse & 1);
      if (!want_gs) {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_GS_EMPTY + gslot));
        return;
      }
      tc_fence_after();
      uint32_t v[KQ];
      if constexpr (KH == 1 && PMX_GS_STACK) {
        tmem_ld16(lane_addr + TM_GS + grp * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
#pragma unroll
        for (int h = 0; h < 2; ++h) {   // the hl half in two steps of 8 columns (register pressure: R^T is still live)
          uint32_t w[8];
          tmem_ld8(lane_addr + TM_GS + 64 + grp * 16 + h * 8, w);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) v[h * 8 + k] = __float_as_uint(__uint_as_float(v[h * 8 + k]) + __uint_as_float(w[k]));
        }
      } else {
#pragma unroll
        for (int h = 0; h < KH; ++h) tmem_ld16(lane_addr + TM_GS + gslot * 64 + grp * KQ + h * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[h * 16]));
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_GS_EMPTY + gslot));   // values are in registers: the accumulator is free
      if (ABL(32)) return;
      const int n = q.st * TILE_N + row;
      float* dst = GSb + (size_t)(grp * KQ) * N + n;
      const uint32_t pitch = (uint32_t)N * 4u;
      const uint64_t a0 = reinterpret_cast<uint64_t>(dst);
      uint32_t lo = (uint32_t)a0;
      const uint32_t hi = (uint32_t)(a0 >> 32);
      if (K == KPT && (q.st + 1) * TILE_N <= N && pitch <= (1u << 26) &&
          __all_sync(0xffffffffu, lo <= 0xffffffffu - (uint32_t)KQ * pitch)) {
#pragma unroll
        for (int k = 0; k < KQ; ++k) {
          red_add_f32_lohi(lo, hi, __uint_as_float(v[k]));
          lo += pitch;
        }
      } else if (n < N) {
#pragma unroll
        for (int k = 0; k < KQ; ++k)
          if (grp * KQ + k < K) red_add_f32(dst + (size_t)k * N, __uint_as_float(v[k]));
      }
    };
    // G_A[m, KQ grp .. KQ grp + KQ - 1] of the row segment of m-block q.mb (KH = 1: hh + hl halves)
    auto flush_ga = [&](const TilePos& q, uint32_t seg) {
      mbar_wait(bar(B_GA_FULL), seg & 1);
      if (!want_ga) {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_GA_EMPTY));
        return;
      }
      tc_fence_after();
      float s[KQ];
      if constexpr (KH == 1) {
        uint32_t v[16], w[16];
        tmem_ld16(lane_addr + TM_GA + grp * 16, v);
        tmem_ld16(lane_addr + TM_GA + 64 + grp * 16, w);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) s[k] = __uint_as_float(v[k]) + __uint_as_float(w[k]);
      } else {
        uint32_t v[KQ];
#pragma unroll
        for (int h = 0; h < KH; ++h) tmem_ld16(lane_addr + TM_GA + grp * KQ + h * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[h * 16]));
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < KQ; ++k) s[k] = __uint_as_float(v[k]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_GA_EMPTY));
      const int m = q.mb * TILE_M + row;
      if (m < M) {
        float* dst = GA + (size_t)m * K + grp * KQ;
        if ((K & 3) == 0) {
#pragma unroll
          for (int k = 0; k < KQ; k += 4)
            if (grp * KQ + k < K) red_add_v4(dst + k, s[k], s[k + 1], s[k + 2], s[k + 3]);
        } else {
#pragma unroll
          for (int k = 0; k < KQ; ++k)
            if (grp * KQ + k < K) red_add_f32(dst + k, s[k]);
        }
      }
    };
    uint32_t seg = 0;
    TilePos pos = pos0, prev = pos0;
    bool prev_last = false;
    using IC0 = std::integral_constant<int, 0>;
    using IC1 = std::integral_constant<int, 1>;
    using IC2 = std::integral_constant<int, 2>;
    // WGT: no rotation -- buffers 0 and 1 hold the two Y halves of the coming tile, buffer 2 the first half of W; the
    // second half of W is fetched inside the conversion (its latency is exposed: the weighted pass streams twice the
    // bytes and is not the tuned path)
    if (ntiles > 0) {
      issue_y_half(IC0{}, pos0, 0, p.Y);
      issue_y_half(IC1{}, pos0, 1, p.Y);
      if constexpr (WGT) {
        issue_y_half(IC2{}, pos0, 0, p.W);
      } else if (ntiles > 1) {
        TilePos q1 = pos0;
        q1.next(NS);
        issue_y_half(IC2{}, q1, 0, p.Y);
      }
    }
    // one tile; ROT = t % 3 selects the register buffers: first half in buffer YA, second half in buffer YB
    auto tile_body = [&](auto ROT, const int t) {
      constexpr int YA = WGT ? 0 : (2 * decltype(ROT)::value) % 3, YB = WGT ? 1 : (2 * decltype(ROT)::value + 1) % 3;
      const uint32_t slot = t & 1;
      mbar_wait(bar(B_ACC_FULL + slot), (t >> 1) & 1);
      tc_fence_after();
      if (warp == 4) TR(3, t, 0);
      // ---- residual -> bf16 (hi, lo), written back over the accumulator columns it came from.  The accumulator is
      // read from tensor memory exactly once (TMEM reads run at ~64 B/clk per SM); the shared-memory copy of R^T is
      // written from the same registers.  hl layout per 16 columns: [hi 8 pairs | lo 8 pairs]
      uint32_t hl[32];
      TilePos nxt = pos;
      nxt.next(NS);
      TilePos nxt2 = nxt;
      nxt2.next(NS);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t acc[16];
        tmem_ld16(lane_addr + TM_ACC + slot * 128 + grp * 32 + h * 16, acc);
        tmem_ld_wait();
        if constexpr (WGT) {
          if (h == 1) issue_y_half(IC2{}, pos, 1, p.W);   // buffer 2 (first half of W) was consumed by h = 0
        }
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float r0 = __uint_as_float(acc[j]) - y[(h == 0 ? YA : YB) * 16 + j];          // nmf.py:40  (A S - Y)
          float r1 = __uint_as_float(acc[j + 1]) - y[(h == 0 ? YA : YB) * 16 + j + 1];
          if constexpr (WGT) {                                                            // D = W (A S - Y)
            const float d0 = y[32 + j] * r0, d1 = y[32 + j + 1] * r1;
            if (LOSS) {                                                                   // sum W (Y - A S)^2 / 2
              loss_part = fmaf(d0, r0, loss_part);
              loss_part = fmaf(d1, r1, loss_part);
            }
            r0 = d0;
            r1 = d1;
          } else if (LOSS) {
            loss_part = fmaf(r0, r0, loss_part);
            loss_part = fmaf(r1, r1, loss_part);
          }
          const uint32_t hh = pack_bf16x2(r0, r1);                           // x = hi + lo, hi = rn_bf16(x)
          const uint32_t ll = pack_bf16x2(r0 - __uint_as_float(hh << 16), r1 - __uint_as_float(hh & 0xffff0000u));
          hl[h * 16 + (j >> 1)] = hh;
          hl[h * 16 + 8 + (j >> 1)] = ll;
        }
#if PMX_Y_REFILL_EARLY
        // refill the buffer this half just consumed right away (second half of tile t + 1 / first half of tile t + 2)
        if constexpr (!WGT) {
          if (h == 0) {
            if (t + 1 < ntiles) issue_y_half(std::integral_constant<int, YA>{}, nxt, 1, p.Y);
          } else {
            if (t + 2 < ntiles) issue_y_half(std::integral_constant<int, YB>{}, nxt2, 0, p.Y);
          }
        }
#endif
        tmem_st16(lane_addr + TM_ACC + slot * 128 + grp * 32 + h * 16, &hl[h * 16]);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_RT_FULL + slot));       // MMA2(t) may start
      if (warp == 4) TR(3, t, 1);
      // KH = 2: the G_S^T accumulator is single-buffered: issuer 2 may only start on this tile once the previous tile's
      // G_S^T has left tensor memory.  Flushing it here -- after this tile's conversion, which therefore overlaps the
      // G_S MMAs of the previous tile -- keeps the serial chain per tile at (G_S MMAs + one TMEM read).
      if constexpr (GS1) {
        if (t > 0) flush_gs(prev, t - 1);
      }
      // ---- R^T from registers to shared memory (the MN-major operand of the G_A GEMM) once MMA3(t-1) released it
      mbar_wait(bar(B_RS_EMPTY), (t & 1) ^ 1);
      if (warp == 4) TR(3, t, 2);
      if (want_ga) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (ABL(16)) continue;
          const int chunk = ((grp & 1) * 4 + c) ^ (row & 7);
          const int o = (c >> 1) * 16 + (c & 1) * 4;     // hi pairs of elements 8c .. 8c+7
          *reinterpret_cast<uint4*>(rh + (chunk << 4)) = make_uint4(hl[o], hl[o + 1], hl[o + 2], hl[o + 3]);
          *reinterpret_cast<uint4*>(rl + (chunk << 4)) = make_uint4(hl[o + 8], hl[o + 9], hl[o + 10], hl[o + 11]);
        }
        fence_async_smem();   // generic-proxy writes of R^T -> visible to the tensor-core (async) proxy
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(B_RS_FULL));             // MMA3(t) may start
        mbar_arrive(bar(B_ACC_EMPTY + slot));    // our reads of this accumulator are done
      }
      if (warp == 4) TR(3, t, 3);
      // both Y buffers of this tile are consumed: refill them (second half of tile t + 1, first half of tile t + 2).
      // Issued after the hand-offs above so that load-queue back pressure never delays the MMA issuers
      // (re-loading every Y register right after its use was measured 8 % slower in round 1: the load issue then
      // sits on the path to the R^T hand-off)
      if constexpr (WGT) {
        if (t + 1 < ntiles) {
          issue_y_half(IC0{}, nxt, 0, p.Y);
          issue_y_half(IC1{}, nxt, 1, p.Y);
          issue_y_half(IC2{}, nxt, 0, p.W);
        }
      } else {
#if !PMX_Y_REFILL_EARLY
        if (t + 1 < ntiles) issue_y_half(std::integral_constant<int, YA>{}, nxt, 1, p.Y);
        if (t + 2 < ntiles) issue_y_half(std::integral_constant<int, YB>{}, nxt2, 0, p.Y);
#endif
      }
      // ---- gradient flushes, one tile behind so that they never wait for the tensor pipe in steady state
      if (t > 0) {
        if constexpr (!GS1) flush_gs(prev, t - 1);
        if (prev_last) {
          flush_ga(prev, seg);
          ++seg;
        }
      }
      if (warp == 4) TR(3, t, 4);
      prev = pos;
      prev_last = last_in_seg(t, pos);
      pos = nxt;
    };
    for (int t = 0; t < ntiles; t += 3) {
      tile_body(IC0{}, t);
      if (t + 1 < ntiles) tile_body(IC1{}, t + 1);
      if (t + 2 < ntiles) tile_body(IC2{}, t + 2);
    }
    if (ntiles > 0) {
      flush_gs(prev, ntiles - 1);
      flush_ga(prev, seg);
    }
    if (LOSS) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss_part += __shfl_xor_sync(0xffffffffu, loss_part, o);
      if (lane == 0) atomicAdd(p.loss, 0.5 * (double)loss_part);
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
             uint64_t pitch_bytes, uint32_t box_cols, uint32_t box_rows,
             CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
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
                   swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    pmx_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PMX_ERR_CUDA;
  }
  return PMX_OK;
}

}  // namespace

struct UmmaPlan {
  int M, N, K, Mp, Np;
  int KH, KPT;   // k-halves (1: K <= 64, 2: K <= 128) and the padded K = 64 KH of the bf16 operand buffers
  const float* Y;
  int ldY;
  void *Ahi, *Alo, *Shi, *Slo;  // bf16 operand buffers (zero padded) the kernel reads
  void *Ahi_own, *Alo_own;      // the plan's own A buffers (Ahi/Alo may point at external ones, umma_plan_use_A)
  CUtensorMap tmAhi, tmAlo, tmShi, tmSlo;
  int y_blocked;
  const float* W;               // weights in the tiled layout of Y, or nullptr (umma_plan_set_W)
};

bool umma_supported(int M, int N, int K) { return K >= 1 && K <= 2 * KP && M >= 1 && N >= 1; }

int umma_plan_create(pmx_ctx* ctx, const float* Y, int ldY, int M, int N, int K, UmmaPlan** out, int y_blocked) {
  PMX_REQUIRE(umma_supported(M, N, K), "unsupported shape for the tcgen05 kernel");
  if (y_blocked) {
    PMX_REQUIRE(ldY == (N + TILE_N - 1) / TILE_N && (reinterpret_cast<uintptr_t>(Y) & 15) == 0,
                "tiled Y: ldY must be the number of 128-column tiles per row block and the base 16-byte aligned");
  } else {
    PMX_REQUIRE(ldY >= N && (long long)ldY * 4 < (1LL << 31), "Y row pitch must cover N and stay below 2 GiB");
  }
  UmmaPlan* pl = new UmmaPlan();
  memset(pl, 0, sizeof(*pl));
  pl->M = M; pl->N = N; pl->K = K;
  pl->Y = Y; pl->ldY = ldY; pl->y_blocked = y_blocked;
  pl->Mp = pmx_div_up(M, TILE_M) * TILE_M;
  pl->Np = pmx_div_up(N, TILE_N) * TILE_N;
  pl->KH = K <= KP ? 1 : 2;
  pl->KPT = KP * pl->KH;
  const int KPT = pl->KPT;
  PMX_CUDA(cudaSetDevice(ctx->device));
  PMX_CHECK(pmx_dev_alloc(ctx, &pl->Ahi, (size_t)pl->Mp * KPT * 2));
  PMX_CHECK(pmx_dev_alloc(ctx, &pl->Alo, (size_t)pl->Mp * KPT * 2));
  PMX_CHECK(pmx_dev_alloc(ctx, &pl->Shi, (size_t)KPT * pl->Np * 2));
  PMX_CHECK(pmx_dev_alloc(ctx, &pl->Slo, (size_t)KPT * pl->Np * 2));
  pl->Ahi_own = pl->Ahi;
  pl->Alo_own = pl->Alo;
  PMX_CUDA(cudaMemsetAsync(pl->Ahi, 0, (size_t)pl->Mp * KPT * 2, ctx->stream));
  PMX_CUDA(cudaMemsetAsync(pl->Alo, 0, (size_t)pl->Mp * KPT * 2, ctx->stream));
  PMX_CUDA(cudaMemsetAsync(pl->Shi, 0, (size_t)KPT * pl->Np * 2, ctx->stream));
  PMX_CUDA(cudaMemsetAsync(pl->Slo, 0, (size_t)KPT * pl->Np * 2, ctx->stream));
  // boxes: A = 64 k-columns x 128 m-rows at (64 kh, m0); S = 64 n-columns x 64 k-rows at (n0, 64 kh)
  PMX_CHECK(make_map(&pl->tmAhi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl->Ahi, KPT, (uint64_t)pl->Mp, KPT * 2, KP, TILE_M));
  PMX_CHECK(make_map(&pl->tmAlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl->Alo, KPT, (uint64_t)pl->Mp, KPT * 2, KP, TILE_M));
  PMX_CHECK(make_map(&pl->tmShi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl->Shi, (uint64_t)pl->Np, KPT, (uint64_t)pl->Np * 2, 64, KP));
  PMX_CHECK(make_map(&pl->tmSlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl->Slo, (uint64_t)pl->Np, KPT, (uint64_t)pl->Np * 2, 64, KP));
  static bool attr = false;
  if (!attr) {
    const int s1 = (int)SmemMap<1>::BYTES, s2 = (int)SmemMap<2>::BYTES;
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma<1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1));
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma<1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1));
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma<1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1));
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma<1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1));
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma<2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2));
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma<2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2));
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2));
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma<2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2));
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma<1, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1));
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma<1, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s1));
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma<2, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2));
    PMX_CUDA(cudaFuncSetAttribute(k_grad_umma<2, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2));
    attr = true;
  }
  *out = pl;
  return PMX_OK;
}

int umma_plan_use_A(UmmaPlan* pl, void* Ahi, void* Alo) {
  pl->Ahi = Ahi ? Ahi : pl->Ahi_own;
  pl->Alo = Alo ? Alo : pl->Alo_own;
  PMX_CHECK(make_map(&pl->tmAhi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl->Ahi, pl->KPT, (uint64_t)pl->Mp, pl->KPT * 2, KP, TILE_M));
  PMX_CHECK(make_map(&pl->tmAlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pl->Alo, pl->KPT, (uint64_t)pl->Mp, pl->KPT * 2, KP, TILE_M));
  return PMX_OK;
}

int umma_plan_set_W(UmmaPlan* pl, const float* W) {
  if (W && !pl->y_blocked) {
    pmx_set_error("weighted likelihood needs the tiled layout of Y and W");
    return PMX_ERR_UNSUPPORTED;
  }
  pl->W = W;
  return PMX_OK;
}

void umma_plan_destroy(pmx_ctx* ctx, UmmaPlan* pl) {
  if (!pl) return;
  pmx_dev_free(ctx, pl->Ahi_own);
  pmx_dev_free(ctx, pl->Alo_own);
  pmx_dev_free(ctx, pl->Shi);
  pmx_dev_free(ctx, pl->Slo);
  delete pl;
}

void umma_plan_buffers(UmmaPlan* pl, void** Ahi, void** Alo, void** Shi, void** Slo, int* ldA, int* ldS) {
  *Ahi = pl->Ahi; *Alo = pl->Alo; *Shi = pl->Shi; *Slo = pl->Slo; *ldA = pl->KPT; *ldS = pl->Np;
}

int launch_grad_umma(pmx_ctx* ctx, UmmaPlan* pl, const float* A, const float* S, float* GA, float* GS, double* loss,
                     const int* done, int skip_split, const unsigned* ga_epoch, size_t ga_stride, int want,
                     size_t gs_stride, int reserve_sms) {
  const bool want_ga = (want & 1) && GA, want_gs = (want & 2) && GS;
  if (!skip_split) {
    PMX_CHECK(launch_split_bf16(ctx, A, pl->M, pl->K, pl->Ahi, pl->Alo, pl->Mp, pl->KPT, done));
    PMX_CHECK(launch_split_bf16(ctx, S, pl->K, pl->N, pl->Shi, pl->Slo, pl->KPT, pl->Np, done));
  }
  // (with ga_epoch the G_A pair lives in the peer arena and is cleared by the consumer of the previous epoch)
  // (with gs_stride the G_S pair is cleared by the fused tail of the previous iteration as well)
  PMX_CHECK(launch_zero3(ctx, ctx->stream, GS, (want_gs && gs_stride == 0) ? (size_t)pl->K * pl->N : 0, GA,
                         (ga_epoch || !want_ga) ? 0 : (size_t)pl->M * pl->K, reinterpret_cast<float*>(loss), loss ? 2 : 0,
                         done));
  Params p;
  p.M = pl->M; p.N = pl->N; p.K = pl->K;
  p.NS = pl->Np / TILE_N;
  p.total_tiles = (long long)(pl->Mp / TILE_M) * p.NS;
  p.Y = pl->Y; p.W = pl->W; p.ldY = pl->ldY;
  p.GA = GA; p.GS = GS; p.loss = loss; p.done = done;
  p.want_ga = want_ga ? 1 : 0; p.want_gs = want_gs ? 1 : 0;
  p.ga_epoch = ga_epoch; p.ga_stride = (long long)ga_stride; p.gs_stride = (long long)gs_stride;
  p.y_blocked = pl->y_blocked;
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
  }
  // reserve_sms: SMs left free for a small kernel that runs next to this one on another stream (the persistent CTAs
  // own their SM: 190+ KB of shared memory each)
  const int sms = ctx->sm_count - reserve_sms > 0 ? ctx->sm_count - reserve_sms : 1;
  int grid = (int)(p.total_tiles < sms ? p.total_tiles : sms);
  const bool prof = ctx->profile && ctx->prof_n < PMX_PROF_MAX;
  {
    // the setmaxnreg split must balance: registers released by warpgroup 0 >= registers requested by warpgroups 1-4
    static int regs_ok = -1;
    if (regs_ok < 0) {
      cudaFuncAttributes fa;
      PMX_CUDA(cudaFuncGetAttributes(&fa, k_grad_umma<1, false, false>));
      regs_ok = (fa.numRegs >= PMX_REG_DEC && 128 * (fa.numRegs - PMX_REG_DEC) >= (NUM_THREADS - 128) * (PMX_REG_INC - fa.numRegs)) ? 1 : 0;
      if (!regs_ok) pmx_set_error("k_grad_umma: compiled for %d registers, the setmaxnreg split %d/%d cannot be served", fa.numRegs, PMX_REG_DEC, PMX_REG_INC);
    }
    if (!regs_ok) return PMX_ERR_UNSUPPORTED;
  }
  if (prof) PMX_CUDA(cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n], ctx->stream));
  {
    // instantiations: k-halves (K <= 64 / K <= 128) x with / without the loss reduction x production / debug
    // (timing ablations + trace)
    const bool dbg = p.ablate != 0 || p.trace != nullptr;
    if (pl->W) {   // weighted likelihood (production instantiations only)
      auto kern = pl->KH == 1 ? (loss ? k_grad_umma<1, true, false, true> : k_grad_umma<1, false, false, true>)
                              : (loss ? k_grad_umma<2, true, false, true> : k_grad_umma<2, false, false, true>);
      kern<<<grid, NUM_THREADS, pl->KH == 1 ? SmemMap<1>::BYTES : SmemMap<2>::BYTES, ctx->stream>>>(pl->tmAhi, pl->tmAlo, pl->tmShi,
                                                                                                 pl->tmSlo, p);
    } else if (pl->KH == 1) {
      auto kern = dbg ? (loss ? k_grad_umma<1, true, true> : k_grad_umma<1, false, true>)
                      : (loss ? k_grad_umma<1, true, false> : k_grad_umma<1, false, false>);
      kern<<<grid, NUM_THREADS, SmemMap<1>::BYTES, ctx->stream>>>(pl->tmAhi, pl->tmAlo, pl->tmShi, pl->tmSlo, p);
    } else {
      auto kern = dbg ? (loss ? k_grad_umma<2, true, true> : k_grad_umma<2, false, true>)
                      : (loss ? k_grad_umma<2, true, false> : k_grad_umma<2, false, false>);
      kern<<<grid, NUM_THREADS, SmemMap<2>::BYTES, ctx->stream>>>(pl->tmAhi, pl->tmAlo, pl->tmShi, pl->tmSlo, p);
    }
  }
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
