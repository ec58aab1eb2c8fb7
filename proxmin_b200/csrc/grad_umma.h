// tcgen05 / TMA / TMEM gradient kernel (grad_umma.cu): host-side plan and launch wrapper.
#pragma once
#include "common.cuh"

struct UmmaPlan;

// shapes the tcgen05 kernel takes: K <= 128 (operands are zero-padded to K = 64 or, for K > 64, to 128 and handled
// as two 64-deep k-halves), any M, any N
bool umma_supported(int M, int N, int K);
// Y: fp32, two layouts:
//   y_blocked = 0: row-major with a row pitch of ldY floats (any alignment: plain coalesced 4-byte loads);
//   y_blocked = 1: tiled -- the layout of a solver handle's device copy of Y (pmx_nmf_set_Y builds it).  The matrix is
//     cut into 128 x 128 tiles stored one after the other in the order the gradient kernel walks them (row block major);
//     inside a tile: [32 row quads][128 columns][4 rows].  ldY = number of tiles per row block = ceil(N / 128); the
//     buffer holds ceil(M / 128) * ldY tiles of 64 KB, padding zero, base 16-byte aligned.  A thread of the gradient
//     kernel fetches 4 rows of its column with one 16-byte load, a warp instruction covers 512 contiguous bytes and a
//     CTA streams one contiguous region (pmx_y_index below).
int umma_plan_create(pmx_ctx* ctx, const float* Y, int ldY, int M, int N, int K, UmmaPlan** out, int y_blocked = 0);
__host__ __device__ inline size_t pmx_y_index(int m, int n, int ldY, int y_blocked) {
  if (!y_blocked) return (size_t)m * ldY + n;
  return ((((size_t)(m >> 7) * ldY + (n >> 7)) * 32 + ((m & 127) >> 2)) * 128 + (n & 127)) * 4 + (m & 3);
}
void umma_plan_destroy(pmx_ctx* ctx, UmmaPlan* plan);
// weights of the weighted likelihood (nmf.py:25, 40): M x N in the tiled layout of Y (y_blocked = 1 plans only), or
// nullptr for W = 1
int umma_plan_set_W(UmmaPlan* plan, const float* W);
// skip_split != 0: the plan's bf16 (hi, lo) operand buffers already hold the split of (A, S) -- the fused update
// kernels wrote them -- so the two split passes are skipped
// ga_epoch != nullptr (sharded runs over peer memory): GA is the base of a pair of buffers ga_stride elements apart
// and the partials go to buffer (*ga_epoch + 1) & 1, which its consumer cleared (comm.cu)
// want: bit 0 = G_A, bit 1 = G_S; a gradient that is not wanted costs no MMAs, no flush and its buffer is not
// touched (bsdmm needs one gradient per pass, nmf.py:181-185; the loss needs none, nmf.py:13-25)
int launch_grad_umma(pmx_ctx* ctx, UmmaPlan* plan, const float* A, const float* S, float* GA, float* GS, double* loss,
                     const int* done, int skip_split = 0, const unsigned* ga_epoch = nullptr, size_t ga_stride = 0,
                     int want = 3, size_t gs_stride = 0, int reserve_sms = 0);
// the A operand buffers the kernel reads (tensor maps re-encoded): external Mp x ldA bf16 buffers, e.g. regions of the
// peer arena that every rank's fused PGM tail writes into; nullptr = back to the plan's own buffers
int umma_plan_use_A(UmmaPlan* plan, void* Ahi, void* Alo);
// bf16 operand buffers of the plan: A_hi/A_lo are Mp x *ldA (row pitch *ldA = padded K), S_hi/S_lo are *ldA x Np
// (row pitch *ldS)
void umma_plan_buffers(UmmaPlan* plan, void** Ahi, void** Alo, void** Shi, void** Slo, int* ldA, int* ldS);
