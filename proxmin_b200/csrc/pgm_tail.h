// Fused tail of a PGM iteration (pgm_tail.cu): argument block and launch wrappers.
#pragma once
#include "prox.cuh"

#define PMX_TAIL_RA 64     // rows of A per A block (upper bound)
#define PMX_TAIL_NREP 8    // replicated fp32 accumulators of the per-block Gram partials

struct PgmTailArgs {
  int M, N, K;              // N = local columns of this rank's stripe
  int world, rank;
  pmx_ctl* ctl;
  // ---- S block (column stripe, local)
  float* S;
  float* GS2;               // pair of K x N gradient buffers, gs_stride elements apart
  long long gs_stride;
  unsigned short *Shi, *Slo;   // bf16 operands of the next gradient kernel
  int ldS;
  ProxChain chS;
  int n_tiles_S, nS;        // 32-column tiles and the number of S blocks (= first block index of the A blocks)
  // ---- A block (replicated; this rank updates rows [m_lo, m_hi))
  int m_lo, m_hi, nA;
  int ra;                   // rows of A per A block (<= PMX_TAIL_RA)
  ProxChain chA;
  int ldA;                  // row pitch (elements) of the bf16 A operand buffers
  long long ga_stride;      // elements between the two G_A buffers
  // single GPU: local buffers
  float* A_loc;
  float* GA2_loc;
  unsigned short *Ahi_loc, *Alo_loc;
  // multi GPU: symmetric arena of every rank (comm.cu) and the byte offsets of its regions
  pmx_peer_ptrs arena;
  size_t off_GA, off_A, off_Ahi, off_Alo, off_inbox;
  pmx_peer_ptrs flags;      // flag blocks of every rank
  unsigned* epoch;          // local epoch counters [PMX_PEER_SETS]
  const unsigned* my_flags; // local flag block
  // ---- scratch (local, zero at the start of a solve)
  float* gram_rep;          // [2 (A, S)][PMX_TAIL_NREP][K * K]
  double* acc;              // [2][4] norm partials (|dX|^2, |X|^2, |X_old|^2)
  unsigned* tickets;        // [3]
  float e2A, e2S;           // e_rel^2 per block
  unsigned long long* trace;   // debug (env PMX_TAIL_TRACE): globaltimer stamps of the phases, see pgm_tail.cu
};

size_t pgm_tail_smem_bytes(int K);
// number of S blocks for a stripe of n_cols columns; *n_tiles = number of 32-column tiles
int pgm_tail_s_blocks(pmx_ctx* ctx, int n_cols, int* n_tiles);
int launch_pgm_tail(pmx_ctx* ctx, const PgmTailArgs& a);
// Gram totals, lambda_max, convergence test of the iteration the roles kernel just completed (one block; no-op when
// nothing is pending); st = the stream it runs on (the side stream: it overlaps the next gradient kernel)
int launch_tail_final(pmx_ctx* ctx, cudaStream_t st, const PgmTailArgs& a);
// step2[it & 1][j] = step[j]: seeds the double-buffered steps from the stand-alone Lipschitz kernels (first iteration)
int launch_tail_seed_steps(pmx_ctx* ctx, pmx_ctl* ctl);
