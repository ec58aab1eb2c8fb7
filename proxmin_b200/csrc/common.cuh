// Shared internals of libproxmin_b200.so: context, error reporting, launch bookkeeping.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/proxmin_b200.h"

void pmx_set_error(const char* fmt, ...);

#define PMX_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      pmx_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));     \
      return PMX_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define PMX_CHECK(st)                 \
  do {                                \
    int _s = (st);                    \
    if (_s != PMX_OK) return _s;      \
  } while (0)

#define PMX_REQUIRE(cond, msg)                                        \
  do {                                                                \
    if (!(cond)) {                                                    \
      pmx_set_error("%s:%d: %s (%s)", __FILE__, __LINE__, msg, #cond); \
      return PMX_ERR_ARG;                                             \
    }                                                                 \
  } while (0)

// Device-side control block of a solver run.  Every solver kernel starts with
// `if (ctl->done) return;` so that the iterate freezes at exactly the iteration where
// the reference would `break` (algorithms.py:134-135) even though the host has already
// enqueued further iterations.
struct pmx_ctl {
  int done;          // set when every block converged (or a non-finite step was met)
  int it;            // iterations completed
  int conv[2];       // convergence flags of the last completed iteration
  int nonfinite;     // Gram matrix / lambda_max not finite
  int sub_done;      // adaprox: proximal sub-iteration loop of the current block finished
  int sub_tau;       // adaprox: sub-iterations used by the current block in this iteration
  int sub_parity;    // adaprox: which z buffer holds the result
  long long sub_total[2];
  double norms[16];  // [0..2] block A: |dX|^2, |X|^2, |Xprev|^2 ; [3..5] block S ; [6] loss ; [8..10] adaprox sub-iteration
  float step[2];     // 1/lambda_max for A and S (algorithms.py:106)
  float lip[2];      // lambda_max(S S^T), lambda_max(A^T A)
  float psi_max[2];  // adaprox: max(Psi) per block (algorithms.py:384)
  float step2[2][2]; // fused PGM tail: steps double-buffered by iteration parity ([it & 1] = steps of iteration `it`)
  int fault;         // a peer did not answer within the spin limit (multi-GPU exchange): the solve is aborted
  unsigned par_ctr;  // fused PGM tail: iterations whose factor updates are complete (parity of the gradient buffers)
  int final_pending; // fused PGM tail: the roles kernel finished an iteration that k_tail_final has not closed yet
  int paused_block;  // adaprox: block (1 = A, 2 = S) whose speculative sub-iterations did not converge; done == 2 then
                     // freezes every following kernel until the host has finished the block (nmf_solver.cu)
  int sub_last[2];   // adaprox: sub-iterations the last completed block update took, per block
  int pad2[2];
};

// ---- peer-memory exchange (comm.cu): symmetric device regions mapped into every rank of the box over CUDA IPC
#define PMX_MAX_WORLD 8
#define PMX_PEER_SETS 4            // independent flag sets (one per exchange point of an iteration)
#define PMX_SMALL_MAX 256          // 8-byte slots per rank of a small all-reduce inbox (comm.cu)
struct pmx_peer_region {
  void* local;                     // this rank's allocation (cudaMalloc)
  void* peer[PMX_MAX_WORLD];       // the same region of every rank, peer[rank] == local
  size_t bytes;
};
struct pmx_peer_ptrs {             // by-value kernel argument
  void* p[PMX_MAX_WORLD];
};

struct pmx_ctx {
  int device;
  int sm_count;
  cudaStream_t stream;   // main stream: every kernel of the hot path
  cudaStream_t aux;      // side stream: Gram / lambda_max overlap with the gradient kernel
  cudaEvent_t ev_fork, ev_join, ev_fork2, ev_join2, ev_t0, ev_t1;
  long long launches;
  // NCCL (dlopen'ed lazily)
  void* nccl_comm;
  void* nccl_comm_aux;   // second communicator (ncclCommSplit) for collectives on the side stream
  // peer-memory exchange: flags[set][rank] written by the peers (st.release.sys), epoch counters local
  int peer_ok;
  pmx_peer_region peer_flags;
  unsigned* peer_epoch;  // [PMX_PEER_SETS] device counters, never reset
  pmx_peer_region peer_arena;   // symmetric scratch for the one-shot all-reduces (grown collectively)
  int world, rank;
  // optional per-launch timing of the dominant kernel (bench.py roofline leg)
  int profile;                 // 0 off, 1 on
  int prof_n;                  // recorded launches
  cudaEvent_t* prof_ev;        // 2 * PMX_PROF_MAX events (start, stop)
  // scratch
  int* h_flags;          // pinned host mirror for polled device flags
  char dev_name[128];
  size_t total_mem;
  void* pool;            // cache of freed device blocks (pmx_dev_alloc / pmx_dev_free), see api.cu
  // scratch of launch_gram (per-block partial Gram matrices): one buffer per (tall/wide, main/side stream)
  float* gram_scratch[4];
  size_t gram_scratch_bytes[4];
};

static inline int pmx_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

#define PMX_LAUNCHED(ctx) ((ctx)->launches++)
#define PMX_PROF_MAX 4096

int pmx_check_launch(pmx_ctx* ctx, const char* what);

// Device memory for solver handles goes through a per-context cache of freed blocks: a solve on host arrays
// (create -> iterate -> destroy) then costs no cudaMalloc/cudaFree after the first one of a given shape (freeing and
// re-mapping the 2 GB Y buffer was measured at 5-700 ms).  pmx_ctx_trim releases the cached blocks.
int pmx_dev_alloc(pmx_ctx* ctx, void** p, size_t bytes);
void pmx_dev_free(pmx_ctx* ctx, void* p);
void pmx_dev_trim(pmx_ctx* ctx);
