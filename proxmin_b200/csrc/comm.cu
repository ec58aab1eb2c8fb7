// Multi-GPU plumbing: one process per GPU, NCCL all-reduce over NVLink 5 / NVSwitch on the
// context's stream.  NCCL is dlopen'ed on first use so that the single-GPU path has no link-time
// dependency on it.  The reference has no distributed code at all; the only exchange
// step the column-sharded NMF needs is sum(G_A partials) (+ K x K Gram and a few scalars).
#include <dlfcn.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8, ncclInt32 = 2 };
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
  void* lib;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*);   // optional (NCCL >= 2.18)
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
} g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

int load_nccl() {
  if (g_nccl.lib) return PMX_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  void* lib = nullptr;
  for (int i = 0; names[i] && !lib; ++i) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    pmx_set_error("cannot dlopen libnccl.so.2: %s", dlerror());
    return PMX_ERR_NCCL;
  }
  g_nccl.GetUniqueId = (ncclResult_t(*)(ncclUniqueId*))dlsym(lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (ncclResult_t(*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (ncclResult_t(*)(ncclComm_t))dlsym(lib, "ncclCommDestroy");
  g_nccl.AllReduce = (ncclResult_t(*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllReduce");
  g_nccl.GetErrorString = (const char* (*)(ncclResult_t))dlsym(lib, "ncclGetErrorString");
  g_nccl.CommSplit = (ncclResult_t(*)(ncclComm_t, int, int, ncclComm_t*, void*))dlsym(lib, "ncclCommSplit");
  g_nccl.AllGather = (ncclResult_t(*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(lib, "ncclAllGather");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
    pmx_set_error("libnccl is missing a required symbol");
    return PMX_ERR_NCCL;
  }
  g_nccl.lib = lib;
  return PMX_OK;
}

#define PMX_NCCL(call)                                                                              \
  do {                                                                                              \
    ncclResult_t _r = (call);                                                                       \
    if (_r != 0) {                                                                                  \
      pmx_set_error("%s -> %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error"); \
      return PMX_ERR_NCCL;                                                                          \
    }                                                                                               \
  } while (0)

}  // namespace

int pmx_peer_teardown_internal(pmx_ctx* ctx);
int pmx_comm_destroy_internal(pmx_ctx* ctx) {
  pmx_peer_teardown_internal(ctx);
  if (ctx->nccl_comm_aux && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm_aux);
  ctx->nccl_comm_aux = nullptr;
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  return PMX_OK;
}

// dtype: 0 fp32 sum, 1 fp64 sum, 2 fp32 max (bit pattern of non-negative floats == int max)
int pmx_comm_allreduce_internal(pmx_ctx* ctx, void* buf, size_t count, int kind, cudaStream_t st) {
  if (ctx->world <= 1) return PMX_OK;
  if (!ctx->nccl_comm) {
    pmx_set_error("communicator not initialised (call pmx_comm_init)");
    return PMX_ERR_NCCL;
  }
  const int dt = kind == 1 ? ncclFloat64 : (kind == 2 ? ncclInt32 : ncclFloat32);
  const int op = kind == 2 ? ncclMax : ncclSum;
  // collectives on the side stream go through their own communicator so that they can overlap the ones of the
  // main stream (one communicator serialises its operations)
  void* comm = (st == ctx->aux && ctx->nccl_comm_aux) ? ctx->nccl_comm_aux : ctx->nccl_comm;
  PMX_NCCL(g_nccl.AllReduce(buf, buf, count, dt, op, (ncclComm_t)comm, st));
  return PMX_OK;
}

bool pmx_comm_has_aux(pmx_ctx* ctx) { return ctx->world > 1 && ctx->nccl_comm_aux != nullptr; }

// =====================================================================================================
// Peer-memory exchange: one-shot all-reduce over NVLink with CUDA-IPC mapped buffers.
//
// The column-sharded solvers exchange small replicated quantities every iteration (the 2 MB G_A partials, a K x K
// Gram matrix, a few norms).  An NCCL all-reduce of that size is latency bound (tens of microseconds per call next
// to a ~70 us gradient kernel at 8 GPUs), so the hot path reads the partials of every rank straight out of peer
// memory instead:
//   * every rank owns a "symmetric" region (cudaMalloc + cudaIpcGetMemHandle) that all ranks map;
//   * producer side: the partial lands in the local region (buffer parity = epoch & 1), then k_peer_signal bumps
//     the local epoch counter of the flag set and stores it (st.release.sys) into flags[set][rank] of every peer;
//   * consumer side: k_peer_sum waits (ld.acquire.sys on LOCAL memory) until every rank has reached the epoch,
//     sums the world partials in rank order -- every rank computes bit-identical results -- and clears the other
//     parity buffer for the next iteration (its readers are done: they signalled this epoch after finishing the
//     previous one).
// No host involvement: the kernels replay inside the iteration's CUDA graph.  NCCL stays for setup (exchange of
// the IPC handles) and as the fallback when IPC mapping is unavailable.
// =====================================================================================================
namespace {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// copy_src != nullptr: the partial (n doubles) is first copied into the local parity buffer at copy_dst (base of
// the 2 x n buffer pair) -- used for small quantities that the producing kernels keep in their own buffers
__global__ void k_peer_signal(unsigned* epoch, pmx_peer_ptrs flags, int set, int world, int rank, const int* done,
                              const double* copy_src, double* copy_dst, size_t n) {
  if (done && *done) return;
  __shared__ unsigned e;
  if (threadIdx.x == 0) {
    e = epoch[set] + 1;
    epoch[set] = e;
  }
  __syncthreads();
  if (copy_src) {
    double* dst = copy_dst + (size_t)(e & 1u) * n;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = copy_src[i];
  }
  // the partial sits in LOCAL memory (peers read it from here): the fence orders it before the flag for
  // system-scope observers
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world)
    st_release_sys(reinterpret_cast<unsigned*>(flags.p[threadIdx.x]) + set * PMX_MAX_WORLD + rank, e);
}

// dst[i] = sum_r part_r[parity][i] (rank order), local other-parity buffer cleared; T = float or double
template <typename T>
__global__ void k_peer_sum(pmx_peer_ptrs parts, size_t offset_bytes, size_t n, T* dst, const unsigned* epoch,
                           const unsigned* my_flags, int set, int world, int rank, const int* done, int* fault) {
  if (done && *done) return;
  const unsigned e = epoch[set];
  if ((int)threadIdx.x < world) {
    const unsigned* f = my_flags + set * PMX_MAX_WORLD + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(f) - e) < 0) {
      if (clock64() - t0 > 20000000000LL) {   // ~10 s of SM clocks: a lost peer must not hang the GPU
        if (fault) *fault = 1;
        break;
      }
    }
  }
  __syncthreads();
  const size_t par = e & 1u, stride = n * sizeof(T);
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
  if (sizeof(T) == 4 && (n & 3) == 0) {
    const size_t n4 = n >> 2;
    for (size_t i = i0; i < n4; i += step) {
      // all remote loads first (one NVLink round trip instead of `world` dependent ones), then the sum in rank order
      float4 v[PMX_MAX_WORLD];
#pragma unroll
      for (int r = 0; r < PMX_MAX_WORLD; ++r) {
        if (r < world) {
          const float4* src = reinterpret_cast<const float4*>(static_cast<const char*>(parts.p[r]) + offset_bytes + par * stride) + i;
          asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[r].x), "=f"(v[r].y), "=f"(v[r].z), "=f"(v[r].w) : "l"(src));
        }
      }
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < PMX_MAX_WORLD; ++r) {
        if (r < world) {
          acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w;
        }
      }
      reinterpret_cast<float4*>(dst)[i] = acc;
      reinterpret_cast<float4*>(static_cast<char*>(parts.p[rank]) + offset_bytes + (par ^ 1) * stride)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    for (size_t i = i0; i < n; i += step) {
      T acc = 0;
      for (int r = 0; r < world; ++r) {
        const volatile T* src = reinterpret_cast<const volatile T*>(static_cast<const char*>(parts.p[r]) + offset_bytes + par * stride) + i;
        acc += *src;
      }
      dst[i] = acc;
      reinterpret_cast<T*>(static_cast<char*>(parts.p[rank]) + offset_bytes + (par ^ 1) * stride)[i] = 0;
    }
  }
}

int peer_close(pmx_ctx* ctx, pmx_peer_region* r) {
  for (int i = 0; i < ctx->world && i < PMX_MAX_WORLD; ++i)
    if (r->peer[i] && i != ctx->rank) cudaIpcCloseMemHandle(r->peer[i]);
  if (r->local) cudaFree(r->local);
  memset(r, 0, sizeof(*r));
  return PMX_OK;
}

// collective over the communicator: every rank allocates `bytes` (zero filled) and maps the regions of all ranks
int peer_alloc(pmx_ctx* ctx, size_t bytes, pmx_peer_region* out) {
  memset(out, 0, sizeof(*out));
  if (!g_nccl.AllGather) return PMX_ERR_NCCL;
  const int world = ctx->world;
  PMX_CUDA(cudaMalloc(&out->local, bytes));
  PMX_CUDA(cudaMemset(out->local, 0, bytes));
  cudaIpcMemHandle_t mine;
  int fail = cudaIpcGetMemHandle(&mine, out->local) != cudaSuccess;
  cudaGetLastError();
  // exchange [handle | fail flag] through NCCL (device staging buffers)
  const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
  char h_send[sizeof(cudaIpcMemHandle_t) + 8] = {0};
  memcpy(h_send, &mine, sizeof(mine));
  h_send[sizeof(mine)] = (char)fail;
  char *d_send = nullptr, *d_recv = nullptr;
  PMX_CUDA(cudaMalloc((void**)&d_send, rec));
  PMX_CUDA(cudaMalloc((void**)&d_recv, rec * world));
  PMX_CUDA(cudaMemcpyAsync(d_send, h_send, rec, cudaMemcpyHostToDevice, ctx->stream));
  PMX_NCCL(g_nccl.AllGather(d_send, d_recv, rec, /*ncclInt8*/ 0, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  char* h_recv = (char*)malloc(rec * world);
  PMX_CUDA(cudaMemcpyAsync(h_recv, d_recv, rec * world, cudaMemcpyDeviceToHost, ctx->stream));
  PMX_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_send);
  cudaFree(d_recv);
  for (int r = 0; r < world; ++r) fail |= h_recv[r * rec + sizeof(mine)];
  int open_fail = 0;
  if (!fail) {
    for (int r = 0; r < world; ++r) {
      if (r == ctx->rank) {
        out->peer[r] = out->local;
        continue;
      }
      cudaIpcMemHandle_t hnd;
      memcpy(&hnd, h_recv + r * rec, sizeof(hnd));
      if (cudaIpcOpenMemHandle(&out->peer[r], hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        out->peer[r] = nullptr;
        open_fail = 1;
      }
    }
  }
  free(h_recv);
  // every rank must take the same decision: all-reduce(max) of the failure flag (also the barrier that keeps a
  // rank from using the mapping before everybody has it)
  int* d_flag = nullptr;
  PMX_CUDA(cudaMalloc((void**)&d_flag, sizeof(int)));
  int hf = fail | open_fail;
  PMX_CUDA(cudaMemcpyAsync(d_flag, &hf, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  PMX_NCCL(g_nccl.AllReduce(d_flag, d_flag, 1, ncclInt32, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  PMX_CUDA(cudaMemcpyAsync(&hf, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  PMX_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_flag);
  if (hf) {
    peer_close(ctx, out);
    return PMX_ERR_UNSUPPORTED;
  }
  out->bytes = bytes;
  return PMX_OK;
}

}  // namespace

bool pmx_peer_available(pmx_ctx* ctx) { return ctx->world > 1 && ctx->peer_ok; }

// symmetric scratch of at least `bytes` (collective when it has to grow: every rank asks for the same size)
int pmx_peer_arena(pmx_ctx* ctx, size_t bytes, pmx_peer_region** out) {
  if (!pmx_peer_available(ctx)) return PMX_ERR_UNSUPPORTED;
  if (ctx->peer_arena.bytes < bytes) {
    PMX_CUDA(cudaStreamSynchronize(ctx->stream));
    PMX_CUDA(cudaStreamSynchronize(ctx->aux));
    if (ctx->peer_arena.local) {
      // nobody may still read the old mapping: barrier through NCCL before unmapping
      int* d = nullptr;
      PMX_CUDA(cudaMalloc((void**)&d, sizeof(int)));
      PMX_CUDA(cudaMemsetAsync(d, 0, sizeof(int), ctx->stream));
      PMX_NCCL(g_nccl.AllReduce(d, d, 1, ncclInt32, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
      PMX_CUDA(cudaStreamSynchronize(ctx->stream));
      cudaFree(d);
      peer_close(ctx, &ctx->peer_arena);
    }
    size_t want = bytes < ((size_t)8 << 20) ? ((size_t)8 << 20) : bytes;
    int st = peer_alloc(ctx, want, &ctx->peer_arena);
    if (st != PMX_OK) {
      ctx->peer_ok = 0;
      return st;
    }
  }
  *out = &ctx->peer_arena;
  return PMX_OK;
}

// "my partial of this epoch is complete" (it lives in buffer parity (new epoch) & 1 of its 2 x n pair in the arena);
// with copy_src the n doubles are first copied there from a private buffer
int pmx_peer_signal(pmx_ctx* ctx, int set, cudaStream_t st, const int* done, const double* copy_src,
                    size_t copy_offset_bytes, size_t n) {
  pmx_peer_ptrs f;
  for (int r = 0; r < PMX_MAX_WORLD; ++r) f.p[r] = ctx->peer_flags.peer[r];
  double* dstb = copy_src ? reinterpret_cast<double*>(static_cast<char*>(ctx->peer_arena.local) + copy_offset_bytes) : nullptr;
  k_peer_signal<<<1, copy_src ? 256 : 32, 0, st>>>(ctx->peer_epoch, f, set, ctx->world, ctx->rank, done, copy_src, dstb, n);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_peer_signal");
}

// start of a sharded solve (collective, host synchronous): every rank has finished its previous kernels, then the
// local arena is cleared -- whatever layout the previous solve used
int pmx_peer_reset(pmx_ctx* ctx) {
  if (!pmx_peer_available(ctx) || !ctx->peer_arena.local) return PMX_OK;
  PMX_CUDA(cudaStreamSynchronize(ctx->stream));
  PMX_CUDA(cudaStreamSynchronize(ctx->aux));
  int* d = nullptr;
  PMX_CUDA(cudaMalloc((void**)&d, sizeof(int)));
  PMX_CUDA(cudaMemsetAsync(d, 0, sizeof(int), ctx->stream));
  PMX_NCCL(g_nccl.AllReduce(d, d, 1, ncclInt32, ncclMax, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  PMX_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d);
  PMX_CUDA(cudaMemsetAsync(ctx->peer_arena.local, 0, ctx->peer_arena.bytes, ctx->stream));
  return PMX_OK;
}

// dst = sum over ranks of the partials; kind 0 fp32, 1 fp64
int pmx_peer_sum(pmx_ctx* ctx, int set, size_t offset_bytes, size_t n, void* dst, int kind, cudaStream_t st, const int* done,
                 int* fault) {
  pmx_peer_ptrs parts;
  for (int r = 0; r < PMX_MAX_WORLD; ++r) parts.p[r] = ctx->peer_arena.peer[r];
  const unsigned* my_flags = reinterpret_cast<const unsigned*>(ctx->peer_flags.local);
  const size_t work = kind == 0 && (n & 3) == 0 ? n / 4 : n;
  int blocks = (int)((work + 255) / 256);
  if (blocks > 4 * ctx->sm_count) blocks = 4 * ctx->sm_count;
  if (blocks < 1) blocks = 1;
  if (kind == 0)
    k_peer_sum<float><<<blocks, 256, 0, st>>>(parts, offset_bytes, n, (float*)dst, ctx->peer_epoch, my_flags, set, ctx->world, ctx->rank, done, fault);
  else
    k_peer_sum<double><<<blocks, 256, 0, st>>>(parts, offset_bytes, n, (double*)dst, ctx->peer_epoch, my_flags, set, ctx->world, ctx->rank, done, fault);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_peer_sum");
}

// ---- small all-reduce in ONE kernel (push model): every rank writes its n values into slot `rank` of every rank's
// inbox, signals, waits for all ranks and reduces the slots in rank order (bit-identical results everywhere).
// Replaces the latency-bound NCCL all-reduces of a few scalars inside the adaprox / bsdmm iterations (max Psi, the
// sub-iteration norms, the row sums of S, the constraint norms): ~6 us instead of 40-60 us at 8 ranks.
// Inbox layout at off_bytes of the arena: [2 parity][PMX_MAX_WORLD ranks][PMX_SMALL_MAX] 8-byte slots.  The parity
// double-buffering is safe for any sequence of exchanges on one flag set: a rank can only be one epoch ahead of the
// slowest one (it needs everybody's signal to finish an exchange).
// kind 1: doubles, sum.  kind 2: 32-bit words compared as signed integers, max (non-negative floats order like ints).
namespace {
__global__ void __launch_bounds__(256) k_small_allreduce(pmx_peer_ptrs arena, size_t off_bytes, void* buf, int n, int kind,
                                                         unsigned* epoch, pmx_peer_ptrs flags, const unsigned* my_flags,
                                                         int set, int world, int rank, const int* done, int* fault) {
  if (done && *done) return;
  __shared__ unsigned s_e;
  __shared__ int s_fault;
  if (threadIdx.x == 0) {
    s_e = epoch[set] + 1;
    epoch[set] = s_e;
    s_fault = 0;
  }
  __syncthreads();
  const unsigned e = s_e;
  const size_t par = e & 1u;
  for (int r = 0; r < world; ++r) {
    double* dst = reinterpret_cast<double*>(static_cast<char*>(arena.p[r]) + off_bytes) + (par * PMX_MAX_WORLD + rank) * PMX_SMALL_MAX;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      if (kind == 1) dst[i] = static_cast<const double*>(buf)[i];
      else reinterpret_cast<long long*>(dst)[i] = (long long)static_cast<const int*>(buf)[i];
    }
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world)
    st_release_sys(reinterpret_cast<unsigned*>(flags.p[threadIdx.x]) + set * PMX_MAX_WORLD + rank, e);
  if ((int)threadIdx.x < world) {
    const unsigned* f = my_flags + set * PMX_MAX_WORLD + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(f) - e) < 0) {
      if (clock64() - t0 > 20000000000LL) {   // ~10 s: a lost peer must not hang the GPU
        s_fault = 1;
        break;
      }
    }
  }
  __syncthreads();
  if (s_fault) {
    if (threadIdx.x == 0 && fault) *fault = 1;
    return;
  }
  const volatile double* in = reinterpret_cast<const volatile double*>(static_cast<char*>(arena.p[rank]) + off_bytes) +
                              par * PMX_MAX_WORLD * PMX_SMALL_MAX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (kind == 1) {
      double acc = 0.0;
      for (int r = 0; r < world; ++r) acc += in[(size_t)r * PMX_SMALL_MAX + i];
      static_cast<double*>(buf)[i] = acc;
    } else {
      long long m = reinterpret_cast<const volatile long long*>(in)[i];
      for (int r = 1; r < world; ++r) {
        const long long v = reinterpret_cast<const volatile long long*>(in)[(size_t)r * PMX_SMALL_MAX + i];
        m = v > m ? v : m;
      }
      static_cast<int*>(buf)[i] = (int)m;
    }
  }
}
}  // namespace

size_t pmx_peer_small_bytes() { return sizeof(double) * 2 * PMX_MAX_WORLD * PMX_SMALL_MAX; }

// in-place all-reduce of n <= PMX_SMALL_MAX values at `buf` (device) through the inbox at off_bytes of the arena
int pmx_peer_small_allreduce(pmx_ctx* ctx, int set, size_t off_bytes, void* buf, int n, int kind, cudaStream_t st,
                             const int* done, int* fault) {
  if (n > PMX_SMALL_MAX || (kind != 1 && kind != 2)) {
    pmx_set_error("pmx_peer_small_allreduce: n = %d, kind = %d not supported", n, kind);
    return PMX_ERR_ARG;
  }
  pmx_peer_ptrs ar, fl;
  for (int r = 0; r < PMX_MAX_WORLD; ++r) {
    ar.p[r] = ctx->peer_arena.peer[r];
    fl.p[r] = ctx->peer_flags.peer[r];
  }
  k_small_allreduce<<<1, 256, 0, st>>>(ar, off_bytes, buf, n, kind, ctx->peer_epoch, fl,
                                       reinterpret_cast<const unsigned*>(ctx->peer_flags.local), set, ctx->world, ctx->rank,
                                       done, fault);
  PMX_LAUNCHED(ctx);
  return pmx_check_launch(ctx, "k_small_allreduce");
}

int pmx_peer_setup_internal(pmx_ctx* ctx) {
  ctx->peer_ok = 0;
  if (ctx->world <= 1 || ctx->world > PMX_MAX_WORLD || getenv("PMX_NO_PEER")) return PMX_OK;
  if (peer_alloc(ctx, sizeof(unsigned) * PMX_PEER_SETS * PMX_MAX_WORLD, &ctx->peer_flags) != PMX_OK) return PMX_OK;
  PMX_CUDA(cudaMalloc((void**)&ctx->peer_epoch, sizeof(unsigned) * PMX_PEER_SETS));
  PMX_CUDA(cudaMemset(ctx->peer_epoch, 0, sizeof(unsigned) * PMX_PEER_SETS));
  ctx->peer_ok = 1;
  return PMX_OK;
}

int pmx_peer_teardown_internal(pmx_ctx* ctx) {
  if (ctx->peer_arena.local) peer_close(ctx, &ctx->peer_arena);
  if (ctx->peer_flags.local) peer_close(ctx, &ctx->peer_flags);
  if (ctx->peer_epoch) cudaFree(ctx->peer_epoch);
  ctx->peer_epoch = nullptr;
  ctx->peer_ok = 0;
  return PMX_OK;
}

extern "C" {

int pmx_comm_unique_id(void* unique_id_128) {
  PMX_REQUIRE(unique_id_128 != nullptr, "NULL id buffer");
  PMX_CHECK(load_nccl());
  ncclUniqueId id;
  PMX_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(unique_id_128, &id, sizeof(id));
  return PMX_OK;
}

int pmx_comm_init(pmx_ctx* ctx, const void* unique_id_128, int world, int rank) {
  PMX_REQUIRE(ctx && unique_id_128, "NULL argument");
  PMX_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad world/rank");
  ctx->world = world;
  ctx->rank = rank;
  if (world == 1) return PMX_OK;
  PMX_CHECK(load_nccl());
  PMX_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, unique_id_128, sizeof(id));
  ncclComm_t comm;
  PMX_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  ctx->nccl_comm = comm;
  ctx->nccl_comm_aux = nullptr;
  if (g_nccl.CommSplit && !getenv("PMX_NO_AUX_COMM")) {
    ncclComm_t aux = nullptr;
    if (g_nccl.CommSplit(comm, 0, rank, &aux, nullptr) == 0) ctx->nccl_comm_aux = aux;
  }
  return pmx_peer_setup_internal(ctx);
}

int pmx_comm_peer_enabled(pmx_ctx* ctx, int* enabled) {
  PMX_REQUIRE(ctx && enabled, "NULL argument");
  *enabled = pmx_peer_available(ctx) ? 1 : 0;
  return PMX_OK;
}

int pmx_comm_allreduce_sum(pmx_ctx* ctx, float* dev_buf, size_t count) {
  PMX_REQUIRE(ctx && dev_buf, "NULL argument");
  return pmx_comm_allreduce_internal(ctx, dev_buf, count, 0, ctx->stream);
}

}  // extern "C"
