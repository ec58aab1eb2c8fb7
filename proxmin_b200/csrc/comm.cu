// Multi-GPU plumbing: one process per GPU, NCCL all-reduce over NVLink 5 / NVSwitch on the
// context's stream.  NCCL is dlopen'ed on first use so that the single-GPU path has no link-time
// dependency on it.  The reference has no distributed code at all; the only exchange
// step the column-sharded NMF needs is sum(G_A partials) (+ K x K Gram and a few scalars).
#include <dlfcn.h>

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
} g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

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

int pmx_comm_destroy_internal(pmx_ctx* ctx) {
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
  PMX_NCCL(g_nccl.AllReduce(buf, buf, count, dt, op, (ncclComm_t)ctx->nccl_comm, st));
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
  return PMX_OK;
}

int pmx_comm_allreduce_sum(pmx_ctx* ctx, float* dev_buf, size_t count) {
  PMX_REQUIRE(ctx && dev_buf, "NULL argument");
  return pmx_comm_allreduce_internal(ctx, dev_buf, count, 0, ctx->stream);
}

}  // extern "C"
