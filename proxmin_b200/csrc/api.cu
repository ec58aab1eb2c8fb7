// extern "C" surface of libproxmin_b200.so: context, memory, standalone operators.
#include <stdarg.h>

#include <vector>

#include "kernels.h"

static thread_local char g_err[1024] = "";

void pmx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int pmx_check_launch(pmx_ctx* ctx, const char* what) {
  (void)ctx;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    pmx_set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return PMX_ERR_CUDA;
  }
  return PMX_OK;
}

int pmx_comm_destroy_internal(pmx_ctx* ctx);

// ------------------------------------------------------------------ cached device allocations
namespace {
struct DevBlock {
  void* p;
  size_t bytes;
  bool used;
};
struct DevPool {
  std::vector<DevBlock> blocks;
};
}  // namespace

int pmx_dev_alloc(pmx_ctx* ctx, void** out, size_t bytes) {
  if (bytes == 0) bytes = 4;
  if (!ctx->pool) ctx->pool = new DevPool();
  DevPool* pool = static_cast<DevPool*>(ctx->pool);
  for (DevBlock& b : pool->blocks)
    if (!b.used && b.bytes == bytes) {   // exact size: repeated solves of one shape reuse their buffers
      b.used = true;
      *out = b.p;
      return PMX_OK;
    }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {   // give the cached blocks back and retry once
    cudaGetLastError();
    pmx_dev_trim(ctx);
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    pmx_set_error("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return PMX_ERR_CUDA;
  }
  pool->blocks.push_back({p, bytes, true});
  *out = p;
  return PMX_OK;
}

void pmx_dev_free(pmx_ctx* ctx, void* p) {
  if (!p) return;
  DevPool* pool = static_cast<DevPool*>(ctx->pool);
  if (pool)
    for (DevBlock& b : pool->blocks)
      if (b.p == p) {
        b.used = false;
        return;
      }
  cudaFree(p);   // not ours
}

void pmx_dev_trim(pmx_ctx* ctx) {
  DevPool* pool = static_cast<DevPool*>(ctx->pool);
  if (!pool) return;
  std::vector<DevBlock> keep;
  for (DevBlock& b : pool->blocks) {
    if (b.used) keep.push_back(b);
    else cudaFree(b.p);
  }
  pool->blocks.swap(keep);
}

extern "C" {

const char* pmx_last_error(void) { return g_err; }
int pmx_version(void) { return 100; }

int pmx_device_count(int* count) {
  PMX_REQUIRE(count != nullptr, "count is NULL");
  PMX_CUDA(cudaGetDeviceCount(count));
  return PMX_OK;
}

int pmx_ctx_create(int device, pmx_ctx** out) {
  PMX_REQUIRE(out != nullptr, "out is NULL");
  PMX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PMX_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    pmx_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return PMX_ERR_UNSUPPORTED;
  }
  pmx_ctx* c = new pmx_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->total_mem = prop.totalGlobalMem;
  snprintf(c->dev_name, sizeof(c->dev_name), "%s", prop.name);
  c->world = 1;
  PMX_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  PMX_CUDA(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
  PMX_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  PMX_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  PMX_CUDA(cudaEventCreateWithFlags(&c->ev_fork2, cudaEventDisableTiming));
  PMX_CUDA(cudaEventCreateWithFlags(&c->ev_join2, cudaEventDisableTiming));
  PMX_CUDA(cudaEventCreate(&c->ev_t0));
  PMX_CUDA(cudaEventCreate(&c->ev_t1));
  PMX_CUDA(cudaMallocHost((void**)&c->h_flags, 256));
  *out = c;
  return PMX_OK;
}

int pmx_ctx_destroy(pmx_ctx* ctx) {
  if (!ctx) return PMX_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->aux);
  pmx_comm_destroy_internal(ctx);
  cudaEventDestroy(ctx->ev_fork);
  cudaEventDestroy(ctx->ev_join);
  cudaEventDestroy(ctx->ev_fork2);
  cudaEventDestroy(ctx->ev_join2);
  cudaEventDestroy(ctx->ev_t0);
  cudaEventDestroy(ctx->ev_t1);
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->aux);
  cudaFreeHost(ctx->h_flags);
  for (int i = 0; i < 4; ++i)
    if (ctx->gram_scratch[i]) cudaFree(ctx->gram_scratch[i]);
  pmx_dev_trim(ctx);
  delete static_cast<DevPool*>(ctx->pool);
  if (ctx->prof_ev) {
    for (int i = 0; i < 2 * PMX_PROF_MAX; ++i) cudaEventDestroy(ctx->prof_ev[i]);
    delete[] ctx->prof_ev;
  }
  delete ctx;
  return PMX_OK;
}

int pmx_ctx_trim(pmx_ctx* ctx) {
  PMX_REQUIRE(ctx != nullptr, "NULL context");
  PMX_CUDA(cudaSetDevice(ctx->device));
  pmx_dev_trim(ctx);
  return PMX_OK;
}

int pmx_ctx_sync(pmx_ctx* ctx) {
  PMX_REQUIRE(ctx != nullptr, "ctx is NULL");
  PMX_CUDA(cudaStreamSynchronize(ctx->aux));
  PMX_CUDA(cudaStreamSynchronize(ctx->stream));
  return PMX_OK;
}

int pmx_ctx_launch_count(pmx_ctx* ctx, long long* count) {
  PMX_REQUIRE(ctx && count, "NULL argument");
  *count = ctx->launches;
  return PMX_OK;
}

int pmx_ctx_profile(pmx_ctx* ctx, int enable) {
  PMX_REQUIRE(ctx != nullptr, "ctx is NULL");
  if (enable && !ctx->prof_ev) {
    ctx->prof_ev = new cudaEvent_t[2 * PMX_PROF_MAX];
    for (int i = 0; i < 2 * PMX_PROF_MAX; ++i) PMX_CUDA(cudaEventCreate(&ctx->prof_ev[i]));
  }
  ctx->profile = enable ? 1 : 0;
  ctx->prof_n = 0;
  return PMX_OK;
}

int pmx_ctx_profile_read(pmx_ctx* ctx, float* total_ms, int* launches) {
  PMX_REQUIRE(ctx && total_ms && launches, "NULL argument");
  PMX_CUDA(cudaStreamSynchronize(ctx->stream));
  float tot = 0.f;
  for (int i = 0; i < ctx->prof_n; ++i) {
    float ms = 0.f;
    PMX_CUDA(cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
    tot += ms;
  }
  *total_ms = tot;
  *launches = ctx->prof_n;
  return PMX_OK;
}

int pmx_ctx_device_info(pmx_ctx* ctx, char* name, int name_len, int* sm_count, size_t* total_mem) {
  PMX_REQUIRE(ctx != nullptr, "ctx is NULL");
  if (name && name_len > 0) snprintf(name, name_len, "%s", ctx->dev_name);
  if (sm_count) *sm_count = ctx->sm_count;
  if (total_mem) *total_mem = ctx->total_mem;
  return PMX_OK;
}

int pmx_malloc(pmx_ctx* ctx, size_t bytes, void** dev_ptr) {
  PMX_REQUIRE(ctx && dev_ptr, "NULL argument");
  PMX_CUDA(cudaSetDevice(ctx->device));
  PMX_CUDA(cudaMalloc(dev_ptr, bytes ? bytes : 16));
  return PMX_OK;
}

int pmx_free(pmx_ctx* ctx, void* dev_ptr) {
  PMX_REQUIRE(ctx != nullptr, "ctx is NULL");
  if (dev_ptr) {
    PMX_CUDA(cudaStreamSynchronize(ctx->stream));
    PMX_CUDA(cudaFree(dev_ptr));
  }
  return PMX_OK;
}

int pmx_memset(pmx_ctx* ctx, void* dev_ptr, int byte, size_t bytes) {
  PMX_REQUIRE(ctx && dev_ptr, "NULL argument");
  PMX_CUDA(cudaMemsetAsync(dev_ptr, byte, bytes, ctx->stream));
  return PMX_OK;
}

int pmx_h2d(pmx_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes) {
  PMX_REQUIRE(ctx && dev_dst && host_src, "NULL argument");
  PMX_CUDA(cudaMemcpyAsync(dev_dst, host_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  PMX_CUDA(cudaStreamSynchronize(ctx->stream));  // the host buffer may be pageable / reused by the caller
  return PMX_OK;
}

int pmx_d2h(pmx_ctx* ctx, void* host_dst, const void* dev_src, size_t bytes) {
  PMX_REQUIRE(ctx && host_dst && dev_src, "NULL argument");
  PMX_CUDA(cudaMemcpyAsync(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PMX_CUDA(cudaStreamSynchronize(ctx->stream));
  return PMX_OK;
}

int pmx_d2d(pmx_ctx* ctx, void* dev_dst, const void* dev_src, size_t bytes) {
  PMX_REQUIRE(ctx && dev_dst && dev_src, "NULL argument");
  PMX_CUDA(cudaMemcpyAsync(dev_dst, dev_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return PMX_OK;
}

int pmx_host_alloc(size_t bytes, void** host_ptr) {
  PMX_REQUIRE(host_ptr != nullptr, "host_ptr is NULL");
  PMX_CUDA(cudaMallocHost(host_ptr, bytes ? bytes : 16));
  return PMX_OK;
}

int pmx_host_free(void* host_ptr) {
  if (host_ptr) PMX_CUDA(cudaFreeHost(host_ptr));
  return PMX_OK;
}

int pmx_timer_start(pmx_ctx* ctx) {
  PMX_REQUIRE(ctx != nullptr, "ctx is NULL");
  PMX_CUDA(cudaEventRecord(ctx->ev_t0, ctx->stream));
  return PMX_OK;
}

int pmx_timer_stop(pmx_ctx* ctx, float* ms) {
  PMX_REQUIRE(ctx && ms, "NULL argument");
  PMX_CUDA(cudaEventRecord(ctx->ev_t1, ctx->stream));
  PMX_CUDA(cudaEventSynchronize(ctx->ev_t1));
  PMX_CUDA(cudaEventElapsedTime(ms, ctx->ev_t0, ctx->ev_t1));
  return PMX_OK;
}

static int check_prox(const pmx_prox* p) {
  PMX_REQUIRE(p != nullptr, "prox is NULL");
  PMX_REQUIRE(p->n_ops >= 0 && p->n_ops <= PMX_MAX_OPS, "prox chain length out of range");
  for (int i = 0; i < p->n_ops; ++i) {
    PMX_REQUIRE(p->ops[i].op >= PMX_OP_ID && p->ops[i].op <= PMX_OP_MAXENT64, "unknown prox op code");
    if (p->ops[i].op == PMX_OP_UNITY) PMX_REQUIRE(p->ops[i].axis == 0 || p->ops[i].axis == 1, "UNITY axis must be 0 or 1");
  }
  return PMX_OK;
}

int pmx_prox_apply(pmx_ctx* ctx, const pmx_prox* prox, float* dev_X, int rows, int cols, float step) {
  PMX_REQUIRE(ctx && dev_X, "NULL argument");
  PMX_REQUIRE(rows >= 0 && cols >= 0, "negative shape");
  PMX_CHECK(check_prox(prox));
  ProxChain ch = make_chain(prox);
  UpdIO io;
  memset(&io, 0, sizeof(io));
  io.Xin = dev_X;
  io.Xout = dev_X;
  io.rows = rows;
  io.cols = cols;
  io.step.mode = 0;
  io.step.value = step;
  io.step.scale = 1.f;
  return launch_update(ctx, IN_PLAIN, ch, io);
}

int pmx_pgm_update(pmx_ctx* ctx, const pmx_prox* prox, const float* dev_Xe, const float* dev_G, float* dev_X, int rows,
                   int cols, float step, double* norm_diff_sq, double* norm_new_sq) {
  PMX_REQUIRE(ctx && dev_Xe && dev_G && dev_X, "NULL argument");
  PMX_CHECK(check_prox(prox));
  ProxChain ch = make_chain(prox);
  double* d_norms = nullptr;
  float* d_old = nullptr;
  const size_t n = (size_t)rows * cols;
  PMX_CUDA(cudaMalloc((void**)&d_norms, 3 * sizeof(double)));
  PMX_CUDA(cudaMemsetAsync(d_norms, 0, 3 * sizeof(double), ctx->stream));
  PMX_CUDA(cudaMalloc((void**)&d_old, (n ? n : 1) * sizeof(float)));
  UpdIO io;
  memset(&io, 0, sizeof(io));
  io.Xin = dev_Xe;
  io.G = dev_G;
  io.Xprev = dev_X;
  io.Xout = dev_X;
  io.Xold_out = d_old;
  io.norms = d_norms;
  io.rows = rows;
  io.cols = cols;
  io.step.mode = 0;
  io.step.value = step;
  io.step.scale = 1.f;
  int st = launch_update(ctx, IN_PGM, ch, io);
  double h[3] = {0, 0, 0};
  if (st == PMX_OK) {
    cudaError_t e = cudaMemcpyAsync(h, d_norms, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      pmx_set_error("pmx_pgm_update: %s", cudaGetErrorString(e));
      st = PMX_ERR_CUDA;
    }
  }
  cudaFree(d_norms);
  cudaFree(d_old);
  if (norm_diff_sq) *norm_diff_sq = h[0];
  if (norm_new_sq) *norm_new_sq = h[1];
  return st;
}

}  // extern "C"
