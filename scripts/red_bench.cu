// Micro-benchmark: throughput of fp32 reductions into an L2-resident array, the flush pattern of k_grad_umma
// (64 x 128 tile per step, 148 CTAs, 4 flush warps) done with (A) red.global.add.v4.f32, (B) scalar red,
// (C) cp.reduce.async.bulk of 512-byte rows staged in shared memory, (D) one 32 KB cp.reduce.async.bulk per tile
// into a tile-major array.       nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench.bin red_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int K = 64, N = 65536, NS = N / 128, MB = 64;   // 64 m-blocks x 512 stripes of partial tiles

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(128, 1) k_red(float* G, int tiles_per_cta) {
  extern __shared__ __align__(128) float stage[];   // 64 x 128 fp32
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long g0 = (long long)blockIdx.x * tiles_per_cta;
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = 1e-3f * (float)(i + lane);
  for (int t = 0; t < tiles_per_cta; ++t) {
    const int stripe = (int)((g0 + t) % NS);
    if (MODE == 0) {
      const int nq = stripe * 128 + warp * 32 + (lane & ~3), r = lane & 3;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = 4 * i + r;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(G + (size_t)k * N + nq), "f"(v[4 * i]),
                     "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                     : "memory");
      }
    } else if (MODE == 1) {
      const int n = stripe * 128 + warp * 32 + lane;
#pragma unroll
      for (int k = 0; k < 64; ++k) atomicAdd(G + (size_t)k * N + n, v[k]);
    } else {
      // stage the tile (row k, 128 n) then bulk-reduce
      if (t > 0) {
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
      }
      const int n = warp * 32 + lane;
#pragma unroll
      for (int k = 0; k < 64; ++k) stage[k * 128 + n] = v[k];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (MODE == 2) {
        if (threadIdx.x < 64) {
          const int k = threadIdx.x;
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                           G + (size_t)k * N + stripe * 128),
                       "r"(smem_u32(stage + k * 128)), "r"(512)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        if (threadIdx.x == 0) {
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                           G + (size_t)stripe * 64 * 128),
                       "r"(smem_u32(stage)), "r"(32768)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (MODE == 2 && threadIdx.x < 64 && t == tiles_per_cta - 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      if (MODE == 3 && threadIdx.x == 0 && t == tiles_per_cta - 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
}

// W flush warps (the 64 k-rows are split among W/4 warps per 32-column group) + BG warps streaming a big buffer
// through L2 (ld.global.nc.v4, evict_first-like streaming) to model the Y stream
template <int W, int BG>
__global__ void __launch_bounds__((W + BG) * 32, 1) k_red_bg(float* G, int tiles_per_cta, const float4* Y, size_t y_per_cta,
                                                             float* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= W) {
    const float4* y = Y + (size_t)blockIdx.x * y_per_cta;
    float4 acc = make_float4(0, 0, 0, 0);
    for (size_t i = (warp - W) * 32 + lane; i < y_per_cta; i += BG * 32) {
      float4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(y + i));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.f) sink[0] = acc.x;
    return;
  }
  const long long g0 = (long long)blockIdx.x * tiles_per_cta;
  float v[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = 1e-3f * (float)(i + lane);
  constexpr int SPLIT = W / 4;            // warps per 32-column group
  const int grp = warp & 3, part = warp >> 2;
  for (int t = 0; t < tiles_per_cta; ++t) {
    const int stripe = (int)((g0 + t) % NS);
    const int nq = stripe * 128 + grp * 32 + (lane & ~3), r = lane & 3;
#pragma unroll
    for (int i = 0; i < 16 / SPLIT; ++i) {
      const int k = 4 * (i + part * (16 / SPLIT)) + r;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(G + (size_t)k * N + nq), "f"(v[4 * i]),
                   "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                   : "memory");
    }
  }
}

template <int W, int BG>
void run_bg(const char* name, float* G, int sms, const float4* Y, size_t y_elems, float* sink, bool stream) {
  const long long tiles = (long long)MB * NS;
  const int per = (int)(tiles / sms);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t ypc = stream ? y_elems / sms : 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(G, 0, sizeof(float) * K * N);
    cudaEventRecord(e0);
    k_red_bg<W, BG><<<sms, (W + BG) * 32>>>(G, per, Y, ypc, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)per * sms * 32768.0;
    if (rep == 2)
      printf("%-44s %8.3f ms  red %7.1f GB/s  stream %7.1f GB/s (%s)\n", name, ms, bytes / ms * 1e-6,
             (double)ypc * sms * 16 / ms * 1e-6, cudaGetErrorString(e));
  }
}

template <int MODE>
void run(const char* name, float* G, int sms) {
  const long long tiles = (long long)MB * NS;
  const int per = (int)(tiles / sms);
  cudaFuncSetAttribute(k_red<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(G, 0, sizeof(float) * K * N);
    cudaEventRecord(e0);
    k_red<MODE><<<sms, 128, 32768>>>(G, per);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)per * sms * 32768.0;
    if (rep == 2)
      printf("%-34s %8.3f ms  %7.1f GB/s of reduced data   (%s)\n", name, ms, bytes / ms * 1e-6, cudaGetErrorString(e));
  }
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* G;
  cudaMalloc(&G, sizeof(float) * K * N);
  run<0>("red.global.add.v4.f32", G, sms);
  run<1>("red.global.add.f32 (scalar)", G, sms);
  run<2>("cp.reduce.async.bulk 512 B rows", G, sms);
  run<3>("cp.reduce.async.bulk 32 KB tile", G, sms);
  float4* Y; float* sink;
  const size_t y_elems = (size_t)2147483648ull / 16;
  cudaMalloc(&Y, y_elems * 16); cudaMalloc(&sink, 16);
  cudaMemset(Y, 0, y_elems * 16);
  run_bg<4, 8>("4 red warps, no stream", G, sms, Y, y_elems, sink, false);
  run_bg<8, 8>("8 red warps, no stream", G, sms, Y, y_elems, sink, false);
  run_bg<16, 8>("16 red warps, no stream", G, sms, Y, y_elems, sink, false);
  run_bg<4, 8>("no reds (0 tiles) + 2 GB stream", G, sms, Y, y_elems, sink, true);
  run_bg<4, 8>("4 red warps + 2 GB stream (8 warps)", G, sms, Y, y_elems, sink, true);
  run_bg<8, 8>("8 red warps + 2 GB stream (8 warps)", G, sms, Y, y_elems, sink, true);
  // sanity: value of one element after mode 3
  float h = 0;
  cudaMemcpy(&h, G, 4, cudaMemcpyDeviceToHost);
  printf("G[0] after the last run = %g\n", h);
  return 0;
}
