// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128) for the operand configurations used by the
// gradient kernel.  One CTA; 64 back-to-back MMAs per measurement; garbage operands (timing only).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int amn, int bmn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)amn << 15) | ((uint32_t)bmn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}

// mode: 0 SS K/K N=128 | 1 SS A=K,B=MN N=128 | 2 SS K/K N=64 | 3 SS MN/MN N=64 | 4 TS B=K N=64 | 5 TS B=MN N=128
//       6 SS MN/MN N=128 | 7 SS A=MN,B=K N=64 | 8 TS B=K N=128 | 9 SS K/K N=256
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_bench(int reps, long long* out) {
  constexpr int mode = MODE;
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar_storage;
  __shared__ uint32_t tmem_slot;
  const uint32_t bar = smem_u32(&bar_storage);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3f803f80u;  // bf16 1.0 pairs
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t A = base, B = base + 64 * 1024;
    uint32_t parity = 0;
    for (int rep = 0; rep < reps; ++rep) {
      const long long t0 = clock64();
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const uint32_t ks = i & 7;
        switch (mode) {
          case 0: mma_ss(tmem, make_desc(A + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), make_desc(B + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), make_idesc(128, 128, 0, 0), i); break;
          case 1: mma_ss(tmem, make_desc(A + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), make_desc(B + ks * 2048, 16384, 1024), make_idesc(128, 128, 0, 1), i); break;
          case 2: mma_ss(tmem, make_desc(A + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), make_desc(B + (ks >> 2) * 8192 + (ks & 3) * 32, 16, 1024), make_idesc(128, 64, 0, 0), i); break;
          case 3: mma_ss(tmem, make_desc(A + ks * 2048, 16384, 1024), make_desc(B + ks * 2048, 1024, 1024), make_idesc(128, 64, 1, 1), i); break;
          case 4: mma_ts(tmem, tmem + 256 + ks * 8, make_desc(B + (ks >> 2) * 8192 + (ks & 3) * 32, 16, 1024), make_idesc(128, 64, 0, 0), i); break;
          case 5: mma_ts(tmem, tmem + 256 + ks * 8, make_desc(B + ks * 2048, 16384, 1024), make_idesc(128, 128, 0, 1), i); break;
          case 6: mma_ss(tmem, make_desc(A + ks * 2048, 16384, 1024), make_desc(B + ks * 2048, 16384, 1024), make_idesc(128, 128, 1, 1), i); break;
          case 7: mma_ss(tmem, make_desc(A + ks * 2048, 16384, 1024), make_desc(B + (ks >> 2) * 8192 + (ks & 3) * 32, 16, 1024), make_idesc(128, 64, 1, 0), i); break;
          case 8: mma_ts(tmem, tmem + 256 + ks * 8, make_desc(B + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), make_idesc(128, 128, 0, 0), i); break;
          case 9: mma_ss(tmem, make_desc(A + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024), make_desc(B + (ks >> 2) * 32768 + (ks & 3) * 32, 16, 1024), make_idesc(128, 256, 0, 0), i); break;
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      mbar_wait(bar, parity);
      parity ^= 1;
      const long long t1 = clock64();
      if (rep == reps - 1) out[mode] = t1 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 16 * sizeof(long long));
  cudaMemset(d, 0, 16 * sizeof(long long));
  const char* names[] = {"SS K/K N=128", "SS A=K B=MN N=128", "SS K/K N=64", "SS MN/MN N=64", "TS B=K N=64", "TS B=MN N=128",
                         "SS MN/MN N=128", "SS A=MN B=K N=64", "TS B=K N=128", "SS K/K N=256"};
  void (*kern[10])(int, long long*) = {k_bench<0>, k_bench<1>, k_bench<2>, k_bench<3>, k_bench<4>, k_bench<5>, k_bench<6>, k_bench<7>, k_bench<8>, k_bench<9>};
  for (int mode = 0; mode < 10; ++mode) {
    cudaFuncSetAttribute(kern[mode], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    kern[mode]<<<1, 128, 200 * 1024>>>(5, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
  }
  long long h[16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int mode = 0; mode < 10; ++mode) printf("%-22s %6lld cycles / 64 MMAs = %.1f per MMA\n", names[mode], h[mode], h[mode] / 64.0);
  return 0;
}
