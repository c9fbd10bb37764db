// Microbenchmark: cycles per tcgen05.mma for SS (A,B in smem) and TS (A in TMEM) operand modes, tf32 vs bf16,
// N = 64/128/256, one CTA per SM, back-to-back accumulating MMAs on fixed shared-memory data.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_floor tools/mma_floor.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  uint64_t d = (uint64_t)((addr & 0x3ffffu) >> 4);
  d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int m, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int KIND /*0 tf32, 1 bf16*/, int N, int TS, int KSTEPS>
__global__ void __launch_bounds__(128, 1) floor_kernel(long long* cycles, int iters) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  // zero the operand area (A: 16 KB x KSTEPS/4, B: 32 KB ...) -- contents are irrelevant, but avoid NaN slow paths
  for (uint32_t i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0x3f800000u));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc(KIND == 0 ? 2 : 1, 128, N);
    const uint32_t a_addr = base, b_addr = base + 64 * 1024;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < KSTEPS; ++k) {
        // walk through distinct k-blocks like a real mainloop: 4 k-steps of 32 B per 128-B swizzle row, then next 16/32 KB block
        const uint64_t adesc = smem_desc(a_addr + (k / 4) * 16384) + (uint64_t)((k % 4) * 2);
        const uint64_t bdesc = smem_desc(b_addr + (k / 4) * (N * 128)) + (uint64_t)((k % 4) * 2);
        const uint32_t d = tmem + ((it & 1) ? 0 : 0);
        if (TS) {
          const uint32_t a_t = tmem + 256 + (k % 32) * 8;   // A in TMEM: 128 lanes x 8 columns (tf32) per k-step
          if (KIND == 0)
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}"
                         ::"r"(d), "r"(a_t), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}"
                         ::"r"(d), "r"(a_t), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
        } else {
          if (KIND == 0)
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}"
                         ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}"
                         ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int KIND, int N, int TS>
void run(const char* name) {
  constexpr int KSTEPS = 8;
  fprintf(stderr, "[%s] start\n", name); fflush(stderr);
  long long* d = nullptr;
  cudaError_t e = cudaMalloc(&d, 8);
  fprintf(stderr, "[%s] malloc %d\n", name, (int)e); fflush(stderr);
  auto kern = floor_kernel<KIND, N, TS, KSTEPS>;
  const int smem = 162 * 1024;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  fprintf(stderr, "[%s] attr %d\n", name, (int)e); fflush(stderr);
  const int iters = 2000;
  kern<<<148, 128, smem>>>(d, 10);
  e = cudaDeviceSynchronize();
  fprintf(stderr, "[%s] warm %d %s\n", name, (int)e, cudaGetErrorString(e)); fflush(stderr);
  if (e != cudaSuccess) return;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<148, 128, smem>>>(d, iters);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  const double n_mma = (double)iters * KSTEPS;
  const double kk = KIND == 0 ? 8 : 16;
  const double flops = 2.0 * 128 * N * kk * n_mma * 148;
  printf("%-28s err=%d  cycles/MMA=%7.1f  ms=%8.3f  TFLOP/s=%8.1f  (clock %.0f MHz)\n", name, (int)err, c / n_mma, ms,
         flops / ms / 1e9, c / (ms * 1e3));
  cudaFree(d);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int ndev = 0; cudaError_t e0 = cudaGetDeviceCount(&ndev);
  fprintf(stderr, "devices %d err %d\n", ndev, (int)e0); fflush(stderr);
  run<0, 256, 0>("tf32 SS M128 N256");
  run<0, 128, 0>("tf32 SS M128 N128");
  run<0, 64, 0>("tf32 SS M128 N64");
  run<1, 256, 0>("bf16 SS M128 N256");
  run<1, 128, 0>("bf16 SS M128 N128");
  run<0, 256, 1>("tf32 TS M128 N256");
  run<0, 128, 1>("tf32 TS M128 N128");
  run<1, 256, 1>("bf16 TS M128 N256");
  return 0;
}
