// Micro-timings (single CTA, clock64): TMA bulk-copy round trip (L2-resident source) and tcgen05.mma kind::tf32
// latency / throughput for M=128, N=128|64, K=8 with SWIZZLE_128B A and SWIZZLE_64B B operands.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) { return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61); }
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) { return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61); }
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__global__ void timing(const float* wsrc, long long* out, int N) {
  extern __shared__ unsigned char smem[];
  const uint32_t raw = smem_u32(smem);
  const uint32_t base = raw + ((1024 - (raw & 1023)) & 1023);
  __shared__ uint64_t bars[8];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 48 * 1024; i += blockDim.x) reinterpret_cast<float*>(smem + (base - raw))[i] = 0.001f * (i % 97);
  if (tid == 0) { for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (tid == 0) {
    int k = 0;
    // ---- TMA bulk copy round trips: 16 KB and 8 KB, first touch then L2-warm, then 4 in flight
    const uint32_t dst = base + 96 * 1024;
    for (int rep = 0; rep < 3; ++rep) {
      for (int bytes = 16384; bytes >= 8192; bytes >>= 1) {
        const uint32_t bar = smem_u32(&bars[0]);
        static int ph = 0;
        long long t0 = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(wsrc), "r"(bytes), "r"(bar) : "memory");
        mbar_wait(bar, ph & 1);
        ++ph;
        out[k++] = clock64() - t0;
      }
    }
    {   // 4 x 16 KB in flight on 4 barriers
      long long t0 = clock64();
      for (int i = 0; i < 4; ++i) {
        const uint32_t bar = smem_u32(&bars[1 + i]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(16384) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + i * 16384), "l"(wsrc + i * 4096), "r"(16384), "r"(bar) : "memory");
      }
      for (int i = 0; i < 4; ++i) mbar_wait(smem_u32(&bars[1 + i]), 0);
      out[k++] = clock64() - t0;
    }
    // ---- MMA: latency of 1, 6, 12, 48, 96 back-to-back MMAs (issue -> commit -> wait)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t bar = smem_u32(&bars[5]);
    int ph = 0;
    const int counts[6] = {1, 6, 12, 48, 96, 96};
    for (int c = 0; c < 6; ++c) {
      long long t0 = clock64();
      for (int i = 0; i < counts[c]; ++i)
        mma_tf32(tmem, desc_sw128(base + (i % 4) * 32 + (i % 3) * 2048), desc_sw64(base + 32768 + (i % 2) * 32), idesc, i > 0);
      long long t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
      mbar_wait(bar, ph & 1);
      ++ph;
      out[k++] = t1 - t0;           // issue time
      out[k++] = clock64() - t0;    // until complete
    }
    out[k++] = -1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}
int main() {
  float* w; long long* out; long long h[64];
  CK(cudaMalloc(&w, 1 << 20)); CK(cudaMemset(w, 0, 1 << 20)); CK(cudaMalloc(&out, 64 * 8));
  for (int N : {256, 192, 128, 64, 32, 16}) {
    CK(cudaMemset(out, 0, 64 * 8));
    CK(cudaFuncSetAttribute(timing, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    timing<<<1, 128, 200 * 1024>>>(w, out, N);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out, 64 * 8, cudaMemcpyDeviceToHost));
    printf("N=%d\n  TMA bulk round trip clks (16K,8K x3 reps): %lld %lld | %lld %lld | %lld %lld ; 4x16K in flight: %lld\n", N, h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    const int counts[6] = {1, 6, 12, 48, 96, 96};
    for (int c = 0; c < 6; ++c) printf("  %3d MMAs: issue %lld clks, complete %lld clks (%.1f clk/MMA)\n", counts[c], h[7 + 2 * c], h[8 + 2 * c], (double)h[8 + 2 * c] / counts[c]);
  }
  return 0;
}
