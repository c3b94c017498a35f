// Can two warps issue tcgen05.mma concurrently?  NW issuing warps, each 96 MMAs (M=128,N,K=8 tf32) into its own accumulator.
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
__device__ __forceinline__ void mma64(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t dA(uint32_t s) { return (uint64_t)((s & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61); }
__device__ __forceinline__ uint64_t dB(uint32_t s) { return (uint64_t)((s & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61); }
template <int UNROLL>
__global__ void timing(long long* out, int NW, int N) {
  extern __shared__ unsigned char smem[];
  const uint32_t raw = smem_u32(smem);
  const uint32_t base = raw + ((1024 - (raw & 1023)) & 1023);
  __shared__ uint64_t bars[8];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 40 * 1024; i += blockDim.x) reinterpret_cast<float*>(smem + (base - raw))[i] = 0.001f * (i % 97);
  if (tid == 0) { for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp < NW && lane == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t acc = tmem + warp * 128;
    const uint64_t a0 = dA(base + warp * 20480), b0 = dB(base + 81920 + warp * 16384);
    long long t0 = clock64();
#pragma unroll UNROLL
    for (int i = 0; i < 96; ++i) mma64(acc, a0 + (uint64_t)((i & 3) * 2), b0 + (uint64_t)((i & 1) * 2), idesc, i > 0);
    long long t1 = clock64();
    commit(smem_u32(&bars[warp]));
    mbar_wait(smem_u32(&bars[warp]), 0);
    out[warp * 2] = t1 - t0; out[warp * 2 + 1] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
  long long* out; long long h[8];
  CK(cudaMalloc(&out, 64));
  CK(cudaFuncSetAttribute(timing<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(timing<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int N : {128, 64})
    for (int nw : {1, 2, 4})
      for (int u : {1, 8}) {
        for (int rep = 0; rep < 2; ++rep) { if (u == 1) timing<1><<<1, 128, 200 * 1024>>>(out, nw, N); else timing<8><<<1, 128, 200 * 1024>>>(out, nw, N); CK(cudaDeviceSynchronize()); }
        CK(cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost));
        printf("N=%d warps=%d unroll=%d:", N, nw, u);
        for (int w = 0; w < nw; ++w) printf("  w%d issue %lld complete %lld", w, h[2 * w], h[2 * w + 1]);
        printf("   => %.1f clk per MMA overall\n", (double)h[1] / (96.0 * nw));
      }
  return 0;
}
