// Micro-timing of the MMA issue loop structure of the pair kernels (single CTA, single issuing thread):
// 96 x tcgen05.mma kind::tf32 M=128 N=128 K=8 in groups of 6, with different per-group bookkeeping.
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
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}
__device__ __forceinline__ void mma_w(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// mode bits: 1 = commit per group, 2 = arrive+try_wait on an (already complete) barrier per group, 4 = test_wait per group,
//            8 = tcgen05.fence::after_thread_sync per group, 16 = modulo arithmetic per group, 32 = whole warp walks the loop (lane 0 issues)
__global__ void timing(long long* out, int mode, int N) {
  extern __shared__ unsigned char smem[];
  const uint32_t raw = smem_u32(smem);
  const uint32_t base = raw + ((1024 - (raw & 1023)) & 1023);
  __shared__ uint64_t bars[16];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 40 * 1024; i += blockDim.x) reinterpret_cast<float*>(smem + (base - raw))[i] = 0.001f * (i % 97);
  if (tid == 0) { for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const bool whole = (mode & 32) != 0;
  if (tid < 32 && (whole || tid == 0)) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hiw = (1024u >> 4) | (1u << 14) | (2u << 29), b_hiw = (512u >> 4) | (1u << 14) | (4u << 29);
    const uint32_t a0 = ((base & 0x3FFFF) >> 4) | (1u << 16), b0 = (((base + 81920) & 0x3FFFF) >> 4) | (1u << 16);
    const uint32_t bdone = smem_u32(&bars[15]);
    // pre-complete barrier 8 phase 0 so waits on it succeed immediately
    if (tid == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[8])) : "memory");
    if (whole) __syncwarp();
    uint32_t k = 0;
    long long t0 = clock64();
    for (int g = 0; g < 16; ++g) {
      uint32_t st = g & 3;
      if (mode & 16) st = k % 3 + (k / 3) % 2;
      if (mode & 2) mbar_wait(smem_u32(&bars[8]), 0);
      if (mode & 8) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t nr = 0;
      if (mode & 4) nr = mbar_test(smem_u32(&bars[8]), 0);
      const uint32_t ah = a0 + ((g & 3) * 2048 >> 4), al = ah + (40960 >> 4), bh = b0 + st * (16384 >> 4), bl = bh + (8192 >> 4);
      if (tid == 0 && (mode & 64)) {          // A-row-major order: both K steps of a 16-channel chunk back to back per A image
        mma_w(tmem, ah, a_hiw, bh, b_hiw, idesc, g > 0);
        mma_w(tmem, ah + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
        mma_w(tmem, ah, a_hiw, bl, b_hiw, idesc, 1u);
        mma_w(tmem, ah + 2, a_hiw, bl + 2, b_hiw, idesc, 1u);
        mma_w(tmem, al, a_hiw, bh, b_hiw, idesc, 1u);
        mma_w(tmem, al + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
        if (mode & 1) commit(smem_u32(&bars[st]));
      } else if (tid == 0 && (mode & 128)) {  // four K steps of one 128-byte row back to back (32-channel groups), hi A then lo A
        const uint32_t s2 = (g & 1) * 4;
        if ((g & 1) == 0 || true) {
          mma_w(tmem, ah + s2 + 0, a_hiw, bh + 0, b_hiw, idesc, g > 0);
          mma_w(tmem, ah + s2 + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
          mma_w(tmem, ah + s2 + 0, a_hiw, bl + 0, b_hiw, idesc, 1u);
          mma_w(tmem, ah + s2 + 2, a_hiw, bl + 2, b_hiw, idesc, 1u);
          mma_w(tmem, al + s2 + 0, a_hiw, bh + 0, b_hiw, idesc, 1u);
          mma_w(tmem, al + s2 + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
        }
      } else if (tid == 0 && (mode & 256)) {  // same A descriptor six times (upper bound of any row reuse)
        for (int r = 0; r < 6; ++r) mma_w(tmem, ah, a_hiw, bh + 2 * (r & 1), b_hiw, idesc, (g | r) > 0);
      } else if (tid == 0) {
        mma_w(tmem, ah, a_hiw, bh, b_hiw, idesc, g > 0);
        mma_w(tmem, al, a_hiw, bh, b_hiw, idesc, 1u);
        mma_w(tmem, ah, a_hiw, bl, b_hiw, idesc, 1u);
        mma_w(tmem, ah + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
        mma_w(tmem, al + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
        mma_w(tmem, ah + 2, a_hiw, bl + 2, b_hiw, idesc, 1u);
        if (mode & 1) commit(smem_u32(&bars[st]));
      }
      if (whole) __syncwarp();
      k += nr + 1;
    }
    long long t1 = clock64();
    if (tid == 0) commit(bdone);
    mbar_wait(bdone, 0);
    if (tid == 0) { out[0] = t1 - t0; out[1] = clock64() - t0; out[2] = k; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}
int main() {
  long long* out; long long h[4];
  CK(cudaMalloc(&out, 64));
  CK(cudaFuncSetAttribute(timing, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int modes[] = {0, 64, 128, 256, 65, 71};
  for (int N : {128, 64})
    for (int m : modes) {
      for (int rep = 0; rep < 2; ++rep) { timing<<<1, 128, 200 * 1024>>>(out, m, N); CK(cudaDeviceSynchronize()); }
      CK(cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost));
      printf("N=%d mode=%2d: issue %lld clks, complete %lld clks  (%.1f clk/MMA)\n", N, m, h[0], h[1], (double)h[1] / 96);
    }
  return 0;
}
