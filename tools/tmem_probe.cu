// Which (lane, column) does each register of tcgen05.ld.16x256b.x2 hold?  Write lane*1000+col with 32x32b stores, read back.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(float* out) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  // each thread writes its lane's 16 columns: value = lane_global*1000 + col
  {
    uint32_t v[16];
    for (int c = 0; c < 16; ++c) v[c] = __float_as_uint((float)((warp * 32 + lane) * 1000 + c));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tmem + lane_addr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int h = 0; h < 2; ++h) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(tmem + lane_addr + ((uint32_t)(16 * h) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[((warp * 2 + h) * 32 + lane) * 8 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}
int main() {
  float* d; float h[4 * 2 * 32 * 8];
  cudaMalloc(&d, sizeof(h));
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int w = 0; w < 2; ++w)
    for (int hh = 0; hh < 2; ++hh) {
      printf("warp %d half %d:\n", w, hh);
      for (int l = 0; l < 32; l += (l < 8 ? 1 : 8)) {
        printf("  lane %2d:", l);
        for (int i = 0; i < 8; ++i) printf(" %7.0f", h[((w * 2 + hh) * 32 + l) * 8 + i]);
        printf("\n");
      }
    }
  return 0;
}
