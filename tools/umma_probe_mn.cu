// Probe: tcgen05.mma kind::tf32 with MN-major SWIZZLE_128B operands (the NHWC activation layout used as the
// A (=act^T) and B (=grad^T) operands of the weight-gradient GEMM  dW[ci][co] = sum_pixels A[pix][ci] * G[pix][co]).
// smem layout per operand: [slab = channel/32][pixel row][128 B], SWIZZLE_128B on the absolute address.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe_mn tools/umma_probe_mn.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ int g_variant = 0;   // 0: SWIZZLE_128B (16-byte chunks)   1: SWIZZLE_128B_BASE32B, SBO 512   2: BASE32B, SBO 1024
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint64_t lt = g_variant == 0 ? 2 : 1;
  if (g_variant == 1) sbo_bytes = 512;
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | (lt << 61);
}
__device__ __forceinline__ uint32_t elem_addr(uint32_t ra, int c) {   // ra = row base address, c = channel in the 32-wide slab
  if (g_variant == 0) return ra + (((c / 4) ^ ((ra >> 7) & 7)) << 4) + (c % 4) * 4;
  return ra + (((c / 8) ^ ((ra >> 7) & 3)) << 5) + (c % 8) * 4;      // Swizzle<2,5,2>: 32-byte chunks, key = row % 4
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_rna(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }

// A: [rowsA pixels][CA channels], B: [rowsB pixels][CB]; D[ca][cb] = sum_{k<K} A[rowA0+k][ca] * B[rowB0+k][cb]
template <int CA, int CB>
__global__ void probe(const float* A, const float* B, float* D, int rowsA, int rowsB, int K, int rowA0, int rowB0) {
  extern __shared__ unsigned char smem[];
  const uint32_t raw = smem_u32(smem);
  const uint32_t base = raw + ((1024 - (raw & 1023)) & 1023);
  unsigned char* gen = smem + (base - raw);
  const uint32_t slabA = rowsA * 128, slabB = rowsB * 128;
  const uint32_t a_off = 0, b_off = (CA / 32) * slabA;
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < rowsA * CA; i += blockDim.x) {
    const int r = i / CA, c = i % CA;
    const uint32_t ra = base + a_off + (c / 32) * slabA + r * 128;
    const uint32_t ad = elem_addr(ra, c % 32);
    *reinterpret_cast<float*>(gen + (ad - base)) = tf32_rna(A[i]);
  }
  for (int i = tid; i < rowsB * CB; i += blockDim.x) {
    const int r = i / CB, c = i % CB;
    const uint32_t ra = base + b_off + (c / 32) * slabB + r * 128;
    const uint32_t ad = elem_addr(ra, c % 32);
    *reinterpret_cast<float*>(gen + (ad - base)) = tf32_rna(B[i]);
  }
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(128) : "memory"); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (tid == 0) {
    // M = CA (128), N = CB, both operands MN-major
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(CB >> 3) << 17) | ((uint32_t)(CA >> 4) << 24);
    for (int k = 0; k < K; k += 8) {
      const uint64_t ad = desc_mn_sw128(base + a_off + (rowA0 + k) * 128, slabA, 1024);
      const uint64_t bd = desc_mn_sw128(base + b_off + (rowB0 + k) * 128, slabB, 1024);
      mma_tf32(tmem, ad, bd, idesc, k > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  { uint32_t done = 0; while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory"); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid < 128) {
    for (int c0 = 0; c0 < CB; c0 += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)((tid >> 5) * 32) << 16) + c0, v);
      if (CA == 128) { for (int j = 0; j < 32; ++j) D[(size_t)tid * CB + c0 + j] = v[j]; }
      else if ((tid & 31) < 16) {   // M = 64: accumulator row r lives in TMEM lane (r % 16) + 32 * (r / 16)
        const int row = (tid >> 5) * 16 + (tid & 31);
        for (int j = 0; j < 32; ++j) D[(size_t)row * CB + c0 + j] = v[j];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}
static float tf32h(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }
template <int CA, int CB>
static int run(int rowsA, int rowsB, int K, int rowA0, int rowB0) {
  std::vector<float> A((size_t)rowsA * CA), B((size_t)rowsB * CB), D((size_t)CA * CB);
  srand(7 + rowsA + K + rowA0 * 3 + rowB0 + CA + CB);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, D.size() * 4));
  size_t smem = (size_t)(CA / 32) * rowsA * 128 + (size_t)(CB / 32) * rowsB * 128 + 2048;
  CK(cudaFuncSetAttribute(probe<CA, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<CA, CB><<<1, 256, smem>>>(dA, dB, dD, rowsA, rowsB, K, rowA0, rowB0);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int i = 0; i < CA; ++i) for (int j = 0; j < CB; ++j) {
    double ref = 0;
    for (int k = 0; k < K; ++k) ref += (double)tf32h(A[(size_t)(rowA0 + k) * CA + i]) * tf32h(B[(size_t)(rowB0 + k) * CB + j]);
    maxref = fmax(maxref, fabs(ref)); maxerr = fmax(maxerr, fabs(D[(size_t)i * CB + j] - ref));
  }
  const double e = maxerr / maxref;
  printf("   D[0][0..3] = %g %g %g %g  D[5][7]=%g\n", D[0], D[1], D[2], D[3], D[5 * CB + 7]);
  printf("MN-major CA=%3d CB=%3d rowsA=%3d rowsB=%3d K=%3d rowA0=%2d rowB0=%2d : rel err %.3e %s\n", CA, CB, rowsA, rowsB, K, rowA0, rowB0, e, e < 2e-5 ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return e < 2e-5 ? 0 : 1;
}
int main() {
  int f = 0;
  for (int variant = 1; variant < 2; ++variant) {
  printf("==== variant %d\n", variant);
  CK(cudaMemcpyToSymbol(g_variant, &variant, sizeof(int)));
  f += run<128, 128>(32, 32, 32, 0, 0);
  f += run<128, 128>(64, 32, 32, 16, 0);
  f += run<128, 128>(64, 32, 32, 1, 0);
  f += run<128, 128>(64, 48, 32, 17, 5);
  f += run<128, 64>(64, 32, 32, 2, 0);
  f += run<64, 64>(48, 32, 32, 3, 0);
  f += run<64, 128>(48, 32, 16, 3, 8);
  }
  printf(f ? "MN PROBE FAILED (%d)\n" : "MN PROBE OK\n", f);
  return f;
}
