// Shared helpers for the sm_100a kernels of libmdil_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace mdil {

int set_error(int code, const char* what, const char* file, int line);
void note_launch();  // host-side count of kernels launched by this library

#define MDIL_LAUNCH_CHECK()                                                                         \
  do {                                                                                              \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess) return ::mdil::set_error((int)_e, cudaGetErrorString(_e), __FILE__, __LINE__); \
    ::mdil::note_launch();                                                                          \
  } while (0)

#define MDIL_CUDA(expr)                                                                             \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return ::mdil::set_error((int)_e, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define MDIL_REQUIRE(cond, msg)                                                                     \
  do {                                                                                              \
    if (!(cond)) return ::mdil::set_error(-1, msg, __FILE__, __LINE__);                             \
  } while (0)

#define MDIL_TRY(expr)                                                                              \
  do {                                                                                              \
    int _r = (expr);                                                                                \
    if (_r != 0) return _r;                                                                         \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200

// Launch-time state that belongs to a DEVICE (function attributes, occupancy answers): the library may be driven from
// several host threads for several devices of one process (nn.DataParallel, train_new_task_step2.py:473-475), so a
// plain function-level static would leave every device but the first unconfigured.  Races are benign (idempotent).
constexpr int kMaxDevices = 64;
static inline int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 make4(float v) { return make_float4(v, v, v, v); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace mdil
