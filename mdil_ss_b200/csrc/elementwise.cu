// Bandwidth-bound helper kernels: layout conversion, per-domain BatchNorm statistics / apply /
// backward, max-pool branch of the DownsamplerBlock, scaling, Adam.
// All activations NHWC fp32, 128-bit accesses along C, grids sized in multiples of the SM count.
#include "kernels.cuh"

namespace mdil {

static inline int ew_grid(size_t work_items, int per_block) {
  size_t need = (work_items + per_block - 1) / per_block;
  size_t cap = (size_t)kNumSMs * 16;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// ------------------------------------------------------------------ NCHW -> NHWC4
__global__ void nchw_to_nhwc4_kernel(const float* __restrict__ x, float* __restrict__ y, int C, size_t HW, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t n = i / HW, p = i % HW;
    const float* src = x + n * C * HW + p;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    v.x = __ldg(src);
    if (C > 1) v.y = __ldg(src + HW);
    if (C > 2) v.z = __ldg(src + 2 * HW);
    if (C > 3) v.w = __ldg(src + 3 * HW);
    reinterpret_cast<float4*>(y)[i] = v;
  }
}

int launch_nchw_to_nhwc4(const float* x, float* y, int N, int C, int H, int W, cudaStream_t s) {
  MDIL_REQUIRE(C >= 1 && C <= 4, "nchw_to_nhwc4: C must be 1..4");
  size_t HW = (size_t)H * W, total = (size_t)N * HW;
  nchw_to_nhwc4_kernel<<<ew_grid(total, 256), 256, 0, s>>>(x, y, C, HW, total);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ channel statistics
// Thread layout: C4 = C/4 threads span the channels of one pixel (one float4 each), 256/C4 pixel lanes.
// Per-thread fp32 partials over a short strided pixel walk, fp64 across lanes / CTAs.
template <int MODE>  // 0: sum x, sum x^2 ; 1: bn-backward sums (dz, dz*uhat)
__global__ void __launch_bounds__(256)
stats_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ drop,
             const float* __restrict__ u, const float* __restrict__ stats, size_t P, size_t HW, int C,
             double* __restrict__ sums) {
  extern __shared__ double sh[];  // [2][lanes][C]
  const int C4 = C >> 2;
  const int lanes = 256 / C4;
  const int c4 = threadIdx.x % C4;
  const int lane = threadIdx.x / C4;
  float4 s1 = make4(0.f), s2 = make4(0.f);
  float4 mean = make4(0.f), istd = make4(0.f);
  if (MODE == 1) {
    mean = ldg4(stats + c4 * 4);
    istd = ldg4(stats + C + c4 * 4);
  }
  // two pixels per iteration, all of their loads issued before the first use: the launch runs 4 CTAs per SM (more CTAs
  // mean more same-address fp64 atomics at the end) and needs the bytes in flight to cover the HBM latency
  const size_t stride = (size_t)gridDim.x * lanes;
  for (size_t p = (size_t)blockIdx.x * lanes + lane; p < P; p += 2 * stride) {
    const size_t q = p + stride;
    const bool two = q < P;
    float4 v = ldg4(x + p * C + c4 * 4);
    float4 v2 = two ? ldg4(x + q * C + c4 * 4) : make4(0.f);
    if (MODE == 0) {
      s1.x += v.x + v2.x; s1.y += v.y + v2.y; s1.z += v.z + v2.z; s1.w += v.w + v2.w;
      s2.x += v.x * v.x + v2.x * v2.x; s2.y += v.y * v.y + v2.y * v2.y;
      s2.z += v.z * v.z + v2.z * v2.z; s2.w += v.w * v.w + v2.w * v2.w;
    } else {
      float4 yy = make4(1.f), yy2 = make4(1.f), d = make4(1.f), d2 = make4(1.f);
      if (y != nullptr) { yy = ldg4(y + p * C + c4 * 4); if (two) yy2 = ldg4(y + q * C + c4 * 4); }
      if (drop != nullptr) { d = ldg4(drop + (p / HW) * C + c4 * 4); if (two) d2 = ldg4(drop + (q / HW) * C + c4 * 4); }
      const float4 uu = ldg4(u + p * C + c4 * 4);
      const float4 uu2 = two ? ldg4(u + q * C + c4 * 4) : make4(0.f);
      v.x = yy.x > 0.f ? v.x * d.x : 0.f; v.y = yy.y > 0.f ? v.y * d.y : 0.f;
      v.z = yy.z > 0.f ? v.z * d.z : 0.f; v.w = yy.w > 0.f ? v.w * d.w : 0.f;
      v2.x = yy2.x > 0.f ? v2.x * d2.x : 0.f; v2.y = yy2.y > 0.f ? v2.y * d2.y : 0.f;
      v2.z = yy2.z > 0.f ? v2.z * d2.z : 0.f; v2.w = yy2.w > 0.f ? v2.w * d2.w : 0.f;
      s1.x += v.x + v2.x; s1.y += v.y + v2.y; s1.z += v.z + v2.z; s1.w += v.w + v2.w;
      s2.x += v.x * ((uu.x - mean.x) * istd.x) + v2.x * ((uu2.x - mean.x) * istd.x);
      s2.y += v.y * ((uu.y - mean.y) * istd.y) + v2.y * ((uu2.y - mean.y) * istd.y);
      s2.z += v.z * ((uu.z - mean.z) * istd.z) + v2.z * ((uu2.z - mean.z) * istd.z);
      s2.w += v.w * ((uu.w - mean.w) * istd.w) + v2.w * ((uu2.w - mean.w) * istd.w);
    }
  }
  double* a = sh + (size_t)lane * C + c4 * 4;
  double* b = sh + (size_t)lanes * C + (size_t)lane * C + c4 * 4;
  a[0] = s1.x; a[1] = s1.y; a[2] = s1.z; a[3] = s1.w;
  b[0] = s2.x; b[1] = s2.y; b[2] = s2.z; b[3] = s2.w;
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    int which = c / C, ch = c % C;
    double t = 0.0;
    for (int l = 0; l < lanes; ++l) t += sh[(size_t)which * lanes * C + (size_t)l * C + ch];
    atomicAdd(sums + which * C + ch, t);
  }
}

static int stats_grid(size_t P, int lanes) {
  size_t need = (P + lanes - 1) / lanes;
  size_t cap = (size_t)kNumSMs * 4;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

int launch_channel_stats(const float* x, size_t P, int ld, int coff, int cnt, double* sums, int ldsum, cudaStream_t s) {
  MDIL_REQUIRE(coff == 0 && ld == cnt && ldsum == cnt, "channel_stats: only full-width tensors supported");
  int C = cnt;
  MDIL_REQUIRE(C % 4 == 0 && 256 % (C / 4) == 0, "channel_stats: unsupported channel count");
  int lanes = 256 / (C / 4);
  size_t smem = (size_t)2 * lanes * C * sizeof(double);
  stats_kernel<0><<<stats_grid(P, lanes), 256, smem, s>>>(x, nullptr, nullptr, nullptr, nullptr, P, 1, C, sums);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_bn_bwd_stats(const float* dy, const float* y, const float* drop, const float* u, const float* stats,
                        double* sums, int N, size_t HW, int C, cudaStream_t s) {
  MDIL_REQUIRE(C % 4 == 0 && 256 % (C / 4) == 0, "bn_bwd_stats: unsupported channel count");
  int lanes = 256 / (C / 4);
  size_t P = (size_t)N * HW;
  size_t smem = (size_t)2 * lanes * C * sizeof(double);
  stats_kernel<1><<<stats_grid(P, lanes), 256, smem, s>>>(dy, y, drop, u, stats, P, HW, C, sums);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ BN finalize (forward)
// fold > 1 (packed-4 view of the C = 16 blocks): the sums arrive per (pixel slot, channel) = [2][fold * C]; channel c
// sums its `fold` slots, and the statistics are also written replicated per slot into stats_rep [4][fold * C]
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int ldsum, double count, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float* rm, float* rv,
                                   float eps, float momentum, int train, float* __restrict__ stats, int fold,
                                   float* __restrict__ stats_rep, long long* nbt) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && train && nbt != nullptr) *nbt += 1;      // nn.BatchNorm2d.num_batches_tracked
  if (c >= C) return;
  float mean, var;
  if (train) {
    double s1 = 0.0, s2 = 0.0;
    for (int p = 0; p < fold; ++p) { s1 += sums[p * C + c]; s2 += sums[ldsum + p * C + c]; }
    double m = s1 / count;
    double v = s2 / count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    double unbiased = count > 1.0 ? v * count / (count - 1.0) : v;
    rm[c] = (1.f - momentum) * rm[c] + momentum * mean;
    rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unbiased;
  } else {
    mean = rm[c];
    var = rv[c];
  }
  float invstd = 1.0f / sqrtf(var + eps);
  float scale = gamma[c] * invstd;
  const float shift = beta[c] - mean * scale;
  stats[c] = mean;
  stats[C + c] = invstd;
  stats[2 * C + c] = scale;
  stats[3 * C + c] = shift;
  if (stats_rep != nullptr) {
    const int CF = fold * C;
    for (int p = 0; p < fold; ++p) {
      stats_rep[p * C + c] = mean;
      stats_rep[CF + p * C + c] = invstd;
      stats_rep[2 * CF + p * C + c] = scale;
      stats_rep[3 * CF + p * C + c] = shift;
    }
  }
}

int launch_bn_finalize(const double* sums, int ldsum, double count, int C, const float* gamma, const float* beta,
                       float* rm, float* rv, float eps, float momentum, int train, float* stats, cudaStream_t s,
                       int fold, float* stats_rep, long long* nbt) {
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, s>>>(sums, ldsum, count, C, gamma, beta, rm, rv, eps, momentum, train, stats,
                                                  fold, stats_rep, nbt);
  MDIL_LAUNCH_CHECK();
  return 0;
}

__global__ void replicate_stats_kernel(const float* __restrict__ stats, float* __restrict__ rep, int C, int fold) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4 * fold * C) return;
  const int k = i / (fold * C), c = i % C;
  rep[i] = stats[k * C + c];
}
int launch_replicate_stats(const float* stats, float* rep, int C, int fold, cudaStream_t s) {
  replicate_stats_kernel<<<cdiv(4 * fold * C, 128), 128, 0, s>>>(stats, rep, C, fold);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ BN apply + dropout + residual + ReLU
__global__ void __launch_bounds__(256)
bn_act_kernel(const float* __restrict__ u, const float* __restrict__ stats, const float* __restrict__ drop,
              const float* __restrict__ res, float* __restrict__ y, size_t total4, size_t HWC4, int C) {
  const int C4 = C >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4);
    float4 v = ldg4(u + i * 4);
    float4 sc = ldg4(stats + 2 * C + c4 * 4), sh = ldg4(stats + 3 * C + c4 * 4);
    v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    if (drop != nullptr) {
      float4 d = ldg4(drop + (i / HWC4) * C + c4 * 4);
      v.x *= d.x; v.y *= d.y; v.z *= d.z; v.w *= d.w;
    }
    if (res != nullptr) {
      float4 r = ldg4(res + i * 4);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    reinterpret_cast<float4*>(y)[i] = v;
  }
}

// bn_finalize + bn_act in one launch (C <= 128): every CTA derives scale / shift from the fp64 sums (train) or the
// running statistics (eval) in its prologue; CTA 0 also writes the statistics block and updates the running statistics
// (it reads rm / rv before it writes them; the other CTAs of a train-mode launch do not read them at all).
__global__ void __launch_bounds__(256)
bn_act_fused_kernel(const float* __restrict__ u, const double* __restrict__ sums, int ldsum, double count,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float* rm, float* rv, float eps,
                    float momentum, int train, int fold, float* __restrict__ stats, const float* __restrict__ drop,
                    const float* __restrict__ res, float* __restrict__ y, size_t total4, size_t HWC4, int C, long long* nbt) {
  __shared__ __align__(16) float cf[2 * 128];      // scale, shift
  if (blockIdx.x == 0 && threadIdx.x == 0 && train && nbt != nullptr) *nbt += 1;      // nn.BatchNorm2d.num_batches_tracked
  const double inv_count = 1.0 / count, bessel = count > 1.0 ? count / (count - 1.0) : 1.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, var;
    double unbiased = 0.0;
    if (train) {
      double s1 = 0.0, s2 = 0.0;
      for (int p = 0; p < fold; ++p) { s1 += sums[p * C + c]; s2 += sums[ldsum + p * C + c]; }
      const double m = s1 * inv_count;
      double v = s2 * inv_count - m * m;
      if (v < 0.0) v = 0.0;
      mean = (float)m;
      var = (float)v;
      unbiased = v * bessel;
    } else {
      mean = rm[c];
      var = rv[c];
    }
    const float invstd = 1.0f / sqrtf(var + eps);
    const float scale = gamma[c] * invstd;
    const float shift = beta[c] - mean * scale;
    cf[c] = scale;
    cf[C + c] = shift;
    if (blockIdx.x == 0) {
      if (train) {
        rm[c] = (1.f - momentum) * rm[c] + momentum * mean;
        rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unbiased;
      }
      stats[c] = mean;
      stats[C + c] = invstd;
      stats[2 * C + c] = scale;
      stats[3 * C + c] = shift;
    }
  }
  __syncthreads();
  const int C4 = C >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    float4 v = ldg4(u + i * 4);
    const float4 sc = *reinterpret_cast<const float4*>(cf + c4 * 4), sh = *reinterpret_cast<const float4*>(cf + C + c4 * 4);
    v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    if (drop != nullptr) {
      const float4 d = ldg4(drop + (i / HWC4) * C + c4 * 4);
      v.x *= d.x; v.y *= d.y; v.z *= d.z; v.w *= d.w;
    }
    if (res != nullptr) {
      const float4 r = ldg4(res + i * 4);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    reinterpret_cast<float4*>(y)[i] = v;
  }
}

int launch_bn_act_fused(const float* u, const double* sums, int ldsum, double count, const float* gamma, const float* beta,
                        float* rm, float* rv, float eps, float momentum, int train, int fold, float* stats, const float* drop,
                        const float* res, float* y, int N, size_t HW, int C, cudaStream_t s, long long* nbt) {
  MDIL_REQUIRE(C % 4 == 0 && C <= 128, "bn_act_fused: C % 4, C <= 128");
  size_t total4 = (size_t)N * HW * (C / 4);
  bn_act_fused_kernel<<<ew_grid(total4, 256), 256, 0, s>>>(u, sums, ldsum, count, gamma, beta, rm, rv, eps, momentum, train, fold,
                                                           stats, drop, res, y, total4, HW * (C / 4), C, nbt);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_bn_act(const float* u, const float* stats, const float* drop, const float* res, float* y, int N, size_t HW,
                  int C, cudaStream_t s) {
  MDIL_REQUIRE(C % 4 == 0, "bn_act: C % 4");
  size_t total4 = (size_t)N * HW * (C / 4);
  bn_act_kernel<<<ew_grid(total4, 256), 256, 0, s>>>(u, stats, drop, res, y, total4, HW * (C / 4), C);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ BN backward finalize / apply
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, double count, int C,
                                       const float* __restrict__ gamma, const float* __restrict__ stats,
                                       float* __restrict__ coef, float* dgamma, float* dbeta, int fold) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double sdz = 0.0, sdzu = 0.0;      // fold > 1: sums are [2][fold * C] per (pixel slot, channel)
  for (int p = 0; p < fold; ++p) { sdz += sums[p * C + c]; sdzu += sums[fold * C + p * C + c]; }
  coef[c] = gamma[c] * stats[C + c];
  coef[C + c] = (float)(sdz / count);
  coef[2 * C + c] = (float)(sdzu / count);
  if (dgamma != nullptr) dgamma[c] = (float)sdzu;
  if (dbeta != nullptr) dbeta[c] = (float)sdz;
}

int launch_bn_bwd_finalize(const double* sums, double count, int C, const float* gamma, const float* stats, float* coef,
                           float* dgamma, float* dbeta, cudaStream_t s, int fold) {
  bn_bwd_finalize_kernel<<<cdiv(C, 128), 128, 0, s>>>(sums, count, C, gamma, stats, coef, dgamma, dbeta, fold);
  MDIL_LAUNCH_CHECK();
  return 0;
}

__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ drop,
                    const float* __restrict__ u, const float* __restrict__ stats, const float* __restrict__ coef,
                    float* __restrict__ du, size_t total4, size_t HWC4, int C, int split) {
  const int C4 = C >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    int c4 = (int)(i % C4);
    float4 v = ldg4(dy + i * 4);
    if (y != nullptr) {
      float4 yy = ldg4(y + i * 4);
      v.x = yy.x > 0.f ? v.x : 0.f; v.y = yy.y > 0.f ? v.y : 0.f;
      v.z = yy.z > 0.f ? v.z : 0.f; v.w = yy.w > 0.f ? v.w : 0.f;
    }
    if (drop != nullptr) {
      float4 d = ldg4(drop + (i / HWC4) * C + c4 * 4);
      v.x *= d.x; v.y *= d.y; v.z *= d.z; v.w *= d.w;
    }
    float4 uu = ldg4(u + i * 4);
    float4 mean = ldg4(stats + c4 * 4), istd = ldg4(stats + C + c4 * 4);
    float4 g = ldg4(coef + c4 * 4), k1 = ldg4(coef + C + c4 * 4), k2 = ldg4(coef + 2 * C + c4 * 4);
    float4 o;
    o.x = g.x * (v.x - k1.x - (uu.x - mean.x) * istd.x * k2.x);
    o.y = g.y * (v.y - k1.y - (uu.y - mean.y) * istd.y * k2.y);
    o.z = g.z * (v.z - k1.z - (uu.z - mean.z) * istd.z * k2.z);
    o.w = g.w * (v.w - k1.w - (uu.w - mean.w) * istd.w * k2.w);
    if (split) {
      // "S16" internal format of the gradient tensors the tensor-core kernels consume: every group of four channels
      // (16 bytes, where the fp32 values would be) holds its four bf16 hi halves, then its four bf16 lo halves
      // (x = hi + lo) -- same addresses and access widths as fp32, but loaders / producers copy instead of converting
      uint4 w;
      split_bf16x2(o.x, o.y, w.x, w.z);
      split_bf16x2(o.z, o.w, w.y, w.w);
      reinterpret_cast<uint4*>(du)[i] = w;
    } else {
      reinterpret_cast<float4*>(du)[i] = o;
    }
  }
}

// bn_bwd_finalize + bn_bwd_apply in one launch: every CTA derives the per-channel coefficients from the fp64 sums in its
// prologue (<= 128 channels: a handful of fp64 operations per CTA) and keeps them, with mean / invstd, in shared memory;
// CTA 0 also writes the affine gradients.  One dependent launch less per BatchNorm backward (39 per training step).
__global__ void __launch_bounds__(256)
bn_bwd_apply_fused_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ drop,
                          const float* __restrict__ u, const float* __restrict__ stats, const double* __restrict__ sums,
                          double count, const float* __restrict__ gamma, int fold, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, float* __restrict__ du, size_t total4, size_t HWC4, int C, int split) {
  __shared__ __align__(16) float cf[5 * 128];      // g, k1, k2, mean, invstd
  const double inv_count = 1.0 / count;             // one fp64 division per thread, off the per-channel dependency chains
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double sdz = 0.0, sdzu = 0.0;      // fold > 1: sums are [2][fold * C] per (pixel slot, channel)
    for (int p = 0; p < fold; ++p) { sdz += sums[p * C + c]; sdzu += sums[fold * C + p * C + c]; }
    cf[c] = __ldg(gamma + c) * __ldg(stats + C + c);
    cf[C + c] = (float)(sdz * inv_count);
    cf[2 * C + c] = (float)(sdzu * inv_count);
    cf[3 * C + c] = __ldg(stats + c);
    cf[4 * C + c] = __ldg(stats + C + c);
    if (blockIdx.x == 0) {
      if (dgamma != nullptr) dgamma[c] = (float)sdzu;
      if (dbeta != nullptr) dbeta[c] = (float)sdz;
    }
  }
  __syncthreads();
  const int C4 = C >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    float4 v = ldg4(dy + i * 4);
    if (y != nullptr) {
      const float4 yy = ldg4(y + i * 4);
      v.x = yy.x > 0.f ? v.x : 0.f; v.y = yy.y > 0.f ? v.y : 0.f;
      v.z = yy.z > 0.f ? v.z : 0.f; v.w = yy.w > 0.f ? v.w : 0.f;
    }
    if (drop != nullptr) {
      const float4 d = ldg4(drop + (i / HWC4) * C + c4 * 4);
      v.x *= d.x; v.y *= d.y; v.z *= d.z; v.w *= d.w;
    }
    const float4 uu = ldg4(u + i * 4);
    const float4 g = *reinterpret_cast<const float4*>(cf + c4 * 4), k1 = *reinterpret_cast<const float4*>(cf + C + c4 * 4),
                 k2 = *reinterpret_cast<const float4*>(cf + 2 * C + c4 * 4), mean = *reinterpret_cast<const float4*>(cf + 3 * C + c4 * 4),
                 istd = *reinterpret_cast<const float4*>(cf + 4 * C + c4 * 4);
    float4 o;
    o.x = g.x * (v.x - k1.x - (uu.x - mean.x) * istd.x * k2.x);
    o.y = g.y * (v.y - k1.y - (uu.y - mean.y) * istd.y * k2.y);
    o.z = g.z * (v.z - k1.z - (uu.z - mean.z) * istd.z * k2.z);
    o.w = g.w * (v.w - k1.w - (uu.w - mean.w) * istd.w * k2.w);
    if (split) {       // S16 format: see bn_bwd_apply_kernel
      uint4 w;
      split_bf16x2(o.x, o.y, w.x, w.z);
      split_bf16x2(o.z, o.w, w.y, w.w);
      reinterpret_cast<uint4*>(du)[i] = w;
    } else {
      reinterpret_cast<float4*>(du)[i] = o;
    }
  }
}

int launch_bn_bwd_apply_fused(const float* dy, const float* y, const float* drop, const float* u, const float* stats,
                              const double* sums, double count, const float* gamma, int fold, float* dgamma, float* dbeta,
                              float* du, int N, size_t HW, int C, cudaStream_t s, int split) {
  MDIL_REQUIRE(C % 4 == 0 && C <= 128, "bn_bwd_apply_fused: C % 4, C <= 128");
  size_t total4 = (size_t)N * HW * (C / 4);
  bn_bwd_apply_fused_kernel<<<ew_grid(total4, 256), 256, 0, s>>>(dy, y, drop, u, stats, sums, count, gamma, fold, dgamma, dbeta,
                                                                 du, total4, HW * (C / 4), C, split);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_bn_bwd_apply(const float* dy, const float* y, const float* drop, const float* u, const float* stats,
                        const float* coef, float* du, int N, size_t HW, int C, cudaStream_t s, int split) {
  MDIL_REQUIRE(C % 4 == 0, "bn_bwd_apply: C % 4");
  size_t total4 = (size_t)N * HW * (C / 4);
  bn_bwd_apply_kernel<<<ew_grid(total4, 256), 256, 0, s>>>(dy, y, drop, u, stats, coef, du, total4, HW * (C / 4), C, split);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ max-pool branch
__global__ void __launch_bounds__(256)
pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ u, int H, int W, int Cin, int ldin, int ldu, int coff,
                size_t total) {
  const int OH = H >> 1, OW = W >> 1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % Cin);
    size_t p = i / Cin;
    int ox = (int)(p % OW);
    size_t t = p / OW;
    int oy = (int)(t % OH);
    size_t n = t / OH;
    const float* b = x + ((n * H + 2 * oy) * (size_t)W + 2 * ox) * ldin + c;
    float m = __ldg(b);
    m = fmaxf(m, __ldg(b + ldin));
    m = fmaxf(m, __ldg(b + (size_t)W * ldin));
    m = fmaxf(m, __ldg(b + (size_t)W * ldin + ldin));
    u[p * ldu + coff + c] = m;
  }
}

int launch_pool_fwd(const float* x, float* u, int N, int H, int W, int Cin, int ldin, int ldu, int coff, cudaStream_t s) {
  size_t total = (size_t)N * (H / 2) * (W / 2) * Cin;
  pool_fwd_kernel<<<ew_grid(total, 256), 256, 0, s>>>(x, u, H, W, Cin, ldin, ldu, coff, total);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// First maximum in row-major window order wins (torch max_pool2d backward).
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ du, float* __restrict__ dx, int H, int W, int Cin,
                int ldin, int ldu, int coff, int accumulate, size_t total) {
  const int OH = H >> 1, OW = W >> 1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % Cin);
    size_t p = i / Cin;
    int ox = (int)(p % OW);
    size_t t = p / OW;
    int oy = (int)(t % OH);
    size_t n = t / OH;
    size_t base = ((n * H + 2 * oy) * (size_t)W + 2 * ox);
    const float* b = x + base * ldin + c;
    float v0 = __ldg(b), v1 = __ldg(b + ldin), v2 = __ldg(b + (size_t)W * ldin), v3 = __ldg(b + (size_t)W * ldin + ldin);
    int arg = 0;
    float m = v0;
    if (v1 > m) { m = v1; arg = 1; }
    if (v2 > m) { m = v2; arg = 2; }
    if (v3 > m) { m = v3; arg = 3; }
    float g = __ldg(du + p * ldu + coff + c);
    float* d = dx + base * Cin + c;
    size_t offs[4] = {0, (size_t)Cin, (size_t)W * Cin, (size_t)W * Cin + Cin};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float val = (k == arg) ? g : 0.f;
      if (accumulate) d[offs[k]] += val; else d[offs[k]] = val;
    }
  }
}

int launch_pool_bwd(const float* x, const float* du, float* dx, int N, int H, int W, int Cin, int ldin, int ldu,
                    int coff, int accumulate, cudaStream_t s) {
  size_t total = (size_t)N * (H / 2) * (W / 2) * Cin;
  pool_bwd_kernel<<<ew_grid(total, 256), 256, 0, s>>>(x, du, dx, H, W, Cin, ldin, ldu, coff, accumulate, total);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ scaling / Adam
__global__ void __launch_bounds__(256)
scale_kernel(float* __restrict__ x, size_t n4, size_t n, const double* __restrict__ den, const float* __restrict__ mul) {
  float f = 1.0f;
  if (mul != nullptr) f *= __ldg(mul);
  if (den != nullptr) f = (float)((double)f / den[0]);
  if (f == 1.0f) return;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    v.x *= f; v.y *= f; v.z *= f; v.w *= f;
    reinterpret_cast<float4*>(x)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) x[n4 * 4 + threadIdx.x] *= f;
}

int launch_scale(float* x, size_t n, const double* den, const float* mul, cudaStream_t s) {
  MDIL_REQUIRE(((uintptr_t)x & 15) == 0, "scale: pointer must be 16-byte aligned");
  size_t n4 = n / 4;
  scale_kernel<<<ew_grid(n4, 256), 256, 0, s>>>(x, n4, n, den, mul);
  MDIL_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
            float step_size, float b1, float b2, float eps, float wd, float bc2_sqrt, float grad_scale) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * grad_scale + wd * pi;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

// CUDA-graph-safe variant: the step counter and the learning rate live in device memory (state[0] = steps taken so far,
// state[1] = lr; state[2], state[3] receive step_size and sqrt(bias_correction2)), so a captured optimiser step advances
// its own bias corrections on every replay.  Corrections in double, as torch.optim.Adam computes them.
__global__ void adam_prepare_kernel(float* __restrict__ state, float b1, float b2) {
  const double step = (double)state[0] + 1.0;
  state[0] = (float)step;
  state[2] = (float)((double)state[1] / (1.0 - pow((double)b1, step)));
  state[3] = (float)sqrt(1.0 - pow((double)b2, step));
}
__global__ void __launch_bounds__(256)
adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
                const float* __restrict__ state, float b1, float b2, float eps, float wd, float grad_scale) {
  const float step_size = state[2], bc2_sqrt = state[3];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * grad_scale + wd * pi;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}
int launch_adam_dev(float* param, const float* grad, float* m, float* v, size_t n, float* state, float b1, float b2, float eps,
                    float wd, float grad_scale, cudaStream_t s) {
  adam_prepare_kernel<<<1, 1, 0, s>>>(state, b1, b2);
  MDIL_LAUNCH_CHECK();
  adam_dev_kernel<<<ew_grid(n, 256), 256, 0, s>>>(param, grad, m, v, n, state, b1, b2, eps, wd, grad_scale);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_adam(float* param, const float* grad, float* m, float* v, size_t n, float lr, float b1, float b2, float eps,
                float wd, int step, float grad_scale, cudaStream_t s) {
  // bias corrections in double on the host, as torch.optim.Adam computes them in Python floats (1 - 0.999f carries a
  // 1.3e-5 relative error at step 1); the kernel receives step_size = lr / bc1 and sqrt(bc2)
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  adam_kernel<<<ew_grid(n, 256), 256, 0, s>>>(param, grad, m, v, n, (float)((double)lr / bc1), b1, b2, eps, wd,
                                             (float)sqrt(bc2), grad_scale);
  MDIL_LAUNCH_CHECK();
  return 0;
}

}  // namespace mdil
