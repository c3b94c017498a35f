// Network head and losses (all HBM-bound, one pass over the logits each):
//   output_conv  ConvTranspose2d(16, Ccls, 2, stride 2)      models/erfnet_RA_parallel.py:179-180,188
//   CrossEntropyLoss2d (weighted NLL of log_softmax)          train_new_task_step2.py:84-92
//   output distillation KLDivLoss(softmax(s), softmax(t))     train_new_task_step2.py:241,296-297
//   argmax + confusion matrix for iouEval                     iouEval.py:21-70 (validation path)
// Logits are NCHW fp32 (what the reference's consumers index); threads run along W so every
// class-plane access is a coalesced 128/256-byte row segment.
#include "kernels.cuh"

namespace mdil {

constexpr int kMaxCls = 32;

// ------------------------------------------------------------------ output_conv forward
// smem weights: Ws[co][ky][kx][ci] (ci contiguous, broadcast float4 reads)
__global__ void __launch_bounds__(256)
outconv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ logits, int N, int H, int W, int Ccls) {
  __shared__ __align__(16) float Ws[kMaxCls * 4 * 16];
  __shared__ float Bs[kMaxCls];
  for (int i = threadIdx.x; i < Ccls * 64; i += 256) {
    const int ci = i & 15, j = i >> 4;  // j = co*4 + ky*2 + kx
    Ws[i] = __ldg(w + (size_t)ci * Ccls * 4 + j);
  }
  for (int i = threadIdx.x; i < Ccls; i += 256) Bs[i] = bias != nullptr ? __ldg(bias + i) : 0.f;
  __syncthreads();
  const size_t P = (size_t)N * H * W;
  const size_t OW = 2 * (size_t)W, OHW = 4 * (size_t)H * W;
  for (size_t p = blockIdx.x * (size_t)256 + threadIdx.x; p < P; p += (size_t)gridDim.x * 256) {
    const int j = (int)(p % W);
    const size_t t = p / W;
    const int i = (int)(t % H);
    const size_t n = t / H;
    float xv[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = ldg4(x + p * 16 + q * 4);
      xv[q * 4 + 0] = v.x; xv[q * 4 + 1] = v.y; xv[q * 4 + 2] = v.z; xv[q * 4 + 3] = v.w;
    }
    float* obase = logits + n * Ccls * OHW + (size_t)(2 * i) * OW + 2 * j;
    for (int co = 0; co < Ccls; ++co) {
      const float b = Bs[co];
#pragma unroll
      for (int ky = 0; ky < 2; ++ky) {
        float o0 = b, o1 = b;
        const float* w0 = Ws + ((co * 2 + ky) * 2 + 0) * 16;
        const float* w1 = w0 + 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 a = *reinterpret_cast<const float4*>(w0 + q * 4);
          const float4 c = *reinterpret_cast<const float4*>(w1 + q * 4);
          o0 = fmaf(xv[q * 4 + 0], a.x, o0); o0 = fmaf(xv[q * 4 + 1], a.y, o0);
          o0 = fmaf(xv[q * 4 + 2], a.z, o0); o0 = fmaf(xv[q * 4 + 3], a.w, o0);
          o1 = fmaf(xv[q * 4 + 0], c.x, o1); o1 = fmaf(xv[q * 4 + 1], c.y, o1);
          o1 = fmaf(xv[q * 4 + 2], c.z, o1); o1 = fmaf(xv[q * 4 + 3], c.w, o1);
        }
        *reinterpret_cast<float2*>(obase + (size_t)co * OHW + (size_t)ky * OW) = make_float2(o0, o1);
      }
    }
  }
}

int launch_outconv_fwd(const float* x, const float* w, const float* bias, float* logits, int N, int H, int W, int Ccls,
                       cudaStream_t s) {
  MDIL_REQUIRE(Ccls >= 1 && Ccls <= kMaxCls, "outconv: Ccls must be 1..32");
  size_t P = (size_t)N * H * W;
  int grid = (int)((P + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  outconv_fwd_kernel<<<grid, 256, 0, s>>>(x, w, bias, logits, N, H, W, Ccls);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ output_conv backward (data)
__global__ void __launch_bounds__(256)
outconv_dgrad_kernel(const float* __restrict__ dl, const float* __restrict__ w, float* __restrict__ dx, int N, int H,
                     int W, int Ccls) {
  __shared__ __align__(16) float Ws[kMaxCls * 4 * 16];
  for (int i = threadIdx.x; i < Ccls * 64; i += 256) {
    const int ci = i & 15, j = i >> 4;
    Ws[i] = __ldg(w + (size_t)ci * Ccls * 4 + j);
  }
  __syncthreads();
  const size_t P = (size_t)N * H * W;
  const size_t OW = 2 * (size_t)W, OHW = 4 * (size_t)H * W;
  for (size_t p = blockIdx.x * (size_t)256 + threadIdx.x; p < P; p += (size_t)gridDim.x * 256) {
    const int j = (int)(p % W);
    const size_t t = p / W;
    const int i = (int)(t % H);
    const size_t n = t / H;
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.f;
    const float* gbase = dl + n * Ccls * OHW + (size_t)(2 * i) * OW + 2 * j;
    for (int co = 0; co < Ccls; ++co) {
#pragma unroll
      for (int ky = 0; ky < 2; ++ky) {
        const float2 g = __ldg(reinterpret_cast<const float2*>(gbase + (size_t)co * OHW + (size_t)ky * OW));
        const float* w0 = Ws + ((co * 2 + ky) * 2 + 0) * 16;
        const float* w1 = w0 + 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 a = *reinterpret_cast<const float4*>(w0 + q * 4);
          const float4 c = *reinterpret_cast<const float4*>(w1 + q * 4);
          acc[q * 4 + 0] = fmaf(g.x, a.x, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(g.x, a.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(g.x, a.z, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(g.x, a.w, acc[q * 4 + 3]);
          acc[q * 4 + 0] = fmaf(g.y, c.x, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(g.y, c.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(g.y, c.z, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(g.y, c.w, acc[q * 4 + 3]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(dx + p * 16 + q * 4) = make_float4(acc[q * 4 + 0], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
  }
}

// ------------------------------------------------------------------ output_conv backward (weights, bias)
constexpr int OC_PT = 64;  // pixels per tile
__global__ void __launch_bounds__(256)
outconv_wgrad_kernel(const float* __restrict__ dl, const float* __restrict__ x, float* __restrict__ dw,
                     float* __restrict__ db, int N, int H, int W, int Ccls) {
  __shared__ __align__(16) float xs[OC_PT][16];
  __shared__ __align__(16) float gs[OC_PT][kMaxCls * 4 + 2];
  const int J = Ccls * 4;
  const int j = threadIdx.x & 127, half = threadIdx.x >> 7;
  const size_t P = (size_t)N * H * W;
  const size_t OW = 2 * (size_t)W, OHW = 4 * (size_t)H * W;
  const size_t tiles = (P + OC_PT - 1) / OC_PT;
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  float bsum = 0.f;
  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const size_t p0 = tile * OC_PT;
    {
      const int pp = threadIdx.x >> 2, q = threadIdx.x & 3;
      float4 v = make4(0.f);
      if (p0 + pp < P) v = ldg4(x + (p0 + pp) * 16 + q * 4);
      *reinterpret_cast<float4*>(&xs[pp][q * 4]) = v;
    }
    for (int idx = threadIdx.x; idx < OC_PT * Ccls * 2; idx += 256) {
      const int pp = idx % OC_PT, r = idx / OC_PT;  // r = co*2 + ky
      const size_t p = p0 + pp;
      float2 g = make_float2(0.f, 0.f);
      if (p < P) {
        const int jj = (int)(p % W);
        const size_t t = p / W;
        const int ii = (int)(t % H);
        const size_t n = t / H;
        g = __ldg(reinterpret_cast<const float2*>(dl + n * Ccls * OHW + (size_t)(r >> 1) * OHW +
                                                   (size_t)(2 * ii + (r & 1)) * OW + 2 * jj));
      }
      *reinterpret_cast<float2*>(&gs[pp][r * 2]) = g;
    }
    __syncthreads();
    if (j < J) {
#pragma unroll 4
      for (int pp = 0; pp < OC_PT; ++pp) {
        const float g = gs[pp][j];
        const float4 a = *reinterpret_cast<const float4*>(&xs[pp][half * 8]);
        const float4 c = *reinterpret_cast<const float4*>(&xs[pp][half * 8 + 4]);
        acc[0] = fmaf(a.x, g, acc[0]); acc[1] = fmaf(a.y, g, acc[1]);
        acc[2] = fmaf(a.z, g, acc[2]); acc[3] = fmaf(a.w, g, acc[3]);
        acc[4] = fmaf(c.x, g, acc[4]); acc[5] = fmaf(c.y, g, acc[5]);
        acc[6] = fmaf(c.z, g, acc[6]); acc[7] = fmaf(c.w, g, acc[7]);
        bsum += g;
      }
    }
    __syncthreads();
  }
  if (j < J) {
    if (dw != nullptr) {
#pragma unroll
      for (int q = 0; q < 8; ++q) atomicAdd(dw + (size_t)(half * 8 + q) * J + j, acc[q]);
    }
    if (db != nullptr && half == 0) atomicAdd(db + (j >> 2), bsum);
  }
}

int launch_outconv_bwd(const float* dlogits, const float* x, const float* w, float* dx, float* dw, float* db, int N,
                       int H, int W, int Ccls, cudaStream_t s) {
  MDIL_REQUIRE(Ccls >= 1 && Ccls <= kMaxCls, "outconv: Ccls must be 1..32");
  size_t P = (size_t)N * H * W;
  if (dx != nullptr) {
    int grid = (int)((P + 255) / 256);
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    outconv_dgrad_kernel<<<grid, 256, 0, s>>>(dlogits, w, dx, N, H, W, Ccls);
    MDIL_LAUNCH_CHECK();
  }
  if (dw != nullptr || db != nullptr) {
    if (dw != nullptr) MDIL_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 16 * Ccls * 4, s));
    if (db != nullptr) MDIL_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Ccls, s));
    size_t tiles = (P + OC_PT - 1) / OC_PT;
    int grid = (int)(tiles < (size_t)kNumSMs * 4 ? tiles : (size_t)kNumSMs * 4);
    outconv_wgrad_kernel<<<grid, 256, 0, s>>>(dlogits, x, dw, db, N, H, W, Ccls);
    MDIL_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------ block reduction helper
__device__ __forceinline__ void block_atomic_add2(double a, double b, double* acc, double* sh /*[2*8]*/) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[warp] = a; sh[8 + warp] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int i = 0; i < 8; ++i) { ta += sh[i]; tb += sh[8 + i]; }
    atomicAdd(acc, ta);
    if (acc + 1 != nullptr) atomicAdd(acc + 1, tb);
  }
}

// ------------------------------------------------------------------ CrossEntropyLoss2d forward + dlogits
__global__ void __launch_bounds__(256)
ce2d_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ class_w, int C,
            size_t HW, size_t P, double* __restrict__ acc, float* __restrict__ dlogits) {
  __shared__ double sh[16];
  __shared__ float ws[kMaxCls];
  if (threadIdx.x < kMaxCls) ws[threadIdx.x] = threadIdx.x < C ? __ldg(class_w + threadIdx.x) : 0.f;
  __syncthreads();
  double num = 0.0, den = 0.0;
  for (size_t p = blockIdx.x * (size_t)256 + threadIdx.x; p < P; p += (size_t)gridDim.x * 256) {
    const size_t n = p / HW, hw = p % HW;
    const float* lp = logits + n * C * HW + hw;
    float x[kMaxCls];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c) {
      if (c < C) { x[c] = __ldg(lp + (size_t)c * HW); m = fmaxf(m, x[c]); }
    }
    float z = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c) {
      if (c < C) { x[c] = expf(x[c] - m); z += x[c]; }
    }
    const long long y = labels[p];
    const bool ok = y >= 0 && y < C;
    float w = 0.f, xy = 0.f;
    if (ok) {
      w = ws[y];
      xy = __ldg(lp + (size_t)y * HW);
    }
    const float logz = logf(z);
    if (ok) {
      num += (double)(w * (logz + m - xy));
      den += (double)w;
    }
    if (dlogits != nullptr) {
      float* dp = dlogits + n * C * HW + hw;
      const float inv = w / z;
#pragma unroll
      for (int c = 0; c < kMaxCls; ++c) {
        if (c < C) dp[(size_t)c * HW] = x[c] * inv - ((ok && c == (int)y) ? w : 0.f);
      }
    }
  }
  block_atomic_add2(num, den, acc, sh);
}

__global__ void ce2d_finish_kernel(const double* acc, float* loss) { loss[0] = (float)(acc[0] / acc[1]); }

int launch_ce2d(const float* logits, const int64_t* labels, const float* class_w, int N, int C, int H, int W,
                float* loss, double* acc, float* dlogits, cudaStream_t s) {
  MDIL_REQUIRE(C >= 1 && C <= kMaxCls, "ce2d: C must be 1..32");
  size_t HW = (size_t)H * W, P = (size_t)N * HW;
  MDIL_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), s));
  int grid = (int)((P + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  ce2d_kernel<<<grid, 256, 0, s>>>(logits, labels, class_w, C, HW, P, acc, dlogits);
  MDIL_LAUNCH_CHECK();
  ce2d_finish_kernel<<<1, 1, 0, s>>>(acc, loss);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ output distillation forward + dstudent
__global__ void __launch_bounds__(256)
kd_kernel(const float* __restrict__ student, const float* __restrict__ teacher, int C, size_t HW, size_t P,
          float inv_count, double* __restrict__ acc, float* __restrict__ dstudent) {
  __shared__ double sh[16];
  double tot = 0.0;
  for (size_t p = blockIdx.x * (size_t)256 + threadIdx.x; p < P; p += (size_t)gridDim.x * 256) {
    const size_t n = p / HW, hw = p % HW;
    const float* sp = student + n * C * HW + hw;
    const float* tp = teacher + n * C * HW + hw;
    float sx[kMaxCls], tx[kMaxCls];
    float ms = -INFINITY, mt = -INFINITY;
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c) {
      if (c < C) {
        sx[c] = __ldg(sp + (size_t)c * HW); ms = fmaxf(ms, sx[c]);
        tx[c] = __ldg(tp + (size_t)c * HW); mt = fmaxf(mt, tx[c]);
      }
    }
    float zs = 0.f, zt = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c) {
      if (c < C) {
        tx[c] -= mt;                       // keep t - max for log T
        sx[c] = expf(sx[c] - ms); zs += sx[c];
        zt += expf(tx[c]);
      }
    }
    const float izs = 1.f / zs, izt = 1.f / zt, logzt = logf(zt);
    float dot = 0.f, part = 0.f;  // dot = sum_c T_c S_c
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c) {
      if (c < C) {
        const float T = expf(tx[c]) * izt;
        const float S = sx[c] * izs;
        const float tlogt = T > 0.f ? T * (tx[c] - logzt) : 0.f;  // xlogy(T, T)
        part += tlogt - T * S;
        dot += T * S;
        sx[c] = S;
        tx[c] = T;
      }
    }
    tot += (double)part;
    if (dstudent != nullptr) {
      // g_c = -T_c / count ; dlogit_c = S_c * (g_c - sum_j g_j S_j) = -S_c * (T_c - dot) / count
      float* dp = dstudent + n * C * HW + hw;
#pragma unroll
      for (int c = 0; c < kMaxCls; ++c) {
        if (c < C) dp[(size_t)c * HW] = -sx[c] * (tx[c] - dot) * inv_count;
      }
    }
  }
  tot = warp_sum(tot);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(acc, t);
  }
}

__global__ void kd_finish_kernel(const double* acc, double inv_count, float* loss) { loss[0] = (float)(acc[0] * inv_count); }

int launch_kd(const float* student, const float* teacher, int N, int C, int H, int W, float* loss, double* acc,
              float* dstudent, cudaStream_t s) {
  MDIL_REQUIRE(C >= 1 && C <= kMaxCls, "kd: C must be 1..32");
  size_t HW = (size_t)H * W, P = (size_t)N * HW;
  double inv = 1.0 / ((double)P * C);
  MDIL_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), s));
  int grid = (int)((P + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  kd_kernel<<<grid, 256, 0, s>>>(student, teacher, C, HW, P, (float)inv, acc, dstudent);
  MDIL_LAUNCH_CHECK();
  kd_finish_kernel<<<1, 1, 0, s>>>(acc, inv, loss);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------ argmax + confusion matrix
__global__ void __launch_bounds__(256)
argmax_confusion_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int C, size_t HW, size_t P,
                        int64_t* __restrict__ pred, unsigned long long* __restrict__ conf) {
  __shared__ unsigned int hist[kMaxCls * kMaxCls];
  const bool do_conf = conf != nullptr && labels != nullptr;
  if (do_conf) {
    for (int i = threadIdx.x; i < C * C; i += 256) hist[i] = 0u;
    __syncthreads();
  }
  for (size_t p = blockIdx.x * (size_t)256 + threadIdx.x; p < P; p += (size_t)gridDim.x * 256) {
    const size_t n = p / HW, hw = p % HW;
    const float* lp = logits + n * C * HW + hw;
    float best = __ldg(lp);
    int arg = 0;
    for (int c = 1; c < C; ++c) {
      const float v = __ldg(lp + (size_t)c * HW);
      if (v > best) { best = v; arg = c; }
    }
    if (pred != nullptr) pred[p] = arg;
    if (do_conf) {
      const long long y = labels[p];
      if (y >= 0 && y < C) atomicAdd(&hist[(int)y * C + arg], 1u);
    }
  }
  if (do_conf) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += 256)
      if (hist[i] != 0u) atomicAdd(conf + i, (unsigned long long)hist[i]);
  }
}

int launch_argmax_confusion(const float* logits, const int64_t* labels, int N, int C, int H, int W, int64_t* pred,
                            long long* conf, cudaStream_t s) {
  MDIL_REQUIRE(C >= 1 && C <= kMaxCls, "argmax: C must be 1..32");
  size_t HW = (size_t)H * W, P = (size_t)N * HW;
  int grid = (int)((P + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  argmax_confusion_kernel<<<grid, 256, 0, s>>>(logits, labels, C, HW, P, pred, reinterpret_cast<unsigned long long*>(conf));
  MDIL_LAUNCH_CHECK();
  return 0;
}

}  // namespace mdil
