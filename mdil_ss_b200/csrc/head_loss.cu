// Network head and losses (all HBM-bound, one pass over the logits each):
//   output_conv  ConvTranspose2d(16, Ccls, 2, stride 2)      models/erfnet_RA_parallel.py:179-180,188
//   CrossEntropyLoss2d (weighted NLL of log_softmax)          train_new_task_step2.py:84-92
//   output distillation KLDivLoss(softmax(s), softmax(t))     train_new_task_step2.py:241,296-297
//   argmax + confusion matrix for iouEval                     iouEval.py:21-70 (validation path)
// Logits are NCHW fp32 (what the reference's consumers index); threads run along W so every
// class-plane access is a coalesced 128/256-byte row segment.
#include "kernels.cuh"

#include <stdlib.h>
#include <string.h>

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

// Even W: one thread per PAIR of input pixels -> per class plane and output row four consecutive logits (one 128-bit
// store; a warp writes 512 contiguous bytes), index arithmetic once per 160 stores.
__global__ void __launch_bounds__(256)
outconv_fwd2_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ logits, int N, int H, int W, int Ccls) {
  __shared__ __align__(16) float Ws[kMaxCls * 4 * 16];
  __shared__ float Bs[kMaxCls];
  for (int i = threadIdx.x; i < Ccls * 64; i += 256) {
    const int ci = i & 15, j = i >> 4;  // j = co*4 + ky*2 + kx
    Ws[i] = __ldg(w + (size_t)ci * Ccls * 4 + j);
  }
  for (int i = threadIdx.x; i < Ccls; i += 256) Bs[i] = bias != nullptr ? __ldg(bias + i) : 0.f;
  __syncthreads();
  const unsigned W2 = (unsigned)W >> 1, total = (unsigned)N * (unsigned)H * W2;
  const size_t OW = 2 * (size_t)W, OHW = 4 * (size_t)H * W;
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    const unsigned j2 = idx % W2, row = idx / W2, i = row % (unsigned)H, n = row / (unsigned)H;
    const float* xp = x + ((size_t)row * W + 2 * j2) * 16;
    float xv[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = ldg4(xp + q * 4);
      xv[q * 4 + 0] = v.x; xv[q * 4 + 1] = v.y; xv[q * 4 + 2] = v.z; xv[q * 4 + 3] = v.w;
    }
    float* obase = logits + (size_t)n * Ccls * OHW + (size_t)(2 * i) * OW + 4 * j2;
    for (int co = 0; co < Ccls; ++co) {
      const float b = Bs[co];
#pragma unroll
      for (int ky = 0; ky < 2; ++ky) {
        float o0 = b, o1 = b, o2 = b, o3 = b;
        const float* w0 = Ws + ((co * 2 + ky) * 2 + 0) * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 a = *reinterpret_cast<const float4*>(w0 + q * 4);
          const float4 c = *reinterpret_cast<const float4*>(w0 + 16 + q * 4);
          o0 = fmaf(xv[q * 4 + 0], a.x, o0); o0 = fmaf(xv[q * 4 + 1], a.y, o0);
          o0 = fmaf(xv[q * 4 + 2], a.z, o0); o0 = fmaf(xv[q * 4 + 3], a.w, o0);
          o1 = fmaf(xv[q * 4 + 0], c.x, o1); o1 = fmaf(xv[q * 4 + 1], c.y, o1);
          o1 = fmaf(xv[q * 4 + 2], c.z, o1); o1 = fmaf(xv[q * 4 + 3], c.w, o1);
          o2 = fmaf(xv[16 + q * 4 + 0], a.x, o2); o2 = fmaf(xv[16 + q * 4 + 1], a.y, o2);
          o2 = fmaf(xv[16 + q * 4 + 2], a.z, o2); o2 = fmaf(xv[16 + q * 4 + 3], a.w, o2);
          o3 = fmaf(xv[16 + q * 4 + 0], c.x, o3); o3 = fmaf(xv[16 + q * 4 + 1], c.y, o3);
          o3 = fmaf(xv[16 + q * 4 + 2], c.z, o3); o3 = fmaf(xv[16 + q * 4 + 3], c.w, o3);
        }
        *reinterpret_cast<float4*>(obase + (size_t)co * OHW + (size_t)ky * OW) = make_float4(o0, o1, o2, o3);
      }
    }
  }
}

int launch_outconv_fwd(const float* x, const float* w, const float* bias, float* logits, int N, int H, int W, int Ccls,
                       cudaStream_t s) {
  MDIL_REQUIRE(Ccls >= 1 && Ccls <= kMaxCls, "outconv: Ccls must be 1..32");
  size_t P = (size_t)N * H * W;
  if (W % 2 == 0 && P < (1ull << 31) && ((uintptr_t)logits & 15) == 0) {
    int grid = (int)((P / 2 + 255) / 256);
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    outconv_fwd2_kernel<<<grid, 256, 0, s>>>(x, w, bias, logits, N, H, W, Ccls);
    MDIL_LAUNCH_CHECK();
    return 0;
  }
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

// ------------------------------------------------------------------ output_conv backward: ONE pass over dlogits
// Even W.  A tile = 128 input pixels of one row = 256 logit columns x 2 rows x Ccls planes, staged once in shared memory
// (128-bit coalesced loads, no per-element index arithmetic) and consumed twice: the data gradient on the FMA pipe
// (thread = pixel pair x 4 input channels) and the [16 x 4 Ccls] weight gradient on the warp-level tensor-core path
// (mma.sync m16n8k8, error-compensated 3xTF32, K = pixels, fragments read conflict-free from the staged tiles); the bias
// gradient is a row sum of the staged tile.  dlogits (the largest tensor of the step) is read once instead of twice.
constexpr int OB_PX = 64;    // input pixels per tile
constexpr int OB_THREADS = 128;   // small CTAs, 4 per SM: their load / compute phases interleave
constexpr int OB_GS = 136;   // row stride of the staged gradient tile in floats (8 mod 32: conflict-free B fragments)
constexpr int OB_XS = 24;    // row stride of the staged activation tile (conflict-free A fragments)

__device__ __forceinline__ void ob_mma(float (&dd)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(dd[0]), "+f"(dd[1]), "+f"(dd[2]), "+f"(dd[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ob_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xFFFFE000u;
}

__device__ __forceinline__ void ob_cp_async16(float* smem_dst, const float* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;          // 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}

template <int NT>   // tiles of 8 (co, ky, kx) columns and staging iterations: NT >= ceil(Ccls / 2)
__global__ void __launch_bounds__(OB_THREADS)
outconv_bwd2_kernel(const float* __restrict__ dl, const float* __restrict__ x, const float* __restrict__ w,
                    float* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int N, int H, int W, int Ccls) {
  extern __shared__ __align__(16) float ob_sm[];
  const int R = 2 * Ccls, J = 4 * Ccls;
  const int GSZ = R * OB_GS, XSZ = OB_PX * OB_XS;
  float* gs = ob_sm;                       // [2][R][OB_GS]: row r = co*2 + ky, column = 2 * pixel + kx   (double buffer)
  float* xs = gs + 2 * GSZ;                // [2][OB_PX][OB_XS]
  float* Ws = xs + 2 * XSZ;                // [(co*2 + ky)*2 + kx][16 ci]
  float* red = Ws + Ccls * 64;             // [16 ci][J]
  float* bred = red + 16 * J;              // [R]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, t4 = lane & 3;
  for (int i = tid; i < Ccls * 64; i += OB_THREADS) {
    const int ci = i & 15, j = i >> 4;
    Ws[i] = __ldg(w + (size_t)ci * J + j);
  }
  for (int i = tid; i < 16 * J + R; i += OB_THREADS) red[i] = 0.f;      // red and bred are contiguous
  const size_t OW = 2 * (size_t)W, OHW = 4 * (size_t)H * W;
  const int tpr = (W + OB_PX - 1) / OB_PX;
  const int tiles = N * H * tpr;
  // tile `tile` -> shared-memory buffer `buf`, asynchronously (cp.async, 16 bytes each, zero fill past the row end):
  // the loads of tile t+1 fly while tile t is consumed
  auto stage = [&](int tile, int buf) {
    const int row = tile / tpr, j0 = (tile - row * tpr) * OB_PX;
    const int i = row % H, n = row / H;
    const int npx = W - j0 < OB_PX ? W - j0 : OB_PX;      // even
    float* gsb = gs + buf * GSZ;
    float* xsb = xs + buf * XSZ;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int idx = tid + k * OB_THREADS, px = idx >> 2, q = idx & 3;
      const bool ok = px < npx;
      ob_cp_async16(xsb + px * OB_XS + q * 4, ok ? x + ((size_t)row * W + j0 + px) * 16 + q * 4 : x, ok);
    }
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      const int idx = tid + k * OB_THREADS, r = idx >> 5, c4 = idx & 31;
      if (r < R) {
        const bool ok = 2 * c4 < npx;
        const float* src = dl + ((size_t)n * Ccls + (r >> 1)) * OHW + (size_t)(2 * i + (r & 1)) * OW + 2 * j0 + 4 * c4;
        ob_cp_async16(gsb + r * OB_GS + 4 * c4, ok ? src : dl, ok);
      }
    }
  };
  float acc[NT][4];
  float bsum[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; bsum[i] = 0.f; }
  int buf = 0;
  if ((int)blockIdx.x < tiles) stage(blockIdx.x, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1) {
    if (tile + (int)gridDim.x < tiles) stage(tile + gridDim.x, buf ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");       // everything but the newest group: this tile has landed
    __syncthreads();
    const int row = tile / tpr, j0 = (tile - row * tpr) * OB_PX;
    const int npx = W - j0 < OB_PX ? W - j0 : OB_PX;
    const float* gsb = gs + buf * GSZ;
    const float* xsb = xs + buf * XSZ;
    if (db != nullptr) {
#pragma unroll
      for (int k = 0; k < NT; ++k) {       // warp `warp` sums rows warp, warp + 4, ...: one 128-bit read per lane and row
        const int r = warp + 4 * k;
        if (r < R) {
          const float4 v = *reinterpret_cast<const float4*>(gsb + r * OB_GS + 4 * lane);
          bsum[k] += (v.x + v.y) + (v.z + v.w);
        }
      }
    }
    if (dx != nullptr) {
      const int pp = tid >> 2, q = tid & 3;
      float4 a0 = make4(0.f), a1 = make4(0.f);
#pragma unroll 4
      for (int r = 0; r < R; ++r) {
        const float4 g = *reinterpret_cast<const float4*>(gsb + r * OB_GS + 4 * pp);
        const float4 w0 = *reinterpret_cast<const float4*>(Ws + (r * 2 + 0) * 16 + q * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(Ws + (r * 2 + 1) * 16 + q * 4);
        a0.x = fmaf(g.x, w0.x, a0.x); a0.y = fmaf(g.x, w0.y, a0.y); a0.z = fmaf(g.x, w0.z, a0.z); a0.w = fmaf(g.x, w0.w, a0.w);
        a0.x = fmaf(g.y, w1.x, a0.x); a0.y = fmaf(g.y, w1.y, a0.y); a0.z = fmaf(g.y, w1.z, a0.z); a0.w = fmaf(g.y, w1.w, a0.w);
        a1.x = fmaf(g.z, w0.x, a1.x); a1.y = fmaf(g.z, w0.y, a1.y); a1.z = fmaf(g.z, w0.z, a1.z); a1.w = fmaf(g.z, w0.w, a1.w);
        a1.x = fmaf(g.w, w1.x, a1.x); a1.y = fmaf(g.w, w1.y, a1.y); a1.z = fmaf(g.w, w1.z, a1.z); a1.w = fmaf(g.w, w1.w, a1.w);
      }
      if (2 * pp < npx) {
        float* d = dx + ((size_t)row * W + j0 + 2 * pp) * 16 + q * 4;
        *reinterpret_cast<float4*>(d) = a0;
        *reinterpret_cast<float4*>(d + 16) = a1;
      }
    }
    if (dw != nullptr) {
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int px0 = (warp * 2 + kk) * 8;
        if (px0 >= npx) continue;
        uint32_t ah[4], al[4];
        ob_split(xsb[(px0 + t4) * OB_XS + gq], ah[0], al[0]);
        ob_split(xsb[(px0 + t4) * OB_XS + gq + 8], ah[1], al[1]);
        ob_split(xsb[(px0 + t4 + 4) * OB_XS + gq], ah[2], al[2]);
        ob_split(xsb[(px0 + t4 + 4) * OB_XS + gq + 8], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int j = nt * 8 + gq;       // column of the B fragment: (co, ky, kx)
          if (nt * 8 < J) {
            const float* gr = gsb + (j >> 1) * OB_GS + (j & 1);
            uint32_t bh[2], bl[2];
            ob_split(j < J ? gr[2 * (px0 + t4)] : 0.f, bh[0], bl[0]);
            ob_split(j < J ? gr[2 * (px0 + t4 + 4)] : 0.f, bh[1], bl[1]);
            ob_mma(acc[nt], ah, bh);
            ob_mma(acc[nt], al, bh);
            ob_mma(acc[nt], ah, bl);
          }
        }
      }
    }
    __syncthreads();        // the buffer is free for the loads of tile t + 2
  }
  if (dw != nullptr) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int c0 = nt * 8 + 2 * t4;
      if (c0 < J) {       // J is a multiple of 4: c0 + 1 < J as well
        atomicAdd(red + gq * J + c0, acc[nt][0]);
        atomicAdd(red + gq * J + c0 + 1, acc[nt][1]);
        atomicAdd(red + (gq + 8) * J + c0, acc[nt][2]);
        atomicAdd(red + (gq + 8) * J + c0 + 1, acc[nt][3]);
      }
    }
  }
  if (db != nullptr) {
#pragma unroll
    for (int k = 0; k < NT; ++k) {
      const int r = warp + 4 * k;
      const float v = warp_sum(bsum[k]);
      if (lane == 0 && r < R) atomicAdd(bred + r, v);
    }
  }
  __syncthreads();
  if (dw != nullptr)
    for (int e = tid; e < 16 * J; e += OB_THREADS) atomicAdd(dw + e, red[e]);      // dw is [16][Ccls][2][2] = [ci][J]
  if (db != nullptr && tid < Ccls) atomicAdd(db + tid, bred[2 * tid] + bred[2 * tid + 1]);
}

template <int NT>
static int outconv_bwd2_launch(const float* dl, const float* x, const float* w, float* dx, float* dw, float* db, int N, int H,
                               int W, int Ccls, cudaStream_t s) {
  const size_t smem = sizeof(float) * ((size_t)2 * (2 * Ccls * OB_GS + OB_PX * OB_XS) + Ccls * 64 + 16 * 4 * Ccls + 2 * Ccls);
  MDIL_CUDA(cudaFuncSetAttribute(outconv_bwd2_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = N * H * ((W + OB_PX - 1) / OB_PX);
  const int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));
  int grid = kNumSMs * (per_sm < 1 ? 1 : (per_sm > 3 ? 3 : per_sm));
  if (grid > tiles) grid = tiles;
  outconv_bwd2_kernel<NT><<<grid, OB_THREADS, smem, s>>>(dl, x, w, dx, dw, db, N, H, W, Ccls);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_outconv_bwd(const float* dlogits, const float* x, const float* w, float* dx, float* dw, float* db, int N,
                       int H, int W, int Ccls, cudaStream_t s) {
  MDIL_REQUIRE(Ccls >= 1 && Ccls <= kMaxCls, "outconv: Ccls must be 1..32");
  size_t P = (size_t)N * H * W;
  static const bool fused = [] { const char* e = getenv("MDIL_HEAD_FUSED"); return !(e != nullptr && strcmp(e, "0") == 0); }();
  if (fused && W % 2 == 0 && P < (1ull << 31) && ((uintptr_t)dlogits & 15) == 0) {
    if (dw != nullptr) MDIL_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 16 * Ccls * 4, s));
    if (db != nullptr) MDIL_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Ccls, s));
    if (Ccls <= 20) return outconv_bwd2_launch<10>(dlogits, x, w, dx, dw, db, N, H, W, Ccls, s);
    if (Ccls <= 28) return outconv_bwd2_launch<14>(dlogits, x, w, dx, dw, db, N, H, W, Ccls, s);
    return outconv_bwd2_launch<16>(dlogits, x, w, dx, dw, db, N, H, W, Ccls, s);
  }
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

// Two-phase form used by training (the normaliser sum(w) is known only after the whole batch has been seen):
//   MODE 0  loss sums only: logits are read once, nothing is written;
//   MODE 1  logit gradient w[y] * (softmax - onehot) * (*grad_out) / acc[1], softmax recomputed from the logits.
// 252 MB fewer written and 252 MB fewer read per step at 6 x 20 x 512 x 1024 than "unnormalised gradient in the forward
// pass + in-place rescale in the backward pass".  VEC adjacent pixels per thread (64-bit plane accesses when H*W is even).
// Class planes are loaded with a CLAMPED plane index and masked afterwards (-inf -> exp = 0): no branch sits between
// the loads, so all CMAX of them are in flight together (a per-class "if (c < C)" compiles to branches that serialise
// the loads: one DRAM round trip per class).
template <int MODE, int VEC, int CMAX>
__global__ void __launch_bounds__(256)
ce2d_phase_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ class_w,
                  int C, size_t HWV, size_t PV, double* __restrict__ acc, const float* __restrict__ grad_out,
                  float* __restrict__ dlogits) {
  __shared__ double sh[16];
  __shared__ float ws[kMaxCls];
  if (threadIdx.x < kMaxCls) ws[threadIdx.x] = threadIdx.x < C ? __ldg(class_w + threadIdx.x) : 0.f;
  __syncthreads();
  float scale = 1.f;
  if (MODE == 1) scale = (float)((grad_out != nullptr ? (double)__ldg(grad_out) : 1.0) / acc[1]);
  const size_t HW = HWV * VEC;
  double num = 0.0, den = 0.0;
  for (size_t pv = blockIdx.x * (size_t)256 + threadIdx.x; pv < PV; pv += (size_t)gridDim.x * 256) {
    const size_t n = pv / HWV, hw = (pv - n * HWV) * VEC;
    const float* lp = logits + n * C * HW + hw;
    float x[CMAX][VEC];
    float m[VEC], z[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { m[v] = -INFINITY; z[v] = 0.f; }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      const float* pc = lp + (size_t)(c < C ? c : C - 1) * HW;
      if (VEC == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(pc));
        x[c][0] = t.x; x[c][VEC - 1] = t.y;
      } else {
        x[c][0] = __ldg(pc);
      }
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c)
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        x[c][v] = c < C ? x[c][v] : -INFINITY;
        m[v] = fmaxf(m[v], x[c][v]);
      }
    long long y[VEC];
    float wy[VEC], xy[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      y[v] = labels[n * HW + hw + v];
      const bool ok = y[v] >= 0 && y[v] < C;
      wy[v] = ok ? ws[ok ? y[v] : 0] : 0.f;
      xy[v] = 0.f;
      if (!ok) y[v] = -1;
    }
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (MODE == 0) xy[v] = c == (int)y[v] ? x[c][v] : xy[v];
        x[c][v] = expf(x[c][v] - m[v]);
        z[v] += x[c][v];
      }
    }
    if (MODE == 0) {
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        if (y[v] >= 0) { num += (double)(wy[v] * (logf(z[v]) + m[v] - xy[v])); den += (double)wy[v]; }
    } else {
      float* dp = dlogits + n * C * HW + hw;
      float inv[VEC], sub[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) { sub[v] = wy[v] * scale; inv[v] = sub[v] / z[v]; }
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
          float o[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) o[v] = x[c][v] * inv[v] - (c == (int)y[v] ? sub[v] : 0.f);
          if (VEC == 2) *reinterpret_cast<float2*>(dp + (size_t)c * HW) = make_float2(o[0], o[VEC - 1]);
          else dp[(size_t)c * HW] = o[0];
        }
      }
    }
  }
  if (MODE == 0) block_atomic_add2(num, den, acc, sh);
}

int launch_ce2d_bwd(const float* logits, const int64_t* labels, const float* class_w, int N, int C, int H, int W,
                    const double* acc, const float* grad_out, float* dlogits, cudaStream_t s) {
  MDIL_REQUIRE(C >= 1 && C <= kMaxCls, "ce2d: C must be 1..32");
  const size_t HW = (size_t)H * W, P = (size_t)N * HW;
  const bool v2 = HW % 2 == 0 && ((uintptr_t)logits & 7) == 0 && ((uintptr_t)dlogits & 7) == 0;
  const size_t PV = v2 ? P / 2 : P;
  int grid = (int)((PV + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  double* acc_rw = const_cast<double*>(acc);     // MODE 1 only reads it
  if (v2 && C <= 20) ce2d_phase_kernel<1, 2, 20><<<grid, 256, 0, s>>>(logits, labels, class_w, C, HW / 2, PV, acc_rw, grad_out, dlogits);
  else if (v2) ce2d_phase_kernel<1, 2, kMaxCls><<<grid, 256, 0, s>>>(logits, labels, class_w, C, HW / 2, PV, acc_rw, grad_out, dlogits);
  else ce2d_phase_kernel<1, 1, kMaxCls><<<grid, 256, 0, s>>>(logits, labels, class_w, C, HW, PV, acc_rw, grad_out, dlogits);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_ce2d(const float* logits, const int64_t* labels, const float* class_w, int N, int C, int H, int W,
                float* loss, double* acc, float* dlogits, cudaStream_t s) {
  MDIL_REQUIRE(C >= 1 && C <= kMaxCls, "ce2d: C must be 1..32");
  size_t HW = (size_t)H * W, P = (size_t)N * HW;
  MDIL_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(double), s));
  int grid = (int)((P + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (dlogits == nullptr && HW % 2 == 0 && ((uintptr_t)logits & 7) == 0) {
    grid = (int)((P / 2 + 255) / 256);
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    if (C <= 20) ce2d_phase_kernel<0, 2, 20><<<grid, 256, 0, s>>>(logits, labels, class_w, C, HW / 2, P / 2, acc, nullptr, nullptr);
    else ce2d_phase_kernel<0, 2, kMaxCls><<<grid, 256, 0, s>>>(logits, labels, class_w, C, HW / 2, P / 2, acc, nullptr, nullptr);
  } else {
    ce2d_kernel<<<grid, 256, 0, s>>>(logits, labels, class_w, C, HW, P, acc, dlogits);
  }
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
    for (int c = 0; c < kMaxCls; ++c) {     // clamped plane index, masked below: all loads in flight together
      const size_t off = (size_t)(c < C ? c : C - 1) * HW;
      sx[c] = __ldg(sp + off);
      tx[c] = __ldg(tp + off);
    }
#pragma unroll
    for (int c = 0; c < kMaxCls; ++c) {
      if (c < C) { ms = fmaxf(ms, sx[c]); mt = fmaxf(mt, tx[c]); }
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
