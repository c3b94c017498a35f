// Generic gather implicit-GEMM kernels over NHWC fp32 tensors ("tap" formulation):
//   conv_taps  : out[g(v)] = bias + sum_t A[a_t(v)] . W[widx_t]        (forward / data-gradient)
//   wgrad_taps : dW[widx_t] += sum_v A[a_t(v)] (x) G[g(v)]             (weight-gradient, + bias gradient)
// where v walks a virtual pixel grid, a_t(v) = v*a_s + a_d[t] and g(v) = v*g_s + o_d(class).
// One geometry description covers: the DownsamplerBlock 3x3 stride-2 conv and its dgrad
// (models/erfnet_RA_parallel.py:17,23), the UpsamplerBlock 3x3 stride-2 transposed conv as four
// sub-pixel parity classes with 1/2/2/4 taps and its dgrad (:155-156), and every weight gradient.
// FP32 FFMA register-tiled GEMM (8x4 / TMxTN per thread), operands staged through shared memory.
#include <atomic>

#include "kernels.cuh"

#include <stdlib.h>
#include <string.h>

namespace mdil {

// ============================================================================ forward / dgrad
constexpr int CT_M = 128;  // pixels per CTA
constexpr int CT_N = 64;   // output channels per CTA
constexpr int CT_K = 16;   // input channels per chunk

__global__ void __launch_bounds__(256)
conv_taps_kernel(const __grid_constant__ ConvGeom g, const float* __restrict__ A, const float* __restrict__ Wp,
                 const float* __restrict__ bias, float* __restrict__ out) {
  __shared__ __align__(16) float As[CT_K][CT_M + 4];
  __shared__ __align__(16) float Bs[CT_K][CT_N];
  const TapClass& tc = g.cls[blockIdx.z];
  const int tid = threadIdx.x;
  const int tn = tid & 15, tm = tid >> 4;
  const int lm = tid & 127, lq = tid >> 7;
  const size_t P = (size_t)g.N * g.VH * g.VW;
  const size_t lp = (size_t)blockIdx.x * CT_M + lm;
  const bool lvalid = lp < P;
  int ln = 0, lvy = 0, lvx = 0;
  if (lvalid) {
    const unsigned lq32 = (unsigned)lp;
    lvx = (int)(lq32 % (unsigned)g.VW);
    const unsigned t = lq32 / (unsigned)g.VW;
    lvy = (int)(t % (unsigned)g.VH);
    ln = (int)(t / (unsigned)g.VH);
  }
  const int n0 = blockIdx.y * CT_N;
  const int brow = tid >> 4, bcol = (tid & 15) * 4;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < tc.ntaps; ++t) {
    const int ay = lvy * g.a_sy + tc.a_dy[t], ax = lvx * g.a_sx + tc.a_dx[t];
    const bool inb = lvalid && ay >= 0 && ay < g.AH && ax >= 0 && ax < g.AW;
    const float* arow = A + (((size_t)ln * g.AH + (inb ? ay : 0)) * g.AW + (inb ? ax : 0)) * g.lda + g.a_coff;
    const float* wslab = Wp + (size_t)tc.widx[t] * g.CIN * g.COUT_PAD;
    for (int c0 = 0; c0 < g.CIN; c0 += CT_K) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int q = lq + 2 * h;
        const int k = c0 + 4 * q;
        float4 v = make4(0.f);
        if (inb && k < g.CIN) v = ldg4(arow + k);
        As[4 * q + 0][lm] = v.x;
        As[4 * q + 1][lm] = v.y;
        As[4 * q + 2][lm] = v.z;
        As[4 * q + 3][lm] = v.w;
      }
      {
        const int kb = c0 + brow, co = n0 + bcol;
        float4 v = make4(0.f);
        if (kb < g.CIN && co < g.COUT_PAD) v = ldg4(wslab + (size_t)kb * g.COUT_PAD + co);
        *reinterpret_cast<float4*>(&Bs[brow][bcol]) = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < CT_K; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[k][tm * 8 + 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i][0] = fmaf(a[i], b.x, acc[i][0]);
          acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
          acc[i][2] = fmaf(a[i], b.z, acc[i][2]);
          acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
        }
      }
      __syncthreads();
    }
  }

  const int co = n0 + tn * 4;
  if (co >= g.COUT) return;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (co + j < g.COUT) bv[j] = __ldg(bias + co + j);
  }
  const bool vec = ((g.ldg | g.g_coff) & 3) == 0 && co + 3 < g.COUT;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const size_t p = (size_t)blockIdx.x * CT_M + tm * 8 + i;
    if (p >= P) continue;
    const unsigned p32 = (unsigned)p;
    const int vx = (int)(p32 % (unsigned)g.VW);
    const unsigned tt = p32 / (unsigned)g.VW;
    const int vy = (int)(tt % (unsigned)g.VH);
    const int n = (int)(tt / (unsigned)g.VH);
    const int oy = vy * g.g_sy + tc.o_dy, ox = vx * g.g_sx + tc.o_dx;
    if (oy < 0 || oy >= g.GH || ox < 0 || ox >= g.GW) continue;
    float* o = out + (((size_t)n * g.GH + oy) * g.GW + ox) * g.ldg + g.g_coff + co;
    if (vec) {
      *reinterpret_cast<float4*>(o) = make_float4(acc[i][0] + bv[0], acc[i][1] + bv[1], acc[i][2] + bv[2], acc[i][3] + bv[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (co + j < g.COUT) o[j] = acc[i][j] + bv[j];
    }
  }
}

// ---- the same contract on the warp-level tensor-core path (mma.sync m16n8k8, 3xTF32, fp32 accumulate).
// One warp = 16 virtual pixels x 16 output channels; K walks (tap, 16-channel group): one 128-bit load per pixel row
// and group feeds two K steps (lane t owns channels 4t..4t+3 of the group; the weight fragment uses the same channel
// permutation).  The CTA's [taps][CIN][16] weight slice is staged once in shared memory (row pitch 18 floats:
// conflict-free fragment reads).  CIN = 4 (the network input, NHWC padded to 4): one K step = two taps.
__device__ __forceinline__ void mma_16x8x8_c(float (&dd)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(dd[0]), "+f"(dd[1]), "+f"(dd[2]), "+f"(dd[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_tf32_c(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xFFFFE000u;
}
constexpr int CM_WP = 18;   // shared-memory weight row pitch (floats)

__global__ void __launch_bounds__(256, 3)
conv_mma_kernel(const __grid_constant__ ConvGeom g, const float* __restrict__ A, const float* __restrict__ Wp,
                const float* __restrict__ bias, float* __restrict__ out, int tiles_per_cta) {
  extern __shared__ __align__(16) float wsm[];   // [total taps][CIN][CM_WP]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, t = lane & 3;
  const int co0 = blockIdx.y * 16;
  int tapbase[kMaxClasses + 1];
  tapbase[0] = 0;
  for (int c = 0; c < g.nclasses; ++c) tapbase[c + 1] = tapbase[c] + g.cls[c].ntaps;
  // stage this CTA's weight slice: wsm[(slot*CIN + ci)*CM_WP + c] = Wp[widx(slot)][ci][co0 + c]
  for (int c = 0; c < g.nclasses; ++c)
    for (int tp = 0; tp < g.cls[c].ntaps; ++tp) {
      const float* src = Wp + (size_t)g.cls[c].widx[tp] * g.CIN * g.COUT_PAD;
      float* dst = wsm + (size_t)(tapbase[c] + tp) * g.CIN * CM_WP;
      for (int i = tid; i < g.CIN * 16; i += 256) {
        const int ci = i >> 4, cc = i & 15;
        dst[ci * CM_WP + cc] = (co0 + cc < g.COUT_PAD) ? __ldg(src + (size_t)ci * g.COUT_PAD + co0 + cc) : 0.f;
      }
    }
  __syncthreads();
  float bv[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  if (bias != nullptr) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int co = co0 + nt * 8 + 2 * t + e;
        if (co < g.COUT) bv[nt][e] = __ldg(bias + co);
      }
  }
  const size_t P = (size_t)g.N * g.VH * g.VW;
  const size_t ntiles = (P + 15) / 16;
  const float* Ab = A + g.a_coff;
  for (int it = 0; it < tiles_per_cta; ++it) {
    const size_t tile = ((size_t)blockIdx.x * tiles_per_cta + it) * 8 + warp;
    if (tile >= ntiles) break;
    // this lane's two pixel rows of the 16-pixel tile: r0 = gq, r1 = gq + 8
    int pn[2], py[2], px[2];
    bool pv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const size_t p = tile * 16 + gq + 8 * r;
      pv[r] = p < P;
      const unsigned p32 = pv[r] ? (unsigned)p : 0u;
      px[r] = (int)(p32 % (unsigned)g.VW);
      const unsigned tt = p32 / (unsigned)g.VW;
      py[r] = (int)(tt % (unsigned)g.VH);
      pn[r] = (int)(tt / (unsigned)g.VH);
    }
    for (int c = 0; c < g.nclasses; ++c) {
      const TapClass& tc = g.cls[c];
      float acc[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
      if (g.CIN == 4) {
        for (int tp = 0; tp < tc.ntaps; tp += 2) {
          float av[4] = {0.f, 0.f, 0.f, 0.f};   // a0: row0 tap A, a1: row1 tap A, a2: row0 tap B, a3: row1 tap B (channel t)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (tp + h >= tc.ntaps) continue;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const int ay = py[r] * g.a_sy + tc.a_dy[tp + h], ax = px[r] * g.a_sx + tc.a_dx[tp + h];
              if (pv[r] && (unsigned)ay < (unsigned)g.AH && (unsigned)ax < (unsigned)g.AW)
                av[2 * h + r] = __ldg(Ab + (size_t)((unsigned)(pn[r] * g.AH + ay) * (unsigned)g.AW + (unsigned)ax) * (unsigned)g.lda + t);
            }
          }
          uint32_t ah[4], al[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split_tf32_c(av[e], ah[e], al[e]);
          const float* w0 = wsm + (size_t)((tapbase[c] + tp) * 4 + t) * CM_WP;
          const float* w1 = w0 + 4 * CM_WP;
          const bool has1 = tp + 1 < tc.ntaps;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            uint32_t bh[2], bl[2];
            split_tf32_c(w0[nt * 8 + gq], bh[0], bl[0]);
            split_tf32_c(has1 ? w1[nt * 8 + gq] : 0.f, bh[1], bl[1]);
            mma_16x8x8_c(acc[nt], ah, bh);
            mma_16x8x8_c(acc[nt], al, bh);
            mma_16x8x8_c(acc[nt], ah, bl);
          }
        }
      } else {
        for (int tp = 0; tp < tc.ntaps; ++tp) {
          const float* ar[2];
          bool inb[2];
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int ay = py[r] * g.a_sy + tc.a_dy[tp], ax = px[r] * g.a_sx + tc.a_dx[tp];
            inb[r] = pv[r] && (unsigned)ay < (unsigned)g.AH && (unsigned)ax < (unsigned)g.AW;
            ar[r] = Ab + (size_t)((unsigned)(pn[r] * g.AH + (inb[r] ? ay : 0)) * (unsigned)g.AW + (unsigned)(inb[r] ? ax : 0)) * (unsigned)g.lda + 4 * t;
          }
          const float* wt = wsm + (size_t)(tapbase[c] + tp) * g.CIN * CM_WP;
#pragma unroll 2
          for (int cg = 0; cg < g.CIN; cg += 16) {
            const float4 x0 = inb[0] ? ldg4(ar[0] + cg) : make4(0.f);
            const float4 x1 = inb[1] ? ldg4(ar[1] + cg) : make4(0.f);
            const float xa[2][4] = {{x0.x, x1.x, x0.y, x1.y}, {x0.z, x1.z, x0.w, x1.w}};   // per K step: a0..a3
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              uint32_t ah[4], al[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) split_tf32_c(xa[ks][e], ah[e], al[e]);
              const float* w0 = wt + (size_t)(cg + 4 * t + 2 * ks) * CM_WP;
#pragma unroll
              for (int nt = 0; nt < 2; ++nt) {
                uint32_t bh[2], bl[2];
                split_tf32_c(w0[nt * 8 + gq], bh[0], bl[0]);
                split_tf32_c(w0[CM_WP + nt * 8 + gq], bh[1], bl[1]);
                mma_16x8x8_c(acc[nt], ah, bh);
                mma_16x8x8_c(acc[nt], al, bh);
                mma_16x8x8_c(acc[nt], ah, bl);
              }
            }
          }
        }
      }
      // ---- epilogue of this class: rows gq / gq+8, columns co0 + nt*8 + 2t, +1
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (!pv[r]) continue;
        const int oy = py[r] * g.g_sy + tc.o_dy, ox = px[r] * g.g_sx + tc.o_dx;
        if ((unsigned)oy >= (unsigned)g.GH || (unsigned)ox >= (unsigned)g.GW) continue;
        float* o = out + (size_t)((unsigned)(pn[r] * g.GH + oy) * (unsigned)g.GW + (unsigned)ox) * (unsigned)g.ldg + g.g_coff;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int co = co0 + nt * 8 + 2 * t;
          const float v0 = acc[nt][2 * r] + bv[nt][0], v1 = acc[nt][2 * r + 1] + bv[nt][1];
          if (co + 1 < g.COUT && ((g.ldg | g.g_coff) & 1) == 0) *reinterpret_cast<float2*>(o + co) = make_float2(v0, v1);
          else {
            if (co < g.COUT) o[co] = v0;
            if (co + 1 < g.COUT) o[co + 1] = v1;
          }
        }
      }
    }
  }
}

static int conv_mma_launch(const ConvGeom& g, const float* A, const float* Wp, const float* bias, float* out, cudaStream_t s) {
  int total_taps = 0;
  for (int c = 0; c < g.nclasses; ++c) total_taps += g.cls[c].ntaps;
  const size_t smem = (size_t)total_taps * g.CIN * CM_WP * sizeof(float);
  MDIL_REQUIRE(smem <= 100 * 1024, "conv_mma: weight slice too large");
  static std::atomic<bool> attr_set[kMaxDevices];
  std::atomic<bool>& done = attr_set[current_device_slot()];
  if (!done.load(std::memory_order_acquire)) {
    MDIL_CUDA(cudaFuncSetAttribute(conv_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    done.store(true, std::memory_order_release);
  }
  const size_t P = (size_t)g.N * g.VH * g.VW;
  const size_t ntiles = (P + 15) / 16;
  const int coblocks = cdiv(g.COUT, 16);
  // persistent-ish: ~4 CTAs per SM in total, each walking tiles_per_cta groups of 8 warp tiles
  size_t groups = (ntiles + 7) / 8;
  int want = cdiv(6 * kNumSMs, coblocks);
  int tpc = (int)((groups + want - 1) / want);
  if (tpc < 1) tpc = 1;
  dim3 grid((unsigned)((groups + tpc - 1) / tpc), (unsigned)coblocks);
  conv_mma_kernel<<<grid, 256, smem, s>>>(g, A, Wp, bias, out, tpc);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_conv_taps(const ConvGeom& g, const float* A, const float* Wp, const float* bias, float* out, cudaStream_t s) {
  MDIL_REQUIRE((size_t)g.N * g.VH * g.VW < (1ull << 31), "conv_taps: more than 2^31 pixels");
  MDIL_REQUIRE(g.CIN % 4 == 0 && g.lda % 4 == 0 && g.a_coff % 4 == 0 && g.COUT_PAD % 4 == 0, "conv_taps: alignment");
  MDIL_REQUIRE(g.nclasses >= 1 && g.nclasses <= kMaxClasses, "conv_taps: classes");
  static const bool use_mma = [] { const char* e = getenv("MDIL_CONV_IMPL"); return !(e != nullptr && strcmp(e, "ffma") == 0); }();
  if (use_mma && (g.CIN == 4 || g.CIN % 16 == 0)) return conv_mma_launch(g, A, Wp, bias, out, s);
  size_t P = (size_t)g.N * g.VH * g.VW;
  dim3 grid((unsigned)((P + CT_M - 1) / CT_M), (unsigned)cdiv(g.COUT, CT_N), (unsigned)g.nclasses);
  conv_taps_kernel<<<grid, 256, 0, s>>>(g, A, Wp, bias, out);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ============================================================================ weight gradient
template <int T>
__device__ __forceinline__ void load_frag(const float* row, int t, float (&f)[T]) {
  if constexpr (T == 1) {
    f[0] = row[t];
  } else if constexpr (T == 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + t * 4);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  } else {
    const float4 v = *reinterpret_cast<const float4*>(row + t * 4);
    const float4 w = *reinterpret_cast<const float4*>(row + 64 + t * 4);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    f[4] = w.x; f[5] = w.y; f[6] = w.z; f[7] = w.w;
  }
}
template <int T>
__device__ __forceinline__ int frag_index(int t, int i) {
  if constexpr (T == 8) return i < 4 ? t * 4 + i : 64 + t * 4 + (i - 4);
  return t * T + i;
}

template <int TMW, int TNW, int KP>
__global__ void __launch_bounds__(256)
wgrad_taps_kernel(const __grid_constant__ ConvGeom g, const float* __restrict__ A, const float* __restrict__ a_scale,
                  const float* __restrict__ a_shift, const float* __restrict__ G, float* __restrict__ dW, long s_ci,
                  long s_co, long s_t, float* __restrict__ db, int pixels_per_cta) {
  constexpr int CI_TILE = 16 * TMW, CO_TILE = 16 * TNW;
  __shared__ __align__(16) float As[KP][CI_TILE];
  __shared__ __align__(16) float Gs[KP][CO_TILE];
  int cls = 0, t = blockIdx.y;
  while (t >= g.cls[cls].ntaps) { t -= g.cls[cls].ntaps; ++cls; }
  const TapClass& tc = g.cls[cls];
  const int ci_tiles = (g.CIN + CI_TILE - 1) / CI_TILE;
  const int ci0 = (blockIdx.z % ci_tiles) * CI_TILE, co0 = (blockIdx.z / ci_tiles) * CO_TILE;
  const int tid = threadIdx.x, tn = tid & 15, tm = tid >> 4;
  const size_t P = (size_t)g.N * g.VH * g.VW;
  const size_t p_begin = (size_t)blockIdx.x * pixels_per_cta;
  const size_t p_end = p_begin + pixels_per_cta < P ? p_begin + pixels_per_cta : P;
  const bool do_bias = db != nullptr && t == 0 && ci0 == 0 && tm == 0;

  float acc[TMW][TNW];
  float bsum[TNW];
#pragma unroll
  for (int i = 0; i < TMW; ++i)
#pragma unroll
    for (int j = 0; j < TNW; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int j = 0; j < TNW; ++j) bsum[j] = 0.f;

  for (size_t p0 = p_begin; p0 < p_end; p0 += KP) {
    for (int idx = tid; idx < KP * (CI_TILE / 4); idx += 256) {
      const int kp = idx / (CI_TILE / 4), f4 = idx % (CI_TILE / 4);
      const unsigned p = (unsigned)p0 + kp;
      const int ci = ci0 + f4 * 4;
      float4 v = make4(0.f);
      if (p < (unsigned)p_end && ci < g.CIN) {
        const int vx = (int)(p % (unsigned)g.VW);
        const unsigned tt = p / (unsigned)g.VW;
        const int vy = (int)(tt % (unsigned)g.VH);
        const int n = (int)(tt / (unsigned)g.VH);
        const int ay = vy * g.a_sy + tc.a_dy[t], ax = vx * g.a_sx + tc.a_dx[t];
        if (ay >= 0 && ay < g.AH && ax >= 0 && ax < g.AW) {
          v = ldg4(A + (((size_t)n * g.AH + ay) * g.AW + ax) * g.lda + g.a_coff + ci);
          if (a_scale != nullptr) {
            const float4 sc = ldg4(a_scale + ci), sh = ldg4(a_shift + ci);
            v.x = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f);
            v.y = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
            v.z = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f);
            v.w = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
          }
        }
      }
      *reinterpret_cast<float4*>(&As[kp][f4 * 4]) = v;
    }
    for (int idx = tid; idx < KP * (CO_TILE / 4); idx += 256) {
      const int kp = idx / (CO_TILE / 4), f4 = idx % (CO_TILE / 4);
      const unsigned p = (unsigned)p0 + kp;
      const int co = co0 + f4 * 4;
      float4 v = make4(0.f);
      if (p < (unsigned)p_end && co < g.COUT) {
        const int vx = (int)(p % (unsigned)g.VW);
        const unsigned tt = p / (unsigned)g.VW;
        const int vy = (int)(tt % (unsigned)g.VH);
        const int n = (int)(tt / (unsigned)g.VH);
        const int gy = vy * g.g_sy + tc.o_dy, gx = vx * g.g_sx + tc.o_dx;
        if (gy >= 0 && gy < g.GH && gx >= 0 && gx < g.GW)
          v = ldg4(G + (((size_t)n * g.GH + gy) * g.GW + gx) * g.ldg + g.g_coff + co);
      }
      *reinterpret_cast<float4*>(&Gs[kp][f4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int kp = 0; kp < KP; ++kp) {
      float a[TMW], gg[TNW];
      load_frag<TMW>(&As[kp][0], tm, a);
      load_frag<TNW>(&Gs[kp][0], tn, gg);
#pragma unroll
      for (int i = 0; i < TMW; ++i)
#pragma unroll
        for (int j = 0; j < TNW; ++j) acc[i][j] = fmaf(a[i], gg[j], acc[i][j]);
      if (do_bias) {
#pragma unroll
        for (int j = 0; j < TNW; ++j) bsum[j] += gg[j];
      }
    }
    __syncthreads();
  }

  float* slab = dW + (long)tc.widx[t] * s_t;
#pragma unroll
  for (int i = 0; i < TMW; ++i) {
    const int ci = ci0 + frag_index<TMW>(tm, i);
    if (ci >= g.CIN_VALID) continue;
#pragma unroll
    for (int j = 0; j < TNW; ++j) {
      const int co = co0 + frag_index<TNW>(tn, j);
      if (co < g.COUT) atomicAdd(slab + (long)ci * s_ci + (long)co * s_co, acc[i][j]);
    }
  }
  if (do_bias) {
#pragma unroll
    for (int j = 0; j < TNW; ++j) {
      const int co = co0 + frag_index<TNW>(tn, j);
      if (co < g.COUT) atomicAdd(db + co, bsum[j]);
    }
  }
}

template <int TMW, int TNW, int KP>
static int wgrad_launch(const ConvGeom& g, const float* A, const float* a_scale, const float* a_shift, const float* G,
                        float* dW, long s_ci, long s_co, long s_t, float* db, cudaStream_t s) {
  constexpr int CI_TILE = 16 * TMW, CO_TILE = 16 * TNW;
  int total_taps = 0;
  for (int c = 0; c < g.nclasses; ++c) total_taps += g.cls[c].ntaps;
  int tiles = cdiv(g.CIN, CI_TILE) * cdiv(g.COUT, CO_TILE);
  size_t P = (size_t)g.N * g.VH * g.VW;
  int want = cdiv(2 * kNumSMs, total_taps * tiles);
  if (want < 1) want = 1;
  size_t ppc = (P + want - 1) / want;
  ppc = (ppc + KP - 1) / KP * KP;
  if (ppc < (size_t)KP) ppc = KP;
  int ksplit = (int)((P + ppc - 1) / ppc);
  dim3 grid((unsigned)ksplit, (unsigned)total_taps, (unsigned)tiles);
  wgrad_taps_kernel<TMW, TNW, KP><<<grid, 256, 0, s>>>(g, A, a_scale, a_shift, G, dW, s_ci, s_co, s_t, db, (int)ppc);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ---- small-channel weight gradient (CIN <= 16, COUT <= 16: the C = 16 decoder blocks and the 3->13 initial conv).
// HBM/L1-bound: 16 threads share a pixel (4x4 register tiles over ci x co), 16 pixel lanes per CTA, up to 3 taps per
// pass; operands come straight from global memory (the 64-byte pixel rows are L1-resident across the 16 threads).
template <int NT>
__global__ void __launch_bounds__(256, 2)
wgrad_small_kernel(const __grid_constant__ ConvGeom g, const float* __restrict__ A, const float* __restrict__ a_scale,
                   const float* __restrict__ a_shift, const float* __restrict__ G, float* __restrict__ dW, long s_ci,
                   long s_co, long s_t, float* __restrict__ db, int pixels_per_cta) {
  __shared__ float red[8][16][NT * 16 + 4];
  const TapClass& tc = g.cls[0];
  const int tid = threadIdx.x, tile = tid & 15, pl = tid >> 4;
  const int ci0 = (tile >> 2) * 4, co0 = (tile & 3) * 4;
  const int tap0 = blockIdx.y * NT;
  const size_t P = (size_t)g.N * g.VH * g.VW;
  const size_t p_begin = (size_t)blockIdx.x * pixels_per_cta;
  const size_t p_end = p_begin + pixels_per_cta < P ? p_begin + pixels_per_cta : P;
  const bool ci_ok = ci0 < g.CIN, co_ok = co0 < g.COUT;
  float4 sc = make4(1.f), sh = make4(0.f);
  if (a_scale != nullptr && ci_ok) { sc = ldg4(a_scale + ci0); sh = ldg4(a_shift + ci0); }
  float acc[NT][4][4];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;
  float4 bsum = make4(0.f);
  // running (n, vy, vx) of this thread's pixel: one division at entry, increments afterwards (32-bit: P < 2^31).
  // Software-pipelined: the loads of U pixels (U * (1 + NT) independent 128-bit loads) are issued before the first FMA.
  constexpr int U = 4;
  unsigned p = (unsigned)p_begin + pl;
  int vx = (int)(p % (unsigned)g.VW);
  int vy = (int)((p / (unsigned)g.VW) % (unsigned)g.VH);
  int n = (int)(p / ((unsigned)g.VW * (unsigned)g.VH));
  const float* Gc = G + g.g_coff + co0;
  const float* Ac = A + g.a_coff + ci0;
  for (; p < (unsigned)p_end; p += 16 * U) {
    float4 gv[U], av[U][NT];
    unsigned inimg = 0;   // bit u*NT+t: the activation tap is inside the image (the BN+ReLU prologue applies)
#pragma unroll
    for (int u = 0; u < U; ++u) {
      gv[u] = make4(0.f);
#pragma unroll
      for (int t = 0; t < NT; ++t) av[u][t] = make4(0.f);
      if (p + 16 * u < (unsigned)p_end) {
        const int gy = vy * g.g_sy + tc.o_dy, gx = vx * g.g_sx + tc.o_dx;
        if (co_ok && (unsigned)gy < (unsigned)g.GH && (unsigned)gx < (unsigned)g.GW)
          gv[u] = ldg4(Gc + (size_t)((unsigned)(n * g.GH + gy) * (unsigned)g.GW + (unsigned)gx) * (unsigned)g.ldg);
        const int ayb = vy * g.a_sy, axb = vx * g.a_sx;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int tap = tap0 + t;
          if (tap < tc.ntaps) {
            const int ay = ayb + tc.a_dy[tap], ax = axb + tc.a_dx[tap];
            if (ci_ok && (unsigned)ay < (unsigned)g.AH && (unsigned)ax < (unsigned)g.AW) {
              av[u][t] = ldg4(Ac + (size_t)((unsigned)(n * g.AH + ay) * (unsigned)g.AW + (unsigned)ax) * (unsigned)g.lda);
              inimg |= 1u << (u * NT + t);
            }
          }
        }
      }
      vx += 16;
      while (vx >= g.VW) { vx -= g.VW; ++vy; }
      while (vy >= g.VH) { vy -= g.VH; ++n; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      bsum.x += gv[u].x; bsum.y += gv[u].y; bsum.z += gv[u].z; bsum.w += gv[u].w;
      const float gg[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        float4 a4 = av[u][t];
        if (a_scale != nullptr && ((inimg >> (u * NT + t)) & 1u)) {
          a4.x = fmaxf(fmaf(a4.x, sc.x, sh.x), 0.f); a4.y = fmaxf(fmaf(a4.y, sc.y, sh.y), 0.f);
          a4.z = fmaxf(fmaf(a4.z, sc.z, sh.z), 0.f); a4.w = fmaxf(fmaf(a4.w, sc.w, sh.w), 0.f);
        }
        const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[t][i][j] = fmaf(aa[i], gg[j], acc[t][i][j]);
      }
    }
  }
  // reduce over the 16 pixel lanes: lanes l and l^16 of a warp share a tile, then across the 8 warps through smem
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[t][i][j];
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (lane < 16) red[warp][tile][t * 16 + i * 4 + j] = v;
      }
  float bs[4] = {bsum.x, bsum.y, bsum.z, bsum.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bs[j] += __shfl_xor_sync(0xffffffffu, bs[j], 16);
    if (lane < 16) red[warp][tile][NT * 16 + j] = bs[j];
  }
  __syncthreads();
  for (int e = tid; e < 16 * (NT * 16 + 4); e += 256) {
    const int tl = e / (NT * 16 + 4), k = e % (NT * 16 + 4);
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[w][tl][k];
    const int ci_b = (tl >> 2) * 4, co_b = (tl & 3) * 4;
    if (k < NT * 16) {
      const int t = k / 16, i = (k % 16) / 4, j = k % 4;
      const int tap = tap0 + t, ci = ci_b + i, co = co_b + j;
      if (tap < tc.ntaps && ci < g.CIN_VALID && co < g.COUT)
        atomicAdd(dW + (long)tc.widx[tap] * s_t + (long)ci * s_ci + (long)co * s_co, v);
    } else if (db != nullptr && blockIdx.y == 0 && ci_b == 0) {
      const int co = co_b + (k - NT * 16);
      if (co < g.COUT) atomicAdd(db + co, v);
    }
  }
}

static int wgrad_small_launch(const ConvGeom& g, const float* A, const float* a_scale, const float* a_shift, const float* G,
                              float* dW, long s_ci, long s_co, long s_t, float* db, cudaStream_t s) {
  const int ntaps = g.cls[0].ntaps;
  const int passes = cdiv(ntaps, 3);
  size_t P = (size_t)g.N * g.VH * g.VW;
  int want = cdiv(2 * kNumSMs, passes);    // 2 CTAs per SM, one wave
  size_t ppc = (P + want - 1) / want;
  ppc = (ppc + 63) / 64 * 64;
  if (ppc < 64) ppc = 64;
  dim3 grid((unsigned)((P + ppc - 1) / ppc), (unsigned)passes);
  wgrad_small_kernel<3><<<grid, 256, 0, s>>>(g, A, a_scale, a_shift, G, dW, s_ci, s_co, s_t, db, (int)ppc);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ---- skinny weight gradients on the warp-level tensor-core path (mma.sync m16n8k8, error-compensated 3xTF32, fp32
// accumulate): M = 16 input channels, N = 8 * NTL output channels, K = pixels.  These reductions (C = 16 decoder blocks,
// the 3->13 / 16->64 downsamplers, the 64->16 upsampler) move ~100 MB for ~1 GFLOP: they are HBM/L2-bound once the FMA
// and shared-memory work of the generic kernels is gone.  Fragments are loaded straight from the NHWC tensors (each
// 32-bit load instruction covers 4 pixels x 32 contiguous bytes), up to TAPS taps of one tap class per pass.
struct MmaPass { int cls, tap0, ntaps; };
struct MmaPlan { int npass; MmaPass pass[12]; };

__device__ __forceinline__ void mma_16x8x8(float (&dd)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(dd[0]), "+f"(dd[1]), "+f"(dd[2]), "+f"(dd[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// hi = x rounded to TF32 (round-half-away on the integer pipe), lo = TF32 bits of the exact remainder
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xFFFFE000u;
}

template <int NTL, int TAPS>
__global__ void __launch_bounds__(256, NTL <= 2 ? 3 : 2)
wgrad_mma_kernel(const __grid_constant__ ConvGeom g, const __grid_constant__ MmaPlan plan, const float* __restrict__ A,
                 const float* __restrict__ a_scale, const float* __restrict__ a_shift, const float* __restrict__ G,
                 float* __restrict__ dW, long s_ci, long s_co, long s_t, float* __restrict__ db, int pixels_per_cta) {
  __shared__ float red[TAPS][16][8 * NTL];
  __shared__ float bred[8 * NTL];
  const MmaPass ps = plan.pass[blockIdx.y];
  const TapClass& tc = g.cls[ps.cls];
  const int ci_blocks = (g.CIN + 15) / 16;
  const int ci0 = (blockIdx.z % ci_blocks) * 16, co0 = (blockIdx.z / ci_blocks) * (8 * NTL);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, t = lane & 3;
  for (int i = tid; i < TAPS * 16 * 8 * NTL; i += 256) (&red[0][0][0])[i] = 0.f;
  if (tid < 8 * NTL) bred[tid] = 0.f;
  __syncthreads();
  const size_t P = (size_t)g.N * g.VH * g.VW;
  const size_t p_begin = (size_t)blockIdx.x * pixels_per_cta;
  const size_t p_end = p_begin + pixels_per_cta < P ? p_begin + pixels_per_cta : P;
  const int ciA = ci0 + gq, ciB = ci0 + gq + 8;
  const bool okA = ciA < g.CIN, okB = ciB < g.CIN;
  float scA = 1.f, shA = 0.f, scB = 1.f, shB = 0.f;
  const bool pro = a_scale != nullptr;
  if (pro) {
    if (okA) { scA = __ldg(a_scale + ciA); shA = __ldg(a_shift + ciA); }
    if (okB) { scB = __ldg(a_scale + ciB); shB = __ldg(a_shift + ciB); }
  }
  const bool do_bias = db != nullptr && ps.tap0 == 0 && ci0 == 0;
  float acc[TAPS][NTL][4];
  float bsum[NTL];
#pragma unroll
  for (int tp = 0; tp < TAPS; ++tp)
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[tp][nt][e] = 0.f;
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt) bsum[nt] = 0.f;
  const float* Ab = A + g.a_coff;
  const float* Gb = G + g.g_coff + co0 + gq;

  // this lane's two pixels of the warp's 8-pixel K step: q0 = base + t, q1 = q0 + 4; the warp advances 64 pixels per step
  unsigned q0 = (unsigned)p_begin + warp * 8 + t;
  int vx = (int)(q0 % (unsigned)g.VW);
  int vy = (int)((q0 / (unsigned)g.VW) % (unsigned)g.VH);
  int n = (int)(q0 / ((unsigned)g.VW * (unsigned)g.VH));
#pragma unroll 2
  for (; q0 - t < (unsigned)p_end; q0 += 64) {
    int vx1 = vx + 4, vy1 = vy, n1 = n;
    if (vx1 >= g.VW) { vx1 -= g.VW; ++vy1; if (vy1 >= g.VH) { vy1 -= g.VH; ++n1; } }
    const bool v0 = q0 < (unsigned)p_end, v1 = q0 + 4 < (unsigned)p_end;
    // gradient fragments (K x N): b0 = G[q0][co], b1 = G[q1][co]
    float gb[NTL][2];
    {
      const int gy0 = vy * g.g_sy + tc.o_dy, gx0 = vx * g.g_sx + tc.o_dx;
      const int gy1 = vy1 * g.g_sy + tc.o_dy, gx1 = vx1 * g.g_sx + tc.o_dx;
      const bool in0 = v0 && (unsigned)gy0 < (unsigned)g.GH && (unsigned)gx0 < (unsigned)g.GW;
      const bool in1 = v1 && (unsigned)gy1 < (unsigned)g.GH && (unsigned)gx1 < (unsigned)g.GW;
      const float* r0 = Gb + (size_t)((unsigned)(n * g.GH + gy0) * (unsigned)g.GW + (unsigned)gx0) * (unsigned)g.ldg;
      const float* r1 = Gb + (size_t)((unsigned)(n1 * g.GH + gy1) * (unsigned)g.GW + (unsigned)gx1) * (unsigned)g.ldg;
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const bool cok = co0 + nt * 8 + gq < g.COUT;
        gb[nt][0] = (in0 && cok) ? __ldg(r0 + nt * 8) : 0.f;
        gb[nt][1] = (in1 && cok) ? __ldg(r1 + nt * 8) : 0.f;
      }
    }
    // activation fragments (M x K) per tap: a0 = A[q0+tap][ciA], a1 = A[q0+tap][ciB], a2 = A[q1+tap][ciA], a3 = A[q1+tap][ciB]
    float af[TAPS][4];
#pragma unroll
    for (int tp = 0; tp < TAPS; ++tp) {
      af[tp][0] = af[tp][1] = af[tp][2] = af[tp][3] = 0.f;
      if (tp < ps.ntaps) {
        const int tap = ps.tap0 + tp;
        const int ay0 = vy * g.a_sy + tc.a_dy[tap], ax0 = vx * g.a_sx + tc.a_dx[tap];
        const int ay1 = vy1 * g.a_sy + tc.a_dy[tap], ax1 = vx1 * g.a_sx + tc.a_dx[tap];
        if (v0 && (unsigned)ay0 < (unsigned)g.AH && (unsigned)ax0 < (unsigned)g.AW) {
          const float* r = Ab + (size_t)((unsigned)(n * g.AH + ay0) * (unsigned)g.AW + (unsigned)ax0) * (unsigned)g.lda;
          if (okA) { const float x = __ldg(r + ciA); af[tp][0] = pro ? fmaxf(fmaf(x, scA, shA), 0.f) : x; }
          if (okB) { const float x = __ldg(r + ciB); af[tp][1] = pro ? fmaxf(fmaf(x, scB, shB), 0.f) : x; }
        }
        if (v1 && (unsigned)ay1 < (unsigned)g.AH && (unsigned)ax1 < (unsigned)g.AW) {
          const float* r = Ab + (size_t)((unsigned)(n1 * g.AH + ay1) * (unsigned)g.AW + (unsigned)ax1) * (unsigned)g.lda;
          if (okA) { const float x = __ldg(r + ciA); af[tp][2] = pro ? fmaxf(fmaf(x, scA, shA), 0.f) : x; }
          if (okB) { const float x = __ldg(r + ciB); af[tp][3] = pro ? fmaxf(fmaf(x, scB, shB), 0.f) : x; }
        }
      }
    }
    uint32_t bh[NTL][2], bl[NTL][2];
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) {
      split_tf32(gb[nt][0], bh[nt][0], bl[nt][0]);
      split_tf32(gb[nt][1], bh[nt][1], bl[nt][1]);
      bsum[nt] += gb[nt][0] + gb[nt][1];
    }
#pragma unroll
    for (int tp = 0; tp < TAPS; ++tp) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_tf32(af[tp][e], ah[e], al[e]);
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        mma_16x8x8(acc[tp][nt], ah, bh[nt]);
        mma_16x8x8(acc[tp][nt], al, bh[nt]);
        mma_16x8x8(acc[tp][nt], ah, bl[nt]);
      }
    }
    vx += 64;
    while (vx >= g.VW) { vx -= g.VW; ++vy; }
    while (vy >= g.VH) { vy -= g.VH; ++n; }
  }
  // ---- CTA reduction in shared memory, then one atomic per output element
#pragma unroll
  for (int tp = 0; tp < TAPS; ++tp)
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) {
      atomicAdd(&red[tp][gq][nt * 8 + 2 * t], acc[tp][nt][0]);
      atomicAdd(&red[tp][gq][nt * 8 + 2 * t + 1], acc[tp][nt][1]);
      atomicAdd(&red[tp][gq + 8][nt * 8 + 2 * t], acc[tp][nt][2]);
      atomicAdd(&red[tp][gq + 8][nt * 8 + 2 * t + 1], acc[tp][nt][3]);
    }
  if (do_bias) {
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) {
      float v = bsum[nt];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (t == 0) atomicAdd(&bred[nt * 8 + gq], v);
    }
  }
  __syncthreads();
  for (int i = tid; i < TAPS * 16 * 8 * NTL; i += 256) {
    const int tp = i / (16 * 8 * NTL), r = (i / (8 * NTL)) % 16, c = i % (8 * NTL);
    const int ci = ci0 + r, co = co0 + c;
    if (tp < ps.ntaps && ci < g.CIN_VALID && co < g.COUT)
      atomicAdd(dW + (long)tc.widx[ps.tap0 + tp] * s_t + (long)ci * s_ci + (long)co * s_co, red[tp][r][c]);
  }
  if (do_bias && tid < 8 * NTL && co0 + tid < g.COUT) atomicAdd(db + co0 + tid, bred[tid]);
}

template <int NTL, int TAPS>
static int wgrad_mma_launch(const ConvGeom& g, const float* A, const float* a_scale, const float* a_shift, const float* G,
                            float* dW, long s_ci, long s_co, long s_t, float* db, cudaStream_t s) {
  MmaPlan plan;
  plan.npass = 0;
  for (int c = 0; c < g.nclasses; ++c)
    for (int t0 = 0; t0 < g.cls[c].ntaps; t0 += TAPS) {
      MDIL_REQUIRE(plan.npass < 12, "wgrad_mma: too many tap passes");
      plan.pass[plan.npass].cls = c;
      plan.pass[plan.npass].tap0 = t0;
      plan.pass[plan.npass].ntaps = g.cls[c].ntaps - t0 < TAPS ? g.cls[c].ntaps - t0 : TAPS;
      ++plan.npass;
    }
  const int zb = cdiv(g.CIN, 16) * cdiv(g.COUT, 8 * NTL);
  const size_t P = (size_t)g.N * g.VH * g.VW;
  int want = cdiv((NTL <= 2 ? 3 : 2) * kNumSMs, plan.npass * zb);    // one wave of co-resident CTAs
  if (want < 1) want = 1;
  size_t ppc = (P + want - 1) / want;
  ppc = (ppc + 63) / 64 * 64;
  if (ppc < 64) ppc = 64;
  dim3 grid((unsigned)((P + ppc - 1) / ppc), (unsigned)plan.npass, (unsigned)zb);
  wgrad_mma_kernel<NTL, TAPS><<<grid, 256, 0, s>>>(g, plan, A, a_scale, a_shift, G, dW, s_ci, s_co, s_t, db, (int)ppc);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_wgrad_taps(const ConvGeom& g, const float* A, const float* a_scale, const float* a_shift, const float* G,
                      float* dW, long s_ci, long s_co, long s_t, float* db, cudaStream_t s) {
  MDIL_REQUIRE(g.CIN % 4 == 0 && g.lda % 4 == 0 && g.a_coff % 4 == 0 && g.ldg % 4 == 0 && g.g_coff % 4 == 0,
               "wgrad_taps: alignment");
  MDIL_REQUIRE((size_t)g.N * g.VH * g.VW < (1ull << 31), "wgrad_taps: more than 2^31 pixels");
  static const bool use_mma = [] { const char* e = getenv("MDIL_WGRAD_SMALL"); return !(e != nullptr && strcmp(e, "ffma") == 0); }();
  if (use_mma && (g.CIN <= 16 || g.COUT <= 16) && g.COUT <= 48 && g.CIN <= 64) {
    if (g.COUT <= 16) return wgrad_mma_launch<2, 3>(g, A, a_scale, a_shift, G, dW, s_ci, s_co, s_t, db, s);
    return wgrad_mma_launch<6, 3>(g, A, a_scale, a_shift, G, dW, s_ci, s_co, s_t, db, s);
  }
  if (g.CIN <= 16 && g.COUT <= 16 && g.nclasses == 1)
    return wgrad_small_launch(g, A, a_scale, a_shift, G, dW, s_ci, s_co, s_t, db, s);
  const int tm = g.CIN >= 128 ? 8 : (g.CIN >= 64 ? 4 : 1);
  const int tn = g.COUT >= 128 ? 8 : (g.COUT > 16 ? 4 : 1);
#define MDIL_WG(TM_, TN_, KP_) \
  if (tm == TM_ && tn == TN_) return wgrad_launch<TM_, TN_, KP_>(g, A, a_scale, a_shift, G, dW, s_ci, s_co, s_t, db, s)
  MDIL_WG(8, 8, 16);
  MDIL_WG(8, 4, 16);
  MDIL_WG(8, 1, 16);
  MDIL_WG(4, 8, 16);
  MDIL_WG(4, 4, 16);
  MDIL_WG(4, 1, 16);
  MDIL_WG(1, 8, 16);
  MDIL_WG(1, 4, 16);
  MDIL_WG(1, 1, 64);
#undef MDIL_WG
  return set_error(-2, "wgrad_taps: no kernel for this shape", __FILE__, __LINE__);
}

// ============================================================================ weight packing
__global__ void pack_kernel(const float* __restrict__ src, float* __restrict__ dst, int T, int A, int Apad, int B, int Bpad,
                            long sa, long sb, long st, int flip) {
  const long total = (long)T * Apad * Bpad;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int b = (int)(i % Bpad);
    const long r = i / Bpad;
    const int a = (int)(r % Apad);
    const int td = (int)(r / Apad);
    const int t = flip ? T - 1 - td : td;
    dst[i] = (a < A && b < B) ? __ldg(src + a * sa + b * sb + t * st) : 0.f;
  }
}

int launch_pack(const float* src, float* dst, int T, int A, int Apad, int B, int Bpad, long sa, long sb, long st, int flip,
                cudaStream_t s) {
  long total = (long)T * Apad * Bpad;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  if (grid < 1) grid = 1;
  pack_kernel<<<grid, 256, 0, s>>>(src, dst, T, A, Apad, B, Bpad, sa, sb, st, flip);
  MDIL_LAUNCH_CHECK();
  return 0;
}

}  // namespace mdil
