// The 3x3 stride-2 convolutions of the samplers on the 5th-generation tensor cores, in both directions:
//   T  transposed (sub-pixel / parity-class) form: UpsamplerBlock forward (models/erfnet_RA_parallel.py:160-172:
//      ConvTranspose2d(Cin, Cout, 3, stride 2, padding 1, output_padding 1)) and the data gradient of DownsamplerBlock's
//      strided convolution (models/erfnet_RA_parallel.py:24-45);
//   S  strided form: DownsamplerBlock forward (Conv2d(Cin, Cout - Cin, 3, stride 2, padding 1)) and the data gradient of
//      UpsamplerBlock's transposed convolution.
// Same arithmetic as nb1d_pair_h3.cu: fp32 operands split into two 16-bit halves (x = hi + lo; hi*hi + lo*hi + hi*lo in
// fp32 TMEM; fp16 halves forward, bf16 halves for gradients), tcgen05.mma kind::f16, M = 128 pixels, K = 16.
//
//   T: out[n][2y + py][2x + px][co] = bias[co] + sum over the taps (dy, dx) of class (py, px), ci:
//                                     A[n][y + dy][x + dx][ci] * W[widx][ci][co]                     dy, dx in {0, 1}
//      The four parity classes (1, 2, 2, 4 taps) read the SAME input pixels: an input tile is staged ONCE (rows = pixels
//      in [y][x] order, pitch TV + 1 = 16, one halo row / column) and every tap is a row-shifted view of it (shift =
//      16 dy + dx rows of the SWIZZLE_128B K-major operand); nine tap-GEMMs per tile into four TMEM accumulators (one
//      per class), which the epilogue scatters to the four output pixels of every input pixel.
//   S: out[n][oy][ox][co] = bias[co] + sum over ky, kx, ci: A[n][2 oy + ky - 1][2 ox + kx - 1][ci] * W[ky][kx][ci][co]
//      Space-to-depth view: the four pixels (2y' + py, 2x' + px) of a 2 x 2 input cell are ONE operand row of 4 Cin
//      channels (the loader gathers them; no copy in HBM), so the strided convolution becomes four row-shifted views
//      (shift = 16 sy + sx, cell (oy - 1 + sy, ox - 1 + sx)) of a stride-1 tile, each multiplying the planes (py, px)
//      that are taps for it -- only the non-zero (shift, 32-channel chunk) products are issued: the same 9 Cin K-work.
//      Narrow inputs (Cin = 4, 16) thus still fill 64-byte / 128-byte operand rows.
// Epilogue: 256-bit stores of whole NHWC rows; form T accumulates the BatchNorm sums of the output.  It replaces
// conv_mma_kernel (mma.sync 3xTF32, every 16-output-channel block re-reading the input) for every sampler layer.
//
// Warp roles (16 warps, one persistent CTA per SM, as nb1d_pair_h3.cu): 0..7 epilogue, 8..13 loaders (chunk-major, register
// double buffer), 14 MMA issuer, 15 weight producer (16-bit hi/lo chunk images streamed from L2 through a
// cp.async.bulk ring).  Accumulators are double-buffered in TMEM, input tiles too where they fit (<= 128 operand
// channels): the loaders fill tile t+1 and the epilogue drains tile t-1 while the tensor pipe works on tile t.
#include <atomic>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mdil {
namespace ctc {

constexpr int TU = 8, TV = 15, PV = 16;       // tile: TU x TV input pixels, rows of pitch PV (one halo column)
constexpr int INROWS = TU * PV + PV + 1;       // 145: the largest row a tap view touches is 127 + 17
constexpr int IN_MAX = 160;
constexpr int N_EPI = 256, N_LOAD = 192, W_LOAD0 = 8, W_MMA = 14, W_PROD = 15, NTHREADS = 512;
constexpr int KCH = 32;                        // operand channels per weight chunk
constexpr int NSTAGE = 4;                      // weight ring stages
constexpr int MAXE = 36;                       // stream entries (32-channel chunk, tap view, class) per tile

template <int CINP, int NCLS, int NCO> struct Cfg {
  static constexpr int SLABS = (CINP + 63) / 64;                // 128-byte operand rows (64 channels) per pixel and image
  static constexpr uint32_t SLAB_BYTES = IN_MAX * 128;
  static constexpr uint32_t IMG_BYTES = SLABS * SLAB_BYTES;
  static constexpr uint32_t BUF_BYTES = 2 * IMG_BYTES;
  static constexpr int NBUF = CINP <= 128 ? 2 : 1;              // operand buffers
  static constexpr int NKC = CINP / KCH;
  static constexpr uint32_t HALF_STAGE = NCO * 64;
  static constexpr uint32_t STAGE_BYTES = 2 * HALF_STAGE;
  static constexpr uint32_t HDR_BYTES = 3072;
  static constexpr uint32_t SMEM_BYTES = HDR_BYTES + NBUF * BUF_BYTES + NSTAGE * STAGE_BYTES;
  static constexpr uint32_t ACCW = NCLS * NCO;                  // accumulator columns of one tile
  static constexpr uint32_t TMEM_COLS = 2 * ACCW <= 32 ? 32 : 2 * ACCW <= 64 ? 64 : 2 * ACCW <= 128 ? 128 : 2 * ACCW <= 256 ? 256 : 512;
  static constexpr int NPT = NCO / 16;                          // 16-column pieces per class
};

// header offsets
constexpr uint32_t OFF_WFULL = 0;        // [NSTAGE]
constexpr uint32_t OFF_WEMPTY = 32;      // [NSTAGE]
constexpr uint32_t OFF_INFULL = 64;      // [NBUF][NKC]: <= 16
constexpr uint32_t OFF_BUFFREE = 192;    // [NBUF]
constexpr uint32_t OFF_ACCFULL = 208;    // [2]
constexpr uint32_t OFF_ACCFREE = 224;    // [2]
constexpr uint32_t OFF_TMEMSLOT = 240;
constexpr uint32_t OFF_BIAS = 256;       // float [NCO <= 128]
constexpr uint32_t OFF_PIXTAB = 768;     // int [2][IN_MAX]

struct Args {
  const float* A;          // source tensor [N, IH, IW, lda]
  const void* wimg;        // per stream entry a stage (hi image, lo image), [NCO rows][32 k] 16-bit, SWIZZLE_64B
  const float* bias;       // nullable [NCO]
  float* out;              // T: [N, 2H, 2W, ldg]; S: [N, H, W, ldg] (channels g_coff .. g_coff + NCO - 1)
  double* sums;            // nullable (form T only) [2][NCO]: sum and sum of squares of the written values
  int mode;                // 0: T, 1: S
  int N, H, W;             // tile grid: T input pixels, S output pixels
  int IH, IW;              // source tensor pixels (T: H, W; S: 2H, 2W)
  int lda, a_coff, cin;    // source row stride, channel offset, channels (T: loaded channels; S: channels per plane)
  int ldg, g_coff;
  int nent;                // stream entries per tile, chunk-major
  unsigned short ent[MAXE];   // chunk j | row shift << 4 | class << 9 | first-of-class << 11
  int tiles_u, tiles_v, total_tiles;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d), "r"(a_lo),
      "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}
__device__ __forceinline__ void stg8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]));
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int FMT>
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  if (FMT == 0) {
    float h0, h1;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
  } else {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
  }
}
// Sum v[0..16) over the 32 lanes of the warp: afterwards the lane holds the total of channel (lane >> 1) & 15
__device__ __forceinline__ float transpose_reduce16(const float (&v)[16], int lane) {
  float w8[8], w4[4], w2[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = b4 ? v[i] : v[i + 8], keep = b4 ? v[i + 8] : v[i];
    w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = b3 ? w8[i] : w8[i + 4], keep = b3 ? w8[i + 4] : w8[i];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = b2 ? w4[i] : w4[i + 2], keep = b2 ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = b1 ? w2[0] : w2[1], keep = b1 ? w2[1] : w2[0];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

template <int CINP, int NCLS, int NCO, int FMT>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ Args a) {
  using K = Cfg<CINP, NCLS, NCO>;
  constexpr int NKC = K::NKC, NBUF = K::NBUF;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t hdr = smem_u32(smem_raw);
  if ((hdr & 1023u) != 0) __trap();
  unsigned char* gen = smem_raw;
  const uint32_t act0 = hdr + K::HDR_BYTES;
  const uint32_t ring = act0 + NBUF * K::BUF_BYTES;
  float* bias_s = reinterpret_cast<float*>(gen + OFF_BIAS);
  int* pixtab = reinterpret_cast<int*>(gen + OFF_PIXTAB);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // this CTA's tiles: blockIdx.x, + gridDim.x, ...
  const int ntiles = ((int)blockIdx.x < a.total_tiles) ? (a.total_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int S = a.mode;      // strided form: the tile origin is one cell up / left of the first output pixel

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(hdr + OFF_WFULL + 8 * i, 1); mbar_init(hdr + OFF_WEMPTY + 8 * i, 1); }
    for (int i = 0; i < NBUF * NKC; ++i) mbar_init(hdr + OFF_INFULL + 8 * i, N_LOAD);
    for (int i = 0; i < NBUF; ++i) mbar_init(hdr + OFF_BUFFREE + 8 * i, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(hdr + OFF_ACCFULL + 8 * i, 1); mbar_init(hdr + OFF_ACCFREE + 8 * i, N_EPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < NCO) bias_s[tid] = a.bias != nullptr ? __ldg(a.bias + tid) : 0.f;
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hdr + OFF_TMEMSLOT), "r"(K::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + OFF_TMEMSLOT);

  if (warp == W_PROD) {
    // ============================================================ weight producer: nent stages per tile through the ring
    const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(a.wimg);
    const uint32_t nent = (uint32_t)a.nent, total = (uint32_t)ntiles * nent;
    uint32_t e = 0;
    for (uint32_t k = 0; k < total; ++k) {
      const uint32_t st = k % NSTAGE;
      if (k >= (uint32_t)NSTAGE) mbar_wait(hdr + OFF_WEMPTY + 8 * st, ((k / NSTAGE) - 1) & 1);
      if (lane == 0) {
        mbar_expect_tx(hdr + OFF_WFULL + 8 * st, K::STAGE_BYTES);
        bulk_g2s(ring + st * K::STAGE_BYTES, wsrc + (size_t)e * K::STAGE_BYTES, K::STAGE_BYTES, hdr + OFF_WFULL + 8 * st);
      }
      __syncwarp();
      if (++e == nent) e = 0;
    }
  } else if (warp == W_MMA) {
    // ============================================================ MMA issuer (warp-uniform control flow)
    constexpr uint32_t FB = FMT == 0 ? 0u : 1u;
    const uint32_t idesc = (1u << 4) | (FB << 7) | (FB << 10) | ((uint32_t)(NCO >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hiw = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024, version 1, SWIZZLE_128B
    const uint32_t b_hiw = (512u >> 4) | (1u << 14) | (4u << 29);       // SBO 512, version 1, SWIZZLE_64B
    const uint32_t ring0 = ((ring & 0x3FFFF) >> 4) | (1u << 16);
    uint32_t kring = 0;
    for (int t = 0; t < ntiles; ++t) {
      const int b = t % NBUF, use = t / NBUF;
      const int ab = t & 1, ause = t >> 1;
      const uint32_t act_hi = act0 + (uint32_t)b * K::BUF_BYTES;
      const uint32_t ahi0 = ((act_hi & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t alo0 = (((act_hi + K::IMG_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t acc0 = tmem + (uint32_t)ab * K::ACCW;
      if (t >= 2) {      // the epilogue of tile t-2 has drained this accumulator set
        mbar_wait(hdr + OFF_ACCFREE + 8 * ab, (uint32_t)((ause - 1) & 1));
        tc_fence_after();
      }
      int jcur = -1;
#pragma unroll 1
      for (int e = 0; e < a.nent; ++e) {
        const uint32_t en = a.ent[e];
        const int j = (int)(en & 15u);
        if (j != jcur) {   // the loaders fill the tile one 32-channel chunk at a time
          mbar_wait(hdr + OFF_INFULL + 8 * (b * NKC + j), (uint32_t)(use & 1));
          tc_fence_after();
          jcur = j;
        }
        const uint32_t st = kring % NSTAGE;
        mbar_wait(hdr + OFF_WFULL + 8 * st, (kring / NSTAGE) & 1);
        tc_fence_after();
        const uint32_t ad = (uint32_t)(j >> 1) * (K::SLAB_BYTES >> 4) + ((en >> 4) & 31u) * 8u + (uint32_t)(j & 1) * 4u;
        const uint32_t ah = ahi0 + ad, al = alo0 + ad;
        const uint32_t bh = ring0 + st * (K::STAGE_BYTES >> 4), bl = bh + (K::HALF_STAGE >> 4);
        const uint32_t acc = acc0 + ((en >> 9) & 3u) * NCO;
        const uint32_t accumulate = ((en >> 11) & 1u) ? 0u : 1u;
        if (elect_one()) {
          mma_f16(acc, ah, a_hiw, bh, b_hiw, idesc, accumulate);
          mma_f16(acc, al, a_hiw, bh, b_hiw, idesc, 1u);
          mma_f16(acc, ah, a_hiw, bl, b_hiw, idesc, 1u);
          mma_f16(acc, ah + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
          mma_f16(acc, al + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
          mma_f16(acc, ah + 2, a_hiw, bl + 2, b_hiw, idesc, 1u);
          umma_commit(hdr + OFF_WEMPTY + 8 * st);
        }
        __syncwarp();
        ++kring;
      }
      if (elect_one()) {
        umma_commit(hdr + OFF_ACCFULL + 8 * ab);
        umma_commit(hdr + OFF_BUFFREE + 8 * b);
      }
      __syncwarp();
    }
  } else if (warp >= W_LOAD0) {
    // ============================================================ loader warps: fp32 NHWC rows -> 16-bit hi/lo K-major operand rows
    const int lt = tid - W_LOAD0 * 32;
    constexpr int RPP = N_LOAD / 8;                    // rows per pass: 24 (8 threads x 128 bit = one 32-channel chunk of a row)
    constexpr int PB = (IN_MAX + RPP - 1) / RPP;       // passes per chunk: 7
    const int c4 = lt & 7, rsub = lt >> 3;
    const uint32_t half8 = (uint32_t)(c4 & 1) * 8u;
    // pixel table of tile t (double-buffered by t & 1): source pixel index of operand row lt (S: of plane (0, 0) of the
    // row's 2 x 2 cell), -1 = zero padding
    auto new_tile = [&](int t) {
      if (lt < IN_MAX) {
        const int tile = (int)blockIdx.x + t * (int)gridDim.x;
        const int tv = tile % a.tiles_v, r2 = tile / a.tiles_v, tu = r2 % a.tiles_u, n = r2 / a.tiles_u;
        const int y = tu * TU + (lt >> 4) - S, x = tv * TV + (lt & 15) - S;
        int pix = -1;
        if (lt < INROWS && y >= 0 && x >= 0 && y < a.H && x < a.W) pix = S ? (n * a.IH + 2 * y) * a.IW + 2 * x : (n * a.IH + y) * a.IW + x;
        pixtab[(t & 1) * IN_MAX + lt] = pix;
      }
      named_bar_sync(2, N_LOAD);
    };
    auto issue = [&](float4 (&x)[PB], int g) {
      const int t = g / NKC, bt = g - t * NKC;
      if (bt == 0) new_tile(t);
      const int* ptab = pixtab + (t & 1) * IN_MAX;
      // this thread's four operand channels of chunk bt: T: source channels chv..; S: plane chv / cin, its channel chv % cin
      const int chv = bt * KCH + c4 * 4;
      int srcch = chv, poff = 0;
      bool on = chv < a.cin;
      if (S) {
        const int plane = chv / a.cin;
        srcch = chv - plane * a.cin;
        on = plane < 4;
        poff = (plane >> 1) * a.IW + (plane & 1);
      }
      const float* src = a.A + (size_t)poff * a.lda + a.a_coff + srcch;
#pragma unroll
      for (int p = 0; p < PB; ++p) {
        const int row = p * RPP + rsub;
        x[p] = make4(0.f);
        if (row < IN_MAX && on) {
          const int pix = ptab[row];
          if (pix >= 0) x[p] = ldg4(src + (size_t)pix * a.lda);
        }
      }
    };
    auto convert = [&](const float4 (&x)[PB], int g) {
      const int t = g / NKC, bt = g - t * NKC;
      const int b = t % NBUF, use = t / NBUF;
      if (bt == 0 && t >= NBUF) mbar_wait(hdr + OFF_BUFFREE + 8 * b, (uint32_t)((use - 1) & 1));
      unsigned char* buf = gen + K::HDR_BYTES + (size_t)b * K::BUF_BYTES + (size_t)(bt >> 1) * K::SLAB_BYTES;
      const uint32_t chunk16 = (uint32_t)((bt & 1) * 4 + (c4 >> 1));
#pragma unroll
      for (int p = 0; p < PB; ++p) {
        const int row = p * RPP + rsub;
        uint2 hi, lo;
        split2<FMT>(x[p].x, x[p].y, hi.x, lo.x);
        split2<FMT>(x[p].z, x[p].w, hi.y, lo.y);
        const uint32_t off = (uint32_t)row * 128u + (((chunk16 ^ ((uint32_t)row & 7u)) << 4) | half8);
        if (row < INROWS) {
          *reinterpret_cast<uint2*>(buf + off) = hi;
          *reinterpret_cast<uint2*>(buf + off + K::IMG_BYTES) = lo;
        }
      }
      fence_proxy_async();
      mbar_arrive(hdr + OFF_INFULL + 8 * (b * NKC + bt));
    };
    const int nbatch = ntiles * NKC;
    float4 xa[PB], xb[PB];
    if (nbatch > 0) issue(xa, 0);
    for (int g = 0; g < nbatch; g += 2) {
      if (g + 1 < nbatch) issue(xb, g + 1);
      convert(xa, g);
      if (g + 1 >= nbatch) break;
      if (g + 2 < nbatch) issue(xa, g + 2);
      convert(xb, g + 1);
    }
  } else {
    // ============================================================ epilogue warps
    // TMEM lane quadrant q = warp & 3 (accumulator row m = 32 q + lane); the two warps of a quadrant split the tile's
    // 16-column pieces: T: classes {0, 1} / {2, 3}; S: the lower / upper half of the pieces
    const int q = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int m = q * 32 + lane;
    const int mu = m >> 4, mv = m & 15;
    constexpr int NPT = K::NPT;
    constexpr int NPIECES = NCLS * NPT, HP = (NPIECES + 1) / 2;      // pieces of one tile, pieces per warp
    float run1[NPT], run2[NPT];
#pragma unroll
    for (int i = 0; i < NPT; ++i) { run1[i] = 0.f; run2[i] = 0.f; }
    const int GW = S ? a.W : 2 * a.W, GH = S ? a.H : 2 * a.H;
    for (int t = 0; t < ntiles; ++t) {
      const int ab = t & 1, ause = t >> 1;
      const int tile = (int)blockIdx.x + t * (int)gridDim.x;
      const int tv = tile % a.tiles_v, r2 = tile / a.tiles_v, tu = r2 % a.tiles_u, n = r2 / a.tiles_u;
      const int y = tu * TU + mu, x = tv * TV + mv;
      const bool valid = mv < TV && y < a.H && x < a.W;
      mbar_wait(hdr + OFF_ACCFULL + 8 * ab, (uint32_t)(ause & 1));
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < HP; ++k) {
        const int pi = half * HP + k;            // piece of the tile: class pi / NPT, columns 16 (pi % NPT) ..
        if (pi < NPIECES) {
          const int cls = pi / NPT, pc = pi % NPT;
          const int oy = S ? y : 2 * y + (cls >> 1), ox = S ? x : 2 * x + (cls & 1);
          float* dst = a.out + ((size_t)(n * GH + oy) * GW + ox) * a.ldg + a.g_coff + pc * 16;
          uint32_t r[16];
          tmem_ld16(tmem + (uint32_t)ab * K::ACCW + lane_addr + (uint32_t)(pi * 16), r);
          tmem_wait16(r);
          float v[16], v2[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(bias_s + pc * 16 + i);
            v[i] = __uint_as_float(r[i]) + bb.x; v[i + 1] = __uint_as_float(r[i + 1]) + bb.y;
            v[i + 2] = __uint_as_float(r[i + 2]) + bb.z; v[i + 3] = __uint_as_float(r[i + 3]) + bb.w;
          }
          if (valid) {
            stg8(dst, &v[0]);
            stg8(dst + 8, &v[8]);
          }
          if (NCLS == 4 && a.sums != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[i] = valid ? v[i] : 0.f; v2[i] = v[i] * v[i]; }
            const float s1 = transpose_reduce16(v, lane), s2 = transpose_reduce16(v2, lane);
#pragma unroll
            for (int i = 0; i < NPT; ++i)
              if (i == pc) { run1[i] += s1; run2[i] += s2; }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(hdr + OFF_ACCFREE + 8 * ab);
    }
    // ---- BatchNorm partial sums: warp partials -> shared slots -> fixed-order sum over the eight warps -> fp64 atomics
    if (NCLS == 4 && a.sums != nullptr) {
      // the operand buffers are dead: the last MMA has retired (accfull of the last tile was waited for by every epilogue
      // warp) and the loaders finished before it could start
      named_bar_sync(1, N_EPI);
      float* slots = reinterpret_cast<float*>(gen + K::HDR_BYTES);          // [8 warps][2][NCO]
      if ((lane & 1) == 0) {
#pragma unroll
        for (int pc = 0; pc < NPT; ++pc) {
          const int ch = pc * 16 + ((lane >> 1) & 15);
          slots[(warp * 2 + 0) * NCO + ch] = run1[pc];
          slots[(warp * 2 + 1) * NCO + ch] = run2[pc];
        }
      }
      named_bar_sync(1, N_EPI);
      for (int i = tid; i < 2 * NCO; i += N_EPI) {
        const int which = i / NCO, ch = i % NCO;
        float tsum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tsum += slots[(w * 2 + which) * NCO + ch];
        atomicAdd(a.sums + i, (double)tsum);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(K::TMEM_COLS) : "memory");
  }
}

// ---- stream plan: which (32-channel chunk, tap view, class) products a tile needs, and the weights of each
struct Plan {
  int mode, cinp, nco, nent;
  int cin;                     // T: source channels; S: channels per plane
  unsigned short ent[MAXE];
  signed char ew[MAXE][8];     // weight slab (tap) of every min(cin, 32)-channel sub-block of the entry's chunk, -1 = zero
};
struct PackArgs { int nent, cin, sb, mode; unsigned char j[MAXE]; signed char ew[MAXE][8]; };

static bool make_plan(const ConvGeom& g, Plan* pl) {
  memset(pl, 0, sizeof(*pl));
  memset(pl->ew, -1, sizeof(pl->ew));
  pl->nco = g.COUT_PAD;
  if (g.nclasses == 4) {
    // T: parity classes, unit-stride source, stride-2 destination
    if (g.a_sy != 1 || g.a_sx != 1 || g.g_sy != 2 || g.g_sx != 2 || g.GH != 2 * g.VH || g.GW != 2 * g.VW || g.AH != g.VH || g.AW != g.VW)
      return false;
    pl->mode = 0; pl->cin = g.CIN; pl->cinp = (g.CIN + 63) / 64 * 64;
    if (pl->cinp > 128) return false;
    int n = 0;
    for (int j = 0; j < pl->cinp / KCH; ++j)
      for (int c = 0; c < 4; ++c) {
        const TapClass& tc = g.cls[c];
        if (tc.o_dy != (c >> 1) || tc.o_dx != (c & 1)) return false;
        for (int t = 0; t < tc.ntaps; ++t) {
          if (n >= MAXE || tc.a_dy[t] < 0 || tc.a_dy[t] > 1 || tc.a_dx[t] < 0 || tc.a_dx[t] > 1) return false;
          pl->ent[n] = (unsigned short)(j | ((PV * tc.a_dy[t] + tc.a_dx[t]) << 4) | (c << 9) | ((j == 0 && t == 0) ? 1 << 11 : 0));
          pl->ew[n][0] = (signed char)tc.widx[t];
          ++n;
        }
      }
    pl->nent = n;
    return n > 0;
  }
  if (g.nclasses != 1 || g.cls[0].ntaps != 9) return false;
  // S: one class of 3 x 3 taps, stride-2 source, unit-stride destination
  if (g.a_sy != 2 || g.a_sx != 2 || g.g_sy != 1 || g.g_sx != 1 || g.AH != 2 * g.VH || g.AW != 2 * g.VW || g.GH != g.VH || g.GW != g.VW ||
      g.cls[0].o_dy != 0 || g.cls[0].o_dx != 0)
    return false;
  if (g.CIN != 4 && g.CIN != 16 && g.CIN != 64) return false;
  pl->mode = 1; pl->cin = g.CIN; pl->cinp = g.CIN == 4 ? 32 : 4 * g.CIN;
  const TapClass& tc = g.cls[0];
  const int sb = g.CIN < KCH ? g.CIN : KCH, nsub = KCH / sb;
  int n = 0;
  for (int j = 0; j < pl->cinp / KCH; ++j)
    for (int sy = 0; sy < 2; ++sy)
      for (int sx = 0; sx < 2; ++sx) {
        bool any = false;
        signed char ew[8];
        memset(ew, -1, sizeof(ew));
        for (int sub = 0; sub < nsub; ++sub) {
          const int plane = (j * KCH + sub * sb) / g.CIN;
          if (plane >= 4) continue;
          const int py = plane >> 1, px = plane & 1;
          for (int t = 0; t < 9; ++t) {
            if (tc.a_dy[t] < -1 || tc.a_dy[t] > 1 || tc.a_dx[t] < -1 || tc.a_dx[t] > 1) return false;
            // source row 2 oy + dy = 2 (oy - 1 + sy) + py
            const bool my = tc.a_dy[t] == 2 * (sy - 1) + py, mx = tc.a_dx[t] == 2 * (sx - 1) + px;
            if (my && mx) { ew[sub] = (signed char)tc.widx[t]; any = true; }
          }
        }
        if (!any) continue;
        if (n >= MAXE) return false;
        pl->ent[n] = (unsigned short)(j | ((PV * sy + sx) << 4) | (n == 0 ? 1 << 11 : 0));
        memcpy(pl->ew[n], ew, sizeof(ew));
        ++n;
      }
  pl->nent = n;
  return n > 0;
}

// fp32 tap slabs Wp[widx][rows][ld] (what conv_mma_kernel reads: the layouts / flips of the forward and data-gradient uses
// are those of mdil_up_pack / mdil_down_pack) -> the 16-bit hi/lo chunk images of the tensor-core kernel, one per entry
__global__ void pack_conv_tc_kernel(const float* __restrict__ Wp, int rows, int ld, int nco, const PackArgs pa,
                                    unsigned short* __restrict__ img, int fmt) {
  const long total = (long)pa.nent * nco * KCH;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % KCH);
    const int nrow = (int)((i / KCH) % nco);
    const int e = (int)(i / ((long)KCH * nco));
    const int chv = pa.j[e] * KCH + kk;              // operand channel
    const int w = pa.ew[e][kk / pa.sb];
    const int ci = pa.mode ? chv % pa.cin : chv;
    float v = 0.f;
    if (w >= 0 && ci < rows && nrow < ld) v = __ldg(Wp + ((size_t)w * rows + ci) * ld + nrow);
    unsigned short hi, lo;
    if (fmt == 0) {
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      hi = __half_as_ushort(h); lo = __half_as_ushort(l);
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      hi = __bfloat16_as_ushort(h); lo = __bfloat16_as_ushort(l);
    }
    const int off = nrow * 32 + ((((kk >> 3) ^ ((nrow >> 1) & 3)) << 3) | (kk & 7));
    unsigned short* stage = img + (long)e * 2 * nco * KCH;
    stage[off] = hi;
    stage[nco * KCH + off] = lo;
  }
}

template <int CINP, int NCLS, int NCO, int FMT>
static int launch_t(const ConvGeom& g, const Plan& pl, const float* A, const void* wimg, const float* bias, float* out,
                    double* sums, cudaStream_t s) {
  using K = Cfg<CINP, NCLS, NCO>;
  static_assert(K::SMEM_BYTES <= 227 * 1024, "conv_tc shared memory budget");
  static_assert(K::NKC * K::NBUF <= 16 && NCO <= 128 && 8 * 2 * NCO * 4 <= (int)K::BUF_BYTES && K::ACCW * 2 <= 512, "conv_tc layout");
  static_assert(N_LOAD >= IN_MAX && INROWS <= IN_MAX && TU * PV == 128 && OFF_PIXTAB + 8 * IN_MAX <= K::HDR_BYTES &&
                    NTHREADS == (W_PROD + 1) * 32 && NCO % 16 == 0,
                "conv_tc role mapping / tile geometry");
  Args a;
  memset(&a, 0, sizeof(a));
  a.A = A; a.wimg = wimg; a.bias = bias; a.out = out; a.sums = sums;
  a.mode = pl.mode;
  a.N = g.N; a.H = g.VH; a.W = g.VW; a.IH = g.AH; a.IW = g.AW;
  a.lda = g.lda; a.a_coff = g.a_coff; a.cin = pl.cin; a.ldg = g.ldg; a.g_coff = g.g_coff;
  a.nent = pl.nent;
  memcpy(a.ent, pl.ent, sizeof(a.ent));
  a.tiles_u = cdiv(g.VH, TU); a.tiles_v = cdiv(g.VW, TV);
  const long total = (long)g.N * a.tiles_u * a.tiles_v;
  MDIL_REQUIRE(total > 0 && total < (1L << 30) && (size_t)g.N * g.AH * g.AW < (1ull << 31), "conv_tc: tile count");
  a.total_tiles = (int)total;
  static std::atomic<bool> attr_set[kMaxDevices];
  std::atomic<bool>& done = attr_set[current_device_slot()];
  if (!done.load(std::memory_order_acquire)) {
    MDIL_CUDA(cudaFuncSetAttribute(conv_tc_kernel<CINP, NCLS, NCO, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES));
    done.store(true, std::memory_order_release);
  }
  const int grid = total < kNumSMs ? (int)total : kNumSMs;
  conv_tc_kernel<CINP, NCLS, NCO, FMT><<<grid, NTHREADS, K::SMEM_BYTES, s>>>(a);
  MDIL_LAUNCH_CHECK();
  return 0;
}

}  // namespace ctc

// the instantiated shapes: the five sampler layers of the network, forward (fp16 halves) and data gradient (bf16 halves)
static int conv_tc_variant(const ctc::Plan& pl, int grad) {
  const int m = pl.mode, c = pl.cinp, n = pl.nco;
  if (!grad) {
    if (m == 0 && c == 128 && n == 64) return 1;     // upsampler 128 -> 64
    if (m == 0 && c == 64 && n == 16) return 2;      // upsampler 64 -> 16
    if (m == 1 && c == 32 && n == 16) return 3;      // downsampler 3 -> 16 (13 conv channels)
    if (m == 1 && c == 64 && n == 48) return 4;      // downsampler 16 -> 64 (48)
    if (m == 1 && c == 256 && n == 64) return 5;     // downsampler 64 -> 128 (64)
  } else {
    if (m == 0 && c == 64 && n == 64) return 6;      // downsampler 64 -> 128: du (64 of 128) -> dx 64
    if (m == 0 && c == 64 && n == 16) return 7;      // downsampler 16 -> 64: du (48 of 64) -> dx 16
    if (m == 1 && c == 64 && n == 64) return 8;      // upsampler 64 -> 16: du 16 -> dx 64
    if (m == 1 && c == 256 && n == 128) return 9;    // upsampler 128 -> 64: du 64 -> dx 128
  }
  return 0;
}

bool conv_tc_ok(const ConvGeom& g, int grad) {
  static const bool on = [] { const char* e = getenv("MDIL_CONV_TC"); return !(e != nullptr && strcmp(e, "0") == 0); }();
  if (!on || pair_impl_mode() != 4) return false;
  ctc::Plan pl;
  if (!ctc::make_plan(g, &pl)) return false;
  if (g.lda % 4 != 0 || g.a_coff % 4 != 0 || g.ldg % 8 != 0 || g.g_coff % 8 != 0 || g.g_coff + g.COUT_PAD > g.ldg) return false;
  if (pl.mode == 0 && g.a_coff + pl.cin > g.lda) return false;
  return conv_tc_variant(pl, grad) != 0;
}

// 16-bit hi/lo images: at most 9 taps x operand channels x COUT_PAD weights, two halves of 2 bytes each
size_t conv_tc_image_floats(int cin, int cout_pad) {
  const int cinp = cin <= 4 ? 32 : (cin <= 16 ? 64 : (cin + 63) / 64 * 64);
  return (size_t)9 * cinp * cout_pad + 64;
}

int launch_pack_conv_tc(const ConvGeom& g, const float* Wp, int wp_rows, void* img, int grad, cudaStream_t s) {
  ctc::Plan pl;
  MDIL_REQUIRE(ctc::make_plan(g, &pl), "conv_tc: unsupported geometry");
  ctc::PackArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.nent = pl.nent; pa.cin = pl.cin; pa.mode = pl.mode; pa.sb = pl.cin < ctc::KCH ? pl.cin : ctc::KCH;
  for (int e = 0; e < pl.nent; ++e) pa.j[e] = (unsigned char)(pl.ent[e] & 15);
  memcpy(pa.ew, pl.ew, sizeof(pa.ew));
  MDIL_REQUIRE((size_t)pl.nent * pl.nco * ctc::KCH <= conv_tc_image_floats(g.CIN, g.COUT_PAD), "conv_tc: image size");
  const long total = (long)pl.nent * pl.nco * ctc::KCH;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  ctc::pack_conv_tc_kernel<<<grid, 256, 0, s>>>(Wp, wp_rows, g.COUT_PAD, pl.nco, pa, reinterpret_cast<unsigned short*>(img), grad ? 1 : 0);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_conv_tc(const ConvGeom& g, const float* A, const void* wimg, const float* bias, float* out, double* sums, int grad,
                   cudaStream_t s) {
  MDIL_REQUIRE(conv_tc_ok(g, grad), "conv_tc: unsupported geometry");
  MDIL_REQUIRE(wimg != nullptr && ((uintptr_t)wimg & 15) == 0, "conv_tc: weight images");
  ctc::Plan pl;
  ctc::make_plan(g, &pl);
  MDIL_REQUIRE(sums == nullptr || pl.mode == 0, "conv_tc: BatchNorm sums only in the transposed form");
  switch (conv_tc_variant(pl, grad)) {
    case 1: return ctc::launch_t<128, 4, 64, 0>(g, pl, A, wimg, bias, out, sums, s);
    case 2: return ctc::launch_t<64, 4, 16, 0>(g, pl, A, wimg, bias, out, sums, s);
    case 3: return ctc::launch_t<32, 1, 16, 0>(g, pl, A, wimg, bias, out, sums, s);
    case 4: return ctc::launch_t<64, 1, 48, 0>(g, pl, A, wimg, bias, out, sums, s);
    case 5: return ctc::launch_t<256, 1, 64, 0>(g, pl, A, wimg, bias, out, sums, s);
    case 6: return ctc::launch_t<64, 4, 64, 1>(g, pl, A, wimg, bias, out, sums, s);
    case 7: return ctc::launch_t<64, 4, 16, 1>(g, pl, A, wimg, bias, out, sums, s);
    case 8: return ctc::launch_t<64, 1, 64, 1>(g, pl, A, wimg, bias, out, sums, s);
    case 9: return ctc::launch_t<256, 1, 128, 1>(g, pl, A, wimg, bias, out, sums, s);
    default: return set_error(-2, "conv_tc: no kernel for this shape", __FILE__, __LINE__);
  }
}

}  // namespace mdil
