// Sub-pixel (parity-class) 3x3 stride-2 TRANSPOSED convolution on the 5th-generation tensor cores: the forward pass of
// UpsamplerBlock (models/erfnet_RA_parallel.py:160-172: ConvTranspose2d(Cin, Cout, 3, stride 2, padding 1, output_padding
// 1)) and, with the roles of the tensors exchanged, the data gradient of DownsamplerBlock's strided 3x3 convolution
// (models/erfnet_RA_parallel.py:24-45).  Same arithmetic as nb1d_pair_h3.cu: fp32 operands split into two 16-bit halves
// (x = hi + lo; hi*hi + lo*hi + hi*lo accumulated in fp32 TMEM; fp16 halves forward, bf16 halves for gradients),
// tcgen05.mma kind::f16, M = 128 pixels, N = Cout, K = 16.
//
//     out[n][2y + py][2x + px][co] = bias[co] + sum over the taps (dy, dx) of class (py, px), ci:
//                                    A[n][y + dy][x + dx][ci] * W[widx][ci][co]          dy, dx in {0, 1}
//
// The four parity classes have 1, 2, 2 and 4 taps and read the SAME input pixels, so an input tile is staged ONCE
// (rows = pixels in [y][x] order with a pitch of TV + 1 = 16, one halo row / column) and every tap is a row-shifted view
// of it (shift = 16 dy + dx rows of the SWIZZLE_128B K-major operand): nine tap-GEMMs per tile into four TMEM
// accumulators (one per class), which the epilogue warps scatter to the four output pixels of every input pixel
// (256-bit stores of whole NHWC rows) while accumulating the BatchNorm sums of the output.  It replaces conv_mma_kernel
// (mma.sync 3xTF32, every 16-output-channel block re-reading the input) for the layers with CIN >= 64.
//
// Warp roles (16 warps, one persistent CTA per SM, as nb1d_pair_h3.cu): 0..7 epilogue (TMEM lane quadrant = warp & 3, classes
// 2 * (warp >> 2) and + 1), 8..13 loaders (chunk-major, register double buffer), 14 MMA issuer, 15 weight producer
// (16-bit hi/lo chunk images streamed from L2 through a cp.async.bulk ring).  Input tiles and accumulators are double-
// buffered: the loaders fill tile t+1 and the epilogue drains tile t-1 while the tensor pipe works on tile t.
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
constexpr int KCH = 32;                        // input channels per weight chunk
constexpr int NBUF = 2, NSTAGE = 4, NSLOT = 9; // operand buffers, weight ring stages, tap slots (1 + 2 + 2 + 4)

template <int CIN, int NCO> struct Cfg {
  static constexpr int SLABS = CIN / 64;
  static constexpr uint32_t SLAB_BYTES = IN_MAX * 128;
  static constexpr uint32_t IMG_BYTES = SLABS * SLAB_BYTES;
  static constexpr uint32_t BUF_BYTES = 2 * IMG_BYTES;
  static constexpr int NKC = CIN / KCH;
  static constexpr uint32_t HALF_STAGE = NCO * 64;
  static constexpr uint32_t STAGE_BYTES = 2 * HALF_STAGE;
  static constexpr uint32_t HDR_BYTES = 3072;
  static constexpr uint32_t SMEM_BYTES = HDR_BYTES + NBUF * BUF_BYTES + NSTAGE * STAGE_BYTES;
  static constexpr uint32_t TMEM_COLS = 8 * NCO < 32 ? 32 : 8 * NCO;      // [2 tiles][4 classes][NCO]
  static constexpr int NPIECE = NCO / 16;
};

// header offsets
constexpr uint32_t OFF_WFULL = 0;        // [NSTAGE]
constexpr uint32_t OFF_WEMPTY = 32;      // [NSTAGE]
constexpr uint32_t OFF_INFULL = 64;      // [NBUF][NKC <= 4]
constexpr uint32_t OFF_BUFFREE = 128;    // [NBUF]
constexpr uint32_t OFF_ACCFULL = 144;    // [2]
constexpr uint32_t OFF_ACCFREE = 160;    // [2]
constexpr uint32_t OFF_TMEMSLOT = 176;
constexpr uint32_t OFF_BIAS = 256;       // float [NCO <= 64]
constexpr uint32_t OFF_PIXTAB = 512;     // int [2][IN_MAX]

struct Args {
  const float* A;          // [N, H, W, lda] (channels a_coff .. a_coff + CIN - 1)
  const void* wimg;        // chunk images: [NKC][NSLOT] stages of (hi image, lo image), [NCO rows][32 ci] 16-bit, SWIZZLE_64B
  const float* bias;       // nullable [NCO]
  float* out;              // [N, 2H, 2W, ldg] (channels g_coff .. g_coff + NCO - 1)
  double* sums;            // nullable [2][NCO]: sum and sum of squares of the written values (BatchNorm statistics)
  int N, H, W, lda, a_coff, ldg, g_coff;
  int slot_class[NSLOT];   // class (py * 2 + px) of every tap slot, in stream order
  int slot_shift[NSLOT];   // row shift of the tap view: 16 dy + dx
  int slot_first[NSLOT];   // first tap of its class (the accumulator starts from zero at the first K chunk)
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

template <int CIN, int NCO, int FMT>
__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ Args a) {
  using K = Cfg<CIN, NCO>;
  constexpr int NKC = K::NKC;
  constexpr int G = NKC * NSLOT;                 // weight stages per tile
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
    // ============================================================ weight producer: G stages per tile through the ring
    const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(a.wimg);
    const uint32_t total = (uint32_t)ntiles * G;
    for (uint32_t k = 0; k < total; ++k) {
      const uint32_t st = k % NSTAGE;
      if (k >= (uint32_t)NSTAGE) mbar_wait(hdr + OFF_WEMPTY + 8 * st, ((k / NSTAGE) - 1) & 1);
      if (lane == 0) {
        mbar_expect_tx(hdr + OFF_WFULL + 8 * st, K::STAGE_BYTES);
        bulk_g2s(ring + st * K::STAGE_BYTES, wsrc + (size_t)(k % G) * K::STAGE_BYTES, K::STAGE_BYTES, hdr + OFF_WFULL + 8 * st);
      }
      __syncwarp();
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
      const uint32_t acc0 = tmem + (uint32_t)ab * 4 * NCO;
      if (t >= 2) {      // the epilogue of tile t-2 has drained this accumulator set
        mbar_wait(hdr + OFF_ACCFREE + 8 * ab, (uint32_t)((ause - 1) & 1));
        tc_fence_after();
      }
#pragma unroll 1
      for (int j = 0; j < NKC; ++j) {
        mbar_wait(hdr + OFF_INFULL + 8 * (b * NKC + j), (uint32_t)(use & 1));
        tc_fence_after();
#pragma unroll 1
        for (int s = 0; s < NSLOT; ++s) {
          const uint32_t st = kring % NSTAGE;
          mbar_wait(hdr + OFF_WFULL + 8 * st, (kring / NSTAGE) & 1);
          tc_fence_after();
          const uint32_t ad = (uint32_t)(j >> 1) * (K::SLAB_BYTES >> 4) + (uint32_t)a.slot_shift[s] * 8u + (uint32_t)(j & 1) * 4u;
          const uint32_t ah = ahi0 + ad, al = alo0 + ad;
          const uint32_t bh = ring0 + st * (K::STAGE_BYTES >> 4), bl = bh + (K::HALF_STAGE >> 4);
          const uint32_t acc = acc0 + (uint32_t)a.slot_class[s] * NCO;
          const uint32_t accumulate = (j != 0 || a.slot_first[s] == 0) ? 1u : 0u;
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
    const float* src = a.A + a.a_coff + c4 * 4;
    // pixel table of tile t (double-buffered by t & 1): global pixel index of input row lt, -1 = zero padding
    auto new_tile = [&](int t) {
      if (lt < IN_MAX) {
        const int tile = (int)blockIdx.x + t * (int)gridDim.x;
        const int tv = tile % a.tiles_v, r2 = tile / a.tiles_v, tu = r2 % a.tiles_u, n = r2 / a.tiles_u;
        const int y = tu * TU + (lt >> 4), x = tv * TV + (lt & 15);
        int pix = -1;
        if (lt < INROWS && y < a.H && x < a.W) pix = (n * a.H + y) * a.W + x;
        pixtab[(t & 1) * IN_MAX + lt] = pix;
      }
      named_bar_sync(2, N_LOAD);
    };
    auto issue = [&](float4 (&x)[PB], int g) {
      const int t = g / NKC, bt = g - t * NKC;
      if (bt == 0) new_tile(t);
      const int* ptab = pixtab + (t & 1) * IN_MAX;
#pragma unroll
      for (int p = 0; p < PB; ++p) {
        const int row = p * RPP + rsub;
        x[p] = make4(0.f);
        if (row < IN_MAX) {
          const int pix = ptab[row];
          if (pix >= 0) x[p] = ldg4(src + (size_t)pix * a.lda + bt * KCH);
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
    const int q = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int m = q * 32 + lane;
    const int mu = m >> 4, mv = m & 15;
    constexpr int NPIECE = K::NPIECE;
    float run1[NPIECE], run2[NPIECE];
#pragma unroll
    for (int i = 0; i < NPIECE; ++i) { run1[i] = 0.f; run2[i] = 0.f; }
    const int GW = 2 * a.W;
    for (int t = 0; t < ntiles; ++t) {
      const int ab = t & 1, ause = t >> 1;
      const int tile = (int)blockIdx.x + t * (int)gridDim.x;
      const int tv = tile % a.tiles_v, r2 = tile / a.tiles_v, tu = r2 % a.tiles_u, n = r2 / a.tiles_u;
      const int y = tu * TU + mu, x = tv * TV + mv;
      const bool valid = mv < TV && y < a.H && x < a.W;
      mbar_wait(hdr + OFF_ACCFULL + 8 * ab, (uint32_t)(ause & 1));
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int cls = half * 2 + cc, py = cls >> 1, px = cls & 1;
        float* dst = a.out + ((size_t)(n * 2 * a.H + 2 * y + py) * GW + 2 * x + px) * a.ldg + a.g_coff;
        const uint32_t acc = tmem + (uint32_t)(ab * 4 + cls) * NCO;
#pragma unroll
        for (int pc = 0; pc < NPIECE; ++pc) {
          uint32_t r[16];
          tmem_ld16(acc + lane_addr + (uint32_t)(pc * 16), r);
          tmem_wait16(r);
          float v[16], v2[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(bias_s + pc * 16 + i);
            v[i] = __uint_as_float(r[i]) + bb.x; v[i + 1] = __uint_as_float(r[i + 1]) + bb.y;
            v[i + 2] = __uint_as_float(r[i + 2]) + bb.z; v[i + 3] = __uint_as_float(r[i + 3]) + bb.w;
          }
          if (valid) {
            stg8(dst + pc * 16, &v[0]);
            stg8(dst + pc * 16 + 8, &v[8]);
          }
          if (a.sums != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { v[i] = valid ? v[i] : 0.f; v2[i] = v[i] * v[i]; }
            run1[pc] += transpose_reduce16(v, lane);
            run2[pc] += transpose_reduce16(v2, lane);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(hdr + OFF_ACCFREE + 8 * ab);
    }
    // ---- BatchNorm partial sums: warp partials -> shared slots -> fixed-order sum over the eight warps -> fp64 atomics
    if (a.sums != nullptr) {
      // the operand buffers are dead: the last MMA has retired (accfull of the last tile was waited for by every epilogue
      // warp) and the loaders finished before it could start
      named_bar_sync(1, N_EPI);
      float* slots = reinterpret_cast<float*>(gen + K::HDR_BYTES);          // [8 warps][2][NCO]
      if ((lane & 1) == 0) {
#pragma unroll
        for (int pc = 0; pc < NPIECE; ++pc) {
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

// fp32 tap slabs Wp[widx][ci][ld] (what conv_mma_kernel reads: the layouts / flips of the forward and data-gradient uses are
// those of mdil_up_pack / mdil_down_pack) -> the 16-bit hi/lo chunk images of the tensor-core kernel, stage (j, slot)
struct PackArgs { int widx[NSLOT]; };
__global__ void pack_conv_tc_kernel(const float* __restrict__ Wp, int cin, int ld, int nco, const PackArgs pa,
                                    unsigned short* __restrict__ img, int fmt) {
  const int nkc = cin / KCH;
  const long total = (long)nkc * NSLOT * nco * KCH;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % KCH);
    const int nrow = (int)((i / KCH) % nco);
    const int g = (int)(i / ((long)KCH * nco));      // stage = j * NSLOT + slot
    const int j = g / NSLOT, s = g % NSLOT;
    const float v = __ldg(Wp + ((size_t)pa.widx[s] * cin + j * KCH + kk) * ld + nrow);
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
    unsigned short* stage = img + (long)g * 2 * nco * KCH;
    stage[off] = hi;
    stage[nco * KCH + off] = lo;
  }
}

// tap slots in stream order (class-major) from the parity classes of a ConvGeom
static bool make_slots(const ConvGeom& g, int* cls, int* shift, int* first, int* widx) {
  if (g.nclasses != 4) return false;
  int n = 0;
  for (int c = 0; c < 4; ++c) {
    const TapClass& tc = g.cls[c];
    if (tc.o_dy != (c >> 1) || tc.o_dx != (c & 1)) return false;
    for (int t = 0; t < tc.ntaps; ++t) {
      if (n >= NSLOT || tc.a_dy[t] < 0 || tc.a_dy[t] > 1 || tc.a_dx[t] < 0 || tc.a_dx[t] > 1) return false;
      cls[n] = c; shift[n] = PV * tc.a_dy[t] + tc.a_dx[t]; first[n] = t == 0 ? 1 : 0; widx[n] = tc.widx[t];
      ++n;
    }
  }
  return n == NSLOT;
}

template <int CIN, int NCO, int FMT>
static int launch_t(const ConvGeom& g, const float* A, const void* wimg, const float* bias, float* out, double* sums,
                    cudaStream_t s) {
  using K = Cfg<CIN, NCO>;
  static_assert(K::SMEM_BYTES <= 227 * 1024, "conv_tc shared memory budget");
  static_assert(K::NKC <= 4 && NCO <= 64 && 8 * 2 * NCO * 4 <= (int)K::BUF_BYTES, "conv_tc header layout");
  Args a;
  memset(&a, 0, sizeof(a));
  int widx[NSLOT];
  MDIL_REQUIRE(make_slots(g, a.slot_class, a.slot_shift, a.slot_first, widx), "conv_tc: not a parity-class geometry");
  a.A = A; a.wimg = wimg; a.bias = bias; a.out = out; a.sums = sums;
  a.N = g.N; a.H = g.VH; a.W = g.VW; a.lda = g.lda; a.a_coff = g.a_coff; a.ldg = g.ldg; a.g_coff = g.g_coff;
  a.tiles_u = cdiv(g.VH, TU); a.tiles_v = cdiv(g.VW, TV);
  const long total = (long)g.N * a.tiles_u * a.tiles_v;
  MDIL_REQUIRE(total > 0 && total < (1L << 30) && (size_t)g.N * g.VH * g.VW < (1ull << 31), "conv_tc: tile count");
  a.total_tiles = (int)total;
  static std::atomic<bool> attr_set[kMaxDevices];
  std::atomic<bool>& done = attr_set[current_device_slot()];
  if (!done.load(std::memory_order_acquire)) {
    MDIL_CUDA(cudaFuncSetAttribute(conv_tc_kernel<CIN, NCO, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES));
    done.store(true, std::memory_order_release);
  }
  const int grid = total < kNumSMs ? (int)total : kNumSMs;
  conv_tc_kernel<CIN, NCO, FMT><<<grid, NTHREADS, K::SMEM_BYTES, s>>>(a);
  MDIL_LAUNCH_CHECK();
  return 0;
}

}  // namespace ctc

bool conv_tc_ok(const ConvGeom& g, int grad) {
  static const bool on = [] { const char* e = getenv("MDIL_CONV_TC"); return !(e != nullptr && strcmp(e, "0") == 0); }();
  if (!on || pair_impl_mode() != 4) return false;
  int cls[ctc::NSLOT], shift[ctc::NSLOT], first[ctc::NSLOT], widx[ctc::NSLOT];
  if (!ctc::make_slots(g, cls, shift, first, widx)) return false;
  if (g.a_sy != 1 || g.a_sx != 1 || g.g_sy != 2 || g.g_sx != 2 || g.GH != 2 * g.VH || g.GW != 2 * g.VW || g.AH != g.VH || g.AW != g.VW)
    return false;
  if (g.lda % 4 != 0 || g.a_coff % 4 != 0 || g.ldg % 8 != 0 || g.g_coff % 8 != 0 || g.COUT != g.COUT_PAD) return false;
  if (grad) return g.CIN == 64 && g.COUT == 64;
  return (g.CIN == 128 && g.COUT == 64) || (g.CIN == 64 && g.COUT == 16);
}

size_t conv_tc_image_floats(int cin, int cout) { return (size_t)ctc::NSLOT * cin * cout; }   // 2 halves x 16 bit per weight

int launch_pack_conv_tc(const ConvGeom& g, const float* Wp, void* img, int grad, cudaStream_t s) {
  ctc::PackArgs pa;
  int cls[ctc::NSLOT], shift[ctc::NSLOT], first[ctc::NSLOT];
  MDIL_REQUIRE(ctc::make_slots(g, cls, shift, first, pa.widx), "conv_tc: not a parity-class geometry");
  const long total = (long)ctc::NSLOT * g.CIN * g.COUT;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  ctc::pack_conv_tc_kernel<<<grid, 256, 0, s>>>(Wp, g.CIN, g.COUT_PAD, g.COUT, pa, reinterpret_cast<unsigned short*>(img), grad ? 1 : 0);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_conv_tc(const ConvGeom& g, const float* A, const void* wimg, const float* bias, float* out, double* sums, int grad,
                   cudaStream_t s) {
  MDIL_REQUIRE(conv_tc_ok(g, grad), "conv_tc: unsupported geometry");
  MDIL_REQUIRE(wimg != nullptr && ((uintptr_t)wimg & 15) == 0, "conv_tc: weight images");
  if (grad) return ctc::launch_t<64, 64, 1>(g, A, wimg, bias, out, sums, s);
  if (g.CIN == 128) return ctc::launch_t<128, 64, 0>(g, A, wimg, bias, out, sums, s);
  return ctc::launch_t<64, 16, 0>(g, A, wimg, bias, out, sums, s);
}

}  // namespace mdil
