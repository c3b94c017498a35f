// Tensor-core (tcgen05 / TMEM) weight-gradient kernel of the nb1d block's 3-tap and 1x1 convolutions, C = 64 / 128:
//     dW[t][ci][co] = sum_pixels A'[pix + (t-1)*d along the tap axis][ci] * G[pix][co],     db[co] = sum_pixels G[pix][co]
// (B2 of SURVEY appendix: a [C x 3C] reduction over all N*h*w pixels).  A' = A or ReLU(A*scale+shift).
//
// GEMM view: M = ci, N = co, K = pixels.  The NHWC activation layout IS the MN-major operand layout of the UMMA
// (one 128-byte row per pixel and 64-channel slab of 16-bit elements, SWIZZLE_128B), so tiles are staged with plain
// coalesced 128-bit loads, split into hi/lo bf16 halves (x = hi + lo, hi*hi + lo*hi + hi*lo in fp32 TMEM: ~1e-5 relative,
// kind::f16, K = 16 pixels per instruction; gradients need the fp32 exponent range, hence bf16 and not fp16) and never
// transposed.  (Round 1 used 3xTF32 at K = 8: twice the instructions and twice the operand bytes.)  Pixels are walked in
// the d-strided lattice along the tap axis, 16 perpendicular pixels per chunk: the three taps of gradient chunk j
// are then the activation chunks j-1, j, j+1 already in the shared-memory ring.  Accumulators (3 x [C x C] fp32)
// stay in TMEM for the whole life of the persistent CTA and leave through vectorised red.global.add.v4.f32.
//
// Gathered jobs (launch_wgrad_gather_tc, kernel instantiation <64, 18>): the weight gradients of the samplers' strided and
// transposed 3x3 convolutions (models/erfnet_RA_parallel.py:24-45, 160-172) are one-tap jobs of the same kernel whose
// producers gather activation / gradient pixels through ConvGeom's affine maps (pixel = v * stride + tap offset, zeros
// outside the tensor) -- one launch per layer, one job per (tap, 64-channel block of CIN).  Narrow operands are stacked
// into 16- or 4-channel slots of the 64-channel operand row, each slot with its own tap offset: the nine taps of the
// 3 -> 13 convolution form ONE job (stacked along M), the taps of the 64 -> 16 upsampler that read the same activation
// pixel share a job (their gradient pixels stacked along N).  G may arrive pre-split (S16 format, g_split).
//
// Warp roles: warps 0-15 producers (four groups of 4 warps fill ring stages round-robin; the producers are bound by the
// latency of their own instruction stream, so warp count is what buys throughput), warps 16-18: one MMA issuer per tap.
#include <cuda_bf16.h>

#include "kernels.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mdil {
namespace wtc {

constexpr int TP = 16;        // pixels per chunk (= one K step of kind::f16)
constexpr int NWORK = 512;    // producer threads (warps 0..15)
constexpr int NGRP = 4;       // producer groups of 128 threads
constexpr int W_ISSUE0 = NWORK / 32;

template <int C> struct Cfg {
  static constexpr int NST = C == 128 ? 12 : 16;               // ring stages
  static constexpr uint32_t PART = C * TP * 2;                  // one of A_hi, A_lo, G_hi, G_lo (16-bit elements)
  static constexpr uint32_t STAGE = 4 * PART;
  static constexpr uint32_t SLAB = TP * 128;                    // 64-channel slab stride (LBO)
  static constexpr uint32_t HDR = 1024;
  static constexpr uint32_t SMEM = 1024 + HDR + NST * STAGE;
  // C = 64: the hi and lo images are stacked along M (A) and N (G): ONE M=128, N=128 MMA per tap and chunk computes
  // [Ah;Al]^T [Gh|Gl] (all four hi/lo products) where the unstacked form needs three M=64, N=64 MMAs.
  // The four 64x64 blocks of the accumulator are summed by the epilogue's red.global.add.
  static constexpr bool STACK = C == 64;
  static constexpr int ACCW = STACK ? 128 : C;                  // accumulator columns per tap
  static constexpr int TMEM_COLS = 512;                         // 3 accumulators of ACCW columns, power of two
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(acc) : "memory");
}
// MN-major operand, SWIZZLE_128B (16-bit elements): a 128-byte row = 64 channels of one pixel; LBO = stride between
// 64-channel groups, SBO = 1024 (8 pixel rows)
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 16 columns, no wait (call tmem_ld_wait before using the values)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// x0, x1 -> packed bf16 halves (low half-word = x0): hi = round-to-nearest of x, lo = round-to-nearest of x - hi
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(x1 - h1), "f"(x0 - h0));
}
__device__ __forceinline__ void split4(const float4& x, uint2& hi, uint2& lo) {
  split2(x.x, x.y, hi.x, lo.x);
  split2(x.z, x.w, hi.y, lo.y);
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct Plan {
  int U, V, Ul, Vl, vblocks, nseg, L, units, halo;   // halo = 1 for 3 taps
};
template <int NJ> struct JobsT {
  WgradTcArgs a[NJ];
  Plan pl[NJ];
  int cta0[NJ + 1];  // first CTA of every job; unused jobs are empty ranges at the end
};
constexpr int kGatherJobs = 18;   // gathered one-tap jobs per launch (9 taps x two 64-channel blocks of CIN = 128)

// decode unit -> strip + segment
struct Unit { int n, ru, rv, vb, ul0, Lu; };
__device__ __forceinline__ Unit decode_unit(int unit, const Plan& pl, int d) {
  Unit u;
  int t = unit;
  const int seg = t % pl.nseg; t /= pl.nseg;
  u.vb = t % pl.vblocks; t /= pl.vblocks;
  u.rv = t % d; t /= d;
  u.ru = t % d; t /= d;
  u.n = t;
  u.ul0 = seg * pl.L;
  u.Lu = min(pl.L, pl.Ul - u.ul0);
  return u;
}

template <int C, int NJ>
__global__ void __launch_bounds__(NWORK + 96, 1)
wgrad_tc_kernel(const __grid_constant__ JobsT<NJ> jobs) {
  using K = Cfg<C>;
  constexpr bool GEN = NJ > 3;       // gathered one-tap jobs (sampler convolutions)
  // up to three independent weight gradients (the three convolutions of a factorised pair) share one launch: CTAs
  // [cta0[j], cta0[j+1]) work on job j.  Every CTA ends with ONE accumulator flush (red.global.add of [ntaps][C][C]),
  // which at C = 128 costs as much as the main loop of a 148-CTA launch: three launches paid it three times.
  int job = 0;
#pragma unroll
  for (int j = 1; j < NJ; ++j) job += (int)blockIdx.x >= jobs.cta0[j] ? 1 : 0;
  const WgradTcArgs& a = jobs.a[job];
  const Plan& pl = jobs.pl[job];
  const int bid = (int)blockIdx.x - jobs.cta0[job], nbid = jobs.cta0[job + 1] - jobs.cta0[job];
  constexpr int NST = K::NST;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t hdr = raw + ((1024 - (raw & 1023)) & 1023);
  unsigned char* gen = smem_raw + (hdr - raw);
  const uint32_t bar_full = hdr, bar_empty = hdr + 8 * NST, bar_done = hdr + 16 * NST, tmem_slot = bar_done + 16;
  const uint32_t ring = hdr + K::HDR;
  // header: barriers [0, 16 NST + 32) <= 288 | trace counters [320, 448) | bias sums [512, 1024)
  float* bias_red = reinterpret_cast<float*>(gen + 512);   // [C] (C <= 128)
  long long* trc = reinterpret_cast<long long*>(gen + 320);  // 16 trace counters of CTA 0 (a.trace)
  const bool tracing = a.trace != 0 && blockIdx.x == 0;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d = a.dil, ntaps = a.ntaps;
  const long su = a.vert ? (long)a.W * C : C, sv = a.vert ? C : (long)a.W * C;

  if (tid == 0) {
    for (int i = 0; i < NST; ++i) { mbar_init(bar_full + 8 * i, NWORK / NGRP); mbar_init(bar_empty + 8 * i, (uint32_t)ntaps); }
    mbar_init(bar_done, (uint32_t)ntaps);     // one commit per issuing warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < C) bias_red[tid] = 0.f;
  if (tracing && tid < 16) trc[tid] = 0;
  const long long t_start = tracing ? clock64() : 0;
  if (warp == W_ISSUE0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(K::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - hdr));

  if (warp >= W_ISSUE0) {
    // =========================================================== MMA issuers: one warp (one lane) per tap: the three
    // taps own separate TMEM accumulators and are issued by three warps in parallel; each accumulator still receives an
    // ordered sequence of MMAs from a single thread (deterministic).
    const int t = warp - W_ISSUE0;
    if (lane == 0 && t < ntaps) {
      // M = N = ACCW, both operands MN-major (bits 15, 16)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(K::ACCW >> 3) << 17) |
                             ((uint32_t)(K::ACCW >> 4) << 24);      // fp32 accumulate, bf16 x bf16, both MN-major
      const uint32_t acc = tmem + t * K::ACCW;
      uint32_t q = 0;         // running stage-fill counter (identical on the producer side)
      uint32_t waited = 0;    // fills [0, waited) are known to have landed
      uint32_t started = 0;
      for (int unit = bid; unit < pl.units; unit += nbid) {
        const Unit un = decode_unit(unit, pl, d);
        const uint32_t q0 = q;
        // chunk c (c = -halo .. Lu-1+halo) is fill q0 + c + halo.  Tap t of gradient chunk j reads the gradient of fill
        // q0+j+halo and the activation of fill q0+j+t: the newest fill it needs is q0 + j + max(halo, t)
        for (int j = 0; j < un.Lu; ++j) {
          const uint32_t need = q0 + j + (uint32_t)(pl.halo > t ? pl.halo : t) + 1;
          const long long tw0 = tracing ? clock64() : 0;
          while (waited < need) {
            mbar_wait(bar_full + 8 * (waited % NST), (waited / NST) & 1);
            ++waited;
          }
          if (tracing) trc[t] += clock64() - tw0;
          tc_fence_after();
          const uint32_t fg = q0 + j + pl.halo;             // fill holding gradient chunk j (and activation chunk j)
          const uint32_t gbase = ring + (fg % NST) * K::STAGE + 2 * K::PART;
          const uint32_t fa = fg + t - pl.halo;             // activation chunk j + t - 1 (or j for the 1x1)
          const uint32_t abase = ring + (fa % NST) * K::STAGE;
          {
            // STACK (C = 64): the "next 64-channel group" of the M = 128 / N = 128 operand is the lo image (LBO = PART)
            const uint32_t lbo = K::STACK ? K::PART : K::SLAB;
            const uint64_t ah = desc_mn(abase, lbo), gh = desc_mn(gbase, lbo);
            mma_bf16(acc, ah, gh, idesc, started);
            started = 1u;
            if (!K::STACK) {
              const uint64_t al = desc_mn(abase + K::PART, lbo), gl = desc_mn(gbase + K::PART, lbo);
              mma_bf16(acc, al, gh, idesc, 1u);
              mma_bf16(acc, ah, gl, idesc, 1u);
            }
          }
          // fill q0+j (activation chunk j-1 / the 1x1's chunk j) is not read by this tap after these retire:
          // tap 0 read it as its activation now, tap 1 at chunk j-1 (and its gradient then), tap 2 at chunk j-2
          umma_commit(bar_empty + 8 * ((q0 + j) % NST));
        }
        if (pl.halo) {   // the last two fills of the unit (chunks Lu-1 and Lu)
          umma_commit(bar_empty + 8 * ((q0 + un.Lu) % NST));
          umma_commit(bar_empty + 8 * ((q0 + un.Lu + 1) % NST));
        }
        q = q0 + un.Lu + 2 * pl.halo;
      }
      umma_commit(bar_done);
    }
  } else {
    // =========================================================== producers: four groups fill the stages round-robin
    const int grp = warp & (NGRP - 1);
    const int tg = (warp / NGRP) * 32 + lane;        // 0..127 inside the group
    constexpr int C4 = C / 4;
    constexpr int RPT = TP * C4 / 128;               // rows (float4 per part) handled by one thread: 4 (C=128) / 2 (C=64)
    const int c4 = tg % C4, row0 = tg / C4;
    constexpr int RSTEP = 128 / C4;                   // row stride between a thread's rows
    const int ch = c4 * 4;
    const uint32_t in_slab = (uint32_t)(c4 >> 4) * K::SLAB;          // 64-channel slab
    const uint32_t c16 = (uint32_t)((c4 & 15) >> 1), halfo = (uint32_t)(c4 & 1) * 8;  // 16-byte chunk of the row, 8-byte half
    float4 sc = make4(1.f), sh = make4(0.f);
    if (a.a_scale != nullptr) { sc = ldg4(a.a_scale + ch); sh = ldg4(a.a_shift + ch); }
    float4 bsum = make4(0.f);
    // Fills are walked in batches of B per group: the B * 2 * RPT independent 128-bit loads of a batch are all in flight
    // before the first one is consumed (64 KB in flight per SM instead of 16 KB: the producers were latency-bound).
    constexpr int B = 1;
    struct FillDesc { size_t img; int n, u, rv, vb; uint32_t q; bool uok, interior, valid; };
    int unit = bid, f = 0, nfill = 0;
    uint32_t q = 0;
    Unit un;
    if (unit < pl.units) { un = decode_unit(unit, pl, d); nfill = un.Lu + 2 * pl.halo; }
    auto next_fill = [&](FillDesc& fd) {
      fd.valid = false;
      while (unit < pl.units) {
        if (f >= nfill) {
          unit += nbid;
          f = 0;
          if (unit < pl.units) { un = decode_unit(unit, pl, d); nfill = un.Lu + 2 * pl.halo; }
          continue;
        }
        const bool mine = (int)(q & (NGRP - 1)) == grp;
        if (mine) {
          const int cidx = f - pl.halo;                       // chunk index: -1 .. Lu
          const int ul = un.ul0 + cidx;
          fd.interior = cidx >= 0 && cidx < un.Lu;            // carries a gradient chunk
          fd.u = un.ru + d * ul;
          fd.uok = ul >= 0 && fd.u < pl.U;
          fd.rv = un.rv; fd.vb = un.vb;
          fd.img = (size_t)un.n * a.H * a.W * C;
          fd.n = un.n;
          fd.q = q;
          fd.valid = true;
        }
        ++f; ++q;
        if (mine) return;
      }
    };
    // gathered jobs: this thread's four channels belong to one slot of each operand (tap offsets, source channels)
    int gen_ady = 0, gen_adx = 0, gen_gdy = 0, gen_gdx = 0, gen_ach = 0, gen_gch = 0;
    bool gen_aon = false, gen_gon = false;
    if (GEN) {
      const int sa = ch / a.a_sw, sg = ch / a.g_sw;
      const uint32_t na = (uint32_t)(a.a_slots >> (4 * sa)) & 15u, ng = (uint32_t)(a.g_slots >> (4 * sg)) & 15u;
      gen_ady = (int)(na & 3u) - 1; gen_adx = (int)(na >> 2) - 1;
      gen_gdy = (int)(ng & 3u) - 1; gen_gdx = (int)(ng >> 2) - 1;
      gen_ach = a.a_coff + ch % a.a_sw; gen_gch = a.g_coff + ch % a.g_sw;
      gen_aon = sa < a.a_nslots;
      gen_gon = sg < a.g_nslots && ch % a.g_sw < a.g_cout;
    }
    // register double buffer: the loads of batch n+1 are issued before batch n is split and stored
    const bool g_split = a.g_split != 0;
    auto load_batch = [&](FillDesc (&fd)[B], float4 (&av)[B][RPT], float4 (&gv)[B][RPT], uint32_t& inside) {
      inside = 0;   // bit b*RPT+i: the pixel is inside the image (BN+ReLU prologue applies)
#pragma unroll
      for (int b = 0; b < B; ++b) {
        next_fill(fd[b]);
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          av[b][i] = make4(0.f);
          gv[b][i] = make4(0.f);
          if (fd[b].valid) {
            const int row = row0 + i * RSTEP;
            const int v = fd[b].rv + d * (fd[b].vb * TP + row);
            if (GEN) {
              if (fd[b].uok && v < pl.V) {     // (u, v) = (row, column) of the virtual grid
                const int ay = fd[b].u * a.a_sy + gen_ady, ax = v * a.a_sx + gen_adx;
                if (gen_aon && (unsigned)ay < (unsigned)a.AH && (unsigned)ax < (unsigned)a.AW)
                  av[b][i] = ldg4(a.A + ((size_t)(fd[b].n * a.AH + ay) * a.AW + ax) * a.lda + gen_ach);
                const int gy = fd[b].u * a.g_sy + gen_gdy, gx = v * a.g_sx + gen_gdx;
                if (gen_gon && (unsigned)gy < (unsigned)a.GH && (unsigned)gx < (unsigned)a.GW)
                  gv[b][i] = ldg4(a.G + ((size_t)(fd[b].n * a.GH + gy) * a.GW + gx) * a.ldg + gen_gch);
              }
            } else if (fd[b].uok && v < pl.V) {
              const size_t off = fd[b].img + fd[b].u * su + v * sv + ch;
              av[b][i] = ldg4(a.A + off);
              if (fd[b].interior) gv[b][i] = ldg4(a.G + off);   // S16: the same 16 bytes hold 4 hi + 4 lo halves
              inside |= 1u << (b * RPT + i);
            }
          }
        }
      }
    };
    auto store_batch = [&](const FillDesc (&fd)[B], const float4 (&av)[B][RPT], const float4 (&gv)[B][RPT], uint32_t inside) {
#pragma unroll
      for (int b = 0; b < B; ++b) {
        if (!fd[b].valid) continue;
        const uint32_t qq = fd[b].q;
        const long long tw0 = tracing ? clock64() : 0;
        if (qq >= (uint32_t)NST) mbar_wait(bar_empty + 8 * (qq % NST), ((qq / NST) - 1) & 1);   // ring slot free?
        if (tracing && tg == 0 && grp < 2) trc[4 + grp] += clock64() - tw0;
        const uint32_t sbase = ring + (qq % NST) * K::STAGE + in_slab;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int row = row0 + i * RSTEP;
          float4 a4 = av[b][i], g4 = gv[b][i];
          if (a.a_scale != nullptr && ((inside >> (b * RPT + i)) & 1u)) {
            a4.x = fmaxf(fmaf(a4.x, sc.x, sh.x), 0.f);
            a4.y = fmaxf(fmaf(a4.y, sc.y, sh.y), 0.f);
            a4.z = fmaxf(fmaf(a4.z, sc.z, sh.z), 0.f);
            a4.w = fmaxf(fmaf(a4.w, sc.w, sh.w), 0.f);
          }
          const uint32_t ra = sbase + (uint32_t)row * 128;
          const uint32_t ad = ra + (((c16 ^ ((ra >> 7) & 7)) << 4) | halfo);
          uint2 hi, lo;
          split4(a4, hi, lo);
          *reinterpret_cast<uint2*>(gen + (ad - hdr)) = hi;
          *reinterpret_cast<uint2*>(gen + (ad - hdr) + K::PART) = lo;
          if (g_split) {
            hi = make_uint2(__float_as_uint(g4.x), __float_as_uint(g4.y));
            lo = make_uint2(__float_as_uint(g4.z), __float_as_uint(g4.w));
            if (a.db != nullptr) {
              bsum.x += __uint_as_float(hi.x << 16) + __uint_as_float(lo.x << 16);
              bsum.y += __uint_as_float(hi.x & 0xffff0000u) + __uint_as_float(lo.x & 0xffff0000u);
              bsum.z += __uint_as_float(hi.y << 16) + __uint_as_float(lo.y << 16);
              bsum.w += __uint_as_float(hi.y & 0xffff0000u) + __uint_as_float(lo.y & 0xffff0000u);
            }
          } else {
            split4(g4, hi, lo);
            bsum.x += g4.x; bsum.y += g4.y; bsum.z += g4.z; bsum.w += g4.w;
          }
          *reinterpret_cast<uint2*>(gen + (ad - hdr) + 2 * K::PART) = hi;
          *reinterpret_cast<uint2*>(gen + (ad - hdr) + 3 * K::PART) = lo;
        }
        fence_proxy_async();
        mbar_arrive(bar_full + 8 * (qq % NST));
      }
    };
    if (C == 64) {     // register double buffer: the loads of fill n+1 are issued before fill n is split and stored
      FillDesc fd0[B], fd1[B];
      float4 av0[B][RPT], gv0[B][RPT], av1[B][RPT], gv1[B][RPT];
      uint32_t in0, in1;
      load_batch(fd0, av0, gv0, in0);
      for (;;) {
        if (!fd0[0].valid) break;
        load_batch(fd1, av1, gv1, in1);
        store_batch(fd0, av0, gv0, in0);
        if (!fd1[0].valid) break;
        load_batch(fd0, av0, gv0, in0);
        store_batch(fd1, av1, gv1, in1);
      }
    } else {           // C = 128: 8 x 128-bit loads per thread and fill (64 KB in flight per SM) without the second register set
      FillDesc fd0[B];
      float4 av0[B][RPT], gv0[B][RPT];
      uint32_t in0;
      for (;;) {
        load_batch(fd0, av0, gv0, in0);
        if (!fd0[0].valid) break;
        store_batch(fd0, av0, gv0, in0);
      }
    }
    if (a.db != nullptr) {
      atomicAdd(bias_red + ch + 0, bsum.x);
      atomicAdd(bias_red + ch + 1, bsum.y);
      atomicAdd(bias_red + ch + 2, bsum.z);
      atomicAdd(bias_red + ch + 3, bsum.w);
    }
    // =========================================================== epilogue: TMEM -> red.global.add.v4
    if (tracing && tid == 0) trc[6] = clock64() - t_start;
    mbar_wait(bar_done, 0);
    if (tracing && tid == 0) trc[7] = clock64() - t_start;
    tc_fence_after();
    __syncwarp();
    const int qd = warp & 3, cg = warp >> 2;        // TMEM lane quadrant, column group (0..3) of this warp
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    // Every CTA adds its partial [ntaps][C][C] into the same buffer: the walk over taps and 16-byte pieces is rotated by
    // the CTA index so that concurrent CTAs hit different addresses (same-address L2 atomics serialise).
    const int rot = bid;
    if (K::STACK) {
      // accumulator rows: [0,64) = A_hi channels, [64,128) = A_lo channels; columns [0,64) = G_hi, [64,128) = G_lo
      const int ci = (qd * 32 + lane) & 63;
      const int col0 = cg * 16;
      for (int tt = 0; tt < ntaps; ++tt) {
        const int t = (tt + rot) % ntaps;
        float v1[16], v2[16];
        tmem_ld16(tmem + lane_addr + t * K::ACCW + col0, v1);
        tmem_ld16(tmem + lane_addr + t * K::ACCW + 64 + col0, v2);
        tmem_ld_wait();
        float* dst = a.dWacc + ((size_t)t * C + ci) * C + col0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j4 = (jj + (rot >> 2)) & 3;   // (static register indexing is kept by the select chain below)
          float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k == j4) { x0 = v1[k * 4 + 0] + v2[k * 4 + 0]; x1 = v1[k * 4 + 1] + v2[k * 4 + 1]; x2 = v1[k * 4 + 2] + v2[k * 4 + 2]; x3 = v1[k * 4 + 3] + v2[k * 4 + 3]; }
          red_add_v4(dst + j4 * 4, x0, x1, x2, x3);
        }
      }
    } else {
      const int row = qd * 32 + lane;     // M = 128: accumulator row (ci) = TMEM lane
      const int col0 = cg * 32;
      for (int tt = 0; tt < ntaps; ++tt) {
        const int t = (tt + rot) % ntaps;
        float val[32];
        tmem_ld32(tmem + lane_addr + t * K::ACCW + col0, val);
        float* dst = a.dWacc + ((size_t)t * C + row) * C + col0;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j4 = (jj + (rot >> 2)) & 7;
          float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k == j4) { x0 = val[k * 4 + 0]; x1 = val[k * 4 + 1]; x2 = val[k * 4 + 2]; x3 = val[k * 4 + 3]; }
          red_add_v4(dst + j4 * 4, x0, x1, x2, x3);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tracing && tid == 0)
    printf("wgrad_tc<%d> CTA0 taps=%d units=%d: total %lld clk | issuers waited for fills %lld %lld %lld | producers waited for slots %lld %lld | producers done %lld, MMAs done %lld\n",
           C, ntaps, pl.units, clock64() - t_start, trc[0], trc[1], trc[2], trc[4], trc[5], trc[6], trc[7]);
  if (a.db != nullptr && tid < C) atomicAdd(a.db + tid, bias_red[tid]);
  if (warp == W_ISSUE0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(K::TMEM_COLS) : "memory");
  }
}

__global__ void wgrad_unpack_kernel(const float* __restrict__ acc, float* __restrict__ dW, int C, int ntaps, long s_ci,
                                    long s_co, long s_t) {
  const long total = (long)ntaps * C * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int co = (int)(i % C);
    const int ci = (int)((i / C) % C);
    const int t = (int)(i / ((long)C * C));
    dW[ci * s_ci + co * s_co + t * s_t] = acc[i];
  }
}

// all (up to 6) tensor-core weight gradients of a block leave their [ntaps][ci][co] accumulators in one launch
__global__ void wgrad_unpack_multi_kernel(const UnpackList ul, int C) {
  const UnpackItem& it = ul.item[blockIdx.y];
  const long total = (long)it.ntaps * C * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int co = (int)(i % C);
    const int ci = (int)((i / C) % C);
    const int t = (int)(i / ((long)C * C));
    it.dW[ci * it.s_ci + co * it.s_co + t * it.s_t] = it.acc[i];
  }
  if (it.db != nullptr && blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += blockDim.x) it.db[c] = it.dbacc[c];
}

// packed-4 view (a C = 16 block run as C = 64 on [N,H,W/4,64]): fold the [3][64][64] accumulators of wgrad_tc<64> back
// to the 16 x 16 x 3 weight gradient (torch layout [co][ci][k]) and the [64] bias sums to [16]
__global__ void wgrad_unpack_p4_kernel(const UnpackP4List ul) {
  const UnpackP4Item& it = ul.item[blockIdx.x];
  for (int i = threadIdx.x; i < 16 * 16 * 3; i += blockDim.x) {
    const int k = i % 3, ci = (i / 3) % 16, co = i / 48;
    float sum = 0.f;
    if (it.horizontal) {
      for (int gt = 0; gt < 3; ++gt)
        for (int pi = 0; pi < 4; ++pi) {
          const int po = 4 * (gt - 1) + pi + 1 - k;          // 4 (gt - 1) + pi - po + 1 == k
          if (po >= 0 && po < 4) sum += it.acc[(gt * 64 + pi * 16 + ci) * 64 + po * 16 + co];
        }
    } else {
      for (int p = 0; p < 4; ++p) sum += it.acc[(k * 64 + p * 16 + ci) * 64 + p * 16 + co];
    }
    it.dW[(co * 16 + ci) * 3 + k] = sum;
  }
  if (it.db != nullptr && threadIdx.x < 16) {
    const int c = threadIdx.x;
    it.db[c] = (it.dbacc[c] + it.dbacc[16 + c]) + (it.dbacc[32 + c] + it.dbacc[48 + c]);
  }
}

static Plan make_plan(const WgradTcArgs& a, int nctas) {
  Plan pl;
  pl.U = a.vert ? a.H : a.W;
  pl.V = a.vert ? a.W : a.H;
  pl.Ul = cdiv(pl.U, a.dil);
  pl.Vl = cdiv(pl.V, a.dil);
  pl.vblocks = cdiv(pl.Vl, TP);
  pl.halo = a.ntaps == 3 ? 1 : 0;
  const long strips = (long)a.N * a.dil * a.dil * pl.vblocks;
  // split every strip into nseg segments of L chunks so that the job's busiest CTA has the least work:
  // rounds = ceil(units / CTAs), each unit costs L chunks plus the 2 * halo fills of its ends
  int best_nseg = 1;
  double best_cost = 1e30;
  for (int nseg = 1; nseg <= pl.Ul; ++nseg) {
    const int L = cdiv(pl.Ul, nseg);
    if (L < 4 && nseg > 1) break;
    const int ns = cdiv(pl.Ul, L);
    const long units = strips * ns;
    const double cost = (double)cdiv((int)units, nctas) * (L + 2 * pl.halo + 1);
    if (cost < best_cost - 1e-9) { best_cost = cost; best_nseg = ns; }
  }
  pl.L = cdiv(pl.Ul, best_nseg);
  pl.nseg = cdiv(pl.Ul, pl.L);
  pl.units = (int)(strips * pl.nseg);
  return pl;
}

template <int C>
int launch_c(const WgradTcArgs* a, int n, cudaStream_t s) {
  using K = Cfg<C>;
  static_assert(K::SMEM <= 227 * 1024, "wgrad_tc shared memory budget");
  static_assert(16 * K::NST + 32 <= 320, "wgrad_tc barrier header layout");
  // CTA shares: the main loop costs the same for every job (same pixels: the producers bound it), the accumulator flush
  // scales with the tap count (measured at C = 128: flush of three taps ~ one main loop)
  double cost[3] = {0, 0, 0}, tot = 0;
  for (int j = 0; j < n; ++j) { cost[j] = 1.0 + (C == 128 ? 0.85 : 0.3) * a[j].ntaps / 3.0; tot += cost[j]; }
  JobsT<3> jobs;
  memset(&jobs, 0, sizeof(jobs));
  int used = 0;
  for (int j = 0; j < n; ++j) {
    int share = j == n - 1 ? kNumSMs - used : (int)(kNumSMs * cost[j] / tot + 0.5);
    if (share < 1) share = 1;
    jobs.a[j] = a[j];
    jobs.pl[j] = make_plan(a[j], share);
    MDIL_REQUIRE(jobs.pl[j].units > 0, "wgrad_tc: unit count");
    if (jobs.pl[j].units < share) share = jobs.pl[j].units;
    jobs.cta0[j] = used;
    used += share;
  }
  for (int j = n; j <= 3; ++j) jobs.cta0[j] = used;
  MDIL_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<C, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM));
  wgrad_tc_kernel<C, 3><<<used, NWORK + 96, K::SMEM, s>>>(jobs);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// gathered one-tap jobs (C = 64): equal CTA shares (every job walks the same virtual grid)
static int launch_gather(const WgradTcArgs* a, int n, cudaStream_t s) {
  using K = Cfg<64>;
  static_assert(sizeof(JobsT<kGatherJobs>) <= 4000, "wgrad_tc: kernel parameter space");
  JobsT<kGatherJobs> jobs;
  memset(&jobs, 0, sizeof(jobs));
  int used = 0;
  for (int j = 0; j < n; ++j) {
    int share = (kNumSMs * (j + 1)) / n - (kNumSMs * j) / n;
    if (share < 1) share = 1;
    jobs.a[j] = a[j];
    jobs.pl[j] = make_plan(a[j], share);
    MDIL_REQUIRE(jobs.pl[j].units > 0, "wgrad_tc: unit count");
    if (jobs.pl[j].units < share) share = jobs.pl[j].units;
    jobs.cta0[j] = used;
    used += share;
  }
  for (int j = n; j <= kGatherJobs; ++j) jobs.cta0[j] = used;
  MDIL_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<64, kGatherJobs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM));
  wgrad_tc_kernel<64, kGatherJobs><<<used, NWORK + 96, K::SMEM, s>>>(jobs);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// accumulators of the gathered jobs -> dW in the caller's layout; bias sums -> db.  Accumulator row m = slot_a * a_sw +
// ci, column n = slot_g * g_sw + co; one of the operands is stacked (or neither): widx[job][stacked slot] = weight tap.
struct GatherUnpack {
  int n, cin, cout;
  short ci0[kGatherJobs], a_sw[kGatherJobs], a_nslots[kGatherJobs], g_sw[kGatherJobs], g_nslots[kGatherJobs];
  unsigned short bias_slots[kGatherJobs];     // gradient slots whose sums count for db (every class once)
  signed char widx[kGatherJobs][16];
};
__global__ void wgrad_gather_unpack_kernel(const GatherUnpack gu, const float* __restrict__ acc, float* __restrict__ dW,
                                           long s_ci, long s_co, long s_t, const float* __restrict__ dbacc,
                                           float* __restrict__ db) {
  const int total = gu.n * 64 * 64;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int j = e >> 12, m = (e >> 6) & 63, n = e & 63;
    const int sa = m / gu.a_sw[j], sg = n / gu.g_sw[j];
    const int ci = gu.ci0[j] + m % gu.a_sw[j], co = n % gu.g_sw[j];
    if (sa < gu.a_nslots[j] && sg < gu.g_nslots[j] && ci < gu.cin && co < gu.cout)
      dW[(long)gu.widx[j][sa > sg ? sa : sg] * s_t + (long)ci * s_ci + (long)co * s_co] = acc[e];
  }
  if (db != nullptr && blockIdx.x == 0 && threadIdx.x < gu.cout) {
    float sum = 0.f;
    for (int j = 0; j < gu.n; ++j)
      for (int sg = 0; sg < gu.g_nslots[j]; ++sg)
        if ((gu.bias_slots[j] >> sg) & 1) sum += dbacc[j * 64 + sg * gu.g_sw[j] + threadIdx.x];
    db[threadIdx.x] = sum;
  }
}

}  // namespace wtc

int launch_wgrad_tc_multi(const WgradTcArgs* a_in, int n, cudaStream_t s) {
  static const int trace = getenv("MDIL_TC_TRACE") != nullptr ? 1 : 0;
  if (n == 0) return 0;
  MDIL_REQUIRE(n >= 1 && n <= 3, "wgrad_tc: 1 to 3 jobs per launch");
  WgradTcArgs a[3];
  for (int j = 0; j < n; ++j) {
    a[j] = a_in[j];
    a[j].trace = trace;
    MDIL_REQUIRE(a[j].ntaps == 1 || a[j].ntaps == 3, "wgrad_tc: 1 or 3 taps");
    MDIL_REQUIRE(a[j].dWacc != nullptr && ((uintptr_t)a[j].dWacc & 15) == 0, "wgrad_tc: accumulator buffer");
    MDIL_REQUIRE(a[j].C == a[0].C, "wgrad_tc: the jobs of one launch share C");
  }
  switch (a[0].C) {
    case 128: return wtc::launch_c<128>(a, n, s);
    case 64: return wtc::launch_c<64>(a, n, s);
    default: return set_error(-2, "wgrad_tc: C must be 64 or 128", __FILE__, __LINE__);
  }
}

int launch_wgrad_tc(const WgradTcArgs& a, cudaStream_t s) { return launch_wgrad_tc_multi(&a, 1, s); }

static inline unsigned long long gather_nibble(int dy, int dx, int slot) {
  return (unsigned long long)((dy + 1) | ((dx + 1) << 2)) << (4 * slot);
}

bool wgrad_gather_ok(const ConvGeom& g) {
  static const bool on = [] { const char* e = getenv("MDIL_WGRAD_GATHER"); return !(e != nullptr && strcmp(e, "0") == 0); }();
  int taps = 0;
  for (int c = 0; c < g.nclasses; ++c) {
    taps += g.cls[c].ntaps;
    if (g.cls[c].o_dy < -1 || g.cls[c].o_dy > 1 || g.cls[c].o_dx < -1 || g.cls[c].o_dx > 1) return false;
    for (int t = 0; t < g.cls[c].ntaps; ++t)
      if (g.cls[c].a_dy[t] < -1 || g.cls[c].a_dy[t] > 1 || g.cls[c].a_dx[t] < -1 || g.cls[c].a_dx[t] > 1) return false;
  }
  if (!on || g.lda % 4 != 0 || g.a_coff % 4 != 0 || g.ldg % 4 != 0 || g.g_coff % 4 != 0 || g.COUT > 64) return false;
  if (g.g_coff + (g.COUT + 3) / 4 * 4 > g.ldg) return false;
  if (g.CIN % 64 == 0) return g.CIN_VALID == g.CIN && taps * (g.CIN / 64) <= wtc::kGatherJobs;
  return (g.CIN == 4 || g.CIN == 16) && g.nclasses == 1;      // narrow input: taps stacked along M
}
size_t wgrad_gather_scratch_floats() { return (size_t)wtc::kGatherJobs * (64 * 64 + 64); }

int launch_wgrad_gather_tc(const ConvGeom& g, const float* A, const float* G, float* dW, long s_ci, long s_co, long s_t,
                           float* db, float* scratch, cudaStream_t s) {
  static const int trace = getenv("MDIL_TC_TRACE") != nullptr ? 1 : 0;
  MDIL_REQUIRE(wgrad_gather_ok(g), "wgrad_gather: unsupported geometry");
  MDIL_REQUIRE(scratch != nullptr && ((uintptr_t)scratch & 15) == 0, "wgrad_gather: scratch buffer");
  WgradTcArgs jobs[wtc::kGatherJobs];
  wtc::GatherUnpack gu;
  memset(&gu, 0, sizeof(gu));
  float* dbacc = scratch + (size_t)wtc::kGatherJobs * 64 * 64;
  int n = 0;
  auto new_job = [&](int ci0, int a_sw, int g_sw) -> WgradTcArgs& {
    WgradTcArgs& w = jobs[n];
    memset(&w, 0, sizeof(w));
    w.A = A; w.G = G; w.dWacc = scratch + (size_t)n * 64 * 64; w.db = db != nullptr ? dbacc + n * 64 : nullptr;
    w.N = g.N; w.H = g.VH; w.W = g.VW; w.C = 64; w.dil = 1; w.ntaps = 1; w.vert = 1; w.trace = trace;
    w.AH = g.AH; w.AW = g.AW; w.lda = g.lda; w.a_coff = g.a_coff + ci0; w.a_sy = g.a_sy; w.a_sx = g.a_sx; w.a_sw = a_sw;
    w.GH = g.GH; w.GW = g.GW; w.ldg = g.ldg; w.g_coff = g.g_coff; w.g_sy = g.g_sy; w.g_sx = g.g_sx; w.g_sw = g_sw;
    w.g_cout = (g.COUT + 3) / 4 * 4;
    gu.ci0[n] = (short)ci0; gu.a_sw[n] = (short)a_sw; gu.g_sw[n] = (short)g_sw;
    return w;
  };
  if (g.CIN % 64 == 0 && g.COUT <= 16) {
    // narrow output (64 -> 16 upsampler): the taps that read the same activation pixel share a job, their gradient
    // pixels (one parity class each) stacked along N in 16-channel slots
    bool class_counted[kMaxClasses] = {false, false, false, false};
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx)
        for (int cb = 0; cb < g.CIN / 64; ++cb) {
          int slots = 0;
          for (int c = 0; c < g.nclasses; ++c)
            for (int t = 0; t < g.cls[c].ntaps; ++t)
              if (g.cls[c].a_dy[t] == dy && g.cls[c].a_dx[t] == dx) {
                MDIL_REQUIRE(n < wtc::kGatherJobs, "wgrad_gather: job count");
                WgradTcArgs& w = slots == 0 ? new_job(64 * cb, 64, 16) : jobs[n];
                if (slots == 0) { w.a_slots = gather_nibble(dy, dx, 0); w.a_nslots = 1; }
                MDIL_REQUIRE(slots < 4, "wgrad_gather: slot count");
                w.g_slots |= gather_nibble(g.cls[c].o_dy, g.cls[c].o_dx, slots);
                gu.widx[n][slots] = (signed char)g.cls[c].widx[t];
                if (cb == 0 && !class_counted[c]) { gu.bias_slots[n] |= (unsigned short)(1u << slots); class_counted[c] = true; }
                w.g_nslots = ++slots;
              }
          if (slots > 0) { gu.a_nslots[n] = 1; gu.g_nslots[n] = (short)slots; ++n; }
        }
  } else if (g.CIN % 64 == 0) {
    // one job per (tap, 64-channel block of CIN)
    for (int c = 0; c < g.nclasses; ++c)
      for (int t = 0; t < g.cls[c].ntaps; ++t)
        for (int cb = 0; cb < g.CIN / 64; ++cb) {
          MDIL_REQUIRE(n < wtc::kGatherJobs, "wgrad_gather: job count");
          WgradTcArgs& w = new_job(64 * cb, 64, 64);
          w.a_slots = gather_nibble(g.cls[c].a_dy[t], g.cls[c].a_dx[t], 0); w.a_nslots = 1;
          w.g_slots = gather_nibble(g.cls[c].o_dy, g.cls[c].o_dx, 0); w.g_nslots = 1;
          gu.widx[n][0] = (signed char)g.cls[c].widx[t];
          gu.a_nslots[n] = 1; gu.g_nslots[n] = 1;
          if (t == 0 && cb == 0) gu.bias_slots[n] = 1;     // every class walks its gradient pixels once
          ++n;
        }
  } else {
    // narrow input (3 -> 16, 16 -> 64 downsamplers; one tap class): taps stacked along M in CIN-channel slots
    const TapClass& tc = g.cls[0];
    const int per_max = 64 / g.CIN, nj = cdiv(tc.ntaps, per_max), per = cdiv(tc.ntaps, nj);
    for (int t0 = 0; t0 < tc.ntaps; t0 += per) {
      MDIL_REQUIRE(n < wtc::kGatherJobs, "wgrad_gather: job count");
      WgradTcArgs& w = new_job(0, g.CIN, 64);
      const int cnt = tc.ntaps - t0 < per ? tc.ntaps - t0 : per;
      for (int k = 0; k < cnt; ++k) {
        w.a_slots |= gather_nibble(tc.a_dy[t0 + k], tc.a_dx[t0 + k], k);
        gu.widx[n][k] = (signed char)tc.widx[t0 + k];
      }
      w.a_nslots = cnt;
      w.g_slots = gather_nibble(tc.o_dy, tc.o_dx, 0); w.g_nslots = 1;
      gu.a_nslots[n] = (short)cnt; gu.g_nslots[n] = 1;
      if (t0 == 0) gu.bias_slots[n] = 1;
      ++n;
    }
  }
  gu.n = n; gu.cin = g.CIN_VALID; gu.cout = g.COUT;
  MDIL_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * wgrad_gather_scratch_floats(), s));
  MDIL_TRY(wtc::launch_gather(jobs, n, s));
  int grid = (n * 64 * 64 + 255) / 256;
  if (grid > kNumSMs) grid = kNumSMs;
  wtc::wgrad_gather_unpack_kernel<<<grid, 256, 0, s>>>(gu, scratch, dW, s_ci, s_co, s_t, dbacc, db);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_wgrad_unpack_multi(const UnpackList& ul, int C, cudaStream_t s) {
  if (ul.n == 0) return 0;
  const long total = 3L * C * C;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs) grid = kNumSMs;
  wtc::wgrad_unpack_multi_kernel<<<dim3((unsigned)grid, (unsigned)ul.n), 256, 0, s>>>(ul, C);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_wgrad_unpack_p4(const UnpackP4List& ul, cudaStream_t s) {
  if (ul.n == 0) return 0;
  wtc::wgrad_unpack_p4_kernel<<<ul.n, 256, 0, s>>>(ul);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_wgrad_unpack(const float* acc, float* dW, int C, int ntaps, long s_ci, long s_co, long s_t, cudaStream_t s) {
  const long total = (long)ntaps * C * C;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  wtc::wgrad_unpack_kernel<<<grid, 256, 0, s>>>(acc, dW, C, ntaps, s_ci, s_co, s_t);
  MDIL_LAUNCH_CHECK();
  return 0;
}

}  // namespace mdil
