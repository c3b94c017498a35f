// Persistent, warp-specialised, software-pipelined tensor-core (tcgen05 / TMEM) fused factorised-pair kernel
// for C = 64 and C = 128 (same contract as nb1d_pair.cu: see the header comment there for the four uses,
// reference models/erfnet_RA_parallel.py:90-113 and :48-64).
//
// Arithmetic: error-compensated 3xTF32 (hi*hi + lo*hi + hi*lo, fp32 accumulation in TMEM).  Superseded as the default
// by nb1d_pair_h3.cu (16-bit split operands); kept selectable with MDIL_PAIR_IMPL=tc3 for A/B measurements.
//
// Structure:
//   * one CTA per SM walks a strided list of lattice tiles (persistent); per tile the five phases
//     load -> conv 1 (+adapter) -> epilogue 1 -> conv 2 -> epilogue 2 run on different warps and overlap:
//       warps 0..7   epilogue warps (TMEM lane quadrant = warp & 3, interleaved 16-column sub-blocks, 16x256b loads)
//       warps 8..11  loader warps: global -> registers (BN+ReLU prologue) -> hi/lo -> SWIZZLE_128B K-major operand
//       warp 12      MMA issuer (one elected lane issues tcgen05.mma kind::tf32, M=128, N=C, K=8)
//       warp 13      weight producer (cp.async.bulk + mbarrier ring; with a cluster, each CTA fetches 1/CL of every
//                    chunk and MULTICASTS it to all CTAs of the cluster: L2->SM weight traffic / CL)
//   * slab-granular hand-offs (a slab = 32 channels = one 128-byte operand row): the first conv starts when the
//     first 32 input channels of the tile are in shared memory, the second conv when the first 32 `mid` channels are;
//   * the second accumulator is double-buffered in TMEM, so epilogue 2 of tile i overlaps the MMAs of tile i+1;
//     for C = 64 the activation operand is double-buffered too (loads run one tile ahead);
//   * the epilogues read TMEM with the 16x256b shape (four lanes own 32 contiguous bytes of a pixel row): their inputs
//     (ReLU masks, p, dy, y; prefetched before the accumulator is waited for) and results use sector-filling 8-byte
//     global accesses without shared-memory staging; per-channel BatchNorm sums: warp shuffles -> shared -> fp64 atomics.
#include <atomic>

#include "kernels.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace mdil {
namespace tc3 {

constexpr int KC = 16;        // input channels per weight chunk (one 64-byte SWIZZLE_64B row)
constexpr int IN_MAX = 160;   // input-tile rows (pixels) held in shared memory
constexpr int N_EPI = 256;    // warps 0..7
constexpr int N_LOAD = 128;   // warps 8..11
constexpr int W_MMA = 12, W_PROD = 13;
constexpr int NTHREADS = 448;
constexpr int LROWS = IN_MAX / 16;   // rows per loader thread and slab

template <int C> struct Cfg {
  static constexpr int NBUF = C == 64 ? 2 : 1;              // activation operand buffers
  static constexpr int NSTAGE = C == 64 ? 8 : 4;            // weight ring depth
  static constexpr int SLABS = C / 32;
  static constexpr uint32_t SLAB_BYTES = IN_MAX * 128;
  static constexpr uint32_t ACT_BYTES = SLABS * SLAB_BYTES;  // one (hi or lo) operand image
  static constexpr uint32_t BUF_BYTES = 2 * ACT_BYTES;
  static constexpr uint32_t HALF_STAGE = C * 64;             // one (hi or lo) weight image of a chunk
  static constexpr uint32_t STAGE_BYTES = 2 * HALF_STAGE;
  static constexpr uint32_t HDR_BYTES = 2048;                // barriers | tmem slot | trace | [2][C] BatchNorm sums
  static constexpr uint32_t SUM_OFF = 1024;
  static constexpr uint32_t SMEM_BYTES = 1024 + HDR_BYTES + NBUF * BUF_BYTES + NSTAGE * STAGE_BYTES;
  static constexpr int NCH = C / KC;
  static constexpr int NSB = C / 32;                         // 16-column sub-blocks per epilogue warp
  // C = 64: the hi and lo weight images of a chunk are stacked along N (one MMA with N = 2C computes x*Whi into columns
  // [0,C) and x*Wlo into [C,2C)): 4 MMAs per chunk instead of 6 (the issue rate of the single MMA thread, ~100 clocks
  // per instruction whatever N <= 128, is what bounds the tensor pipe here: tools/umma_issue*.cu); the epilogues add the
  // two column halves.  C = 128 would need N = 256 accumulators (768 TMEM columns with the double buffer): not stacked.
  static constexpr bool NSTACK = C == 64;
  static constexpr uint32_t ACCW = NSTACK ? 2 * C : C;       // accumulator width in TMEM columns
  static constexpr uint32_t TMEM_COLS = 512;                 // acc1 [ACCW] + acc2 [2][ACCW] = 384 columns, power of two
};

struct TileShape { int TU, TV, TR; };

struct Geo {   // per-launch tile geometry (kernel argument)
  int TU, TV, TR;
  int total_tiles, tiles_per_cta, cl;
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
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask) : "memory");
}
__device__ __forceinline__ void mma_tf32_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d), "r"(a_lo),
      "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
// 16 TMEM lanes x 16 columns: lane l gets rows (l>>2), (l>>2)+8 and columns 2(l&3), +1 of each 8-column group:
// v[4c + 2k + e] = (row (l>>2) + 8k, column 8c + 2(l&3) + e).  No wait: several loads may be in flight.
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// weights are split once per step at pack time with the hardware rounding
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// hi/lo split of an activation on the integer/FMA pipes (cvt.rna.tf32 runs at a fraction of their rate and was throttling
// the math pipe of the loader and epilogue warps): hi = x rounded to TF32 (add half an ulp of the 10-bit mantissa, clear
// the 13 low bits: round-half-away), lo = x - hi exactly (|lo| <= 2^-11 |x|).  lo is stored as is: the tensor core reads
// only its TF32 bits (truncation, |error| < 2^-10 |lo| <= 2^-21 |x|, sign-symmetric because lo is).
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ void split4(const float4& x, float4& hi, float4& lo) {
  hi.x = tf32_hi(x.x); hi.y = tf32_hi(x.y); hi.z = tf32_hi(x.z); hi.w = tf32_hi(x.w);
  lo.x = x.x - hi.x; lo.y = x.y - hi.y; lo.z = x.z - hi.z; lo.w = x.w - hi.w;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// byte offset of (row, 16-byte chunk c) inside a SWIZZLE_128B slab (slab bases are 1024-byte aligned)
__device__ __forceinline__ uint32_t sw128_off(int row, int c) { return (uint32_t)row * 128 + ((uint32_t)(c ^ (row & 7)) << 4); }
struct TileCoord { int n, cb, ul0, vl0; bool dummy; };

template <int C>
__global__ void __launch_bounds__(NTHREADS, 1)
pair_tc3_kernel(const __grid_constant__ PairArgs a, const Geo geo) {
  using K = Cfg<C>;
  constexpr int NSTAGE = K::NSTAGE, NBUF = K::NBUF, SLABS = K::SLABS;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t hdr = raw + ((1024 - (raw & 1023)) & 1023);
  unsigned char* gen = smem_raw + (hdr - raw);   // generic pointer to hdr
  // ---- barriers
  const uint32_t bar_wfull = hdr, bar_wempty = hdr + 64;
  const uint32_t bar_infull = hdr + 128;      // [NBUF][SLABS]
  const uint32_t bar_midfull = hdr + 160;     // [NBUF][SLABS]
  const uint32_t bar_buffree = hdr + 192;     // [NBUF]
  const uint32_t bar_acc1full = hdr + 208;
  const uint32_t bar_acc2full = hdr + 216;    // [2]
  const uint32_t bar_acc2free = hdr + 232;    // [2]
  const uint32_t tmem_slot = hdr + 248;
  long long* trc = reinterpret_cast<long long*>(gen + 256);   // trace counters of CTA 0 (a.trace)
  const uint32_t act0 = hdr + K::HDR_BYTES;
  const uint32_t ring = act0 + NBUF * K::BUF_BYTES;
  const bool tracing = a.trace != 0 && blockIdx.x == 0;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int TU = geo.TU, TV = geo.TV, TR = geo.TR, TVH = TV + 2;
  const int RT = TR * TVH;             // rows per lattice step along u  (row order: [u][class][v])
  const int INROWS = (TU + 2) * RT;
  const int M1 = TU * RT;
  const int d = a.dil;
  const int U = a.vert_first ? a.H : a.W, V = a.vert_first ? a.W : a.H;
  const int Ul = (U + d - 1) / d, Vl = (V + d - 1) / d;
  const int tiles_u = (Ul + TU - 1) / TU, tiles_v = (Vl + TV - 1) / TV;
  const int ncb = (d * d + TR - 1) / TR;
  const int NAD = a.has_adapter ? 1 : 0;
  const int G = K::NCH * (6 + NAD);          // weight chunks per tile
  const int CL = geo.cl;
  const uint16_t cl_mask = (uint16_t)((1u << CL) - 1);
  const int ntiles = geo.tiles_per_cta;

  auto decode = [&](int it) -> TileCoord {
    TileCoord tcd;
    int b = (int)blockIdx.x + it * (int)gridDim.x;
    tcd.dummy = b >= geo.total_tiles;
    if (tcd.dummy) b = 0;
    const int tvi = b % tiles_v; b /= tiles_v;
    const int tui = b % tiles_u; b /= tiles_u;
    tcd.cb = b % ncb;
    tcd.n = b / ncb;
    tcd.ul0 = tui * TU;
    tcd.vl0 = tvi * TV;
    return tcd;
  };

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar_wfull + 8 * i, 1); mbar_init(bar_wempty + 8 * i, (uint32_t)CL); }
    for (int i = 0; i < NBUF * SLABS; ++i) { mbar_init(bar_infull + 8 * i, N_LOAD); mbar_init(bar_midfull + 8 * i, N_EPI); }
    for (int i = 0; i < NBUF; ++i) mbar_init(bar_buffree + 8 * i, 1);
    mbar_init(bar_acc1full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(bar_acc2full + 8 * i, 1); mbar_init(bar_acc2free + 8 * i, N_EPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (tracing) for (int i = 0; i < 32; ++i) trc[i] = 0;
    if (tracing) trc[0] = clock64();
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(K::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // every CTA's barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 248);
  const uint32_t acc1 = tmem;

  if (warp == W_PROD) {
    // ============================================================ weight producer
    const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(a.wstream_tc);
    const uint32_t slice = K::STAGE_BYTES / (uint32_t)CL;
    const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
    uint32_t k = 0;
    for (int it = 0; it < ntiles; ++it) {
      for (int g = 0; g < G; ++g, ++k) {
        const uint32_t st = k % NSTAGE;
        const long long tw0 = tracing ? clock64() : 0;
        if (k >= NSTAGE) mbar_wait(bar_wempty + 8 * st, ((k / NSTAGE) - 1) & 1);   // every CTA of the cluster released it
        if (tracing && lane == 0) trc[8] += clock64() - tw0;
        if (lane == 0) {
          mbar_expect_tx(bar_wfull + 8 * st, K::STAGE_BYTES);
          const uint32_t dst = ring + st * K::STAGE_BYTES + rank * slice;
          const unsigned char* src = wsrc + (size_t)g * K::STAGE_BYTES + (size_t)rank * slice;
          if (CL > 1) bulk_g2s_mc(dst, src, slice, bar_wfull + 8 * st, cl_mask);
          else bulk_g2s(dst, src, slice, bar_wfull + 8 * st);
        }
        __syncwarp();
      }
    }
  } else if (warp == W_MMA) {
    // ============================================================ MMA issuer (whole warp walks the loop, lane 0 issues)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(K::ACCW >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hiw = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_hiw = (512u >> 4) | (1u << 14) | (4u << 29);
    const uint32_t ring0 = ((ring & 0x3FFFF) >> 4) | (1u << 16);
    uint32_t k = 0;
    uint32_t ahi0 = 0, alo0 = 0;
    auto chunk = [&](int row0, int j, uint32_t acc, uint32_t accumulate) {
      const uint32_t st = k % NSTAGE;
      const long long tw0 = tracing ? clock64() : 0;
      mbar_wait(bar_wfull + 8 * st, (k / NSTAGE) & 1);
      tc_fence_after();
      if (tracing && lane == 0) trc[3] += clock64() - tw0;
      const uint32_t ad = ((uint32_t)(j >> 1) * K::SLAB_BYTES + (uint32_t)row0 * 128 + (uint32_t)(j & 1) * 64) >> 4;
      const uint32_t ah = ahi0 + ad, al = alo0 + ad;
      const uint32_t bh = ring0 + st * (K::STAGE_BYTES >> 4), bl = bh + (K::HALF_STAGE >> 4);
      if (lane == 0) {
        if (K::NSTACK) {     // B = [W_hi ; W_lo] (2C rows: the lo image follows the hi image at the same row pitch)
          mma_tf32_w(acc, ah, a_hiw, bh, b_hiw, idesc, accumulate);
          mma_tf32_w(acc, ah + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);   // second K step: +32 bytes
          mma_tf32_w(acc, al, a_hiw, bh, b_hiw, idesc, 1u);
          mma_tf32_w(acc, al + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
        } else {
          mma_tf32_w(acc, ah, a_hiw, bh, b_hiw, idesc, accumulate);
          mma_tf32_w(acc, al, a_hiw, bh, b_hiw, idesc, 1u);
          mma_tf32_w(acc, ah, a_hiw, bl, b_hiw, idesc, 1u);
          mma_tf32_w(acc, ah + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
          mma_tf32_w(acc, al + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
          mma_tf32_w(acc, ah + 2, a_hiw, bl + 2, b_hiw, idesc, 1u);
        }
        if (CL > 1) umma_commit_mc(bar_wempty + 8 * st, cl_mask);       // ring slot reusable when these retire
        else umma_commit(bar_wempty + 8 * st);
      }
      __syncwarp();
      ++k;
    };
    for (int it = 0; it < ntiles; ++it) {
      const int b = it % NBUF, use = it / NBUF, a2 = it & 1;
      const uint32_t act_hi = act0 + (uint32_t)b * K::BUF_BYTES, act_lo = act_hi + K::ACT_BYTES;
      ahi0 = ((act_hi & 0x3FFFF) >> 4) | (1u << 16);
      alo0 = ((act_lo & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t acc2 = tmem + K::ACCW + (uint32_t)a2 * K::ACCW;
      if (it >= 2) {   // epilogue 2 of tile it-2 has drained this accumulator
        const long long tw0 = tracing ? clock64() : 0;
        mbar_wait(bar_acc2free + 8 * a2, ((it >> 1) - 1) & 1);
        tc_fence_after();
        if (tracing && lane == 0) trc[4] += clock64() - tw0;
      }
      // ---- first conv (tap window = rows shifted by tap*RT) + adapter (centre pixels), slab by slab
      for (int j = 0; j < K::NCH; ++j) {
        if ((j & 1) == 0) {
          const long long tw0 = tracing ? clock64() : 0;
          mbar_wait(bar_infull + 8 * (b * SLABS + (j >> 1)), use & 1);
          tc_fence_after();
          if (tracing && lane == 0) trc[1] += clock64() - tw0;
        }
        for (int tap = 0; tap < 3; ++tap) chunk(tap * RT, j, acc1, (tap | j) != 0);
        if (NAD) chunk(RT + 1, j, acc2, j != 0);
      }
      if (lane == 0) umma_commit(bar_acc1full);
      __syncwarp();
      // ---- second conv over `mid` (tap window = rows shifted by tap), slab by slab
      for (int j = 0; j < K::NCH; ++j) {
        if ((j & 1) == 0) {
          const long long tw0 = tracing ? clock64() : 0;
          mbar_wait(bar_midfull + 8 * (b * SLABS + (j >> 1)), use & 1);
          tc_fence_after();
          if (tracing && lane == 0) trc[2] += clock64() - tw0;
        }
        for (int tap = 0; tap < 3; ++tap) chunk(tap, j, acc2, (NAD != 0) || (tap | j) != 0);
      }
      if (lane == 0) {
        umma_commit(bar_acc2full + 8 * a2);
        umma_commit(bar_buffree + 8 * b);     // every read of this operand buffer has retired
      }
      __syncwarp();
    }
  } else if (warp >= 8) {
    // ============================================================ loader warps
    const int lt = tid - N_EPI;
    const int c16 = lt & 7, rsub = lt >> 3;     // 16-byte chunk of the 128-byte slab row; row within a pass of 16
    int dec[LROWS];                              // tile-independent decomposition of this thread's rows
#pragma unroll
    for (int p = 0; p < LROWS; ++p) {
      const int row = p * 16 + rsub;
      dec[p] = -1;
      if (row < INROWS) {
        const int iu = row / RT, rem = row % RT;
        dec[p] = iu | ((rem / TVH) << 8) | ((rem % TVH) << 16);
      }
    }
    for (int it = 0; it < ntiles; ++it) {
      const int b = it % NBUF, use = it / NBUF;
      const TileCoord tcd = decode(it);
      const size_t img = (size_t)tcd.n * a.H * a.W * C;
      int pix[LROWS];
#pragma unroll
      for (int p = 0; p < LROWS; ++p) {
        pix[p] = -1;
        if (dec[p] >= 0 && !tcd.dummy) {
          const int iu = dec[p] & 0xFF, r = (dec[p] >> 8) & 0xFF, iv = dec[p] >> 16;
          const int cidx = tcd.cb * TR + r;
          const int ul = tcd.ul0 - 1 + iu, vl = tcd.vl0 - 1 + iv;
          const int u = (cidx / d) + d * ul, v = (cidx % d) + d * vl;
          if (cidx < d * d && ul >= 0 && vl >= 0 && u < U && v < V) pix[p] = a.vert_first ? u * a.W + v : v * a.W + u;
        }
      }
      // software pipeline over the slabs: the 10 loads of slab s+1 are in flight while slab s is split and stored;
      // the loads of the first slab are issued before the operand buffer is waited for
      float4 x[LROWS];
#pragma unroll
      for (int p = 0; p < LROWS; ++p) {
        x[p] = make4(0.f);
        if (pix[p] >= 0) x[p] = ldg4(a.in + img + (size_t)pix[p] * C + c16 * 4);
      }
      if (it >= NBUF) {
        const long long tw0 = tracing ? clock64() : 0;
        mbar_wait(bar_buffree + 8 * b, (use - 1) & 1);
        if (tracing && lt == 0) trc[7] += clock64() - tw0;
      }
      const long long tl0 = tracing ? clock64() : 0;
      unsigned char* hi_base = gen + K::HDR_BYTES + (size_t)b * K::BUF_BYTES;
#pragma unroll 1
      for (int slab = 0; slab < SLABS; ++slab) {
        const int ch = slab * 32 + c16 * 4;
        float4 xn[LROWS];
#pragma unroll
        for (int p = 0; p < LROWS; ++p) {
          xn[p] = make4(0.f);
          if (slab + 1 < SLABS && pix[p] >= 0) xn[p] = ldg4(a.in + img + (size_t)pix[p] * C + ch + 32);
        }
        float4 sc = make4(1.f), sh = make4(0.f);
        const bool pro = a.in_scale != nullptr;
        if (pro) { sc = ldg4(a.in_scale + ch); sh = ldg4(a.in_shift + ch); }
        unsigned char* slab_hi = hi_base + (size_t)slab * K::SLAB_BYTES;
#pragma unroll
        for (int p = 0; p < LROWS; ++p) {
          if (dec[p] < 0) continue;
          float4 v4 = x[p];
          if (pro && pix[p] >= 0) {
            v4.x = fmaxf(fmaf(v4.x, sc.x, sh.x), 0.f);
            v4.y = fmaxf(fmaf(v4.y, sc.y, sh.y), 0.f);
            v4.z = fmaxf(fmaf(v4.z, sc.z, sh.z), 0.f);
            v4.w = fmaxf(fmaf(v4.w, sc.w, sh.w), 0.f);
          }
          float4 hi, lo;
          split4(v4, hi, lo);
          const uint32_t off = sw128_off(p * 16 + rsub, c16);
          *reinterpret_cast<float4*>(slab_hi + off) = hi;
          *reinterpret_cast<float4*>(slab_hi + off + K::ACT_BYTES) = lo;
        }
        fence_proxy_async();
        mbar_arrive(bar_infull + 8 * (b * SLABS + slab));
#pragma unroll
        for (int p = 0; p < LROWS; ++p) x[p] = xn[p];
      }
      if (tracing && lt == 0) trc[12] += clock64() - tl0;
    }
  } else {
    // ============================================================ epilogue warps
    // TMEM is read with the 16x256b shape: lane l holds rows (l>>2) and (l>>2)+8 of a 16-lane half and columns
    // 2(l&3), 2(l&3)+1 of every 8-column group, i.e. four lanes own 32 contiguous bytes of an NHWC pixel row: epilogue
    // inputs are loaded and results stored with 8-byte accesses that fill whole 32-byte sectors, without staging
    // through shared memory (tools/tmem_probe.cu prints the fragment layout).
    const int q = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int lr = lane >> 2, lc = (lane & 3) * 2;             // row within a group of 8, first of this lane's 2 columns
    // this lane's four accumulator rows: m(h,k) = 32q + 16h + 8k + lr; their tile-independent decomposition
    int rdec[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = q * 32 + (r >> 1) * 16 + (r & 1) * 8 + lr;
      const int mu = m / RT, rem_m = m % RT;
      rdec[r] = (m < M1) ? (mu | ((rem_m / TVH) << 8) | ((rem_m % TVH) << 16)) : -1;
    }
    float* ssum = reinterpret_cast<float*>(gen + K::SUM_OFF);   // [2][C] per-CTA BatchNorm sums
    if (a.sums != nullptr) {
      for (int i = tid; i < 2 * C; i += N_EPI) ssum[i] = 0.f;
      named_bar_sync(1, N_EPI);
    }

    for (int it = 0; it < ntiles; ++it) {
      const int b = it % NBUF, a2 = it & 1;
      const TileCoord tcd = decode(it);
      const size_t img = (size_t)tcd.n * a.H * a.W * C;
      int pix_mid[4], pix_out[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        pix_mid[r] = -1; pix_out[r] = -1;
        if (rdec[r] >= 0 && !tcd.dummy) {
          const int mu = rdec[r] & 0xFF, rcls = (rdec[r] >> 8) & 0xFF, mv = rdec[r] >> 16;
          const int cidx = tcd.cb * TR + rcls;
          if (cidx < d * d) {
            const int ru = cidx / d, rv = cidx % d;
            const int u = ru + d * (tcd.ul0 + mu);
            const int vm = rv + d * (tcd.vl0 - 1 + mv), vo = rv + d * (tcd.vl0 + mv);
            if ((tcd.vl0 - 1 + mv) >= 0 && u < U && vm < V) pix_mid[r] = a.vert_first ? u * a.W + vm : vm * a.W + u;
            if (mv < TV && u < U && vo < V) pix_out[r] = a.vert_first ? u * a.W + vo : vo * a.W + u;
          }
        }
      }
      unsigned char* hi_base = gen + K::HDR_BYTES + (size_t)b * K::BUF_BYTES;
      const uint32_t acc2 = tmem + K::ACCW + (uint32_t)a2 * K::ACCW;
      // accumulator sub-block i of this warp (columns ch0 .. ch0+15) -> v[r][c][e]: row r (0..3), 8-column group c, column e
      auto load_acc = [&](uint32_t acc, int ch0, float (&v)[4][2][2]) {
        float t0[8], t1[8];
        if (K::NSTACK) {
          float u0[8], u1[8];
          tmem_ld_16x256b_x2(acc + lane_addr + ch0, t0);
          tmem_ld_16x256b_x2(acc + lane_addr + (16u << 16) + ch0, t1);
          tmem_ld_16x256b_x2(acc + lane_addr + C + ch0, u0);
          tmem_ld_16x256b_x2(acc + lane_addr + (16u << 16) + C + ch0, u1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) { t0[j] += u0[j]; t1[j] += u1[j]; }
        } else {
          tmem_ld_16x256b_x2(acc + lane_addr + ch0, t0);
          tmem_ld_16x256b_x2(acc + lane_addr + (16u << 16) + ch0, t1);
          tmem_ld_wait();
        }
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              v[k][c][e] = t0[4 * c + 2 * k + e];        // half 0: rows lr, lr+8
              v[2 + k][c][e] = t1[4 * c + 2 * k + e];    // half 1: rows 16+lr, 24+lr
            }
      };
      // 8-byte loads of this lane's elements of a [pixel][C] tensor for sub-block i (zero where the row has no pixel)
      auto fetch2 = [&](const float* base, const int (&pix)[4], int ch0, float2 (&dst)[4][2]) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            dst[r][c] = make_float2(0.f, 0.f);
            if (pix[r] >= 0) dst[r][c] = __ldg(reinterpret_cast<const float2*>(base + img + (size_t)pix[r] * C + ch0 + 8 * c + lc));
          }
      };

      // ================================================== epilogue 1: mid = f(acc1) -> hi/lo A operand (rows m)
      float2 pre[4][2], pre2[4][2];
      if (a.mid_mask != nullptr) fetch2(a.mid_mask, pix_mid, half * 16, pre);
      {
        const long long tw0 = tracing ? clock64() : 0;
        mbar_wait(bar_acc1full, it & 1);
        tc_fence_after();
        if (tracing && tid == 0) trc[5] += clock64() - tw0;
      }
      const long long te1 = tracing ? clock64() : 0;
#pragma unroll 1
      for (int i = 0; i < K::NSB; ++i) {
        const int ch0 = (2 * i + half) * 16;
        float v[4][2][2];
        float2 mk[4][2];
        if (a.mid_mask != nullptr) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) mk[r][c] = pre[r][c];
          if (i + 1 < K::NSB) fetch2(a.mid_mask, pix_mid, ch0 + 32, pre);
        }
        load_acc(acc1, ch0, v);
        float2 bb[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        if (a.mid_mask == nullptr && a.b1 != nullptr) {
          bb[0] = __ldg(reinterpret_cast<const float2*>(a.b1 + ch0 + lc));
          bb[1] = __ldg(reinterpret_cast<const float2*>(a.b1 + ch0 + 8 + lc));
        }
        unsigned char* slab_hi = hi_base + (size_t)(ch0 >> 5) * K::SLAB_BYTES;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int m = q * 32 + (r >> 1) * 16 + (r & 1) * 8 + lr;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float x0 = v[r][c][0], x1 = v[r][c][1];
            if (pix_mid[r] >= 0) {
              if (a.mid_mask != nullptr) { x0 = mk[r][c].x > 0.f ? x0 : 0.f; x1 = mk[r][c].y > 0.f ? x1 : 0.f; }
              else { x0 = fmaxf(x0 + bb[c].x, 0.f); x1 = fmaxf(x1 + bb[c].y, 0.f); }
            } else { x0 = 0.f; x1 = 0.f; }
            if (rdec[r] >= 0) {
              const float h0 = tf32_hi(x0), h1 = tf32_hi(x1);
              const int col = (ch0 & 31) + 8 * c + lc;                       // channel within the 32-channel slab
              const uint32_t off = (uint32_t)m * 128 + ((uint32_t)((col >> 2) ^ (m & 7)) << 4) + (uint32_t)(col & 3) * 4;
              *reinterpret_cast<float2*>(slab_hi + off) = make_float2(h0, h1);
              *reinterpret_cast<float2*>(slab_hi + off + K::ACT_BYTES) = make_float2(x0 - h0, x1 - h1);
            }
            if (a.mid_out != nullptr && pix_mid[r] >= 0)
              *reinterpret_cast<float2*>(a.mid_out + img + (size_t)pix_mid[r] * C + ch0 + 8 * c + lc) = make_float2(x0, x1);
          }
        }
        // both halves of every quadrant have written their 16 channels of slab i: the second conv may consume it
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(bar_midfull + 8 * (b * SLABS + i));
      }
      if (tracing && tid == 0) trc[10] += clock64() - te1;

      // ================================================== epilogue 2: out = acc2 + biases (+ mask / residual), sums
      if (a.epi != kEpiFwd) {
        fetch2(a.e0, pix_out, half * 16, pre);                                  // p  |  dy
        if (a.epi == kEpiBwdResidual) fetch2(a.e1, pix_out, half * 16, pre2);   //    |  y
      }
      {
        const long long tw0 = tracing ? clock64() : 0;
        mbar_wait(bar_acc2full + 8 * a2, (it >> 1) & 1);
        tc_fence_after();
        if (tracing && tid == 0) trc[6] += clock64() - tw0;
      }
      const long long te2 = tracing ? clock64() : 0;
#pragma unroll 1
      for (int i = 0; i < K::NSB; ++i) {
        const int ch0 = (2 * i + half) * 16;
        float v[4][2][2];
        float2 ev[4][2], ey[4][2];
        if (a.epi != kEpiFwd) {
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) { ev[r][c] = pre[r][c]; ey[r][c] = pre2[r][c]; }
          if (i + 1 < K::NSB) {
            fetch2(a.e0, pix_out, ch0 + 32, pre);
            if (a.epi == kEpiBwdResidual) fetch2(a.e1, pix_out, ch0 + 32, pre2);
          }
        }
        load_acc(acc2, ch0, v);
        if (i == K::NSB - 1) {     // last read of this accumulator: the MMA warp may overwrite it (tile it+2)
          tc_fence_before();
          mbar_arrive(bar_acc2free + 8 * a2);
        }
        float s1[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, s2[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int chc = ch0 + 8 * c + lc;
          float2 bb = make_float2(0.f, 0.f);
          if (a.b2 != nullptr) bb = __ldg(reinterpret_cast<const float2*>(a.b2 + chc));
          if (a.bad != nullptr) { const float2 b3 = __ldg(reinterpret_cast<const float2*>(a.bad + chc)); bb.x += b3.x; bb.y += b3.y; }
          float2 sc = make_float2(0.f, 0.f), sh = sc, mean = sc, istd = sc;
          if (a.epi == kEpiBwdMaskStats) {
            mean = __ldg(reinterpret_cast<const float2*>(a.e_stats + chc));
            istd = __ldg(reinterpret_cast<const float2*>(a.e_stats + C + chc));
            sc = __ldg(reinterpret_cast<const float2*>(a.e_stats + 2 * C + chc));
            sh = __ldg(reinterpret_cast<const float2*>(a.e_stats + 3 * C + chc));
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            float x0 = v[r][c][0] + bb.x, x1 = v[r][c][1] + bb.y;
            float w0 = 0.f, w1 = 0.f;      // second statistic's per-element factor
            if (a.epi == kEpiBwdMaskStats) {
              x0 = fmaf(ev[r][c].x, sc.x, sh.x) > 0.f ? x0 : 0.f;
              x1 = fmaf(ev[r][c].y, sc.y, sh.y) > 0.f ? x1 : 0.f;
              w0 = (ev[r][c].x - mean.x) * istd.x;
              w1 = (ev[r][c].y - mean.y) * istd.y;
            } else if (a.epi == kEpiBwdResidual) {
              x0 += ey[r][c].x > 0.f ? ev[r][c].x : 0.f;
              x1 += ey[r][c].y > 0.f ? ev[r][c].y : 0.f;
            }
            if (pix_out[r] >= 0) {
              *reinterpret_cast<float2*>(a.out + img + (size_t)pix_out[r] * C + chc) = make_float2(x0, x1);
              s1[c][0] += x0; s1[c][1] += x1;
              if (a.epi == kEpiFwd) { s2[c][0] = fmaf(x0, x0, s2[c][0]); s2[c][1] = fmaf(x1, x1, s2[c][1]); }
              else { s2[c][0] = fmaf(x0, w0, s2[c][0]); s2[c][1] = fmaf(x1, w1, s2[c][1]); }
            }
          }
        }
        // ---- per-channel sums over this warp's 32 rows: the 8 lanes with the same (lane & 3) own the same columns
        if (a.sums != nullptr) {
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float t1 = s1[c][e], t2 = s2[c][e];
#pragma unroll
              for (int o = 4; o < 32; o <<= 1) {
                t1 += __shfl_xor_sync(0xffffffffu, t1, o);
                t2 += __shfl_xor_sync(0xffffffffu, t2, o);
              }
              if (lr == 0) {
                atomicAdd(ssum + ch0 + 8 * c + lc + e, t1);
                atomicAdd(ssum + C + ch0 + 8 * c + lc + e, t2);
              }
            }
        }
      }
      if (tracing && tid == 0) trc[11] += clock64() - te2;
    }
    if (a.sums != nullptr) {
      named_bar_sync(1, N_EPI);
      for (int i = tid; i < 2 * C; i += N_EPI) atomicAdd(a.sums + i, (double)ssum[i]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tracing && tid == 0) {
    const long long tot = clock64() - trc[0];
    printf("pair_tc3<%d> CTA0: %d tiles (%d chunks each, CL=%d) in %lld clk | MMA waits: input %lld mid %lld weights %lld acc2free %lld | "
           "epilogue waits: acc1 %lld acc2 %lld, busy epi1 %lld epi2 %lld | loader: wait buffer %lld busy %lld | producer wait slots %lld\n",
           C, ntiles, G, CL, tot, trc[1], trc[2], trc[3], trc[4], trc[5], trc[6], trc[10], trc[11], trc[7], trc[12], trc[8]);
  }
  if (CL > 1) cluster_sync_all();   // no CTA exits while a peer may still multicast into it / arrive on its barriers
  if (warp == W_MMA) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(K::TMEM_COLS) : "memory");
  }
}

// Lattice tile (TU x TV mid pixels of TR residue classes, <= 128 mid rows, <= IN_MAX input rows) that needs the fewest
// rounds of the persistent grid, then the fewest tiles, then the smallest input tile.  Residue classes may be batched
// (TR > 1) whatever part of the class lattice the tile covers: for d = 8 at 64x128 (an 8 x 16 lattice per class) two
// classes x 8 x 6 fill a tile where one class x 8 x 14 leaves every second tile almost empty.
template <int C>
TileShape choose_tile(int Ul, int Vl, int d, int nimg) {
  TileShape best{1, 2, 1};
  long best_rounds = -1, best_tiles = 0, best_load = 0;
  for (int TR = 1; TR <= 8; ++TR) {
    if (TR > d * d) break;
    for (int TU = 1; TU <= 32; ++TU) {
      for (int TV = 2; TV <= 62; TV += 2) {
        const int TVH = TV + 2;
        if (TR * TU * TVH > 128) break;
        if (TR * (TU + 2) * TVH > IN_MAX) break;
        const long tiles = (long)nimg * cdiv(d * d, TR) * cdiv(Ul, TU) * cdiv(Vl, TV);
        const long rounds = (tiles + kNumSMs - 1) / kNumSMs;
        const long load = (long)TR * (TU + 2) * TVH;
        if (best_rounds < 0 || rounds < best_rounds || (rounds == best_rounds && (tiles < best_tiles || (tiles == best_tiles && load < best_load)))) {
          best_rounds = rounds; best_tiles = tiles; best_load = load; best = TileShape{TU, TV, TR};
        }
      }
    }
  }
  return best;
}

static int cluster_size() {
  static const int cl = [] {
    const char* e = getenv("MDIL_TC_CLUSTER");
    int v = e != nullptr ? atoi(e) : 2;
    return (v == 1 || v == 2 || v == 4) ? v : 2;
  }();
  return cl;
}

template <int C>
int launch_c(const PairArgs& a, cudaStream_t s) {
  using K = Cfg<C>;
  static_assert(K::SMEM_BYTES <= 227 * 1024, "tensor-core pair kernel shared memory budget");
  static_assert(8 * K::NSTAGE <= 64 && 8 * K::NBUF * K::SLABS <= 32, "barrier header layout");
  const int d = a.dil;
  const int U = a.vert_first ? a.H : a.W, V = a.vert_first ? a.W : a.H;
  const int Ul = cdiv(U, d), Vl = cdiv(V, d);
  const TileShape ts = choose_tile<C>(Ul, Vl, d, a.N);
  const long total = (long)a.N * cdiv(d * d, ts.TR) * cdiv(Ul, ts.TU) * cdiv(Vl, ts.TV);
  MDIL_REQUIRE(total > 0 && total < (1L << 30), "pair_tc3: tile count");
  MDIL_REQUIRE(a.wstream_tc != nullptr && ((uintptr_t)a.wstream_tc & 15) == 0, "pair_tc3: weight stream");
  const int cl = cluster_size();
  static std::atomic<int> max_ctas_dev[kMaxDevices];   // co-resident CTAs (1 per SM by shared memory), in whole clusters
  std::atomic<int>& max_ctas_slot = max_ctas_dev[current_device_slot()];
  int max_ctas = max_ctas_slot.load(std::memory_order_acquire);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = K::SMEM_BYTES;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_ctas == 0) {
    MDIL_CUDA(cudaFuncSetAttribute(pair_tc3_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES));
    int n = kNumSMs;
    if (cl > 1) {
      cfg.gridDim = dim3((unsigned)(kNumSMs / cl * cl), 1, 1);
      int ncl = 0;
      MDIL_CUDA(cudaOccupancyMaxActiveClusters(&ncl, pair_tc3_kernel<C>, &cfg));
      n = ncl * cl;
    }
    MDIL_REQUIRE(n >= cl, "pair_tc3: no co-resident cluster fits");
    if (n > kNumSMs) n = kNumSMs / cl * cl;
    max_ctas = n;
    max_ctas_slot.store(n, std::memory_order_release);
  }
  long grid = total < max_ctas ? (total + cl - 1) / cl * cl : max_ctas;
  Geo geo;
  geo.TU = ts.TU; geo.TV = ts.TV; geo.TR = ts.TR;
  geo.total_tiles = (int)total;
  geo.tiles_per_cta = (int)((total + grid - 1) / grid);
  geo.cl = cl;
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  MDIL_CUDA(cudaLaunchKernelEx(&cfg, pair_tc3_kernel<C>, a, geo));
  MDIL_LAUNCH_CHECK();
  return 0;
}

// fp32 [chunk][16][C] slabs of the FFMA stream -> hi/lo SWIZZLE_64B [C][16] images in the order the pipelined kernel
// consumes them: per 16-channel chunk j of the first conv its three taps then the adapter, then per chunk j of the
// second conv its three taps
__global__ void pack_tc3_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int has_adapter) {
  const int nch = C / KC;
  const int per1 = 3 + (has_adapter ? 1 : 0);
  const int G = nch * (per1 + 3);
  const long total = (long)G * C * KC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int nrow = (int)(i % C);
    const int kk = (int)((i / C) % KC);
    const int g = (int)(i / ((long)C * KC));
    int gs;
    if (g < nch * per1) {
      const int j = g / per1, slot = g % per1;
      gs = slot < 3 ? slot * nch + j : 6 * nch + j;
    } else {
      const int g2 = g - nch * per1;
      gs = 3 * nch + (g2 % 3) * nch + g2 / 3;
    }
    const float v = __ldg(src + ((long)gs * KC + kk) * C + nrow);
    const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
    const int off = nrow * 16 + ((((kk >> 2) ^ ((nrow >> 1) & 3)) << 2) | (kk & 3));
    float* stage = dst + (long)g * 2 * C * KC;
    stage[off] = hi;
    stage[C * KC + off] = lo;
  }
}

// ---- one launch packs everything a block's four pair launches read: the fp32 [tap][cin][cout] streams of the FFMA
// kernel (optional) and the hi/lo TF32 tensor-core images of this kernel (optional) straight
// from the PyTorch-layout weights (was ~14 launches per block and step)
struct PackSrc { const float* w[6]; };   // w31_1, w13_1, w31_2, w13_2, wp1, wp2  ([co][ci][3] / [co][ci])
__global__ void pack_block_kernel(const PackSrc src, float* __restrict__ packed, int C, int has_adapter, int write_fp32,
                                  int tc_order) {
  const int CC = C * C, nch = C / KC;
  const int per1 = 3 + (has_adapter ? 1 : 0);
  const long total = 4L * 7 * CC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int which = (int)(i / (7 * CC));
    const int r = (int)(i % (7 * CC));
    const int slab = r / CC, a = (r % CC) / C, b = r % C;
    if (slab == 6 && !has_adapter) continue;
    // which: 0 fwd pair 1, 1 fwd pair 2, 2 bwd pair 2, 3 bwd pair 1.  slab: 0..2 first conv taps, 3..5 second conv, 6 adapter
    const bool bwd = which >= 2;
    const int pairsel = (which == 0 || which == 3) ? 0 : 1;                  // weights of pair 1 or pair 2
    float v;
    if (slab == 6) {
      const float* w = src.w[4 + pairsel];
      v = bwd ? __ldg(w + a * C + b) : __ldg(w + b * C + a);                 // fwd [ci][co] = Wp[co][ci]; bwd [co][ci]
    } else {
      const int conv = slab / 3, tap = slab % 3;
      // forward: first conv = 3x1, second = 1x3; backward runs the pair in reverse: first = 1x3 (flipped), second = 3x1
      const int is13 = bwd ? (conv == 0) : (conv == 1);
      const float* w = src.w[pairsel * 2 + is13];
      v = bwd ? __ldg(w + ((long)a * C + b) * 3 + (2 - tap)) : __ldg(w + ((long)b * C + a) * 3 + tap);
    }
    if (write_fp32) packed[i] = v;
    if (tc_order != 0) {
      const int j = a / KC, kk = a % KC, nrow = b;
      int g;
      if (slab < 3) g = j * per1 + slab;
      else if (slab == 6) g = j * per1 + 3;
      else g = nch * per1 + j * 3 + (slab - 3);
      const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
      const int off = nrow * 16 + ((((kk >> 2) ^ ((nrow >> 1) & 3)) << 2) | (kk & 3));
      float* stage = packed + 28L * CC + (long)which * 14 * CC + (long)g * 2 * C * KC;
      stage[off] = hi;
      stage[C * KC + off] = lo;
    }
  }
}

}  // namespace tc3

int launch_pack_block(const float* const* w6, float* packed, int C, int has_adapter, int write_fp32, int tc_order,
                      cudaStream_t s) {
  tc3::PackSrc src;
  for (int i = 0; i < 6; ++i) src.w[i] = w6[i];
  const long total = 4L * 7 * C * C;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  tc3::pack_block_kernel<<<grid, 256, 0, s>>>(src, packed, C, has_adapter, write_fp32, tc_order);
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_pair_tc3(const PairArgs& a, cudaStream_t s) {
  switch (a.C) {
    case 128: return tc3::launch_c<128>(a, s);
    case 64: return tc3::launch_c<64>(a, s);
    default: return set_error(-2, "pair_tc3: C must be 64 or 128", __FILE__, __LINE__);
  }
}

int launch_pack_tc3(const float* src_stream, float* dst_stream, int C, int has_adapter, cudaStream_t s) {
  const long total = (long)(6 + (has_adapter ? 1 : 0)) * C * C;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  tc3::pack_tc3_kernel<<<grid, 256, 0, s>>>(src_stream, dst_stream, C, has_adapter);
  MDIL_LAUNCH_CHECK();
  return 0;
}

}  // namespace mdil
