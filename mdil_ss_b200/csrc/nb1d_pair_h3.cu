// Fused factorised-pair kernel of the non_bottleneck_1d(_RAP) block for C = 64 / 128 on the 5th-generation tensor cores
// with 16-bit split operands ("h3": x = hi + lo, hi*hi + lo*hi + hi*lo accumulated in fp32 TMEM; same contract as
// nb1d_pair.cu -- see the header comment there for the four uses; reference models/erfnet_RA_parallel.py:90-113, :48-64).
//
// Why this kernel replaced the 3xTF32 one (nb1d_pair_tc3.cu, kept for A/B as MDIL_PAIR_IMPL=tc3).  Its ncu source page
// and wait counters (profiles/r2_tc3_stall_analysis.md) show it bound by the MMA-issuing warp's own instruction stream
// (~130 clk per K=8 tcgen05.mma: waterfall loops around every instruction, per-chunk barrier traffic) and, next, by the
// epilogue warps (13.5k clk per tile of serial 8-byte pieces).  This kernel therefore
//   * splits fp32 operands into two 16-bit halves and uses kind::f16 (K = 16 per instruction: half the instructions and
//     half the operand bytes in shared memory).  Forward launches use fp16 halves (11 + 11 significand bits: the error of
//     3xTF32, ~1e-6, activations and weights of a BatchNorm network are far inside the fp16 range; conversions
//     saturate), backward launches use bf16 halves (gradients need the fp32 exponent range; 8 + 8 bits, ~1e-5, and no
//     ReLU decision is taken on backward values);
//   * issues MMAs from warp-uniform code (the warp index comes from a shuffle, operand descriptors are uniform-register
//     arithmetic on the shared-memory base): ptxas emits back-to-back UTCHMMA without R2UR / waterfall loops;
//   * keeps all C = 64 weights of a launch resident in shared memory (112 KB in 16-bit form; no ring, no per-chunk
//     barriers after the first tile); C = 128 streams 16 KB chunks through a 4-deep cp.async.bulk ring, multicast
//     across a 2-CTA cluster;
//   * double-buffers the activation operand for both C (the loader runs a whole tile ahead with every load of the tile
//     in flight);
//   * reads TMEM with the 32x32b shape (one thread = one pixel row): per-row index math once per tile, 256-bit global
//     accesses that fill whole sectors, 128-bit swizzled shared-memory stores of the `mid` operand, BatchNorm partial
//     sums by an in-register transpose-reduce (deterministic: no shared-memory float atomics).
//
// Warp roles (16 warps, one CTA per SM): 0..7 epilogues (TMEM lane quadrant = warp & 3, column half = warp >> 2),
// 8..13 loaders, 14 MMA issuer, 15 weight producer.  Measured alternative (MDIL_TC_TRACE=1 counters): 24 warps with the two
// epilogues on separate warp sets and 80 registers per thread -- slower (spills in every role, loaders starved).
#include <atomic>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mdil {
namespace h3 {

constexpr int IN_MAX_ALL = 160;   // largest input tile (rows = pixels) of any instantiation (pixel table size)
constexpr int N_EPI = 256;    // warps 0..7: epilogues
constexpr int N_LOAD = 192;   // warps 8..13
constexpr int W_LOAD0 = 8, W_MMA = 14, W_PROD = 15;
constexpr int NTHREADS = 512; // 16 warps = 4 per scheduler: 128 registers per thread
constexpr int KCH = 32;       // input channels per weight chunk (one 64-byte SWIZZLE_64B row of 16-bit elements)

template <int C> struct Cfg {
  // activation operand buffers: a buffer is busy from the first load of a tile until its second conv has retired (`mid`
  // overwrites the input in place).  (A third buffer for C = 64 at 144 input rows was measured: forward -4 %, backward +7 %
  // -- smaller tiles mean more epilogue work, and the backward launches are epilogue-bound -- so it is not used.)
  static constexpr int NBUF = 2;
  static constexpr int IN_MAX = 160;                           // input-tile rows (pixels) held in shared memory
  static constexpr int SLABS = C / 64;                         // 128-byte operand rows (64 channels) per pixel and image
  static constexpr uint32_t SLAB_BYTES = IN_MAX * 128;
  static constexpr uint32_t IMG_BYTES = SLABS * SLAB_BYTES;    // one (hi or lo) operand image
  static constexpr uint32_t BUF_BYTES = 2 * IMG_BYTES;
  static constexpr int NKC = C / KCH;                          // weight chunks per tap
  static constexpr uint32_t HALF_STAGE = C * 64;               // one (hi or lo) weight image of a chunk: C rows x 64 bytes
  static constexpr uint32_t STAGE_BYTES = 2 * HALF_STAGE;
  static constexpr bool RESIDENT = C == 64;                    // every chunk of the launch stays in shared memory
  static constexpr int NSTAGE = RESIDENT ? 7 * NKC : 4;
  static constexpr uint32_t HDR_BYTES = 3072;                  // barriers | tmem slot | biases | per-tile pixel table
  // the dynamic shared memory base is 1024-byte aligned (checked at run time): no alignment slack is carried
  static constexpr uint32_t SMEM_BYTES = HDR_BYTES + NBUF * BUF_BYTES + NSTAGE * STAGE_BYTES;
  static constexpr uint32_t ACCW = C;                          // accumulator width in TMEM columns
  static constexpr uint32_t TMEM_COLS = 4 * C;                 // acc1[2] + acc2[2]: 256 / 512 columns
  static constexpr int NPIECE = C / 32;                        // 16-column pieces per epilogue warp and phase
};

// header offsets (bytes from the 1024-aligned header base)
constexpr uint32_t OFF_WFULL = 0;        // [<= 14]
constexpr uint32_t OFF_WEMPTY = 128;     // [<= 4]
constexpr uint32_t OFF_INFULL = 160;     // [NBUF][NKC]: <= 8 (one per 32-channel chunk of the input tile)
constexpr uint32_t OFF_MIDFULL = 224;    // [NBUF][NKC]: <= 8
constexpr uint32_t OFF_BUFFREE = 288;    // [NBUF]: <= 3
constexpr uint32_t OFF_ACC1FULL = 320;   // [2]
constexpr uint32_t OFF_ACC2FULL = 336;   // [2]
constexpr uint32_t OFF_ACC2FREE = 352;   // [2]
constexpr uint32_t OFF_TMEMSLOT = 368;
constexpr uint32_t OFF_B1 = 512;         // float [C]
constexpr uint32_t OFF_B2 = 1024;        // float [C]: b2 + adapter bias
constexpr uint32_t OFF_TRACE = 2816;     // long long [30] (TRACE instantiations only)
constexpr uint32_t OFF_PIXTAB = 1536;    // int [2][IN_MAX_ALL]: pixel index of every input row of the tile being loaded (-1: zero)

struct Geo {   // per-launch tile geometry (kernel argument)
  int TU, TV, TR, TVH, RT, INROWS, M1;
  int U, V, dd, dshift;                  // dshift >= 0: d = 1 << dshift
  int tiles_u, tiles_v, ncb;
  int total_tiles, tiles_per_cta, cl;
  int sv, su, scb, sn;                   // mixed-radix digits of the grid size (tile index step per round)
};

struct TileShape { int TU, TV, TR; };

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
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred;
}
// D[tmem] (+)= A[smem] * B[smem]^T, 16-bit operands (kind::f16), descriptors given as (low word, constant high word)
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d), "r"(a_lo),
      "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
// 32 TMEM lanes x 16 columns: thread t of the warp gets lane (base + t), columns c .. c+15.  No wait.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// tcgen05.wait::ld with the loaded registers as read-write operands: no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :: "memory");
}
// 256-bit read-only load of tensors written by EARLIER kernels (never by this one).  volatile: a plain asm could be
// speculated above the null-pointer test that guards it (a faulting address is a side effect the compiler cannot see);
// no memory clobber: ordinary shared-memory loads may still be scheduled across it
__device__ __forceinline__ void ldg8(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
// 256-bit store of results nothing in this kernel reads back: volatile (must happen) but no memory clobber
__device__ __forceinline__ void stg8u(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void stg8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]));
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

// x0, x1 -> packed 16-bit halves (low half-word = x0): hi = round-to-nearest of x, lo = round-to-nearest of x - hi.
// FMT 0: fp16 (saturating), FMT 1: bf16.
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

// Sum v[0..16) over the 32 lanes of the warp, 16 shuffles: afterwards the lane holds the total of channel (lane >> 1) & 15
// (both lanes of a pair hold the same value).  Deterministic.
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

struct TileIter {   // (n, cb, tu, tv) of the CTA's current tile, advanced by the grid size without divisions
  int tv, tu, cb, n, b;
  __device__ __forceinline__ void init(const Geo& g) {
    b = (int)blockIdx.x;
    int t = b;
    tv = t % g.tiles_v; t /= g.tiles_v;
    tu = t % g.tiles_u; t /= g.tiles_u;
    cb = t % g.ncb;
    n = t / g.ncb;
  }
  __device__ __forceinline__ void next(const Geo& g) {
    b += (int)gridDim.x;
    tv += g.sv; if (tv >= g.tiles_v) { tv -= g.tiles_v; ++tu; }
    tu += g.su; if (tu >= g.tiles_u) { tu -= g.tiles_u; ++cb; }
    cb += g.scb; if (cb >= g.ncb) { cb -= g.ncb; ++n; }
    n += g.sn;
  }
  __device__ __forceinline__ bool dummy(const Geo& g) const { return b >= g.total_tiles; }
};

// TRACE: per-role wait/busy clock counters of CTA 0, printed at exit (MDIL_TC_TRACE=1; separate instantiation)
#define H3_T0() const long long _t0 = TRACE ? clock64() : 0
#define H3_T1(slot) do { if (TRACE && blockIdx.x == 0 && lane == 0 && trole >= 0) trc[(slot) * 6 + trole] += clock64() - _t0; } while (0)
template <int C, int FMT, bool TRACE>
__global__ void __launch_bounds__(NTHREADS, 1)
pair_h3_kernel(const __grid_constant__ PairArgs a, const __grid_constant__ Geo geo) {
  using K = Cfg<C>;
  const long long t_start = TRACE ? clock64() : 0;
  constexpr int NSTAGE = K::NSTAGE, NBUF = K::NBUF, SLABS = K::SLABS, NKC = K::NKC;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t hdr = smem_u32(smem_raw);
  if ((hdr & 1023u) != 0) __trap();              // SWIZZLE_128B operands need 1024-byte aligned slabs
  unsigned char* gen = smem_raw;                 // generic pointer to hdr
  const uint32_t act0 = hdr + K::HDR_BYTES;
  const uint32_t ring = act0 + NBUF * K::BUF_BYTES;
  float* b1s = reinterpret_cast<float*>(gen + OFF_B1);
  float* b2s = reinterpret_cast<float*>(gen + OFF_B2);
  int* pixtab = reinterpret_cast<int*>(gen + OFF_PIXTAB);
  long long* trc = reinterpret_cast<long long*>(gen + OFF_TRACE);   // [5 slots][6 traced warps], TRACE only
  if (TRACE && threadIdx.x < 30) trc[threadIdx.x] = 0;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
  const int RT = geo.RT, TVH = geo.TVH;
  const int d = a.dil;
  const int NAD = a.has_adapter ? 1 : 0;
  const int CL = geo.cl;
  const uint16_t cl_mask = (uint16_t)((1u << CL) - 1);
  const int ntiles = geo.tiles_per_cta;
  // traced warps: epilogue warps 0 and 7, loaders 8, 10 and 13, the MMA warp
  const int trole = warp == 0 ? 0 : warp == 7 ? 1 : warp == 8 ? 2 : warp == 10 ? 3 : warp == 13 ? 4 : warp == W_MMA ? 5 : -1;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) mbar_init(hdr + OFF_WFULL + 8 * i, 1);
    if (!K::RESIDENT) for (int i = 0; i < NSTAGE; ++i) mbar_init(hdr + OFF_WEMPTY + 8 * i, (uint32_t)CL);
    for (int i = 0; i < NBUF * NKC; ++i) mbar_init(hdr + OFF_INFULL + 8 * i, N_LOAD);
    for (int i = 0; i < NBUF * NKC; ++i) mbar_init(hdr + OFF_MIDFULL + 8 * i, 128);   // the 4 warps that own the chunk
    for (int i = 0; i < NBUF; ++i) mbar_init(hdr + OFF_BUFFREE + 8 * i, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(hdr + OFF_ACC1FULL + 8 * i, 1);
      mbar_init(hdr + OFF_ACC2FULL + 8 * i, 1);
      mbar_init(hdr + OFF_ACC2FREE + 8 * i, N_EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < C) {
    b1s[tid] = a.b1 != nullptr ? __ldg(a.b1 + tid) : 0.f;
    b2s[tid] = (a.b2 != nullptr ? __ldg(a.b2 + tid) : 0.f) + (a.bad != nullptr ? __ldg(a.bad + tid) : 0.f);
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hdr + OFF_TMEMSLOT), "r"(K::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // every CTA's barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + OFF_TMEMSLOT);
  // TMEM: first-conv accumulators [2][C] then second-conv accumulators [2][C]; tile t uses index t & 1 of both

  // Software pipeline (all roles use the same static order): conv 1 of tile t+1 is issued BEFORE conv 2 of tile t, so
  // the tensor pipe works on tile t+1 while the epilogue warps turn the first accumulator of tile t into `mid`:
  //     MMA      : c1(0) | c1(1) c2(0) | c1(2) c2(1) | ...
  //     epilogue : e1(0) | e1(1) e2(0) | e1(2) e2(1) | ...
  if (warp == W_PROD) {
    // ============================================================ weight producer
    const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(a.wstream_tc);
    const int G1 = NKC * (3 + NAD), G = NKC * (6 + NAD);
    if (K::RESIDENT) {
      if (lane == 0) {
        for (int g = 0; g < G; ++g) {
          mbar_expect_tx(hdr + OFF_WFULL + 8 * g, K::STAGE_BYTES);
          bulk_g2s(ring + g * K::STAGE_BYTES, wsrc + (size_t)g * K::STAGE_BYTES, K::STAGE_BYTES, hdr + OFF_WFULL + 8 * g);
        }
      }
    } else {
      const uint32_t slice = K::STAGE_BYTES / (uint32_t)CL;
      const uint32_t rank = CL > 1 ? cluster_ctarank() : 0;
      uint32_t k = 0;
      auto fetch = [&](int g0, int g1) {
        for (int g = g0; g < g1; ++g, ++k) {
          const uint32_t st = k % NSTAGE;
          if (k >= (uint32_t)NSTAGE) mbar_wait(hdr + OFF_WEMPTY + 8 * st, ((k / NSTAGE) - 1) & 1);   // every CTA of the cluster released it
          if (lane == 0) {
            mbar_expect_tx(hdr + OFF_WFULL + 8 * st, K::STAGE_BYTES);
            const uint32_t dst = ring + st * K::STAGE_BYTES + rank * slice;
            const unsigned char* src = wsrc + (size_t)g * K::STAGE_BYTES + (size_t)rank * slice;
            if (CL > 1) bulk_g2s_mc(dst, src, slice, hdr + OFF_WFULL + 8 * st, cl_mask);
            else bulk_g2s(dst, src, slice, hdr + OFF_WFULL + 8 * st);
          }
          __syncwarp();
        }
      };
      fetch(0, G1);
      for (int it = 0; it < ntiles; ++it) {
        if (it + 1 < ntiles) fetch(0, G1);
        fetch(G1, G);
      }
    }
  } else if (warp == W_MMA) {
    // ============================================================ MMA issuer (warp-uniform control flow, one elected lane issues)
    constexpr uint32_t FB = FMT == 0 ? 0u : 1u;     // operand format: F16 = 0, BF16 = 1
    const uint32_t idesc = (1u << 4) | (FB << 7) | (FB << 10) | ((uint32_t)(K::ACCW >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hiw = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024, version 1, SWIZZLE_128B
    const uint32_t b_hiw = (512u >> 4) | (1u << 14) | (4u << 29);       // SBO 512, version 1, SWIZZLE_64B
    const uint32_t ring0 = ((ring & 0x3FFFF) >> 4) | (1u << 16);
    const uint32_t rt16 = (uint32_t)RT * 8u;                            // RT rows of 128 bytes, in 16-byte units
    const int G1 = NKC * (3 + NAD);
    uint32_t kring = 0;                                                 // ring position (streamed weights)
    bool first_c1 = true, first_c2 = true;                              // resident weights: barriers are waited for once
    // one weight chunk (32 input channels, two K = 16 steps) of one tap: operand rows shifted by row16 (16-byte units)
    auto chunk = [&](uint32_t ahi0, uint32_t alo0, uint32_t row16, int j, int g, bool first, uint32_t acc, uint32_t accumulate) {
      uint32_t st;
      if (K::RESIDENT) {
        st = (uint32_t)g;
        if (first) { mbar_wait(hdr + OFF_WFULL + 8 * st, 0); tc_fence_after(); }
      } else {
        st = kring % NSTAGE;
        { H3_T0(); mbar_wait(hdr + OFF_WFULL + 8 * st, (kring / NSTAGE) & 1); H3_T1(2); }
        tc_fence_after();
      }
      const uint32_t ad = (uint32_t)(j >> 1) * (K::SLAB_BYTES >> 4) + row16 + (uint32_t)(j & 1) * 4u;
      const uint32_t ah = ahi0 + ad, al = alo0 + ad;
      const uint32_t bh = ring0 + st * (K::STAGE_BYTES >> 4), bl = bh + (K::HALF_STAGE >> 4);
      if (elect_one()) {
        mma_f16(acc, ah, a_hiw, bh, b_hiw, idesc, accumulate);        // x_hi * W_hi
        mma_f16(acc, al, a_hiw, bh, b_hiw, idesc, 1u);                // x_lo * W_hi
        mma_f16(acc, ah, a_hiw, bl, b_hiw, idesc, 1u);                // x_hi * W_lo
        mma_f16(acc, ah + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);        // second K step: +32 bytes
        mma_f16(acc, al + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
        mma_f16(acc, ah + 2, a_hiw, bl + 2, b_hiw, idesc, 1u);
        if (!K::RESIDENT) {
          if (CL > 1) umma_commit_mc(hdr + OFF_WEMPTY + 8 * st, cl_mask);   // ring slot reusable when these retire
          else umma_commit(hdr + OFF_WEMPTY + 8 * st);
        }
      }
      __syncwarp();
      ++kring;
    };
    // first conv of tile t (tap window = rows shifted by tap*RT) + adapter (centre pixels -> second accumulator)
    auto conv1 = [&](int t) {
      const int b = t % NBUF, use = t / NBUF;        // operand buffer and how often it has been used
      const int ab = t & 1, ause = t >> 1;           // accumulator pair
      const uint32_t act_hi = act0 + (uint32_t)b * K::BUF_BYTES;
      const uint32_t ahi0 = ((act_hi & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t alo0 = (((act_hi + K::IMG_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t acc1 = tmem + (uint32_t)ab * K::ACCW, acc2 = tmem + 2 * K::ACCW + (uint32_t)ab * K::ACCW;
      // the three taps first (first accumulator), the adapter chunks last: they write the SECOND accumulator, which
      // epilogue 2 of tile t-2 may still be draining -- only this last quarter of the conv waits for it
      int g = 0;
#pragma unroll
      for (int j = 0; j < NKC; ++j) {
        {   // the loader fills the tile one 32-channel chunk at a time: the conv starts on the first chunk
          H3_T0();
          mbar_wait(hdr + OFF_INFULL + 8 * (b * NKC + j), (uint32_t)(use & 1));
          H3_T1(0);
          tc_fence_after();
        }
#pragma unroll
        for (int tap = 0; tap < 3; ++tap) chunk(ahi0, alo0, (uint32_t)tap * rt16, j, g++, first_c1, acc1, (uint32_t)((tap | j) != 0));
      }
      if (NAD) {
        if (t >= 2) {
          H3_T0();
          mbar_wait(hdr + OFF_ACC2FREE + 8 * ab, (uint32_t)((ause - 1) & 1));
          H3_T1(3);
          tc_fence_after();
        }
#pragma unroll
        for (int j = 0; j < NKC; ++j) chunk(ahi0, alo0, rt16 + 8u, j, g++, first_c1, acc2, (uint32_t)(j != 0));
      }
      if (elect_one()) umma_commit(hdr + OFF_ACC1FULL + 8 * ab);
      __syncwarp();
      first_c1 = false;
    };
    // second conv of tile t over `mid` (tap window = rows shifted by tap)
    auto conv2 = [&](int t) {
      const int b = t % NBUF, use = t / NBUF;
      const int ab = t & 1, ause = t >> 1;
      const uint32_t act_hi = act0 + (uint32_t)b * K::BUF_BYTES;
      const uint32_t ahi0 = ((act_hi & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t alo0 = (((act_hi + K::IMG_BYTES) & 0x3FFFF) >> 4) | (1u << 16);
      const uint32_t acc2 = tmem + 2 * K::ACCW + (uint32_t)ab * K::ACCW;
      if (!NAD && t >= 2) {
        H3_T0();
        mbar_wait(hdr + OFF_ACC2FREE + 8 * ab, (uint32_t)((ause - 1) & 1));
        H3_T1(3);
        tc_fence_after();
      }
      int g = G1;
#pragma unroll
      for (int j = 0; j < NKC; ++j) {
        { H3_T0(); mbar_wait(hdr + OFF_MIDFULL + 8 * (b * NKC + j), (uint32_t)(use & 1)); H3_T1(1); }
        tc_fence_after();
#pragma unroll
        for (int tap = 0; tap < 3; ++tap) chunk(ahi0, alo0, (uint32_t)tap * 8u, j, g++, first_c2, acc2, (uint32_t)((NAD != 0) || (tap | j) != 0));
      }
      if (elect_one()) {
        umma_commit(hdr + OFF_ACC2FULL + 8 * ab);
        umma_commit(hdr + OFF_BUFFREE + 8 * b);     // every read of this operand buffer has retired
      }
      __syncwarp();
      first_c2 = false;
    };
    conv1(0);
    for (int it = 0; it < ntiles; ++it) {
      if (it + 1 < ntiles) conv1(it + 1);
      conv2(it);
    }
  } else if (warp >= W_LOAD0) {
    // ============================================================ loader warps (8..13)
    // Per tile: (1) thread lt < INROWS computes the pixel index of input row lt once (-1: outside the image = zero
    // padding) into a double-buffered shared table; (2) item = (row, 4 channels of one 32-channel chunk); a batch = one
    // chunk of the whole tile; the batches of all tiles form one stream in which the loads of batch g+1 are issued before
    // batch g is converted (register double buffer), across tile boundaries too; (3) fp32 -> BN+ReLU prologue -> 16-bit
    // hi/lo halves -> swizzled K-major operand rows; every chunk has its own "input full" barrier.
    constexpr int IN_MAX = K::IN_MAX;
    const int lt = tid - W_LOAD0 * 32;
    // CHUNK-MAJOR order: a batch = one 32-channel chunk of the whole tile (8 threads x 128-bit per row, 24 rows per pass,
    // 7 passes), so that the first conv can start on chunk 0 while the later chunks are still being loaded
    constexpr int RPP = N_LOAD / 8;           // rows per pass: 24
    constexpr int PB = (IN_MAX + RPP - 1) / RPP;   // passes per batch = per chunk: 7
    constexpr int NB = NKC;                   // batches per tile
    const int c4 = lt & 7, rsub = lt >> 3;    // 4-channel group inside the chunk, row inside the pass
    const uint32_t half8 = (uint32_t)(c4 & 1) * 8u;
    const bool pro = a.in_scale != nullptr;
    int my_iu = 0, my_r = 0, my_iv = 0;          // tile-independent decomposition of input row lt: [iu][class][iv]
    const bool my_row = lt < geo.INROWS;
    if (my_row) { my_iu = lt / RT; const int rem = lt % RT; my_r = rem / TVH; my_iv = rem % TVH; }
    TileIter ti;                                 // tile of the batch being ISSUED
    ti.init(geo);
    const float* src = nullptr;
    const bool in_split = a.in_split != 0;       // S16 input: every 16 bytes hold 4 bf16 hi halves + 4 bf16 lo halves
    // start of tile t on the issue side: pixel table of the tile (-1: outside the image = zero padding) + source pointer.
    // The table is double-buffered by t & 1: a thread writing table t+2 has passed the barrier of tile t+1, which every
    // thread reaches only after its last read of table t.
    auto new_tile = [&](int t) {
      if (lt < IN_MAX) {
        int pix = -1;
        if (my_row && !ti.dummy(geo)) {
          const int cidx = ti.cb * geo.TR + my_r;
          const int ul = ti.tu * geo.TU - 1 + my_iu, vl = ti.tv * geo.TV - 1 + my_iv;
          const int ru = geo.dshift >= 0 ? (cidx >> geo.dshift) : cidx / d;
          const int rv = cidx - ru * d;
          const int u = ru + d * ul, v = rv + d * vl;
          if (cidx < geo.dd && ul >= 0 && vl >= 0 && u < geo.U && v < geo.V) pix = a.vert_first ? u * a.W + v : v * a.W + u;
        }
        pixtab[(t & 1) * IN_MAX_ALL + lt] = pix;
        if (a.mid_mask != nullptr && pix >= 0) {
          // backward launches: the saved activation whose sign masks `mid` is read by epilogue 1 of this tile, about one
          // tile from now, one 16-column piece ahead of its use: pull the pixel's row into L2 already (the mid pixels are
          // the input pixels minus the halo along the first conv's axis)
          const float* mrow_p = a.mid_mask + (size_t)ti.n * a.H * a.W * C + (size_t)pix * C;
#pragma unroll
          for (int l = 0; l < C / 32; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(mrow_p + l * 32));
        }
      }
      { H3_T0(); named_bar_sync(2, N_LOAD); H3_T1(3); }
      src = a.in + (size_t)ti.n * a.H * a.W * C + c4 * 4;      // + 32 * chunk at issue time
    };
    auto issue = [&](float4 (&x)[PB], uint32_t& valid, int g) {
      const int t = g / NB, bt = g - t * NB;
      if (bt == 0) { if (t > 0) ti.next(geo); new_tile(t); }
      const int* ptab = pixtab + (t & 1) * IN_MAX_ALL;
      valid = 0;
#pragma unroll
      for (int p = 0; p < PB; ++p) {
        const int row = p * RPP + rsub;            // every batch walks all rows of the tile
        x[p] = make4(0.f);
        if (row < IN_MAX) {
          const int pix = ptab[row];
          if (pix >= 0) {
            x[p] = ldg4(src + (size_t)pix * C + bt * KCH);    // S16 input: the same 16 bytes hold 4 hi + 4 lo halves
            valid |= 1u << p;
          }
        }
      }
    };
    auto convert = [&](const float4 (&x)[PB], uint32_t valid, int g) {
      const int t = g / NB, bt = g - t * NB;
      const int b = t % NBUF, use = t / NBUF;
      if (bt == 0 && t >= NBUF) { H3_T0(); mbar_wait(hdr + OFF_BUFFREE + 8 * b, (uint32_t)((use - 1) & 1)); H3_T1(0); }
      unsigned char* buf = gen + K::HDR_BYTES + (size_t)b * K::BUF_BYTES + (size_t)(bt >> 1) * K::SLAB_BYTES;
      const uint32_t chunk16 = (uint32_t)((bt & 1) * 4 + (c4 >> 1));     // 16-byte chunk of the 128-byte operand row
      float4 sc = make4(1.f), sh = make4(0.f);
      if (pro) { sc = ldg4(a.in_scale + bt * KCH + c4 * 4); sh = ldg4(a.in_shift + bt * KCH + c4 * 4); }
#pragma unroll
      for (int p = 0; p < PB; ++p) {       // branch-free: the seven items interleave in the instruction stream
        const int row = p * RPP + rsub;            // every batch walks all rows of the tile
        float4 v4 = x[p];
        if (pro) {
          const bool ok = (valid >> p) & 1u;       // padding rows stay zero
          v4.x = ok ? fmaxf(fmaf(v4.x, sc.x, sh.x), 0.f) : 0.f;
          v4.y = ok ? fmaxf(fmaf(v4.y, sc.y, sh.y), 0.f) : 0.f;
          v4.z = ok ? fmaxf(fmaf(v4.z, sc.z, sh.z), 0.f) : 0.f;
          v4.w = ok ? fmaxf(fmaf(v4.w, sc.w, sh.w), 0.f) : 0.f;
        }
        uint2 hi, lo;
        if (in_split) {
          hi = make_uint2(__float_as_uint(v4.x), __float_as_uint(v4.y));
          lo = make_uint2(__float_as_uint(v4.z), __float_as_uint(v4.w));
        } else {
          split2<FMT>(v4.x, v4.y, hi.x, lo.x);
          split2<FMT>(v4.z, v4.w, hi.y, lo.y);
        }
        const uint32_t off = (uint32_t)row * 128u + (((chunk16 ^ ((uint32_t)row & 7u)) << 4) | half8);
        if (row < geo.INROWS) {
          *reinterpret_cast<uint2*>(buf + off) = hi;
          *reinterpret_cast<uint2*>(buf + off + K::IMG_BYTES) = lo;
        }
      }
      fence_proxy_async();
      mbar_arrive(hdr + OFF_INFULL + 8 * (b * NKC + bt));
    };
    const int nbatch = ntiles * NB;
    float4 xa[PB], xb[PB];
    uint32_t va = 0, vb = 0;
    issue(xa, va, 0);
    for (int g = 0; g < nbatch; g += 2) {
      if (g + 1 < nbatch) { H3_T0(); issue(xb, vb, g + 1); H3_T1(1); }
      { H3_T0(); convert(xa, va, g); H3_T1(2); }
      if (g + 1 >= nbatch) break;
      if (g + 2 < nbatch) { H3_T0(); issue(xa, va, g + 2); H3_T1(1); }
      { H3_T0(); convert(xb, vb, g + 1); H3_T1(2); }
    }
  } else {
    // ============================================================ epilogue warps
    // TMEM is read with the 32x32b shape: lane l of warp (q, half) owns accumulator row m = 32q + l and, per piece, 16
    // consecutive channels: 64 bytes of an NHWC pixel row (two 256-bit accesses), 32 bytes of an operand row per image.
    const int q = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int m = q * 32 + lane;
    int mu = 0, rcls = 0, mv = 0;
    const bool mrow = m < geo.M1;
    if (mrow) { mu = m / RT; const int rem = m % RT; rcls = rem / TVH; mv = rem % TVH; }
    constexpr int NPIECE = K::NPIECE;
    const int chw = half * (C / 2);                 // first channel of this warp
    float run1[NPIECE], run2[NPIECE];               // running BatchNorm partial sums (channel chw + 16 pc + ((lane >> 1) & 15))
#pragma unroll
    for (int i = 0; i < NPIECE; ++i) { run1[i] = 0.f; run2[i] = 0.f; }
    const uint32_t sw = (uint32_t)m & 7u;
    const bool has_mask = a.mid_mask != nullptr;
    const bool mid_split = a.mid_out != nullptr && a.mid_out_split != 0;

    // pixel of accumulator row m in tile `ti`: the `mid` pixel (second-conv input, halo included) or the output pixel
    auto row_pixel = [&](const TileIter& ti, bool out) -> long {
      if (!mrow || ti.dummy(geo)) return -1;
      const int cidx = ti.cb * geo.TR + rcls;
      if (cidx >= geo.dd) return -1;
      const int ru = geo.dshift >= 0 ? (cidx >> geo.dshift) : cidx / d;
      const int rv = cidx - ru * d;
      const int u = ru + d * (ti.tu * geo.TU + mu);
      const int vlm = ti.tv * geo.TV - 1 + mv;
      const int vm = rv + d * vlm;
      if (!out) {
        if (vlm >= 0 && u < geo.U && vm < geo.V) return a.vert_first ? (long)u * a.W + vm : (long)vm * a.W + u;
        return -1;
      }
      const int vo = vm + d;
      if (mv < geo.TV && u < geo.U && vo < geo.V) return a.vert_first ? (long)u * a.W + vo : (long)vo * a.W + u;
      return -1;
    };

    // ================================================== epilogue 1: mid = f(acc1) -> hi/lo A operand (row m)
    // one 16-column piece at a time; the global loads of the NEXT piece's inputs are issued as soon as the current
    // piece's have been consumed, so they fly under its conversion and stores (six warps per scheduler hide the rest)
    auto epi1 = [&](int t, const TileIter& ti) {
      const int b = t % NBUF;                        // operand buffer
      const int ab = t & 1, ause = t >> 1;           // accumulator pair
      const size_t img = (size_t)ti.n * a.H * a.W * C;
      const long pix_mid = row_pixel(ti, false);
      unsigned char* buf = gen + K::HDR_BYTES + (size_t)b * K::BUF_BYTES;
      const uint32_t acc1 = tmem + (uint32_t)ab * K::ACCW;
      const float* mid_row = has_mask && pix_mid >= 0 ? a.mid_mask + img + (size_t)pix_mid * C + chw : nullptr;
      float* mid_dst = a.mid_out != nullptr && pix_mid >= 0 ? a.mid_out + img + (size_t)pix_mid * C + chw : nullptr;
      float mk[16];
      if (has_mask) {
#pragma unroll
        for (int i = 0; i < 16; ++i) mk[i] = 0.f;
        if (mid_row != nullptr) { ldg8(mid_row, &mk[0]); ldg8(mid_row + 8, &mk[8]); }
      }
      if (a.epi != kEpiFwd) {
        // the rows epilogue 2 of this tile will read (p / dy, y / block input): pull them into L2 now, one epilogue
        // earlier -- its register prefetch runs only one 16-column piece ahead, which covers an L2 hit, not DRAM
        const long pix_out = row_pixel(ti, true);
        if (pix_out >= 0) {
          const float* e0p = a.e0 + img + (size_t)pix_out * C + chw;
#pragma unroll
          for (int l = 0; l < C / 64; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(e0p + l * 32));
          if (a.epi == kEpiBwdResidual) {
            const float* e1p = a.e1 + img + (size_t)pix_out * C + chw;
#pragma unroll
            for (int l = 0; l < C / 64; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(e1p + l * 32));
          }
        }
      }
      { H3_T0(); mbar_wait(hdr + OFF_ACC1FULL + 8 * ab, (uint32_t)(ause & 1)); H3_T1(0); }
      tc_fence_after();
#pragma unroll 1
      for (int pc = 0; pc < NPIECE; ++pc) {
        const int ch0 = chw + pc * 16;
        uint32_t r[16];
        tmem_ld16(acc1 + lane_addr + (uint32_t)ch0, r);
        tmem_wait16(r);
        float x[16];
        const bool in_mid = pix_mid >= 0;          // rows outside the image feed zero padding to the second conv
        if (has_mask) {
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] = (in_mid && mk[i] > 0.f) ? __uint_as_float(r[i]) : 0.f;
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(b1s + ch0 + i);
            x[i] = in_mid ? fmaxf(__uint_as_float(r[i]) + bb.x, 0.f) : 0.f;
            x[i + 1] = in_mid ? fmaxf(__uint_as_float(r[i + 1]) + bb.y, 0.f) : 0.f;
            x[i + 2] = in_mid ? fmaxf(__uint_as_float(r[i + 2]) + bb.z, 0.f) : 0.f;
            x[i + 3] = in_mid ? fmaxf(__uint_as_float(r[i + 3]) + bb.w, 0.f) : 0.f;
          }
        }
        if (has_mask && pc + 1 < NPIECE && mid_row != nullptr) { ldg8(mid_row + (pc + 1) * 16, &mk[0]); ldg8(mid_row + (pc + 1) * 16 + 8, &mk[8]); }
        if (mid_dst != nullptr && !mid_split) {
          stg8(mid_dst + pc * 16, &x[0]);
          stg8(mid_dst + pc * 16 + 8, &x[8]);
        }
        if (mrow) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) split2<FMT>(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
          if (mid_split && mid_dst != nullptr) {      // S16 `mid_out`: per 4 channels, the 2 hi words then the 2 lo words
            const uint32_t w0[8] = {hi[0], hi[1], lo[0], lo[1], hi[2], hi[3], lo[2], lo[3]};
            const uint32_t w1[8] = {hi[4], hi[5], lo[4], lo[5], hi[6], hi[7], lo[6], lo[7]};
            stg8u(mid_dst + pc * 16, w0);
            stg8u(mid_dst + pc * 16 + 8, w1);
          }
          unsigned char* rowp = buf + (size_t)(ch0 >> 6) * K::SLAB_BYTES + (size_t)m * 128;
          const uint32_t c16 = (uint32_t)(ch0 & 63) >> 3;       // first of the two 16-byte chunks of this piece
          const uint32_t o0 = ((c16 ^ sw) << 4), o1 = (((c16 + 1) ^ sw) << 4);
          *reinterpret_cast<uint4*>(rowp + o0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(rowp + o1) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          *reinterpret_cast<uint4*>(rowp + K::IMG_BYTES + o0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<uint4*>(rowp + K::IMG_BYTES + o1) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        }
        if (pc & 1) {   // the 32-channel weight-chunk column range of `mid` ending here is complete for this warp's rows
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive(hdr + OFF_MIDFULL + 8 * (b * NKC + (ch0 >> 5)));
        }
      }
    };

    // ================================================== epilogue 2: out = acc2 + biases (+ mask / residual), sums
    auto epi2 = [&](int t, const TileIter& ti) {
      const int ab = t & 1, ause = t >> 1;           // accumulator pair
      const size_t img = (size_t)ti.n * a.H * a.W * C;
      const long pix_out = row_pixel(ti, true);
      const uint32_t acc2 = tmem + 2 * K::ACCW + (uint32_t)ab * K::ACCW;
      const float* e0_row = a.epi != kEpiFwd && pix_out >= 0 ? a.e0 + img + (size_t)pix_out * C + chw : nullptr;
      const float* e1_row = a.epi == kEpiBwdResidual && pix_out >= 0 ? a.e1 + img + (size_t)pix_out * C + chw : nullptr;
      float* out_dst = pix_out >= 0 ? a.out + img + (size_t)pix_out * C + chw : nullptr;
      float ev[16];          // p (mask + statistics) or dy * (y > 0) (residual) of the piece in flight
      auto fetch_e = [&](int pc) {
        if (a.epi == kEpiFwd) return;
#pragma unroll
        for (int i = 0; i < 16; ++i) ev[i] = 0.f;
        if (e0_row != nullptr) { ldg8(e0_row + pc * 16, &ev[0]); ldg8(e0_row + pc * 16 + 8, &ev[8]); }
        if (e1_row != nullptr) {
          float y8[8];
          ldg8(e1_row + pc * 16, y8);
#pragma unroll
          for (int i = 0; i < 8; ++i) ev[i] = y8[i] > 0.f ? ev[i] : 0.f;
          ldg8(e1_row + pc * 16 + 8, y8);
#pragma unroll
          for (int i = 0; i < 8; ++i) ev[8 + i] = y8[i] > 0.f ? ev[8 + i] : 0.f;
        }
      };
      fetch_e(0);
      { H3_T0(); mbar_wait(hdr + OFF_ACC2FULL + 8 * ab, (uint32_t)(ause & 1)); H3_T1(1); }
      tc_fence_after();
#pragma unroll 1
      for (int pc = 0; pc < NPIECE; ++pc) {
        const int ch0 = chw + pc * 16;
        uint32_t r[16];
        tmem_ld16(acc2 + lane_addr + (uint32_t)ch0, r);
        tmem_wait16(r);
        if (pc == NPIECE - 1) {     // last read of this accumulator: the MMA warp may overwrite it (tile t+2)
          tc_fence_before();
          mbar_arrive(hdr + OFF_ACC2FREE + 8 * ab);
        }
        float x[16], s2v[16];
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(b2s + ch0 + i);
          x[i] = __uint_as_float(r[i]) + bb.x; x[i + 1] = __uint_as_float(r[i + 1]) + bb.y;
          x[i + 2] = __uint_as_float(r[i + 2]) + bb.z; x[i + 3] = __uint_as_float(r[i + 3]) + bb.w;
        }
        if (a.epi == kEpiBwdMaskStats) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 mean = __ldg(reinterpret_cast<const float4*>(a.e_stats + ch0 + i));
            const float4 istd = __ldg(reinterpret_cast<const float4*>(a.e_stats + C + ch0 + i));
            const float4 scl = __ldg(reinterpret_cast<const float4*>(a.e_stats + 2 * C + ch0 + i));
            const float4 sft = __ldg(reinterpret_cast<const float4*>(a.e_stats + 3 * C + ch0 + i));
            x[i] = fmaf(ev[i], scl.x, sft.x) > 0.f ? x[i] : 0.f;
            x[i + 1] = fmaf(ev[i + 1], scl.y, sft.y) > 0.f ? x[i + 1] : 0.f;
            x[i + 2] = fmaf(ev[i + 2], scl.z, sft.z) > 0.f ? x[i + 2] : 0.f;
            x[i + 3] = fmaf(ev[i + 3], scl.w, sft.w) > 0.f ? x[i + 3] : 0.f;
            s2v[i] = x[i] * ((ev[i] - mean.x) * istd.x);
            s2v[i + 1] = x[i + 1] * ((ev[i + 1] - mean.y) * istd.y);
            s2v[i + 2] = x[i + 2] * ((ev[i + 2] - mean.z) * istd.z);
            s2v[i + 3] = x[i + 3] * ((ev[i + 3] - mean.w) * istd.w);
          }
        } else if (a.epi == kEpiBwdResidual) {
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] += ev[i];
        } else if (a.epi == kEpiFwdBnRes) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 scl = __ldg(reinterpret_cast<const float4*>(a.e_stats + 2 * C + ch0 + i));
            const float4 sft = __ldg(reinterpret_cast<const float4*>(a.e_stats + 3 * C + ch0 + i));
            x[i] = fmaxf(fmaf(x[i], scl.x, sft.x) + ev[i], 0.f);
            x[i + 1] = fmaxf(fmaf(x[i + 1], scl.y, sft.y) + ev[i + 1], 0.f);
            x[i + 2] = fmaxf(fmaf(x[i + 2], scl.z, sft.z) + ev[i + 2], 0.f);
            x[i + 3] = fmaxf(fmaf(x[i + 3], scl.w, sft.w) + ev[i + 3], 0.f);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) s2v[i] = x[i] * x[i];
        }
        if (pc + 1 < NPIECE) fetch_e(pc + 1);
        if (out_dst != nullptr) {
          stg8(out_dst + pc * 16, &x[0]);
          stg8(out_dst + pc * 16 + 8, &x[8]);
        }
        if (a.sums != nullptr) {
          if (pix_out < 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { x[i] = 0.f; s2v[i] = 0.f; }
          }
          run1[pc] += transpose_reduce16(x, lane);
          run2[pc] += transpose_reduce16(s2v, lane);
        }
      }
    };

    TileIter t1, t2;
    t1.init(geo);
    t2.init(geo);
    { H3_T0(); epi1(0, t1); H3_T1(2); }
    for (int it = 0; it < ntiles; ++it) {
      if (it + 1 < ntiles) { t1.next(geo); H3_T0(); epi1(it + 1, t1); H3_T1(2); }
      H3_T0();
      epi2(it, t2);
      H3_T1(3);
      t2.next(geo);
    }
    // ---- BatchNorm partial sums: warp partials -> shared slots -> fixed-order sum over the four row quadrants -> fp64 atomics
    if (a.sums != nullptr) {
      // the operand buffers are dead: every epilogue warp is past its last TMEM read and the last MMA has retired
      // (acc2full of the last tile was waited for by all of them), the loaders finished before that MMA could start
      const int et = tid;
      named_bar_sync(1, N_EPI);
      float* slots = reinterpret_cast<float*>(gen + K::HDR_BYTES);          // [4 quadrants][2][C]
      if ((lane & 1) == 0) {
#pragma unroll
        for (int pc = 0; pc < NPIECE; ++pc) {
          const int ch = chw + pc * 16 + ((lane >> 1) & 15);
          slots[(q * 2 + 0) * C + ch] = run1[pc];
          slots[(q * 2 + 1) * C + ch] = run2[pc];
        }
      }
      named_bar_sync(1, N_EPI);
      for (int i = et; i < 2 * C; i += N_EPI) {
        const float t = (slots[i] + slots[2 * C + i]) + (slots[4 * C + i] + slots[6 * C + i]);
        atomicAdd(a.sums + i, (double)t);
      }
    }
  }

  if (TRACE && blockIdx.x == 0 && lane == 0 && trole >= 0) trc[4 * 6 + trole] = clock64() - t_start;    // role finished
  tc_fence_before();
  __syncthreads();
  if (TRACE && tid == 0 && blockIdx.x == 0) {
    printf("pair_h3<%d,%d> epi=%d CTA0: %d tiles, total %lld clk\n"
           "  MMA   : done %lld | wait input %lld mid %lld weights %lld acc2free %lld\n"
           "  loader: done %lld %lld %lld | wait buffer %lld %lld %lld | issue %lld convert(+wait) %lld table barrier %lld\n"
           "  epi   : done %lld %lld | wait acc1 %lld %lld acc2 %lld %lld | epi1 total %lld %lld epi2 total %lld %lld\n",
           C, FMT, a.epi, ntiles, clock64() - t_start,
           trc[24 + 5], trc[0 + 5], trc[6 + 5], trc[12 + 5], trc[18 + 5],
           trc[24 + 2], trc[24 + 3], trc[24 + 4], trc[2], trc[3], trc[4], trc[6 + 2], trc[12 + 2], trc[18 + 2],
           trc[24 + 0], trc[24 + 1], trc[0], trc[1], trc[6 + 0], trc[6 + 1], trc[12 + 0], trc[12 + 1], trc[18 + 0], trc[18 + 1]);
  }
  if (CL > 1) cluster_sync_all();   // no CTA exits while a peer may still multicast into it / arrive on its barriers
  if (warp == W_MMA) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(K::TMEM_COLS) : "memory");
  }
}

// Lattice tile (TU x TV mid pixels of TR residue classes, <= 128 mid rows, <= IN_MAX input rows) that needs the fewest
// rounds of the persistent grid, then the fewest tiles, then the smallest input tile (as nb1d_pair_tc3.cu).
static TileShape choose_tile(int Ul, int Vl, int d, int nimg, int nctas, int IN_MAX) {
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
        const long rounds = (tiles + nctas - 1) / nctas;
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

template <int C, int FMT, bool TRACE>
int launch_c(const PairArgs& a, cudaStream_t s) {
  using K = Cfg<C>;
  static_assert(K::SMEM_BYTES <= 227 * 1024, "h3 pair kernel shared memory budget");
  static_assert(K::NSTAGE <= 14, "barrier header layout");
  static_assert(OFF_PIXTAB + 8 * IN_MAX_ALL <= OFF_TRACE && K::IN_MAX <= IN_MAX_ALL && 8 * C * 4 <= (int)K::BUF_BYTES, "header layout");
  static_assert(8 * K::NBUF * K::SLABS <= 32 && 8 * K::NBUF * K::NKC <= 64 && K::NBUF <= 3, "barrier header layout");
  static_assert(N_LOAD >= K::IN_MAX && N_LOAD % (C / 4) == 0 && NTHREADS == (W_PROD + 1) * 32, "role mapping");
  const int d = a.dil;
  const int U = a.vert_first ? a.H : a.W, V = a.vert_first ? a.W : a.H;
  const int Ul = cdiv(U, d), Vl = cdiv(V, d);
  MDIL_REQUIRE(a.wstream_tc != nullptr && ((uintptr_t)a.wstream_tc & 15) == 0, "pair_h3: weight stream");
  const int cl = K::RESIDENT ? 1 : cluster_size();
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
    MDIL_CUDA(cudaFuncSetAttribute(pair_h3_kernel<C, FMT, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES));
    int n = kNumSMs;
    if (cl > 1) {
      cfg.gridDim = dim3((unsigned)(kNumSMs / cl * cl), 1, 1);
      int ncl = 0;
      MDIL_CUDA(cudaOccupancyMaxActiveClusters(&ncl, pair_h3_kernel<C, FMT, TRACE>, &cfg));
      n = ncl * cl;
    }
    MDIL_REQUIRE(n >= cl, "pair_h3: no co-resident cluster fits");
    if (n > kNumSMs) n = kNumSMs / cl * cl;
    max_ctas = n;
    max_ctas_slot.store(n, std::memory_order_release);
  }
  const TileShape ts = choose_tile(Ul, Vl, d, a.N, max_ctas, K::IN_MAX);
  const long total = (long)a.N * cdiv(d * d, ts.TR) * cdiv(Ul, ts.TU) * cdiv(Vl, ts.TV);
  MDIL_REQUIRE(total > 0 && total < (1L << 30), "pair_h3: tile count");
  long grid = total < max_ctas ? (total + cl - 1) / cl * cl : max_ctas;
  Geo geo;
  memset(&geo, 0, sizeof(geo));
  geo.TU = ts.TU; geo.TV = ts.TV; geo.TR = ts.TR;
  geo.TVH = ts.TV + 2;
  geo.RT = ts.TR * geo.TVH;
  geo.INROWS = (ts.TU + 2) * geo.RT;
  geo.M1 = ts.TU * geo.RT;
  geo.U = U; geo.V = V; geo.dd = d * d;
  geo.dshift = -1;
  for (int sft = 0; sft < 16; ++sft) if ((1 << sft) == d) geo.dshift = sft;
  geo.tiles_u = cdiv(Ul, ts.TU); geo.tiles_v = cdiv(Vl, ts.TV); geo.ncb = cdiv(d * d, ts.TR);
  geo.total_tiles = (int)total;
  geo.tiles_per_cta = (int)((total + grid - 1) / grid);
  geo.cl = cl;
  {
    long g = grid;
    geo.sv = (int)(g % geo.tiles_v); g /= geo.tiles_v;
    geo.su = (int)(g % geo.tiles_u); g /= geo.tiles_u;
    geo.scb = (int)(g % geo.ncb); g /= geo.ncb;
    geo.sn = (int)g;
  }
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  MDIL_CUDA(cudaLaunchKernelEx(&cfg, pair_h3_kernel<C, FMT, TRACE>, a, geo));
  MDIL_LAUNCH_CHECK();
  return 0;
}

// ---- one launch packs the 16-bit hi/lo SWIZZLE_64B weight images of a block's four pair launches straight from the
// PyTorch-layout weights.  Chunk order inside a launch's stream: per 32-channel chunk j of the first conv its three taps,
// then the adapter's chunks, then per chunk j of the second conv its three taps.  Streams 0, 1 (forward) are fp16, 2, 3
// (backward) bf16 unless fmt_override >= 0.
struct PackSrc { const float* w[6]; };   // w31_1, w13_1, w31_2, w13_2, wp1, wp2  ([co][ci][3] / [co][ci])
__global__ void pack_block_h3_kernel(const PackSrc src, unsigned short* __restrict__ packed, int C, int has_adapter,
                                     int fmt_override) {
  const int CC = C * C, nkc = C / KCH;
  const int per1 = 3 + (has_adapter ? 1 : 0);
  const long total = 4L * 7 * CC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int which = (int)(i / (7 * CC));
    const int r = (int)(i % (7 * CC));
    const int slab = r / CC, aa = (r % CC) / C, bb = r % C;        // aa = input channel of the launch, bb = output channel
    if (slab == 6 && !has_adapter) continue;
    // which: 0 fwd pair 1, 1 fwd pair 2, 2 bwd pair 2, 3 bwd pair 1.  slab: 0..2 first conv taps, 3..5 second conv, 6 adapter
    const bool bwd = which >= 2;
    const int pairsel = (which == 0 || which == 3) ? 0 : 1;                  // weights of pair 1 or pair 2
    float v;
    if (slab == 6) {
      const float* w = src.w[4 + pairsel];
      v = bwd ? __ldg(w + aa * C + bb) : __ldg(w + bb * C + aa);             // fwd [ci][co] = Wp[co][ci]; bwd [co][ci]
    } else {
      const int conv = slab / 3, tap = slab % 3;
      // forward: first conv = 3x1, second = 1x3; backward runs the pair in reverse: first = 1x3 (flipped), second = 3x1
      const int is13 = bwd ? (conv == 0) : (conv == 1);
      const float* w = src.w[pairsel * 2 + is13];
      v = bwd ? __ldg(w + ((long)aa * C + bb) * 3 + (2 - tap)) : __ldg(w + ((long)bb * C + aa) * 3 + tap);
    }
    const int j = aa / KCH, kk = aa % KCH, nrow = bb;
    int g;
    if (slab < 3) g = j * 3 + slab;
    else if (slab == 6) g = nkc * 3 + j;
    else g = nkc * per1 + j * 3 + (slab - 3);
    const int fmt = fmt_override >= 0 ? fmt_override : (bwd ? 1 : 0);
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
    // [image][C rows of 64 bytes], 16-byte chunk index XOR (row >> 1) & 3 (SWIZZLE_64B)
    const int off = nrow * 32 + ((((kk >> 3) ^ ((nrow >> 1) & 3)) << 3) | (kk & 7));
    unsigned short* stage = packed + (long)which * 14 * CC + (long)g * 2 * C * KCH;
    stage[off] = hi;
    stage[C * KCH + off] = lo;
  }
}

// ---- packed-4 view of a C = 16 block (no adapter, dilation 1): four pixels of a row form one 64-"channel" operand row,
// so the block runs on the C = 64 tensor-core kernel (the FFMA kernel it replaces ran at 0.09 of the HBM roof).  The 16x16
// tap matrices become 64x64 ones: block-diagonal for the 3x1 convs (taps along H), block-banded across the three GROUP
// taps for the 1x3 convs (taps along W, the packing axis): group tap gt connects input slot pi to output slot po when
// 4 (gt - 1) + pi - po + 1 is a conv tap (0..2).  Same stream / chunk order as pack_block_h3_kernel without adapter.
struct PackSrcP4 { const float* w[4]; const float* b[4]; };   // w31_1, w13_1, w31_2, w13_2 ([co][ci][3]) and their biases
__global__ void pack_block_p4_kernel(const PackSrcP4 src, unsigned short* __restrict__ packed, float* __restrict__ bias_rep,
                                     int fmt_override) {
  constexpr int C = 64, CW = 16, CC = C * C, nkc = C / KCH;
  const long total = 4L * 6 * CC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int which = (int)(i / (6 * CC));
    const int r = (int)(i % (6 * CC));
    const int slab = r / CC, aa = (r % CC) / C, bb = r % C;
    const bool bwd = which >= 2;
    const int pairsel = (which == 0 || which == 3) ? 0 : 1;
    const int conv = slab / 3, tap = slab % 3;
    const int is13 = bwd ? (conv == 0) : (conv == 1);
    const float* w = src.w[pairsel * 2 + is13];
    const int pi = aa >> 4, ca = aa & 15, po = bb >> 4, cb = bb & 15;
    int teff = tap;                       // effective conv-window position of this (group tap, slot pair)
    bool nz = true;
    if (is13) { teff = 4 * (tap - 1) + pi - po + 1; nz = teff >= 0 && teff <= 2; }
    else nz = pi == po;
    float v = 0.f;
    if (nz) v = bwd ? __ldg(w + (ca * CW + cb) * 3 + (2 - teff)) : __ldg(w + (cb * CW + ca) * 3 + teff);
    const int j = aa / KCH, kk = aa % KCH, nrow = bb;
    const int g = slab < 3 ? j * 3 + slab : nkc * 3 + j * 3 + (slab - 3);
    const int fmt = fmt_override >= 0 ? fmt_override : (bwd ? 1 : 0);
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
    unsigned short* stage = packed + (long)which * 14 * CC + (long)g * 2 * C * KCH;
    stage[off] = hi;
    stage[C * KCH + off] = lo;
    if (i < 4 * C) bias_rep[i] = src.b[i / C] != nullptr ? __ldg(src.b[i / C] + (i & 15)) : 0.f;
  }
}

}  // namespace h3

static int h3_fmt_override() {
  static const int v = [] {
    const char* e = getenv("MDIL_H3_FMT");
    if (e != nullptr && strcmp(e, "fp16") == 0) return 0;
    if (e != nullptr && strcmp(e, "bf16") == 0) return 1;
    return -1;
  }();
  return v;
}

int launch_pack_block_h3(const float* const* w6, void* packed, int C, int has_adapter, cudaStream_t s) {
  h3::PackSrc src;
  for (int i = 0; i < 6; ++i) src.w[i] = w6[i];
  const long total = 4L * 7 * C * C;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  h3::pack_block_h3_kernel<<<grid, 256, 0, s>>>(src, reinterpret_cast<unsigned short*>(packed), C, has_adapter, h3_fmt_override());
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_pack_block_p4(const float* const* w4, const float* const* b4, void* packed16, float* bias_rep, cudaStream_t s) {
  h3::PackSrcP4 src;
  for (int i = 0; i < 4; ++i) { src.w[i] = w4[i]; src.b[i] = b4[i]; }
  const long total = 4L * 6 * 64 * 64;
  h3::pack_block_p4_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(src, reinterpret_cast<unsigned short*>(packed16), bias_rep,
                                                                     h3_fmt_override());
  MDIL_LAUNCH_CHECK();
  return 0;
}

int launch_pair_h3(const PairArgs& a, cudaStream_t s) {
  const int ov = h3_fmt_override();
  const int fmt = ov >= 0 ? ov : ((a.epi == kEpiFwd || a.epi == kEpiFwdBnRes) ? 0 : 1);   // forward: fp16 halves, backward: bf16
  MDIL_REQUIRE(!(a.in_split || a.mid_out_split) || (fmt == 1 && a.in_scale == nullptr),
               "pair_h3: S16 tensors hold bf16 halves (backward launches, no input prologue)");
  if (a.trace) {
    if (a.C == 128) return fmt == 0 ? h3::launch_c<128, 0, true>(a, s) : h3::launch_c<128, 1, true>(a, s);
    if (a.C == 64) return fmt == 0 ? h3::launch_c<64, 0, true>(a, s) : h3::launch_c<64, 1, true>(a, s);
  }
  if (a.C == 128) return fmt == 0 ? h3::launch_c<128, 0, false>(a, s) : h3::launch_c<128, 1, false>(a, s);
  if (a.C == 64) return fmt == 0 ? h3::launch_c<64, 0, false>(a, s) : h3::launch_c<64, 1, false>(a, s);
  return set_error(-2, "pair_h3: C must be 64 or 128", __FILE__, __LINE__);
}

}  // namespace mdil
