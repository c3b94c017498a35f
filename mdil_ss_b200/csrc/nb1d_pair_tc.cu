// Tensor-core (tcgen05 / TMEM) version of the fused factorised-pair kernel for C = 64 and C = 128
// (same contract as nb1d_pair.cu: see the header comment there for the four uses).
//
// The C x C channel contractions of the two 3-tap convolutions and of the 1x1 adapter run on the 5th-gen
// tensor cores as error-compensated 3xTF32:  x*w ~= hi(x)*hi(w) + lo(x)*hi(w) + hi(x)*lo(w), hi = tf32(x),
// lo = tf32(x - hi), fp32 accumulation in TMEM (measured 1e-6 relative on random data, i.e. fp32-class, where a
// single TF32 pass gives 2.5e-4 -- tools/umma_probe.cu).
//
// One CTA = one lattice tile of <= 128 "mid" pixels (M = 128 rows of the MMA):
//   workers (8 warps): input tile global -> registers (BN+ReLU prologue) -> hi/lo split -> K-major
//       SWIZZLE_128B shared-memory operand (one 128-byte row per pixel and 32-channel slab);
//   control thread (warp 8): streams the pre-split, pre-swizzled weight images ([C][16] SWIZZLE_64B, hi+lo)
//       through a 4-deep ring with cp.async.bulk + mbarrier, and issues tcgen05.mma.kind::tf32 (M=128, N=C, K=8):
//       a conv tap is just a shifted row window of the same operand (descriptor start address + rows*128 B);
//   epilogue 1: TMEM -> registers -> bias/ReLU (or mask) -> hi/lo -> shared memory (overwrites the input tile,
//       the adapter contribution having been accumulated first) = A operand of the second conv;
//   epilogue 2: TMEM -> registers -> bias / mask / residual -> global, per-channel sums by warp shuffles.
#include "kernels.cuh"

#include <stdio.h>

namespace mdil {
namespace tc {

constexpr int KC = 16;        // input channels per weight chunk (one 64-byte SWIZZLE_64B row)
constexpr int IN_MAX = 160;   // input-tile rows (pixels) held in shared memory
constexpr int NWORK = 256;    // worker threads (warps 0..7); warp 8 issues the MMAs, warp 9 streams the weights

template <int C> struct Cfg {
  static constexpr int NSTAGE = C == 64 ? 8 : 4;            // weight ring depth (bytes in flight hide the L2 latency)
  static constexpr int SLABS = C / 32;
  static constexpr uint32_t SLAB_BYTES = IN_MAX * 128;
  static constexpr uint32_t ACT_BYTES = SLABS * SLAB_BYTES;
  static constexpr uint32_t HALF_STAGE = C * 64;           // one (hi or lo) weight image of a chunk
  static constexpr uint32_t STAGE_BYTES = 2 * HALF_STAGE;
  static constexpr uint32_t HDR_BYTES = 2048;              // barriers | row tables | trace
  // epilogue-1 staging: one 4 KB piece per worker warp. C = 128: the dead rows 128..159 of the 8 slab arrays;
  // C = 64: an extra region after the weight ring.
  static constexpr uint32_t STG1_EXTRA = C == 128 ? 0 : 8 * 4096;
  static constexpr uint32_t SMEM_BYTES = 1024 + HDR_BYTES + 2 * ACT_BYTES + NSTAGE * STAGE_BYTES + STG1_EXTRA;
  static_assert(16 * NSTAGE + 48 <= 1024, "barrier header");
  static constexpr int NCH = C / KC;
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc),
      "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_w(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d), "r"(a_lo),
      "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
// K-major shared-memory operand descriptors (sm_100 UMMA): start>>4 | LBO | SBO>>4 <<32 | version 1 <<46 | swizzle <<61
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
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
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4& x, float4& hi, float4& lo) {
  hi.x = tf32_rna(x.x); hi.y = tf32_rna(x.y); hi.z = tf32_rna(x.z); hi.w = tf32_rna(x.w);
  lo.x = tf32_rna(x.x - hi.x); lo.y = tf32_rna(x.y - hi.y); lo.z = tf32_rna(x.z - hi.z); lo.w = tf32_rna(x.w - hi.w);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// byte offset of (row, 16-byte chunk c of the 128-byte row) inside a SWIZZLE_128B slab whose base address is `slab`
__device__ __forceinline__ uint32_t sw128_addr(uint32_t slab, int row, int c) {
  const uint32_t ra = slab + (uint32_t)row * 128;
  return ra + ((uint32_t)(c ^ ((ra >> 7) & 7)) << 4);
}

template <int C>
__global__ void __launch_bounds__(NWORK + 64, 1)
pair_tc_kernel(const __grid_constant__ PairArgs a, const TileShape ts) {
  using K = Cfg<C>;
  constexpr int NSTAGE = K::NSTAGE;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t hdr = raw + ((1024 - (raw & 1023)) & 1023);
  const uint32_t bar_full = hdr, bar_empty = hdr + 8 * NSTAGE, bar_act = hdr + 16 * NSTAGE, bar_acc1 = bar_act + 8,
                 bar_acc2 = bar_act + 16, tmem_slot = bar_act + 32;
  const uint32_t act_hi = hdr + K::HDR_BYTES, act_lo = act_hi + K::ACT_BYTES, ring = act_lo + K::ACT_BYTES;
  unsigned char* gen = smem_raw + (hdr - raw);   // generic pointer to hdr
  int* rowpix = reinterpret_cast<int*>(gen + 256);               // [IN_MAX] pixel index of each input-tile row, -1 = zero
  int* midpix = reinterpret_cast<int*>(gen + 256 + 4 * IN_MAX);  // [128] pixel index of each mid row that this tile owns, -1 = none
  long long* tstamp = reinterpret_cast<long long*>(gen + 1536);  // debug trace (a.trace): phase timestamps of CTA 0
#define MDIL_TS(i) do { if (a.trace && blockIdx.x == 0) tstamp[i] = clock64(); } while (0)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int TU = ts.TU, TV = ts.TV, TR = ts.TR, TVH = TV + 2;
  const int RT = TR * TVH;             // rows per lattice step along u  (row order: [u][class][v])
  const int INROWS = (TU + 2) * RT;
  const int M1 = TU * RT;
  const int d = a.dil;
  const int U = a.vert_first ? a.H : a.W, V = a.vert_first ? a.W : a.H;
  const long su = a.vert_first ? (long)a.W * C : C, sv = a.vert_first ? C : (long)a.W * C;
  const int Ul = (U + d - 1) / d, Vl = (V + d - 1) / d;
  const int tiles_u = (Ul + TU - 1) / TU, tiles_v = (Vl + TV - 1) / TV;
  const int ncb = (d * d + TR - 1) / TR;
  int b = blockIdx.x;
  const int tvi = b % tiles_v; b /= tiles_v;
  const int tui = b % tiles_u; b /= tiles_u;
  const int cb = b % ncb;
  const int n = b / ncb;
  const int ul0 = tui * TU, vl0 = tvi * TV;
  const size_t img = (size_t)n * a.H * a.W * C;
  const int NAD = a.has_adapter ? K::NCH : 0;
  const int G = 6 * K::NCH + NAD;

  if (tid == 0) {
    MDIL_TS(0);
    if (a.trace) { tstamp[11] = 0; tstamp[12] = 0; }
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    mbar_init(bar_act, NWORK);
    mbar_init(bar_acc1, 1);
    mbar_init(bar_acc2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * C) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // row -> pixel tables (all the integer divisions of the tile happen here, once per row)
  if (tid < IN_MAX) {
    int pix = -1;
    if (tid < INROWS) {
      const int iu = tid / RT, rem = tid % RT;
      const int r = rem / TVH, iv = rem % TVH;
      const int cidx = cb * TR + r;
      const int ul = ul0 - 1 + iu, vl = vl0 - 1 + iv;
      const int u = (cidx / d) + d * ul, v = (cidx % d) + d * vl;
      if (cidx < d * d && ul >= 0 && vl >= 0 && u < U && v < V) pix = a.vert_first ? u * a.W + v : v * a.W + u;
    }
    rowpix[tid] = pix;
    if (tid < 128) {
      int mp = -1;
      if (tid < M1) {
        const int mu = tid / RT, rem = tid % RT;
        const int r = rem / TVH, mv = rem % TVH;
        const int cidx = cb * TR + r;
        const int u = (cidx / d) + d * (ul0 + mu), v = (cidx % d) + d * (vl0 - 1 + mv);
        if (cidx < d * d && mv >= 1 && mv <= TV && u < U && v < V) mp = a.vert_first ? u * a.W + v : v * a.W + u;
      }
      midpix[tid] = mp;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - hdr));
  const uint32_t acc1 = tmem, acc2 = tmem + C;

  if (warp == 9) {
    // ============================================================ weight producer: TMA bulk copies into the ring
    // (whole warp runs the loop so the address arithmetic stays warp-uniform; lane 0 issues)
    const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(a.wstream_tc);
    for (int g = 0; g < G; ++g) {
      const int st = g % NSTAGE;
      const long long tw0 = a.trace ? clock64() : 0;
      if (g >= NSTAGE) mbar_wait(bar_empty + 8 * st, (uint32_t)(((g / NSTAGE) - 1) & 1));   // MMAs of chunk g-NSTAGE retired
      if (a.trace && blockIdx.x == 0 && lane == 0) tstamp[12] += clock64() - tw0;   // cycles the producer waited for a free slot
      if (lane == 0) {
        mbar_expect_tx(bar_full + 8 * st, K::STAGE_BYTES);
        bulk_g2s(ring + st * K::STAGE_BYTES, wsrc + (size_t)g * K::STAGE_BYTES, K::STAGE_BYTES, bar_full + 8 * st);
      }
      __syncwarp();
    }
  } else if (warp == 8) {
    // ============================================================ MMA issuer.  The whole warp walks the chunk loop
    // (waits, descriptor arithmetic: warp-uniform, so it can live in uniform registers); lane 0 issues the MMAs.
    // The issue rate of this loop bounds the tensor pipe: descriptors are 32-bit adds on precomputed words.
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(C >> 3) << 17) | ((128u >> 4) << 24);
    // descriptor = {lo: start>>4 | LBO(1)<<16, hi: SBO>>4 | version 1<<14 | layout<<29}
    const uint32_t a_hiw = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_hiw = (512u >> 4) | (1u << 14) | (4u << 29);
    const uint32_t ahi0 = ((act_hi & 0x3FFFF) >> 4) | (1u << 16), alo0 = ((act_lo & 0x3FFFF) >> 4) | (1u << 16);
    const uint32_t ring0 = ((ring & 0x3FFFF) >> 4) | (1u << 16);
    int g = 0;
    auto chunk = [&](int row0, int j, uint32_t acc, uint32_t accumulate) {
      const int st = g % NSTAGE;
      const long long tw0 = a.trace ? clock64() : 0;
      mbar_wait(bar_full + 8 * st, (uint32_t)((g / NSTAGE) & 1));
      tc_fence_after();
      if (a.trace && blockIdx.x == 0 && lane == 0) tstamp[11] += clock64() - tw0;   // cycles the issuer waited for weights
      const uint32_t ad = ((uint32_t)(j >> 1) * K::SLAB_BYTES + (uint32_t)row0 * 128 + (uint32_t)(j & 1) * 64) >> 4;
      const uint32_t ah = ahi0 + ad, al = alo0 + ad;
      const uint32_t bh = ring0 + (uint32_t)st * (K::STAGE_BYTES >> 4), bl = bh + (K::HALF_STAGE >> 4);
      if (lane == 0) {
        mma_tf32_w(acc, ah, a_hiw, bh, b_hiw, idesc, accumulate);
        mma_tf32_w(acc, al, a_hiw, bh, b_hiw, idesc, 1u);
        mma_tf32_w(acc, ah, a_hiw, bl, b_hiw, idesc, 1u);
        mma_tf32_w(acc, ah + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);     // second K step: +32 bytes
        mma_tf32_w(acc, al + 2, a_hiw, bh + 2, b_hiw, idesc, 1u);
        mma_tf32_w(acc, ah + 2, a_hiw, bl + 2, b_hiw, idesc, 1u);
        umma_commit(bar_empty + 8 * st);                                // ring slot reusable when these retire
      }
      __syncwarp();
      ++g;
    };
    if (lane == 0) MDIL_TS(1);
    mbar_wait(bar_act, 0);   // input tile is in shared memory
    tc_fence_after();
    if (lane == 0) MDIL_TS(2);
    for (int tap = 0; tap < 3; ++tap)           // first conv: tap window = rows shifted by tap*RT
      for (int j = 0; j < K::NCH; ++j) chunk(tap * RT, j, acc1, (tap | j) != 0);
    for (int j = 0; j < NAD; ++j) chunk(RT + 1, j, acc2, j != 0);   // adapter on the centre pixels
    if (lane == 0) { umma_commit(bar_acc1); MDIL_TS(3); }           // every read of the input tile has retired
    __syncwarp();
    mbar_wait(bar_act, 1);                      // `mid` is in shared memory
    tc_fence_after();
    if (lane == 0) MDIL_TS(4);
    for (int tap = 0; tap < 3; ++tap)           // second conv over `mid`: tap window = rows shifted by tap
      for (int j = 0; j < K::NCH; ++j) chunk(tap, j, acc2, (NAD > 0) || (tap | j) != 0);
    if (lane == 0) { umma_commit(bar_acc2); MDIL_TS(5); }
    __syncwarp();
  } else {
    // ============================================================ workers
    const size_t imgC = img;   // float offset of image n
    // ---- input tile: 4 pixels x one 32-channel slab per warp iteration (128-byte global segments, conflict-free
    //      128-bit shared stores), BN+ReLU prologue and hi/lo split in registers; row decode comes from the table
    {
      const int c16 = lane & 7, psub = lane >> 3;
      const int nitems = ((INROWS + 3) / 4) * K::SLABS;
      constexpr int B = 10;                // loads in flight per thread
      for (int base = warp; base < nitems; base += 8 * B) {
        float4 x[B];
        int meta[B];
#pragma unroll
        for (int bb = 0; bb < B; ++bb) {
          const int it = base + bb * 8;
          x[bb] = make4(0.f);
          meta[bb] = -1;
          if (it < nitems) {
            const int slab = it % K::SLABS, row = (it / K::SLABS) * 4 + psub;
            if (row < INROWS) {
              meta[bb] = (slab << 16) | row;
              const int pix = rowpix[row];
              if (pix >= 0) {
                x[bb] = ldg4(a.in + imgC + (size_t)pix * C + slab * 32 + c16 * 4);
                meta[bb] |= 1 << 30;   // in-image: the BN+ReLU prologue applies
              }
            }
          }
        }
#pragma unroll
        for (int bb = 0; bb < B; ++bb) {
          if (meta[bb] < 0) continue;
          const int slab = (meta[bb] >> 16) & 0xFF, row = meta[bb] & 0xFFFF;
          float4 v4 = x[bb];
          if ((meta[bb] >> 30) && a.in_scale != nullptr) {
            const int ch = slab * 32 + c16 * 4;
            const float4 sc = ldg4(a.in_scale + ch), sh = ldg4(a.in_shift + ch);
            v4.x = fmaxf(fmaf(v4.x, sc.x, sh.x), 0.f);
            v4.y = fmaxf(fmaf(v4.y, sc.y, sh.y), 0.f);
            v4.z = fmaxf(fmaf(v4.z, sc.z, sh.z), 0.f);
            v4.w = fmaxf(fmaf(v4.w, sc.w, sh.w), 0.f);
          }
          float4 hi, lo;
          split4(v4, hi, lo);
          const uint32_t ah = sw128_addr(act_hi + (uint32_t)slab * K::SLAB_BYTES, row, c16);   // key from the absolute address
          *reinterpret_cast<float4*>(gen + (ah - hdr)) = hi;
          *reinterpret_cast<float4*>(gen + (ah - hdr) + K::ACT_BYTES) = lo;   // ACT_BYTES % 1024 == 0: same key
        }
      }
    }
    fence_proxy_async();
    mbar_arrive(bar_act);
    if (tid == 0) MDIL_TS(6);

    // ---- this thread's accumulator row and column half; its warp's private staging buffers
    const int q = warp & 3, half = warp >> 2;
    const int m = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int mu = m / RT, rem_m = m % RT;
    const int rcls = rem_m / TVH, mv = rem_m % TVH;
    const int cidx = cb * TR + rcls;
    const int ru = cidx / d, rv = cidx % d;
    const bool cls_ok = m < M1 && cidx < d * d;
    // staging element (r, j) of a [32 rows][32 ch] fp32 tile: float index r*32 + (((j>>2) ^ (r&7)) << 2) + (j&3)
    auto stg_row4 = [](float* st, int r, int q4) -> float4* { return reinterpret_cast<float4*>(st + r * 32 + ((q4 ^ (r & 7)) << 2)); };
    auto stg_elem = [](const float* st, int r, int j) -> float { return st[r * 32 + ((((j >> 2) ^ (r & 7)) << 2) | (j & 3))]; };
    float* stg1 = C == 128
        ? reinterpret_cast<float*>(gen + K::HDR_BYTES + (size_t)(warp >> 2) * K::ACT_BYTES + (size_t)(warp & 3) * K::SLAB_BYTES + 128 * 128)
        : reinterpret_cast<float*>(gen + K::HDR_BYTES + 2 * K::ACT_BYTES + NSTAGE * K::STAGE_BYTES + (size_t)warp * 4096);
    const int rsub = lane >> 3, q4l = lane & 7;   // coalesced view: 4 rows x 8 float4 per instruction

    // ================================================== epilogue 1: mid = f(acc1) -> hi/lo A operand (rows m)
    mbar_wait(bar_acc1, 0);
    tc_fence_after();
    __syncwarp();
    if (tid == 0) MDIL_TS(7);
    {
      const int u = ru + d * (ul0 + mu), v = rv + d * (vl0 - 1 + mv);
      const bool valid = cls_ok && (vl0 - 1 + mv) >= 0 && u < U && v < V;
      const int mypix = valid ? (a.vert_first ? u * a.W + v : v * a.W + u) : -1;
#pragma unroll 1
      for (int cc = 0; cc < C / 64; ++cc) {
        const int ch0 = half * (C / 2) + cc * 32;
        float val[32];
        if (a.mid_mask != nullptr) {
          // ReLU mask tile of this warp's 32 rows x 32 channels: coalesced global loads -> staging -> own row
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rsub;
            const int pix = __shfl_sync(0xffffffffu, mypix, r);
            float4 mk = make4(0.f);
            if (pix >= 0) mk = ldg4(a.mid_mask + imgC + (size_t)pix * C + ch0 + q4l * 4);
            *stg_row4(stg1, r, q4l) = mk;
          }
          __syncwarp();
        }
        tmem_ld32(acc1 + lane_addr + ch0, val);
        if (valid) {
          if (a.mid_mask != nullptr) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 mk = *stg_row4(stg1, lane, j4);
              val[j4 * 4 + 0] = mk.x > 0.f ? val[j4 * 4 + 0] : 0.f;
              val[j4 * 4 + 1] = mk.y > 0.f ? val[j4 * 4 + 1] : 0.f;
              val[j4 * 4 + 2] = mk.z > 0.f ? val[j4 * 4 + 2] : 0.f;
              val[j4 * 4 + 3] = mk.w > 0.f ? val[j4 * 4 + 3] : 0.f;
            }
          } else {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              float4 bb = make4(0.f);
              if (a.b1 != nullptr) bb = ldg4(a.b1 + ch0 + j4 * 4);
              val[j4 * 4 + 0] = fmaxf(val[j4 * 4 + 0] + bb.x, 0.f);
              val[j4 * 4 + 1] = fmaxf(val[j4 * 4 + 1] + bb.y, 0.f);
              val[j4 * 4 + 2] = fmaxf(val[j4 * 4 + 2] + bb.z, 0.f);
              val[j4 * 4 + 3] = fmaxf(val[j4 * 4 + 3] + bb.w, 0.f);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) val[j] = 0.f;
        }
        if (m < M1) {
          const uint32_t slab_hi = act_hi + (uint32_t)(ch0 >> 5) * K::SLAB_BYTES;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 hi, lo;
            split4(make_float4(val[j4 * 4 + 0], val[j4 * 4 + 1], val[j4 * 4 + 2], val[j4 * 4 + 3]), hi, lo);
            const uint32_t ah = sw128_addr(slab_hi, m, j4);
            *reinterpret_cast<float4*>(gen + (ah - hdr)) = hi;
            *reinterpret_cast<float4*>(gen + (ah - hdr) + K::ACT_BYTES) = lo;
          }
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();
    mbar_arrive(bar_act);
    if (tid == 0) MDIL_TS(8);

    // ---- while the second conv runs on the tensor core: write `mid` (hi + lo) out to global, coalesced
    //      (the saved a / c of the forward pass, dc' / da' of the backward pass: consumers are the ReLU masks and the
    //       weight-gradient operands, for which hi + lo (2^-22 relative) is equivalent to the fp32 value)
    if (a.mid_out != nullptr) {
      // all workers must have finished writing mid before anyone reads other warps' rows
      named_bar_sync(1, NWORK);
      const int c16 = lane & 7;
      for (int it = warp; it < 32 * K::SLABS; it += 8) {      // 128 rows / 4 per instruction, per slab
        const int slab = it % K::SLABS, row = (it / K::SLABS) * 4 + rsub;
        const int pix = midpix[row];
        if (pix >= 0) {
          const uint32_t ah = sw128_addr(act_hi + (uint32_t)slab * K::SLAB_BYTES, row, c16);
          const float4 hi = *reinterpret_cast<const float4*>(gen + (ah - hdr));
          const float4 lo = *reinterpret_cast<const float4*>(gen + (ah - hdr) + K::ACT_BYTES);
          *reinterpret_cast<float4*>(a.mid_out + imgC + (size_t)pix * C + slab * 32 + c16 * 4) =
              make_float4(hi.x + lo.x, hi.y + lo.y, hi.z + lo.z, hi.w + lo.w);
        }
      }
    }

    // ================================================== epilogue 2: out = acc2 + biases (+ mask / residual), sums
    mbar_wait(bar_acc2, 0);
    tc_fence_after();
    __syncwarp();
    if (tid == 0) MDIL_TS(9);
    // the operand region is free now: two private [32][32] staging tiles per warp + the CTA's [2][C] sums
    float* stgA = reinterpret_cast<float*>(gen + K::HDR_BYTES) + (size_t)warp * 2048;
    float* stgB = stgA + 1024;
    float* ssum = reinterpret_cast<float*>(gen + K::HDR_BYTES) + 8 * 2048;   // [2][C]
    if (a.sums != nullptr) {
      for (int i = tid; i < 2 * C; i += NWORK) ssum[i] = 0.f;
      named_bar_sync(1, NWORK);
    }
    {
      const int u = ru + d * (ul0 + mu), v = rv + d * (vl0 + mv);
      const bool valid = cls_ok && mv < TV && u < U && v < V;
      const int mypix = valid ? (a.vert_first ? u * a.W + v : v * a.W + u) : -1;
#pragma unroll 1
      for (int cc = 0; cc < C / 64; ++cc) {
        const int ch0 = half * (C / 2) + cc * 32;
        float val[32];
        __syncwarp();
        // ---- epilogue inputs of this warp's 32 rows x 32 channels, coalesced -> staging
        if (a.epi != kEpiFwd) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + rsub;
            const int pix = __shfl_sync(0xffffffffu, mypix, r);
            float4 e0 = make4(0.f), e1 = make4(0.f);
            if (pix >= 0) {
              e0 = ldg4(a.e0 + imgC + (size_t)pix * C + ch0 + q4l * 4);
              if (a.epi == kEpiBwdResidual) e1 = ldg4(a.e1 + imgC + (size_t)pix * C + ch0 + q4l * 4);
            }
            *stg_row4(stgA, r, q4l) = e0;                                   // p (mask + stats)  |  dy
            if (a.epi == kEpiBwdResidual) *stg_row4(stgB, r, q4l) = e1;     //                   |  y
          }
          __syncwarp();
        }
        tmem_ld32(acc2 + lane_addr + ch0, val);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 bb = make4(0.f);
          if (a.b2 != nullptr) bb = ldg4(a.b2 + ch0 + j4 * 4);
          if (a.bad != nullptr) {
            const float4 b3 = ldg4(a.bad + ch0 + j4 * 4);
            bb.x += b3.x; bb.y += b3.y; bb.z += b3.z; bb.w += b3.w;
          }
          val[j4 * 4 + 0] += bb.x; val[j4 * 4 + 1] += bb.y; val[j4 * 4 + 2] += bb.z; val[j4 * 4 + 3] += bb.w;
        }
        if (a.epi == kEpiBwdMaskStats) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 pv = *stg_row4(stgA, lane, j4);
            const float4 sc = ldg4(a.e_stats + 2 * C + ch0 + j4 * 4), sh = ldg4(a.e_stats + 3 * C + ch0 + j4 * 4);
            val[j4 * 4 + 0] = fmaf(pv.x, sc.x, sh.x) > 0.f ? val[j4 * 4 + 0] : 0.f;
            val[j4 * 4 + 1] = fmaf(pv.y, sc.y, sh.y) > 0.f ? val[j4 * 4 + 1] : 0.f;
            val[j4 * 4 + 2] = fmaf(pv.z, sc.z, sh.z) > 0.f ? val[j4 * 4 + 2] : 0.f;
            val[j4 * 4 + 3] = fmaf(pv.w, sc.w, sh.w) > 0.f ? val[j4 * 4 + 3] : 0.f;
          }
        } else if (a.epi == kEpiBwdResidual) {
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 dy4 = *stg_row4(stgA, lane, j4), y4 = *stg_row4(stgB, lane, j4);
            val[j4 * 4 + 0] += y4.x > 0.f ? dy4.x : 0.f;
            val[j4 * 4 + 1] += y4.y > 0.f ? dy4.y : 0.f;
            val[j4 * 4 + 2] += y4.z > 0.f ? dy4.z : 0.f;
            val[j4 * 4 + 3] += y4.w > 0.f ? dy4.w : 0.f;
          }
        }
        if (!valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) val[j] = 0.f;
        }
        // ---- results -> staging (stgB) -> coalesced global stores; per-channel sums down the staging columns
        __syncwarp();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          *stg_row4(stgB, lane, j4) = make_float4(val[j4 * 4 + 0], val[j4 * 4 + 1], val[j4 * 4 + 2], val[j4 * 4 + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + rsub;
          const int pix = __shfl_sync(0xffffffffu, mypix, r);
          if (pix >= 0) *reinterpret_cast<float4*>(a.out + imgC + (size_t)pix * C + ch0 + q4l * 4) = *stg_row4(stgB, r, q4l);
        }
        if (a.sums != nullptr) {
          // lane j owns channel ch0 + j: walk the 32 rows (conflict-free in the swizzled tile); invalid rows hold zeros
          float s1 = 0.f, s2 = 0.f;
          if (a.epi == kEpiFwd) {
#pragma unroll 8
            for (int r = 0; r < 32; ++r) { const float o = stg_elem(stgB, r, lane); s1 += o; s2 = fmaf(o, o, s2); }
          } else {
            const float mean = __ldg(a.e_stats + ch0 + lane), istd = __ldg(a.e_stats + C + ch0 + lane);
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
              const float o = stg_elem(stgB, r, lane), pv = stg_elem(stgA, r, lane);
              s1 += o;
              s2 = fmaf(o, (pv - mean) * istd, s2);
            }
          }
          atomicAdd(ssum + ch0 + lane, s1);
          atomicAdd(ssum + C + ch0 + lane, s2);
        }
      }
    }
    if (a.sums != nullptr) {
      named_bar_sync(1, NWORK);
      for (int i = tid; i < 2 * C; i += NWORK) atomicAdd(a.sums + i, (double)ssum[i]);
    }
  }

  if (tid == 0) MDIL_TS(10);
  tc_fence_before();
  __syncthreads();
  if (a.trace && blockIdx.x == 0 && tid == 0) {
    const long long t0 = tstamp[0];
    printf("pair_tc<%d> trace (clk since start): ctl_ready %lld act0 %lld | stage1 issued %lld act1 %lld | all issued %lld || worker: tile loaded %lld acc1 %lld epi1 done %lld acc2 %lld epi2 done %lld\n", C,
           tstamp[1] - t0, tstamp[2] - t0, tstamp[3] - t0, tstamp[4] - t0, tstamp[5] - t0, tstamp[6] - t0, tstamp[7] - t0,
           tstamp[8] - t0, tstamp[9] - t0, tstamp[10] - t0);
    printf("      issuer waited %lld clks for weights, producer waited %lld clks for free slots (%d chunks)\n", tstamp[11], tstamp[12], G);
  }
  if (warp == 8) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2 * C) : "memory");
  }
}

template <int C>
TileShape choose_tile(int Ul, int Vl, int d) {
  TileShape best{1, 2, 1};
  long best_ctas = -1, best_load = 0;
  for (int TR = 1; TR <= 8; ++TR) {
    if (TR > d * d) break;
    for (int TU = 1; TU <= 32; ++TU) {
      for (int TV = 2; TV <= 62; TV += 2) {
        const int TVH = TV + 2;
        if (TR * TU * TVH > 128) break;
        if (TR * (TU + 2) * TVH > IN_MAX) break;
        if (TR > 1 && (TU < Ul || TV < Vl)) continue;
        const long ctas = (long)cdiv(d * d, TR) * cdiv(Ul, TU) * cdiv(Vl, TV);
        const long load = (long)TR * (TU + 2) * TVH;
        if (best_ctas < 0 || ctas < best_ctas || (ctas == best_ctas && load < best_load)) {
          best_ctas = ctas; best_load = load; best = TileShape{TU, TV, TR};
        }
      }
    }
  }
  return best;
}

template <int C>
int launch_c(const PairArgs& a, cudaStream_t s) {
  using K = Cfg<C>;
  static_assert(K::SMEM_BYTES <= 227 * 1024, "tensor-core pair kernel shared memory budget");
  const int d = a.dil;
  const int U = a.vert_first ? a.H : a.W, V = a.vert_first ? a.W : a.H;
  const int Ul = cdiv(U, d), Vl = cdiv(V, d);
  const TileShape ts = choose_tile<C>(Ul, Vl, d);
  const long ctas = (long)a.N * cdiv(d * d, ts.TR) * cdiv(Ul, ts.TU) * cdiv(Vl, ts.TV);
  MDIL_REQUIRE(ctas > 0 && ctas < (1L << 31), "pair_tc: grid size");
  MDIL_REQUIRE(a.wstream_tc != nullptr && ((uintptr_t)a.wstream_tc & 15) == 0, "pair_tc: weight stream");
  MDIL_CUDA(cudaFuncSetAttribute(pair_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM_BYTES));
  pair_tc_kernel<C><<<(unsigned)ctas, NWORK + 64, K::SMEM_BYTES, s>>>(a, ts);
  MDIL_LAUNCH_CHECK();
  return 0;
}

// fp32 [chunk][16][C] slabs of the FFMA stream -> hi/lo SWIZZLE_64B [C][16] images in tensor-core stream order
// (first conv taps, adapter, second conv taps)
__global__ void pack_tc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int has_adapter) {
  const int nch = C / KC;
  const int G = 6 * nch + (has_adapter ? nch : 0);
  const long total = (long)G * C * KC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int nrow = (int)(i % C);
    const int kk = (int)((i / C) % KC);
    const int g = (int)(i / ((long)C * KC));
    int gs;
    if (g < 3 * nch) gs = g;
    else if (has_adapter && g < 4 * nch) gs = 6 * nch + (g - 3 * nch);
    else gs = g - (has_adapter ? nch : 0);
    const float v = __ldg(src + ((long)gs * KC + kk) * C + nrow);
    const float hi = tf32_rna(v), lo = tf32_rna(v - hi);
    const int off = nrow * 16 + ((((kk >> 2) ^ ((nrow >> 1) & 3)) << 2) | (kk & 3));
    float* stage = dst + (long)g * 2 * C * KC;
    stage[off] = hi;
    stage[C * KC + off] = lo;
  }
}

}  // namespace tc

int launch_pair_tc(const PairArgs& a, cudaStream_t s) {
  switch (a.C) {
    case 128: return tc::launch_c<128>(a, s);
    case 64: return tc::launch_c<64>(a, s);
    default: return set_error(-2, "pair_tc: C must be 64 or 128", __FILE__, __LINE__);
  }
}

int launch_pack_tc(const float* src_stream, float* dst_stream, int C, int has_adapter, cudaStream_t s) {
  const long total = (long)(6 + (has_adapter ? 1 : 0)) * C * C;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  tc::pack_tc_kernel<<<grid, 256, 0, s>>>(src_stream, dst_stream, C, has_adapter);
  MDIL_LAUNCH_CHECK();
  return 0;
}

}  // namespace mdil
