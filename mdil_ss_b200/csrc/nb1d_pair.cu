// Fused factorised-pair kernel of the non_bottleneck_1d(_RAP) block
// (models/erfnet_RA_parallel.py:90-113 and :48-64).
//
// One launch computes, for a tile of pixels,
//     mid = f( conv_A(in') )            3-tap conv along axis u (dilation d), f = ReLU(.+b1) or .*(mask>0)
//     out = conv_B(mid) + adapter(in') + biases      3-tap conv along axis v (dilation d), 1x1 adapter
// with in' = in or ReLU(in*scale+shift) (the per-domain BatchNorm of the previous pair folded into the
// load), the intermediate `mid` living only in shared memory, and an epilogue that either accumulates the
// per-domain BatchNorm statistics (forward), applies the ReLU mask of BN1 and accumulates the BN-backward
// sums (backward of pair 2), or adds the residual gradient (backward of pair 1).
//
// Forward pair 1:  u = rows, v = cols, in = x,            mid = a, out = p   (+ sum p, sum p^2)
// Forward pair 2:  u = rows, v = cols, in' = relu(bn1(p)), mid = c, out = s   (+ sum s, sum s^2)
// Backward pair 2: u = cols, v = rows, in = ds, mid = dc*(c>0), out = dq = (.)*(r>0)  (+ sum dq, sum dq*phat)
// Backward pair 1: u = cols, v = rows, in = dp, mid = da*(a>0), out = dx = (.) + dy*(y>0)
//
// Dilation is handled by tiling the d-strided sub-lattice (row = ru + d*i, col = rv + d*j): inside a
// residue class the dilated convolution is a dense 3-tap one, so the halo is always one lattice pixel,
// and for large d several residue classes are batched into one CTA tile.
// Activations: NHWC fp32, 128-bit coalesced loads, ReLU/BN applied in registers on the way to shared
// memory.  Weights: one linear stream of [16][C] fp32 slabs per launch, staged into a 3-deep shared
// memory ring with cp.async.bulk (TMA bulk copy, mbarrier complete_tx).  Math: FP32 FFMA, TM x TN
// register tiles, accumulators never leave registers between the two convolutions' K loops.
#include "kernels.cuh"

#include <mutex>
#include <stdlib.h>
#include <vector>

namespace mdil {

// ---- opt-in per-launch timing of the fused pair kernel (bench.py's roofline leg): CUDA events recorded on the
// launching stream around each launch while profiling is enabled.  Off by default; host-side state only.
namespace prof {
struct Rec { cudaEvent_t a, b; int kind; };
static std::mutex mu;
static bool enabled = false;
static std::vector<Rec> recs;
}  // namespace prof

int pair_profile_begin() {
  std::lock_guard<std::mutex> lk(prof::mu);
  for (auto& r : prof::recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  prof::recs.clear();
  prof::enabled = true;
  return 0;
}

// kinds: 0..11 = (C index {16,64,128}) * 4 + {fwd pair1, fwd pair2, bwd pair2, bwd pair1}
int pair_profile_end(float* total_ms, int* counts, int nkinds) {
  std::lock_guard<std::mutex> lk(prof::mu);
  prof::enabled = false;
  for (int i = 0; i < nkinds; ++i) { total_ms[i] = 0.f; counts[i] = 0; }
  for (auto& r : prof::recs) {
    MDIL_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    MDIL_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    if (r.kind >= 0 && r.kind < nkinds) { total_ms[r.kind] += ms; counts[r.kind] += 1; }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  prof::recs.clear();
  return 0;
}

void pair_profile_record_begin(const PairArgs& a, cudaStream_t s, void** out) {
  *out = nullptr;
  if (!prof::enabled) return;
  prof::Rec* rec = new prof::Rec{nullptr, nullptr, -1};
  const int lc = a.view_c != 0 ? a.view_c : a.C;     // logical channel count (packed-4 launches report as C = 16)
  const int ci = lc == 16 ? 0 : (lc == 64 ? 1 : 2);
  const int ph = (a.epi == kEpiFwd || a.epi == kEpiFwdBnRes) ? (a.in_scale == nullptr ? 0 : 1) : (a.epi == kEpiBwdMaskStats ? 2 : 3);
  rec->kind = ci * 4 + ph;
  if (cudaEventCreate(&rec->a) != cudaSuccess || cudaEventCreate(&rec->b) != cudaSuccess) { delete rec; return; }
  cudaEventRecord(rec->a, s);
  *out = rec;
}

void pair_profile_record_end(cudaStream_t s, void* p) {
  if (p == nullptr) return;
  prof::Rec* rec = static_cast<prof::Rec*>(p);
  cudaEventRecord(rec->b, s);
  std::lock_guard<std::mutex> lk(prof::mu);
  prof::recs.push_back(*rec);
  delete rec;
}

namespace {

constexpr int KC = 16;      // input channels per weight slab
constexpr int NSTAGE = 3;   // weight ring depth

template <int C> struct PairCfg;
template <> struct PairCfg<128> { static constexpr int TN = 8, TM = 9, IN_MAX = 184; };
template <> struct PairCfg<64>  { static constexpr int TN = 8, TM = 8, IN_MAX = 320; };
template <> struct PairCfg<16>  { static constexpr int TN = 4, TM = 8, IN_MAX = 576; };

template <int C> struct PairDerived {
  using Cfg = PairCfg<C>;
  static constexpr int TN = Cfg::TN, TM = Cfg::TM;
  static constexpr int NT = C / TN;        // threads along channels
  static constexpr int MT = 256 / NT;      // threads along pixels
  static constexpr int M_MAX = MT * TM;    // pixel slots per CTA
  static constexpr int CP = C + 4;         // padded pixel pitch (floats)
  static constexpr int IN_MAX = Cfg::IN_MAX;
  static constexpr size_t SMEM_FLOATS = (size_t)IN_MAX * CP + (size_t)M_MAX * CP + (size_t)NSTAGE * KC * C + 4 * C;
  static constexpr size_t SMEM_BYTES = SMEM_FLOATS * 4 + NSTAGE * 8 + 16;
};

struct TileShape { int TU, TV, TR; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

template <int C, int TM, int TN>
__device__ __forceinline__ void mma_chunk(float (&acc)[TM][TN], const float* __restrict__ As, const int (&aoff)[TM],
                                          const float* __restrict__ Bs, int tn) {
#pragma unroll
  for (int k4 = 0; k4 < KC; k4 += 4) {
    float4 a[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(As + aoff[i] + k4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 b0 = *reinterpret_cast<const float4*>(Bs + (k4 + kk) * C + tn * 4);
      float4 b1 = make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (TN == 8) b1 = *reinterpret_cast<const float4*>(Bs + (k4 + kk) * C + C / 2 + tn * 4);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
        acc[i][0] = fmaf(av, b0.x, acc[i][0]);
        acc[i][1] = fmaf(av, b0.y, acc[i][1]);
        acc[i][2] = fmaf(av, b0.z, acc[i][2]);
        acc[i][3] = fmaf(av, b0.w, acc[i][3]);
        if constexpr (TN == 8) {
          acc[i][4] = fmaf(av, b1.x, acc[i][4]);
          acc[i][5] = fmaf(av, b1.y, acc[i][5]);
          acc[i][6] = fmaf(av, b1.z, acc[i][6]);
          acc[i][7] = fmaf(av, b1.w, acc[i][7]);
        }
      }
    }
  }
}

template <int C>
__global__ void __launch_bounds__(256, C == 16 ? 2 : 1)   // C = 16: 90 KB of shared memory per CTA, two CTAs per SM
pair_kernel(const __grid_constant__ PairArgs a, const TileShape ts) {
  using D = PairDerived<C>;
  constexpr int TN = D::TN, TM = D::TM, NT = D::NT, MT = D::MT, CP = D::CP, NH = TN / 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* in_s = reinterpret_cast<float*>(smem_raw);
  float* mid_s = in_s + (size_t)D::IN_MAX * CP;
  float* w_s = mid_s + (size_t)D::M_MAX * CP;
  float* vec_s = w_s + NSTAGE * KC * C;  // b1 | b2+bad | e_scale | e_shift ... (4*C)
  uint64_t* bars = reinterpret_cast<uint64_t*>(vec_s + 4 * C);

  const int tid = threadIdx.x;
  const int tn = tid % NT, tm = tid / NT;
  const int TU = ts.TU, TV = ts.TV, TR = ts.TR, TVH = TV + 2;
  const int INR = (TU + 2) * TVH;   // in-tile pixels per residue class
  const int MIDR = TU * TVH;        // mid pixels per residue class
  const int d = a.dil;
  const int U = a.vert_first ? a.H : a.W, V = a.vert_first ? a.W : a.H;
  const long su = a.vert_first ? (long)a.W * C : C, sv = a.vert_first ? C : (long)a.W * C;
  const int Ul = (U + d - 1) / d, Vl = (V + d - 1) / d;
  const int tiles_u = (Ul + TU - 1) / TU, tiles_v = (Vl + TV - 1) / TV;
  const int ncb = (d * d + TR - 1) / TR;  // residue-class blocks
  int b = blockIdx.x;
  const int tvi = b % tiles_v; b /= tiles_v;
  const int tui = b % tiles_u; b /= tiles_u;
  const int cb = b % ncb;
  const int n = b / ncb;
  const int ul0 = tui * TU, vl0 = tvi * TV;
  const size_t img = (size_t)n * a.H * a.W * C;

  // ---- weight stream bookkeeping
  const int G1 = 3 * (C / KC);
  const int G = G1 + 3 * (C / KC) + (a.has_adapter ? C / KC : 0);
  constexpr uint32_t SLAB_BYTES = KC * C * 4;
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int g0 = 0; g0 < NSTAGE - 1 && g0 < G; ++g0) {
      mbar_expect_tx(&bars[g0], SLAB_BYTES);
      bulk_g2s(w_s + g0 * KC * C, a.wstream + (size_t)g0 * KC * C, SLAB_BYTES, &bars[g0]);
    }
  }

  // ---- small vectors
  for (int c = tid; c < C; c += 256) {
    vec_s[c] = a.b1 != nullptr ? __ldg(a.b1 + c) : 0.f;
    float b2 = a.b2 != nullptr ? __ldg(a.b2 + c) : 0.f;
    if (a.bad != nullptr) b2 += __ldg(a.bad + c);
    vec_s[C + c] = b2;
  }

  // ---- input tile: coalesced 128-bit loads, BN+ReLU prologue in registers
  {
    const int total = TR * INR * (C / 4);
    for (int idx = tid; idx < total; idx += 256) {
      const int pix = idx / (C / 4), f4 = idx % (C / 4);
      const int r = pix / INR, rem = pix % INR;
      const int iu = rem / TVH, iv = rem % TVH;
      const int cidx = cb * TR + r;
      const int ru = cidx / d, rv = cidx % d;
      const int ul = ul0 - 1 + iu, vl = vl0 - 1 + iv;
      const int u = ru + d * ul, v = rv + d * vl;
      float4 x = make4(0.f);
      if (cidx < d * d && ul >= 0 && vl >= 0 && u < U && v < V) {
        x = ldg4(a.in + img + u * su + v * sv + f4 * 4);
        if (a.in_scale != nullptr) {
          const float4 sc = ldg4(a.in_scale + f4 * 4), sh = ldg4(a.in_shift + f4 * 4);
          x.x = fmaxf(fmaf(x.x, sc.x, sh.x), 0.f);
          x.y = fmaxf(fmaf(x.y, sc.y, sh.y), 0.f);
          x.z = fmaxf(fmaf(x.z, sc.z, sh.z), 0.f);
          x.w = fmaxf(fmaf(x.w, sc.w, sh.w), 0.f);
        }
      }
      *reinterpret_cast<float4*>(in_s + (size_t)pix * CP + f4 * 4) = x;
    }
  }
  __syncthreads();

  int g = 0;  // next weight slab
  auto advance = [&]() -> const float* {
    const int buf = g % NSTAGE;
    mbar_wait(&bars[buf], (uint32_t)((g / NSTAGE) & 1));
    __syncthreads();  // everyone is done with slab g-1 -> its ring slot may be refilled
    const int nxt = g + NSTAGE - 1;
    if (tid == 0 && nxt < G) {
      const int nb = nxt % NSTAGE;
      mbar_expect_tx(&bars[nb], SLAB_BYTES);
      bulk_g2s(w_s + nb * KC * C, a.wstream + (size_t)nxt * KC * C, SLAB_BYTES, &bars[nb]);
    }
    ++g;
    return w_s + buf * KC * C;
  };

  float acc[TM][TN];

  // =================================================================== stage 1: mid = f(conv_A(in'))
  const int M1 = TR * MIDR;
  int aoff[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = i * MT + tm;
    int off = 0;
    if (m < M1) {
      const int r = m / MIDR, rem = m % MIDR;
      off = (r * INR + rem) * CP;  // tap k adds k*TVH pixels
    }
    aoff[i] = off;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k = 0; k < 3; ++k) {
    for (int c0 = 0; c0 < C; c0 += KC) {
      const float* Bs = advance();
      mma_chunk<C, TM, TN>(acc, in_s + (size_t)k * TVH * CP + c0, aoff, Bs, tn);
    }
  }

  // ---- epilogue 1: bias + ReLU (forward) or ReLU-mask multiply (backward); zero outside the image
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = i * MT + tm;
    if (m >= M1) continue;
    const int r = m / MIDR, rem = m % MIDR;
    const int mu = rem / TVH, mv = rem % TVH;
    const int cidx = cb * TR + r;
    const int ru = cidx / d, rv = cidx % d;
    const int ul = ul0 + mu, vl = vl0 - 1 + mv;
    const int u = ru + d * ul, v = rv + d * vl;
    const bool valid = cidx < d * d && vl >= 0 && u < U && v < V;
    const size_t gaddr = img + u * su + v * sv;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const int ch = h * (C / 2) + tn * 4;
      float4 val = make4(0.f);
      if (valid) {
        if (a.mid_mask != nullptr) {
          const float4 mk = ldg4(a.mid_mask + gaddr + ch);
          val.x = mk.x > 0.f ? acc[i][h * 4 + 0] : 0.f;
          val.y = mk.y > 0.f ? acc[i][h * 4 + 1] : 0.f;
          val.z = mk.z > 0.f ? acc[i][h * 4 + 2] : 0.f;
          val.w = mk.w > 0.f ? acc[i][h * 4 + 3] : 0.f;
        } else {
          val.x = fmaxf(acc[i][h * 4 + 0] + vec_s[ch + 0], 0.f);
          val.y = fmaxf(acc[i][h * 4 + 1] + vec_s[ch + 1], 0.f);
          val.z = fmaxf(acc[i][h * 4 + 2] + vec_s[ch + 2], 0.f);
          val.w = fmaxf(acc[i][h * 4 + 3] + vec_s[ch + 3], 0.f);
        }
        if (a.mid_out != nullptr && mv >= 1 && mv <= TV) *reinterpret_cast<float4*>(a.mid_out + gaddr + ch) = val;
      }
      *reinterpret_cast<float4*>(mid_s + (size_t)m * CP + ch) = val;
    }
  }
  // (the __syncthreads inside the next advance() orders these writes before stage 2 reads)

  // =================================================================== stage 2: out = conv_B(mid) + adapter(in')
  const int OUTR = TU * TV;
  const int M2 = TR * OUTR;
  int aoff3[TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = i * MT + tm;
    int off = 0, off3 = 0;
    if (m < M2) {
      const int r = m / OUTR, rem = m % OUTR;
      const int mu = rem / TV, mv = rem % TV;
      off = (r * MIDR + mu * TVH + mv) * CP;                 // tap k adds k pixels
      off3 = (r * INR + (mu + 1) * TVH + mv + 1) * CP;       // centre pixel of the input tile
    }
    aoff[i] = off;
    aoff3[i] = off3;
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k = 0; k < 3; ++k) {
    for (int c0 = 0; c0 < C; c0 += KC) {
      const float* Bs = advance();
      mma_chunk<C, TM, TN>(acc, mid_s + (size_t)k * CP + c0, aoff, Bs, tn);
    }
  }
  if (a.has_adapter) {
    for (int c0 = 0; c0 < C; c0 += KC) {
      const float* Bs = advance();
      mma_chunk<C, TM, TN>(acc, in_s + c0, aoff3, Bs, tn);
    }
  }

  // ---- epilogue 2
  float s1[TN], s2[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = i * MT + tm;
    if (m >= M2) continue;
    const int r = m / OUTR, rem = m % OUTR;
    const int mu = rem / TV, mv = rem % TV;
    const int cidx = cb * TR + r;
    const int ru = cidx / d, rv = cidx % d;
    const int u = ru + d * (ul0 + mu), v = rv + d * (vl0 + mv);
    if (!(cidx < d * d && u < U && v < V)) continue;
    const size_t gaddr = img + u * su + v * sv;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const int ch = h * (C / 2) + tn * 4;
      float val[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) val[j] = acc[i][h * 4 + j] + vec_s[C + ch + j];
      if (a.epi == kEpiFwd) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { s1[h * 4 + j] += val[j]; s2[h * 4 + j] += val[j] * val[j]; }
      } else if (a.epi == kEpiBwdMaskStats) {
        const float4 pv4 = ldg4(a.e0 + gaddr + ch);
        const float pv[4] = {pv4.x, pv4.y, pv4.z, pv4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float mean = __ldg(a.e_stats + ch + j), istd = __ldg(a.e_stats + C + ch + j);
          const float sc = __ldg(a.e_stats + 2 * C + ch + j), sh = __ldg(a.e_stats + 3 * C + ch + j);
          const float q = fmaf(pv[j], sc, sh);
          val[j] = q > 0.f ? val[j] : 0.f;
          s1[h * 4 + j] += val[j];
          s2[h * 4 + j] += val[j] * ((pv[j] - mean) * istd);
        }
      } else {
        const float4 dy4 = ldg4(a.e0 + gaddr + ch), y4 = ldg4(a.e1 + gaddr + ch);
        val[0] += y4.x > 0.f ? dy4.x : 0.f;
        val[1] += y4.y > 0.f ? dy4.y : 0.f;
        val[2] += y4.z > 0.f ? dy4.z : 0.f;
        val[3] += y4.w > 0.f ? dy4.w : 0.f;
      }
      *reinterpret_cast<float4*>(a.out + gaddr + ch) = make_float4(val[0], val[1], val[2], val[3]);
    }
  }

  // ---- per-channel sums: registers -> shared (per pixel-thread row) -> fp64 atomics
  if (a.sums != nullptr) {
    __syncthreads();  // mid_s is free
    float* red = mid_s;  // [MT][2][C]
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const int ch = h * (C / 2) + tn * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        red[(tm * 2 + 0) * C + ch + j] = s1[h * 4 + j];
        red[(tm * 2 + 1) * C + ch + j] = s2[h * 4 + j];
      }
    }
    __syncthreads();
    for (int c = tid; c < 2 * C; c += 256) {
      const int which = c / C, ch = c % C;
      double t = 0.0;
      for (int l = 0; l < MT; ++l) t += (double)red[(l * 2 + which) * C + ch];
      atomicAdd(a.sums + which * C + ch, t);
    }
  }
}

// Choose the lattice tile (TU x TV, TR residue classes per CTA) that minimises the CTA count.
template <int C>
TileShape choose_tile(int Ul, int Vl, int d) {
  using D = PairDerived<C>;
  TileShape best{1, 2, 1};
  long best_ctas = -1, best_load = 0;
  for (int TR = 1; TR <= 8; ++TR) {
    if (TR > d * d) break;
    for (int TU = 1; TU <= 32; ++TU) {
      for (int TV = 2; TV <= 62; TV += 2) {
        const int TVH = TV + 2;
        if (TR * TU * TVH > D::M_MAX) break;
        if (TR * (TU + 2) * TVH > D::IN_MAX) break;
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
int launch_pair_c(const PairArgs& a, cudaStream_t s) {
  using D = PairDerived<C>;
  static_assert(D::SMEM_BYTES <= 227 * 1024, "pair kernel shared memory budget");
  const int d = a.dil;
  const int U = a.vert_first ? a.H : a.W, V = a.vert_first ? a.W : a.H;
  const int Ul = cdiv(U, d), Vl = cdiv(V, d);
  const TileShape ts = choose_tile<C>(Ul, Vl, d);
  const long ctas = (long)a.N * cdiv(d * d, ts.TR) * cdiv(Ul, ts.TU) * cdiv(Vl, ts.TV);
  MDIL_REQUIRE(ctas > 0 && ctas < (1L << 31), "pair: grid size");
  MDIL_CUDA(cudaFuncSetAttribute(pair_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D::SMEM_BYTES));
  pair_kernel<C><<<(unsigned)ctas, 256, D::SMEM_BYTES, s>>>(a, ts);
  MDIL_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int launch_pair(const PairArgs& a_in, cudaStream_t s) {
  static const int trace = getenv("MDIL_TC_TRACE") != nullptr ? 1 : 0;
  PairArgs a = a_in;
  a.trace = trace;
  void* rec = nullptr;
  pair_profile_record_begin(a, s, &rec);
  const bool tc = a.wstream_tc != nullptr && (a.C == 64 || a.C == 128);
  const int rc = !tc ? launch_pair_ffma(a, s) : (pair_impl_mode() == 3 ? launch_pair_tc3(a, s) : launch_pair_h3(a, s));
  pair_profile_record_end(s, rec);
  return rc;
}

int launch_pair_ffma(const PairArgs& a, cudaStream_t s) {
  MDIL_REQUIRE(a.dil >= 1 && a.N > 0 && a.H > 0 && a.W > 0, "pair: bad dims");
  MDIL_REQUIRE(((uintptr_t)a.wstream & 15) == 0, "pair: weight stream must be 16-byte aligned");
  switch (a.C) {
    case 128: return launch_pair_c<128>(a, s);
    case 64: return launch_pair_c<64>(a, s);
    case 16: return launch_pair_c<16>(a, s);
    default: return set_error(-2, "pair: C must be 16, 64 or 128", __FILE__, __LINE__);
  }
}

}  // namespace mdil
