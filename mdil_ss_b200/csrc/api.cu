// C ABI of libmdil_b200.so (declared in include/mdil_b200.h): argument checking and the per-block
// launch sequences.  No device allocation, no global mutable device state, everything asynchronous
// on the caller's stream.
#include "../../include/mdil_b200.h"
#include "kernels.cuh"

#include <atomic>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mdil {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int set_error(int code, const char* what, const char* file, int line) {
  const char* base = strrchr(file, '/');
  snprintf(g_err, sizeof(g_err), "mdil_b200 error %d: %s (%s:%d)", code, what, base ? base + 1 : file, line);
  return code == 0 ? -1 : code;
}

int pair_impl_mode() {
  static const int mode = [] {
    const char* e = getenv("MDIL_PAIR_IMPL");
    if (e != nullptr && strcmp(e, "ffma") == 0) return 0;
    if (e != nullptr && strcmp(e, "tc3") == 0) return 3;
    return 4;
  }();
  return mode;
}

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

struct Carver {
  unsigned char* base;
  size_t off;
  explicit Carver(void* p) : base(static_cast<unsigned char*>(p)), off(0) {}
  template <typename T> T* take(size_t n) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return r;
  }
};

// 3-tap conv along rows (vert) or columns, dilation d, NHWC [N,H,W,C] -> same shape.
ConvGeom taps3_geom(int N, int H, int W, int C, int d, bool vert) {
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  g.N = N; g.VH = H; g.VW = W;
  g.AH = H; g.AW = W; g.lda = C; g.a_coff = 0; g.a_sy = 1; g.a_sx = 1;
  g.GH = H; g.GW = W; g.ldg = C; g.g_coff = 0; g.g_sy = 1; g.g_sx = 1;
  g.CIN = C; g.COUT = C; g.COUT_PAD = C; g.CIN_VALID = C;
  g.nclasses = 1;
  g.cls[0].ntaps = 3;
  for (int k = 0; k < 3; ++k) {
    g.cls[0].a_dy[k] = vert ? (k - 1) * d : 0;
    g.cls[0].a_dx[k] = vert ? 0 : (k - 1) * d;
    g.cls[0].widx[k] = k;
  }
  return g;
}

ConvGeom pointwise_geom(int N, int H, int W, int C) {
  ConvGeom g = taps3_geom(N, H, W, C, 1, true);
  g.cls[0].ntaps = 1;
  g.cls[0].a_dy[0] = 0; g.cls[0].a_dx[0] = 0; g.cls[0].widx[0] = 0;
  return g;
}

// Sub-pixel parity classes of a 3x3 stride-2 pad-1 (output_padding 1) transposed convolution:
// fine coordinate 2v+p receives kernel index 1 from coarse v (p = 0) or indices 0 / 2 from coarse v+1 / v (p = 1).
void fill_parity_classes(ConvGeom& g) {
  g.nclasses = 4;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      TapClass& tc = g.cls[py * 2 + px];
      tc.o_dy = py; tc.o_dx = px;
      int ny = py ? 2 : 1, nx = px ? 2 : 1;
      const int kys[2] = {py ? 0 : 1, 2}, dys[2] = {py ? 1 : 0, 0};
      const int kxs[2] = {px ? 0 : 1, 2}, dxs[2] = {px ? 1 : 0, 0};
      int t = 0;
      for (int a = 0; a < ny; ++a)
        for (int b = 0; b < nx; ++b) {
          tc.a_dy[t] = dys[a]; tc.a_dx[t] = dxs[b]; tc.widx[t] = kys[a] * 3 + kxs[b];
          ++t;
        }
      tc.ntaps = t;
    }
}

void fill_3x3_taps(ConvGeom& g) {
  g.nclasses = 1;
  TapClass& tc = g.cls[0];
  tc.ntaps = 9; tc.o_dy = 0; tc.o_dx = 0;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) {
      tc.a_dy[ky * 3 + kx] = ky - 1; tc.a_dx[ky * 3 + kx] = kx - 1; tc.widx[ky * 3 + kx] = ky * 3 + kx;
    }
}

int check_nb1d(const mdil_nb1d_desc* d) {
  MDIL_REQUIRE(d != nullptr, "nb1d: null descriptor");
  MDIL_REQUIRE(d->C == 16 || d->C == 64 || d->C == 128, "nb1d: C must be 16, 64 or 128");
  MDIL_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->dil >= 1, "nb1d: bad dims");
  return 0;
}

}  // namespace
}  // namespace mdil

using namespace mdil;

extern "C" {

const char* mdil_version(void) { return "mdil_b200 0.1 (sm_100a)"; }
const char* mdil_last_error_string(void) { return mdil::g_err; }
unsigned long long mdil_launch_count(void) { return mdil::g_launches.load(std::memory_order_relaxed); }

int mdil_profile_begin(void) { return pair_profile_begin(); }
int mdil_profile_end(float* total_ms, int* counts, int nkinds) { return pair_profile_end(total_ms, counts, nkinds); }

int mdil_device_supported(int device) {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}

int mdil_nchw_to_nhwc4(const float* x, float* y, int N, int C, int H, int W, void* stream) {
  return launch_nchw_to_nhwc4(x, y, N, C, H, W, S(stream));
}

// =============================================================================== nb1d
// [0, 28 C^2): fp32 streams of the FFMA kernel; [28 C^2, 84 C^2): hi/lo images of the tensor-core kernel (4 x 14 C^2)
// C = 16: [0, 28 C^2) fp32 streams (FFMA kernel), then the packed-4 images (4 streams x 14 x 64^2 16-bit values) and the
// four conv biases replicated per pixel slot (float [4][64])
static const size_t kP4ImgFloats = (size_t)4 * 14 * 64 * 64 / 2;
size_t mdil_nb1d_packed_floats(int C) { return C == 16 ? (size_t)28 * 256 + kP4ImgFloats + 256 : (size_t)84 * C * C; }

static bool use_tensor_cores(int C) { return pair_impl_mode() != 0 && (C == 64 || C == 128); }
static bool use_tc_wgrad(int C) {
  static const int mode = [] {
    const char* e = getenv("MDIL_WGRAD_IMPL");
    return (e != nullptr && strcmp(e, "ffma") == 0) ? 0 : 1;
  }();
  return mode == 1 && (C == 64 || C == 128);
}
// "S16" format of the backward gradient workspaces (T1 = ds / dp, T2 = dc' / da'): a plane of bf16 hi halves followed by
// a plane of bf16 lo halves (x = hi + lo; the bytes of the fp32 tensor).  Their producers (bn_bwd_apply, the backward
// pair's first epilogue) have the halves at hand or split once; their consumers (the backward pair's loader, the weight-
// gradient producers) copy operand words instead of converting.  Only when every consumer is a tensor-core kernel;
// MDIL_S16=0 keeps fp32 workspaces for A/B measurements.
static bool s16_enabled() {
  static const int on = [] {
    const char* e = getenv("MDIL_S16");
    return (e != nullptr && strcmp(e, "0") == 0) ? 0 : 1;
  }();
  return on == 1 && pair_impl_mode() == 4;
}
static inline const float* tc_stream(const float* packed, int C, int which) {
  if (!use_tensor_cores(C)) return nullptr;
  // tc3: four streams of 14 C^2 floats (hi/lo TF32 images); h3: four streams of 14 C^2 16-bit values (= 7 C^2 floats)
  return packed + (size_t)28 * C * C + (size_t)which * (pair_impl_mode() == 4 ? 7 : 14) * C * C;
}

// Packed-4 view of the C = 16 blocks (decoder non_bottleneck_1d: no adapter, dilation 1): [N,H,W,16] is read as
// [N,H,W/4,64] and the block runs on the C = 64 tensor-core kernels with block-structured 64x64 tap matrices
// (nb1d_pair_h3.cu: pack_block_p4_kernel).  MDIL_P4=0 keeps the FFMA / mma.sync kernels for A/B measurements.
static bool p4_enabled() {
  static const int on = [] {
    const char* e = getenv("MDIL_P4");
    return (e != nullptr && strcmp(e, "0") == 0) ? 0 : 1;
  }();
  return on == 1 && pair_impl_mode() == 4;
}
static bool p4_ok(const mdil_nb1d_desc* d) {
  return d->C == 16 && p4_enabled() && !d->has_adapter && d->dil == 1 && d->W % 4 == 0;
}

size_t mdil_nb1d_fwd_workspace_bytes(const mdil_nb1d_desc* d) {
  const int Cs = d->C == 16 ? 64 : d->C;     // packed-4 view: sums per (pixel slot, channel), two replicated statistics sets
  return 256 + (size_t)4 * Cs * sizeof(double) + (size_t)8 * Cs * sizeof(float) + 1024;
}

size_t mdil_nb1d_bwd_workspace_bytes(const mdil_nb1d_desc* d) {
  size_t T = align_up((size_t)d->N * d->H * d->W * d->C * sizeof(float), 256);
  const int Cw = d->C == 16 ? 64 : d->C;     // packed-4 view: weight-gradient accumulators of the 64-channel view
  return 3 * T + (size_t)4 * Cw * sizeof(double) + (size_t)6 * d->C * sizeof(float) + (size_t)4 * Cw * sizeof(float) +
         ((size_t)18 * Cw * Cw + 6 * Cw) * sizeof(float) + 12 * 256;   // 6 accumulators [3][C][C] + 6 bias sums [C]
}

static inline const float* p4_stream(const float* packed, int which) {
  return packed + (size_t)28 * 256 + (size_t)which * 7 * 64 * 64;
}
static inline const float* p4_bias(const float* packed, int k) { return packed + (size_t)28 * 256 + kP4ImgFloats + (size_t)k * 64; }

int mdil_nb1d_pack(const mdil_nb1d_desc* d, const mdil_nb1d_weights* w, float* packed, void* stream) {
  MDIL_TRY(check_nb1d(d));
  const int C = d->C;
  const long CC = (long)C * C;
  cudaStream_t s = S(stream);
  // streams (28 C^2 floats): [which][7][ci or co][co or ci], which = fwd pair 1, fwd pair 2, bwd pair 2, bwd pair 1;
  // forward slab[k][ci][co] = W[co][ci][k], backward (second conv first, taps flipped) slab[k'][co][ci] = W[co][ci][2-k'];
  // tensor-core images (56 C^2 floats) are written by the same launch
  if (d->has_adapter) MDIL_REQUIRE(w->wp1 != nullptr && w->wp2 != nullptr, "nb1d: adapter weights missing");
  const float* w6[6] = {w->w31_1, w->w13_1, w->w31_2, w->w13_2, w->wp1, w->wp2};
  const bool tc = use_tensor_cores(C);
  (void)CC;
  if (tc && pair_impl_mode() == 4) return launch_pack_block_h3(w6, packed + (size_t)28 * C * C, C, d->has_adapter, s);
  MDIL_TRY(launch_pack_block(w6, packed, C, d->has_adapter, tc ? 0 : 1, tc ? 3 : 0, s));
  if (C == 16 && p4_enabled() && !d->has_adapter) {   // both forms: the launch picks the packed-4 path per shape (W % 4, dil)
    const float* w4[4] = {w->w31_1, w->w13_1, w->w31_2, w->w13_2};
    const float* b4[4] = {w->b31_1, w->b13_1, w->b31_2, w->b13_2};
    MDIL_TRY(launch_pack_block_p4(w4, b4, packed + (size_t)28 * 256, packed + (size_t)28 * 256 + kP4ImgFloats, s));
  }
  return 0;
}

// C = 16 block on the packed-4 view (see p4_ok): same sequence as mdil_nb1d_fwd with C = 64, W/4 pair launches; the
// per-channel BatchNorm sums arrive per (pixel slot, channel) and are folded by the finalize kernel
static int nb1d_fwd_p4(const mdil_nb1d_desc* d, const float* x, const mdil_nb1d_weights* w, const float* packed,
                       const float* drop_mask, float* y, const mdil_nb1d_saved* sv, void* ws, cudaStream_t s) {
  const int C = 16, CP = 64;
  const size_t HW = (size_t)d->H * d->W;
  const double count = (double)d->N * (double)HW;
  Carver cv(ws);
  double* sums1 = cv.take<double>(4 * CP);     // [2][2][64]
  double* sums2 = sums1 + 2 * CP;
  float* rep1 = cv.take<float>(4 * CP);         // BN1 statistics replicated per pixel slot [4][64]
  float* st1 = sv->stats;
  float* st2 = sv->stats + 4 * C;
  if (d->train) MDIL_CUDA(cudaMemsetAsync(sums1, 0, 4 * CP * sizeof(double), s));
  PairArgs a;
  memset(&a, 0, sizeof(a));
  a.N = d->N; a.H = d->H; a.W = d->W / 4; a.C = CP; a.has_adapter = 0; a.vert_first = 1; a.epi = kEpiFwd; a.dil = 1; a.view_c = C;
  a.in = x; a.wstream_tc = p4_stream(packed, 0); a.b1 = p4_bias(packed, 0); a.b2 = p4_bias(packed, 1);
  a.mid_out = d->save ? sv->a : nullptr; a.out = sv->p; a.sums = d->train ? sums1 : nullptr;
  if (!d->train) {
    // eval mode: both BatchNorms use running statistics, known before the block runs: pair 1, then pair 2 whose epilogue
    // applies BN2 + residual + ReLU (no s tensor, no separate elementwise pass)
    float* rep2 = cv.take<float>(4 * CP);
    MDIL_TRY(launch_bn_finalize(sums1, CP, count, C, w->bn1.weight, w->bn1.bias, w->bn1.running_mean, w->bn1.running_var,
                                d->eps, d->momentum, 0, st1, s, 4, rep1));
    MDIL_TRY(launch_bn_finalize(sums2, CP, count, C, w->bn2.weight, w->bn2.bias, w->bn2.running_mean, w->bn2.running_var,
                                d->eps, d->momentum, 0, st2, s, 4, rep2));
    MDIL_TRY(launch_pair(a, s));
    a.in = sv->p; a.in_scale = rep1 + 2 * CP; a.in_shift = rep1 + 3 * CP; a.wstream_tc = p4_stream(packed, 1);
    a.b1 = p4_bias(packed, 2); a.b2 = p4_bias(packed, 3);
    a.mid_out = nullptr; a.out = y; a.sums = nullptr; a.epi = kEpiFwdBnRes; a.e0 = x; a.e_stats = rep2;
    return launch_pair(a, s);
  }
  MDIL_TRY(launch_pair(a, s));
  MDIL_TRY(launch_bn_finalize(sums1, CP, count, C, w->bn1.weight, w->bn1.bias, w->bn1.running_mean, w->bn1.running_var,
                              d->eps, d->momentum, d->train, st1, s, 4, rep1, w->bn1.num_batches_tracked));
  a.in = sv->p; a.in_scale = rep1 + 2 * CP; a.in_shift = rep1 + 3 * CP; a.wstream_tc = p4_stream(packed, 1);
  a.b1 = p4_bias(packed, 2); a.b2 = p4_bias(packed, 3);
  a.mid_out = d->save ? sv->c : nullptr; a.out = sv->s; a.sums = d->train ? sums2 : nullptr;
  MDIL_TRY(launch_pair(a, s));
  // y = relu(bn2(s) * drop + x), the BatchNorm finalisation in the same launch
  return launch_bn_act_fused(sv->s, sums2, CP, count, w->bn2.weight, w->bn2.bias, w->bn2.running_mean, w->bn2.running_var,
                             d->eps, d->momentum, d->train, 4, st2, drop_mask, x, y, d->N, HW, C, s, w->bn2.num_batches_tracked);
}

int mdil_nb1d_fwd(const mdil_nb1d_desc* d, const float* x, const mdil_nb1d_weights* w, const float* packed,
                  const float* drop_mask, float* y, const mdil_nb1d_saved* sv, void* ws, size_t ws_bytes, void* stream) {
  MDIL_TRY(check_nb1d(d));
  MDIL_REQUIRE(ws_bytes >= mdil_nb1d_fwd_workspace_bytes(d), "nb1d_fwd: workspace too small");
  MDIL_REQUIRE(sv != nullptr && sv->p != nullptr && sv->s != nullptr && sv->stats != nullptr, "nb1d_fwd: p/s/stats buffers required");
  if (p4_ok(d)) return nb1d_fwd_p4(d, x, w, packed, drop_mask, y, sv, ws, S(stream));
  const int C = d->C;
  const long CC = (long)C * C;
  const size_t HW = (size_t)d->H * d->W;
  const double count = (double)d->N * (double)HW;
  cudaStream_t s = S(stream);
  Carver cv(ws);
  double* sums1 = cv.take<double>(4 * C);     // [2][2][C]: both pairs' sums, one memset
  double* sums2 = sums1 + 2 * C;
  float* st1 = sv->stats;
  float* st2 = sv->stats + 4 * C;
  if (d->train) MDIL_CUDA(cudaMemsetAsync(sums1, 0, 4 * C * sizeof(double), s));
  PairArgs a;
  memset(&a, 0, sizeof(a));
  a.N = d->N; a.H = d->H; a.W = d->W; a.C = C; a.has_adapter = d->has_adapter; a.vert_first = 1; a.epi = kEpiFwd;
  // pair 1: x -> a -> p
  a.in = x; a.wstream = packed; a.wstream_tc = tc_stream(packed, C, 0); a.b1 = w->b31_1; a.b2 = w->b13_1; a.bad = d->has_adapter ? w->bp1 : nullptr;
  a.mid_out = d->save ? sv->a : nullptr; a.out = sv->p; a.sums = d->train ? sums1 : nullptr; a.dil = 1;
  if (!d->train && use_tensor_cores(C) && pair_impl_mode() == 4) {
    // eval mode (teacher forward of steps 2/3, validation: train_new_task_step2.py:291,398-438): both BatchNorms use
    // running statistics, known before the block runs: pair 1, then pair 2 whose epilogue applies BN2 + residual + ReLU
    // -- two fused launches and 5 tensor passes (R x, W p | R p, R x, W y) instead of five launches and 7
    MDIL_TRY(launch_bn_finalize(sums1, C, count, C, w->bn1.weight, w->bn1.bias, w->bn1.running_mean, w->bn1.running_var,
                                d->eps, d->momentum, 0, st1, s));
    MDIL_TRY(launch_bn_finalize(sums2, C, count, C, w->bn2.weight, w->bn2.bias, w->bn2.running_mean, w->bn2.running_var,
                                d->eps, d->momentum, 0, st2, s));
    MDIL_TRY(launch_pair(a, s));
    a.in = sv->p; a.in_scale = st1 + 2 * C; a.in_shift = st1 + 3 * C; a.wstream = packed + 7 * CC; a.wstream_tc = tc_stream(packed, C, 1);
    a.b1 = w->b31_2; a.b2 = w->b13_2; a.bad = d->has_adapter ? w->bp2 : nullptr;
    a.mid_out = nullptr; a.out = y; a.sums = nullptr; a.dil = d->dil; a.epi = kEpiFwdBnRes; a.e0 = x; a.e_stats = st2;
    return launch_pair(a, s);
  }
  MDIL_TRY(launch_pair(a, s));
  MDIL_TRY(launch_bn_finalize(sums1, C, count, C, w->bn1.weight, w->bn1.bias, w->bn1.running_mean, w->bn1.running_var,
                              d->eps, d->momentum, d->train, st1, s, 1, nullptr, w->bn1.num_batches_tracked));
  // pair 2: r = relu(bn1(p)) -> c -> s
  a.in = sv->p; a.in_scale = st1 + 2 * C; a.in_shift = st1 + 3 * C; a.wstream = packed + 7 * CC; a.wstream_tc = tc_stream(packed, C, 1);
  a.b1 = w->b31_2; a.b2 = w->b13_2; a.bad = d->has_adapter ? w->bp2 : nullptr;
  a.mid_out = d->save ? sv->c : nullptr; a.out = sv->s; a.sums = d->train ? sums2 : nullptr; a.dil = d->dil;
  MDIL_TRY(launch_pair(a, s));
  // y = relu(bn2(s) * drop + x), the BatchNorm finalisation in the same launch
  return launch_bn_act_fused(sv->s, sums2, C, count, w->bn2.weight, w->bn2.bias, w->bn2.running_mean, w->bn2.running_var,
                             d->eps, d->momentum, d->train, 1, st2, drop_mask, x, y, d->N, HW, C, s, w->bn2.num_batches_tracked);
}

// One weight gradient of the block: tensor-core path (C = 64, 128) or the generic FFMA tap kernel.
// taps: 3 (vertical when vert != 0, else horizontal, dilation d) or 1 (1x1 adapter).
// (tensor-core path: `acc_scratch` = this gradient's own zeroed [taps][C][C] slot; its unpack is deferred to `ul`)
struct WgradJobList { int n; WgradTcArgs job[3]; };   // tensor-core weight gradients waiting for their shared launch
static int flush_wgrad_jobs(WgradJobList* jl, cudaStream_t s) {
  const int n = jl->n;
  jl->n = 0;
  return launch_wgrad_tc_multi(jl->job, n, s);
}

static int nb1d_wgrad(const mdil_nb1d_desc* d, int dil, bool vert, int taps, const float* A, const float* sc,
                      const float* sh, const float* G, int g_split, float* dW, float* db, float* acc_scratch, float* db_scratch,
                      UnpackList* ul, WgradJobList* jl, cudaStream_t s) {
  const int C = d->C;
  if (dW == nullptr) {
    MDIL_REQUIRE(db == nullptr, "nb1d_bwd: bias gradient without weight gradient is not supported");
    return 0;
  }
  const long s_ci = taps, s_co = (long)C * taps, s_t = taps == 1 ? 0 : 1;   // torch layout [co][ci][taps]
  if (!use_tc_wgrad(C) && db != nullptr) MDIL_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * C, s));
  if (use_tc_wgrad(C)) {
    WgradTcArgs w;
    memset(&w, 0, sizeof(w));
    w.A = A; w.a_scale = sc; w.a_shift = sh; w.G = G; w.dWacc = acc_scratch; w.db = db != nullptr ? db_scratch : nullptr;
    w.N = d->N; w.H = d->H; w.W = d->W; w.C = C; w.dil = dil; w.ntaps = taps; w.vert = vert ? 1 : 0; w.g_split = g_split;
    jl->job[jl->n++] = w;        // launched together with the pair's other weight gradients (flush_wgrad_jobs)
    UnpackItem& it = ul->item[ul->n++];
    it.acc = acc_scratch; it.dW = dW; it.ntaps = taps; it.s_ci = s_ci; it.s_co = s_co; it.s_t = s_t;
    it.dbacc = db_scratch; it.db = db;
    return 0;
  }
  MDIL_REQUIRE(!g_split, "nb1d_bwd: S16 gradients need the tensor-core weight-gradient kernel");
  MDIL_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)C * C * taps, s));
  ConvGeom g = taps == 1 ? pointwise_geom(d->N, d->H, d->W, C) : taps3_geom(d->N, d->H, d->W, C, dil, vert);
  return launch_wgrad_taps(g, A, sc, sh, G, dW, s_ci, s_co, s_t, db, s);
}

// one weight gradient of the packed-4 view through wgrad_tc<64>; its [3][64][64] accumulator is folded back by `ul`
static int nb1d_wgrad_p4(const mdil_nb1d_desc* d, bool vert, const float* A, const float* sc, const float* sh, const float* G,
                         int g_split, float* dW, float* db, float* acc, float* bacc, UnpackP4List* ul, WgradJobList* jl) {
  if (dW == nullptr) {
    MDIL_REQUIRE(db == nullptr, "nb1d_bwd: bias gradient without weight gradient is not supported");
    return 0;
  }
  WgradTcArgs w;
  memset(&w, 0, sizeof(w));
  w.A = A; w.a_scale = sc; w.a_shift = sh; w.G = G; w.dWacc = acc; w.db = db != nullptr ? bacc : nullptr;
  w.N = d->N; w.H = d->H; w.W = d->W / 4; w.C = 64; w.dil = 1; w.ntaps = 3; w.vert = vert ? 1 : 0; w.g_split = g_split;
  jl->job[jl->n++] = w;
  UnpackP4Item& it = ul->item[ul->n++];
  it.acc = acc; it.dW = dW; it.dbacc = bacc; it.db = db; it.horizontal = vert ? 0 : 1;
  return 0;
}

static int nb1d_bwd_p4(const mdil_nb1d_desc* d, const float* dy, const float* x, const float* y, const mdil_nb1d_weights* w,
                       const float* packed, const float* drop_mask, const mdil_nb1d_saved* sv, float* dx,
                       const mdil_nb1d_grads* gr, void* ws, cudaStream_t s) {
  const int C = 16, CP = 64, N = d->N, H = d->H, W = d->W;
  const size_t HW = (size_t)H * W;
  const size_t T = (size_t)N * HW * C;
  const double count = (double)N * (double)HW;
  Carver cv(ws);
  double* sums2 = cv.take<double>(2 * C + 2 * CP);   // BN2 backward sums [2][16], then BN1's per (slot, channel) [2][64]
  double* sums1 = sums2 + 2 * C;
  (void)cv.take<float>(6 * C);     // (coefficient buffers of the unfused BatchNorm backward: layout kept)
  float* rep1 = cv.take<float>(4 * CP);
  float* T1 = cv.take<float>(T);
  float* T2 = cv.take<float>(T);
  float* T3 = cv.take<float>(T);
  const size_t WS = (size_t)3 * CP * CP;
  float* wacc = cv.take<float>(4 * WS + 4 * CP);     // four accumulators [3][64][64], then four bias sums [64]
  float* bacc = wacc + 4 * WS;
  UnpackP4List ul;
  ul.n = 0;
  MDIL_CUDA(cudaMemsetAsync(wacc, 0, sizeof(float) * (4 * WS + 4 * CP), s));
  MDIL_CUDA(cudaMemsetAsync(sums2, 0, (2 * C + 2 * CP) * sizeof(double), s));
  const float* st1 = sv->stats;
  const float* st2 = sv->stats + 4 * C;
  MDIL_TRY(launch_replicate_stats(st1, rep1, C, 4, s));

  // ---- BN2 backward (+ ReLU mask of y, dropout): ds  (elementwise kernels on the 16-channel view)
  MDIL_TRY(launch_bn_bwd_stats(dy, y, drop_mask, sv->s, st2, sums2, N, HW, C, s));
  const int s16 = s16_enabled() ? 1 : 0;
  MDIL_TRY(launch_bn_bwd_apply_fused(dy, y, drop_mask, sv->s, st2, sums2, count, w->bn2.weight, 1, gr->bn2_w, gr->bn2_b, T1, N,
                                     HW, C, s, s16));

  // ---- pair 2 backward: ds -> dc' -> dq (masked by r>0), sums for BN1 backward
  PairArgs a;
  memset(&a, 0, sizeof(a));
  a.in_split = s16; a.mid_out_split = s16;
  a.N = N; a.H = H; a.W = W / 4; a.C = CP; a.has_adapter = 0; a.vert_first = 0; a.dil = 1; a.view_c = C;
  a.in = T1; a.wstream_tc = p4_stream(packed, 2); a.mid_mask = sv->c; a.mid_out = T2; a.out = T3;
  a.epi = kEpiBwdMaskStats; a.e0 = sv->p; a.e_stats = rep1; a.sums = sums1;
  MDIL_TRY(launch_pair(a, s));
  WgradJobList jl;
  jl.n = 0;
  MDIL_TRY(nb1d_wgrad_p4(d, false, sv->c, nullptr, nullptr, T1, s16, gr->w13_2, gr->b13_2, wacc + 0 * WS, bacc + 0 * CP, &ul, &jl));
  MDIL_TRY(nb1d_wgrad_p4(d, true, sv->p, rep1 + 2 * CP, rep1 + 3 * CP, T2, s16, gr->w31_2, gr->b31_2, wacc + 1 * WS, bacc + 1 * CP, &ul, &jl));
  MDIL_TRY(flush_wgrad_jobs(&jl, s));

  // ---- BN1 backward: dq -> dp (overwrites ds)
  MDIL_TRY(launch_bn_bwd_apply_fused(T3, nullptr, nullptr, sv->p, st1, sums1, count, w->bn1.weight, 4, gr->bn1_w, gr->bn1_b, T1,
                                     N, HW, C, s, s16));

  // ---- pair 1 backward: dp -> da' -> dx (+ residual dy * (y>0))
  a.in = T1; a.wstream_tc = p4_stream(packed, 3); a.mid_mask = sv->a; a.mid_out = T2; a.out = dx;
  a.epi = kEpiBwdResidual; a.e0 = dy; a.e1 = y; a.e_stats = nullptr; a.sums = nullptr;
  MDIL_TRY(launch_pair(a, s));
  MDIL_TRY(nb1d_wgrad_p4(d, false, sv->a, nullptr, nullptr, T1, s16, gr->w13_1, gr->b13_1, wacc + 2 * WS, bacc + 2 * CP, &ul, &jl));
  MDIL_TRY(nb1d_wgrad_p4(d, true, x, nullptr, nullptr, T2, s16, gr->w31_1, gr->b31_1, wacc + 3 * WS, bacc + 3 * CP, &ul, &jl));
  MDIL_TRY(flush_wgrad_jobs(&jl, s));
  return launch_wgrad_unpack_p4(ul, s);
}

int mdil_nb1d_bwd(const mdil_nb1d_desc* d, const float* dy, const float* x, const float* y, const mdil_nb1d_weights* w,
                  const float* packed, const float* drop_mask, const mdil_nb1d_saved* sv, float* dx,
                  const mdil_nb1d_grads* gr, void* ws, size_t ws_bytes, void* stream) {
  MDIL_TRY(check_nb1d(d));
  MDIL_REQUIRE(d->train, "nb1d_bwd: backward is only implemented for train-mode BatchNorm (batch statistics)");
  MDIL_REQUIRE(ws_bytes >= mdil_nb1d_bwd_workspace_bytes(d), "nb1d_bwd: workspace too small");
  MDIL_REQUIRE(sv != nullptr && sv->a && sv->p && sv->c && sv->s && sv->stats, "nb1d_bwd: saved tensors missing");
  MDIL_REQUIRE(dx != nullptr && gr != nullptr, "nb1d_bwd: dx/grads required");
  if (p4_ok(d)) return nb1d_bwd_p4(d, dy, x, y, w, packed, drop_mask, sv, dx, gr, ws, S(stream));
  const int C = d->C, N = d->N, H = d->H, W = d->W;
  const long CC = (long)C * C;
  const size_t HW = (size_t)H * W;
  const size_t T = (size_t)N * HW * C;
  const double count = (double)N * (double)HW;
  cudaStream_t s = S(stream);
  Carver cv(ws);
  double* sums2 = cv.take<double>(4 * C);     // [2][2][C]: both BatchNorms' backward sums, one memset
  double* sums1 = sums2 + 2 * C;
  (void)cv.take<float>(6 * C);     // (coefficient buffers of the unfused BatchNorm backward: layout kept)
  float* T1 = cv.take<float>(T);
  float* T2 = cv.take<float>(T);
  float* T3 = cv.take<float>(T);
  float* wacc = cv.take<float>((size_t)18 * C * C + 6 * C);  // six accumulators [3][C][C], then six bias sums [C]
  const size_t WS = (size_t)3 * C * C;
  float* bacc = wacc + 6 * WS;
  UnpackList ul;
  ul.n = 0;
  if (use_tc_wgrad(C)) MDIL_CUDA(cudaMemsetAsync(wacc, 0, sizeof(float) * (18 * (size_t)C * C + 6 * C), s));
  const float* st1 = sv->stats;
  const float* st2 = sv->stats + 4 * C;
  MDIL_CUDA(cudaMemsetAsync(sums2, 0, 4 * C * sizeof(double), s));

  // ---- BN2 backward (+ ReLU mask of y, dropout): ds
  MDIL_TRY(launch_bn_bwd_stats(dy, y, drop_mask, sv->s, st2, sums2, N, HW, C, s));
  const int s16 = (s16_enabled() && use_tensor_cores(C) && use_tc_wgrad(C)) ? 1 : 0;
  MDIL_TRY(launch_bn_bwd_apply_fused(dy, y, drop_mask, sv->s, st2, sums2, count, w->bn2.weight, 1, gr->bn2_w, gr->bn2_b, T1, N,
                                     HW, C, s, s16));
  static const int stop = [] { const char* dbg = getenv("MDIL_DEBUG_STOP"); return dbg ? atoi(dbg) : 0; }();   // tools/debug_nb1d.py
  if (stop == 1) return 0;

  // ---- pair 2 backward: ds -> dc' -> dq (masked by r>0), sums for BN1 backward
  PairArgs a;
  memset(&a, 0, sizeof(a));
  a.in_split = s16; a.mid_out_split = s16;
  a.N = N; a.H = H; a.W = W; a.C = C; a.has_adapter = d->has_adapter; a.vert_first = 0;
  a.in = T1; a.wstream = packed + 14 * CC; a.wstream_tc = tc_stream(packed, C, 2); a.mid_mask = sv->c; a.mid_out = T2; a.out = T3;
  a.epi = kEpiBwdMaskStats; a.e0 = sv->p; a.e_stats = st1; a.sums = sums1; a.dil = d->dil;
  MDIL_TRY(launch_pair(a, s));
  if (stop == 2) return 0;

  // ---- weight gradients of pair 2
  WgradJobList jl;
  jl.n = 0;
  MDIL_TRY(nb1d_wgrad(d, d->dil, false, 3, sv->c, nullptr, nullptr, T1, s16, gr->w13_2, gr->b13_2, wacc + 0 * WS, bacc + 0 * C, &ul, &jl, s));
  if (d->has_adapter)
    MDIL_TRY(nb1d_wgrad(d, 1, true, 1, sv->p, st1 + 2 * C, st1 + 3 * C, T1, s16, gr->wp2, gr->bp2, wacc + 1 * WS, bacc + 1 * C, &ul, &jl, s));
  MDIL_TRY(nb1d_wgrad(d, d->dil, true, 3, sv->p, st1 + 2 * C, st1 + 3 * C, T2, s16, gr->w31_2, gr->b31_2, wacc + 2 * WS, bacc + 2 * C, &ul, &jl, s));
  MDIL_TRY(flush_wgrad_jobs(&jl, s));     // the pair's three weight gradients: one launch (T1 is overwritten below)

  if (stop == 3) return 0;
  // ---- BN1 backward: dq -> dp (overwrites ds)
  MDIL_TRY(launch_bn_bwd_apply_fused(T3, nullptr, nullptr, sv->p, st1, sums1, count, w->bn1.weight, 1, gr->bn1_w, gr->bn1_b, T1,
                                     N, HW, C, s, s16));

  // ---- pair 1 backward: dp -> da' -> dx (+ residual dy * (y>0))
  a.in = T1; a.wstream = packed + 21 * CC; a.wstream_tc = tc_stream(packed, C, 3); a.mid_mask = sv->a; a.mid_out = T2; a.out = dx;
  a.epi = kEpiBwdResidual; a.e0 = dy; a.e1 = y; a.e_stats = nullptr; a.sums = nullptr; a.dil = 1;
  MDIL_TRY(launch_pair(a, s));

  // ---- weight gradients of pair 1
  MDIL_TRY(nb1d_wgrad(d, 1, false, 3, sv->a, nullptr, nullptr, T1, s16, gr->w13_1, gr->b13_1, wacc + 3 * WS, bacc + 3 * C, &ul, &jl, s));
  if (d->has_adapter) MDIL_TRY(nb1d_wgrad(d, 1, true, 1, x, nullptr, nullptr, T1, s16, gr->wp1, gr->bp1, wacc + 4 * WS, bacc + 4 * C, &ul, &jl, s));
  MDIL_TRY(nb1d_wgrad(d, 1, true, 3, x, nullptr, nullptr, T2, s16, gr->w31_1, gr->b31_1, wacc + 5 * WS, bacc + 5 * C, &ul, &jl, s));
  MDIL_TRY(flush_wgrad_jobs(&jl, s));
  return launch_wgrad_unpack_multi(ul, C, s);
}

// =============================================================================== downsampler
static inline int down_cinp(const mdil_down_desc* d) { return d->ldin; }
static inline int down_cconv(const mdil_down_desc* d) { return d->Cout - d->Cin; }
static inline int down_coutp(const mdil_down_desc* d) { return (down_cconv(d) + 3) / 4 * 4; }

static size_t down_slab_floats(const mdil_down_desc* d) {
  return ((size_t)9 * down_cinp(d) * down_coutp(d) + (size_t)9 * down_cconv(d) * d->Cin + 64 + 3) / 4 * 4;   // images: 16-byte aligned
}
static ConvGeom down_fwd_geom(const mdil_down_desc* d) {
  const int OH = d->H / 2, OW = d->W / 2;
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  g.N = d->N; g.VH = OH; g.VW = OW;
  g.AH = d->H; g.AW = d->W; g.lda = d->ldin; g.a_coff = 0; g.a_sy = 2; g.a_sx = 2;
  g.GH = OH; g.GW = OW; g.ldg = d->Cout; g.g_coff = 0; g.g_sy = 1; g.g_sx = 1;
  g.CIN = down_cinp(d); g.COUT = down_cconv(d); g.COUT_PAD = down_coutp(d); g.CIN_VALID = d->Cin;
  fill_3x3_taps(g);
  return g;
}
// data gradient of the strided conv = a parity-class transposed conv from du (first Cc channels) to dx
static ConvGeom down_dgrad_geom(const mdil_down_desc* d) {
  const int OH = d->H / 2, OW = d->W / 2, Cc = down_cconv(d);
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  g.N = d->N; g.VH = OH; g.VW = OW;
  g.AH = OH; g.AW = OW; g.lda = d->Cout; g.a_coff = 0; g.a_sy = 1; g.a_sx = 1;
  g.GH = d->H; g.GW = d->W; g.ldg = d->Cin; g.g_coff = 0; g.g_sy = 2; g.g_sx = 2;
  g.CIN = Cc; g.COUT = d->Cin; g.COUT_PAD = d->Cin; g.CIN_VALID = Cc;
  fill_parity_classes(g);
  return g;
}
// [fp32 tap slabs of the forward conv and of the data gradient][16-bit chunk images of the tensor-core kernel: forward,
// data gradient]
static size_t down_img_fwd_floats(const mdil_down_desc* d) { return conv_tc_image_floats(down_cinp(d), down_coutp(d)); }
size_t mdil_down_packed_floats(const mdil_down_desc* d) {
  return down_slab_floats(d) + down_img_fwd_floats(d) + conv_tc_image_floats(down_cconv(d), d->Cin);
}

size_t mdil_down_workspace_bytes(const mdil_down_desc* d) {
  size_t du = align_up((size_t)d->N * (d->H / 2) * (d->W / 2) * d->Cout * sizeof(float), 256);
  return du + (size_t)2 * d->Cout * sizeof(double) + (size_t)3 * d->Cout * sizeof(float) + 4 * 256 +
         align_up(wgrad_gather_scratch_floats() * sizeof(float), 256);
}

static int check_down(const mdil_down_desc* d) {
  MDIL_REQUIRE(d != nullptr && d->N > 0 && d->H > 0 && d->W > 0 && d->H % 2 == 0 && d->W % 2 == 0, "down: bad dims");
  MDIL_REQUIRE(d->ldin % 4 == 0 && d->ldin >= d->Cin && d->Cout % 4 == 0 && d->Cout > d->Cin, "down: bad channels");
  MDIL_REQUIRE(256 % (d->Cout / 4) == 0, "down: unsupported Cout");
  return 0;
}

int mdil_down_pack(const mdil_down_desc* d, const float* w, float* packed, void* stream) {
  MDIL_TRY(check_down(d));
  const int Cin = d->Cin, Cc = down_cconv(d), CinP = down_cinp(d), CoP = down_coutp(d);
  cudaStream_t s = S(stream);
  // forward: slab[t][ci][co] = W[co][ci][t]
  MDIL_TRY(launch_pack(w, packed, 9, Cin, CinP, Cc, CoP, 9, 9L * Cin, 1, 0, s));
  // dgrad: slab[t][co][ci] = W[co][ci][t]
  if (Cin % 4 == 0 && Cc % 4 == 0) {
    MDIL_TRY(launch_pack(w, packed + (size_t)9 * CinP * CoP, 9, Cc, Cc, Cin, Cin, 9L * Cin, 9, 1, 0, s));
    if (d->ldin == Cin) {
      const ConvGeom g = down_dgrad_geom(d);
      if (conv_tc_ok(g, 1))
        MDIL_TRY(launch_pack_conv_tc(g, packed + (size_t)9 * CinP * CoP, Cc, packed + down_slab_floats(d) + down_img_fwd_floats(d), 1, s));
    }
  }
  const ConvGeom gf = down_fwd_geom(d);
  if (conv_tc_ok(gf, 0)) MDIL_TRY(launch_pack_conv_tc(gf, packed, CinP, packed + down_slab_floats(d), 0, s));
  return 0;
}

static int bn_forward_tail(const float* u, size_t P, int C, const mdil_bn_params* bn, int train, float eps, float momentum,
                           double* sums, float* stats, float* y, int N, size_t HW, cudaStream_t s, bool have_sums = false) {
  if (train && !have_sums) {     // have_sums: the producing kernel's epilogue accumulated them
    MDIL_CUDA(cudaMemsetAsync(sums, 0, 2 * C * sizeof(double), s));
    MDIL_TRY(launch_channel_stats(u, P, C, 0, C, sums, C, s));
  }
  return launch_bn_act_fused(u, sums, C, (double)P, bn->weight, bn->bias, bn->running_mean, bn->running_var, eps, momentum,
                             train, 1, stats, nullptr, nullptr, y, N, HW, C, s, bn->num_batches_tracked);
}

int mdil_down_fwd(const mdil_down_desc* d, const float* x, const float* packed, const float* bias,
                  const mdil_bn_params* bn, float* u, float* stats, float* y, void* ws, size_t ws_bytes, void* stream) {
  MDIL_TRY(check_down(d));
  MDIL_REQUIRE(ws_bytes >= (size_t)2 * d->Cout * sizeof(double) + 512, "down_fwd: workspace too small");
  cudaStream_t s = S(stream);
  const int OH = d->H / 2, OW = d->W / 2, Cc = down_cconv(d);
  const ConvGeom g = down_fwd_geom(d);
  if (conv_tc_ok(g, 0)) MDIL_TRY(launch_conv_tc(g, x, packed + down_slab_floats(d), bias, u, nullptr, 0, s));
  else MDIL_TRY(launch_conv_taps(g, x, packed, bias, u, s));
  MDIL_TRY(launch_pool_fwd(x, u, d->N, d->H, d->W, d->Cin, d->ldin, d->Cout, Cc, s));
  Carver cv(ws);
  double* sums = cv.take<double>(2 * d->Cout);
  return bn_forward_tail(u, (size_t)d->N * OH * OW, d->Cout, bn, d->train, d->eps, d->momentum, sums, stats, y, d->N,
                         (size_t)OH * OW, s);
}

int mdil_down_bwd(const mdil_down_desc* d, const float* dy, const float* x, const float* u, const float* y,
                  const float* stats, const float* packed, const mdil_bn_params* bn, float* dx, float* dw, float* db,
                  float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream) {
  MDIL_TRY(check_down(d));
  MDIL_REQUIRE(d->train, "down_bwd: backward is only implemented for train-mode BatchNorm");
  MDIL_REQUIRE(ws_bytes >= mdil_down_workspace_bytes(d), "down_bwd: workspace too small");
  cudaStream_t s = S(stream);
  const int OH = d->H / 2, OW = d->W / 2, Cc = down_cconv(d), Cout = d->Cout, Cin = d->Cin;
  const size_t OHW = (size_t)OH * OW;
  Carver cv(ws);
  double* sums = cv.take<double>(2 * Cout);
  (void)cv.take<float>(3 * Cout);
  float* du = cv.take<float>((size_t)d->N * OHW * Cout);
  MDIL_CUDA(cudaMemsetAsync(sums, 0, 2 * Cout * sizeof(double), s));
  MDIL_TRY(launch_bn_bwd_stats(dy, y, nullptr, u, stats, sums, d->N, OHW, Cout, s));
  MDIL_TRY(launch_bn_bwd_apply_fused(dy, y, nullptr, u, stats, sums, (double)d->N * OHW, bn->weight, 1, dgamma, dbeta, du, d->N,
                                     OHW, Cout, s));
  if (dw != nullptr) {
    ConvGeom g;
    memset(&g, 0, sizeof(g));
    g.N = d->N; g.VH = OH; g.VW = OW;
    g.AH = d->H; g.AW = d->W; g.lda = d->ldin; g.a_coff = 0; g.a_sy = 2; g.a_sx = 2;
    g.GH = OH; g.GW = OW; g.ldg = Cout; g.g_coff = 0; g.g_sy = 1; g.g_sx = 1;
    g.CIN = down_cinp(d); g.COUT = Cc; g.COUT_PAD = down_coutp(d); g.CIN_VALID = Cin;
    fill_3x3_taps(g);
    if (wgrad_gather_ok(g)) {     // 64 -> 128: nine gathered one-tap jobs on the tensor-core kernel
      MDIL_TRY(launch_wgrad_gather_tc(g, x, du, dw, 9, 9L * Cin, 1, db, cv.take<float>(wgrad_gather_scratch_floats()), s));
    } else {
      MDIL_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cc * Cin * 9, s));
      if (db != nullptr) MDIL_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Cc, s));
      MDIL_TRY(launch_wgrad_taps(g, x, nullptr, nullptr, du, dw, 9, 9L * Cin, 1, db, s));  // [co][ci][3][3]
    }
  } else {
    MDIL_REQUIRE(db == nullptr, "down_bwd: bias gradient without weight gradient is not supported");
  }
  if (dx != nullptr) {
    MDIL_REQUIRE(Cin % 4 == 0 && Cc % 4 == 0 && d->ldin == Cin, "down_bwd: dx needs Cin % 4 == 0");
    const ConvGeom g = down_dgrad_geom(d);
    if (conv_tc_ok(g, 1))
      MDIL_TRY(launch_conv_tc(g, du, packed + down_slab_floats(d) + down_img_fwd_floats(d), nullptr, dx, nullptr, 1, s));
    else MDIL_TRY(launch_conv_taps(g, du, packed + (size_t)9 * down_cinp(d) * down_coutp(d), nullptr, dx, s));
    MDIL_TRY(launch_pool_bwd(x, du, dx, d->N, d->H, d->W, Cin, d->ldin, Cout, Cc, 1, s));
  }
  return 0;
}

// =============================================================================== upsampler
static size_t up_slab_floats(const mdil_up_desc* d) { return (size_t)18 * d->Cin * d->Cout + 64; }
static size_t up_img_fwd_floats(const mdil_up_desc* d) { return conv_tc_image_floats(d->Cin, d->Cout); }
// [fp32 tap slabs of the forward conv and of the data gradient][16-bit chunk images of the tensor-core kernel: forward,
// data gradient]
size_t mdil_up_packed_floats(const mdil_up_desc* d) {
  return up_slab_floats(d) + up_img_fwd_floats(d) + conv_tc_image_floats(d->Cout, d->Cin);
}
// data gradient of the transposed conv = a strided 3x3 conv from du to dx
static ConvGeom up_dgrad_geom(const mdil_up_desc* d) {
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  g.N = d->N; g.VH = d->H; g.VW = d->W;
  g.AH = 2 * d->H; g.AW = 2 * d->W; g.lda = d->Cout; g.a_coff = 0; g.a_sy = 2; g.a_sx = 2;
  g.GH = d->H; g.GW = d->W; g.ldg = d->Cin; g.g_coff = 0; g.g_sy = 1; g.g_sx = 1;
  g.CIN = d->Cout; g.COUT = d->Cin; g.COUT_PAD = d->Cin; g.CIN_VALID = d->Cout;
  fill_3x3_taps(g);
  return g;
}

size_t mdil_up_workspace_bytes(const mdil_up_desc* d) {
  size_t du = align_up((size_t)d->N * (2 * d->H) * (2 * d->W) * d->Cout * sizeof(float), 256);
  return du + (size_t)2 * d->Cout * sizeof(double) + (size_t)3 * d->Cout * sizeof(float) + 4 * 256 +
         align_up(wgrad_gather_scratch_floats() * sizeof(float), 256);
}

static int check_up(const mdil_up_desc* d) {
  MDIL_REQUIRE(d != nullptr && d->N > 0 && d->H > 0 && d->W > 0, "up: bad dims");
  MDIL_REQUIRE(d->Cin % 4 == 0 && d->Cout % 4 == 0 && 256 % (d->Cout / 4) == 0, "up: bad channels");
  return 0;
}

static ConvGeom up_parity_geom(const mdil_up_desc* d) {
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  g.N = d->N; g.VH = d->H; g.VW = d->W;
  g.AH = d->H; g.AW = d->W; g.lda = d->Cin; g.a_coff = 0; g.a_sy = 1; g.a_sx = 1;
  g.GH = 2 * d->H; g.GW = 2 * d->W; g.ldg = d->Cout; g.g_coff = 0; g.g_sy = 2; g.g_sx = 2;
  g.CIN = d->Cin; g.COUT = d->Cout; g.COUT_PAD = d->Cout; g.CIN_VALID = d->Cin;
  fill_parity_classes(g);
  return g;
}

int mdil_up_pack(const mdil_up_desc* d, const float* w, float* packed, void* stream) {
  MDIL_TRY(check_up(d));
  const int Cin = d->Cin, Cout = d->Cout;
  cudaStream_t s = S(stream);
  // forward: slab[t][ci][co] = W[ci][co][t]   (torch ConvTranspose2d layout [Cin][Cout][3][3])
  MDIL_TRY(launch_pack(w, packed, 9, Cin, Cin, Cout, Cout, 9L * Cout, 9, 1, 0, s));
  // dgrad: slab[t][co][ci] = W[ci][co][t]
  MDIL_TRY(launch_pack(w, packed + (size_t)9 * Cin * Cout, 9, Cout, Cout, Cin, Cin, 9, 9L * Cout, 1, 0, s));
  const ConvGeom g = up_parity_geom(d);
  if (conv_tc_ok(g, 0)) MDIL_TRY(launch_pack_conv_tc(g, packed, Cin, packed + up_slab_floats(d), 0, s));
  const ConvGeom gd = up_dgrad_geom(d);
  if (conv_tc_ok(gd, 1))
    MDIL_TRY(launch_pack_conv_tc(gd, packed + (size_t)9 * Cin * Cout, Cout, packed + up_slab_floats(d) + up_img_fwd_floats(d), 1, s));
  return 0;
}

int mdil_up_fwd(const mdil_up_desc* d, const float* x, const float* packed, const float* bias, const mdil_bn_params* bn,
                float* u, float* stats, float* y, void* ws, size_t ws_bytes, void* stream) {
  MDIL_TRY(check_up(d));
  MDIL_REQUIRE(ws_bytes >= (size_t)2 * d->Cout * sizeof(double) + 512, "up_fwd: workspace too small");
  cudaStream_t s = S(stream);
  ConvGeom g = up_parity_geom(d);
  Carver cv(ws);
  double* sums = cv.take<double>(2 * d->Cout);
  const size_t OHW = (size_t)4 * d->H * d->W;
  const bool tc = conv_tc_ok(g, 0);
  if (tc) {     // tensor-core kernel; its epilogue accumulates the BatchNorm sums of u
    if (d->train) MDIL_CUDA(cudaMemsetAsync(sums, 0, 2 * d->Cout * sizeof(double), s));
    MDIL_TRY(launch_conv_tc(g, x, packed + up_slab_floats(d), bias, u, d->train ? sums : nullptr, 0, s));
  } else {
    MDIL_TRY(launch_conv_taps(g, x, packed, bias, u, s));
  }
  return bn_forward_tail(u, (size_t)d->N * OHW, d->Cout, bn, d->train, d->eps, d->momentum, sums, stats, y, d->N, OHW, s, tc);
}

int mdil_up_bwd(const mdil_up_desc* d, const float* dy, const float* x, const float* u, const float* y,
                const float* stats, const float* packed, const mdil_bn_params* bn, float* dx, float* dw, float* db,
                float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream) {
  MDIL_TRY(check_up(d));
  MDIL_REQUIRE(d->train, "up_bwd: backward is only implemented for train-mode BatchNorm");
  MDIL_REQUIRE(ws_bytes >= mdil_up_workspace_bytes(d), "up_bwd: workspace too small");
  cudaStream_t s = S(stream);
  const int Cin = d->Cin, Cout = d->Cout;
  const size_t OHW = (size_t)4 * d->H * d->W;
  Carver cv(ws);
  double* sums = cv.take<double>(2 * Cout);
  (void)cv.take<float>(3 * Cout);
  float* du = cv.take<float>((size_t)d->N * OHW * Cout);
  MDIL_CUDA(cudaMemsetAsync(sums, 0, 2 * Cout * sizeof(double), s));
  MDIL_TRY(launch_bn_bwd_stats(dy, y, nullptr, u, stats, sums, d->N, OHW, Cout, s));
  MDIL_TRY(launch_bn_bwd_apply_fused(dy, y, nullptr, u, stats, sums, (double)d->N * OHW, bn->weight, 1, dgamma, dbeta, du, d->N,
                                     OHW, Cout, s));
  if (dw != nullptr) {
    ConvGeom g = up_parity_geom(d);
    if (wgrad_gather_ok(g)) {     // 128 -> 64, 64 -> 16: gathered one-tap jobs on the tensor-core kernel
      MDIL_TRY(launch_wgrad_gather_tc(g, x, du, dw, 9L * Cout, 9, 1, db, cv.take<float>(wgrad_gather_scratch_floats()), s));
    } else {
      MDIL_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cin * Cout * 9, s));
      if (db != nullptr) MDIL_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Cout, s));
      MDIL_TRY(launch_wgrad_taps(g, x, nullptr, nullptr, du, dw, 9L * Cout, 9, 1, db, s));  // [ci][co][3][3]
    }
  } else {
    MDIL_REQUIRE(db == nullptr, "up_bwd: bias gradient without weight gradient is not supported");
  }
  if (dx != nullptr) {
    const ConvGeom g = up_dgrad_geom(d);
    if (conv_tc_ok(g, 1))
      MDIL_TRY(launch_conv_tc(g, du, packed + up_slab_floats(d) + up_img_fwd_floats(d), nullptr, dx, nullptr, 1, s));
    else
      MDIL_TRY(launch_conv_taps(g, du, packed + (size_t)9 * Cin * Cout, nullptr, dx, s));
  }
  return 0;
}

// =============================================================================== head + losses
int mdil_outconv_fwd(const float* x, const float* w, const float* bias, float* logits, int N, int H, int W, int Ccls,
                     void* stream) {
  return launch_outconv_fwd(x, w, bias, logits, N, H, W, Ccls, S(stream));
}

int mdil_outconv_bwd(const float* dlogits, const float* x, const float* w, float* dx, float* dw, float* db, int N,
                     int H, int W, int Ccls, void* stream) {
  return launch_outconv_bwd(dlogits, x, w, dx, dw, db, N, H, W, Ccls, S(stream));
}

int mdil_ce2d_fwd_bwd(const float* logits, const int64_t* labels, const float* class_w, int N, int C, int H, int W,
                      float* loss, double* acc, float* dlogits, void* stream) {
  return launch_ce2d(logits, labels, class_w, N, C, H, W, loss, acc, dlogits, S(stream));
}

int mdil_ce2d_bwd(const float* logits, const int64_t* labels, const float* class_w, int N, int C, int H, int W,
                  const double* acc, const float* grad_out, float* dlogits, void* stream) {
  MDIL_REQUIRE(logits != nullptr && labels != nullptr && class_w != nullptr && acc != nullptr && dlogits != nullptr,
               "ce2d_bwd: null argument");
  return launch_ce2d_bwd(logits, labels, class_w, N, C, H, W, acc, grad_out, dlogits, S(stream));
}

int mdil_ce2d_scale(float* dlogits, size_t n, const double* acc, const float* grad_out, void* stream) {
  return launch_scale(dlogits, n, acc + 1, grad_out, S(stream));
}

int mdil_kd_fwd_bwd(const float* student, const float* teacher, int N, int C, int H, int W, float* loss, double* acc,
                    float* dstudent, void* stream) {
  return launch_kd(student, teacher, N, C, H, W, loss, acc, dstudent, S(stream));
}

int mdil_scale_by_device_scalar(float* x, size_t n, const float* grad_out, void* stream) {
  return launch_scale(x, n, nullptr, grad_out, S(stream));
}

int mdil_argmax_confusion(const float* logits, const int64_t* labels, int N, int C, int H, int W, int64_t* pred,
                          long long* conf, void* stream) {
  return launch_argmax_confusion(logits, labels, N, C, H, W, pred, conf, S(stream));
}

int mdil_cotransform(const unsigned char* img, const unsigned char* lab, int N, int Hs, int Ws, int H, int W, const int* xtab,
                     int KX, const int* ytab, int KY, const int* xnear, const int* ynear, const int* params, int num_classes,
                     float* out_img, int64_t* out_lab, void* stream) {
  return launch_cotransform(img, lab, N, Hs, Ws, H, W, xtab, KX, ytab, KY, xnear, ynear, params, num_classes, out_img,
                            reinterpret_cast<long long*>(out_lab), S(stream));
}

int mdil_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
  return launch_adam(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, S(stream));
}

int mdil_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float* state, float beta1,
                       float beta2, float eps, float weight_decay, float grad_scale, void* stream) {
  return launch_adam_dev(param, grad, exp_avg, exp_avg_sq, n, state, beta1, beta2, eps, weight_decay, grad_scale, S(stream));
}

}  // extern "C"
