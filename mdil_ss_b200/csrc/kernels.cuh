// Internal launcher interface between api.cu (C ABI, orchestration) and the kernel files.
#pragma once
#include "common.cuh"

namespace mdil {

// ---------------------------------------------------------------- elementwise.cu
int launch_nchw_to_nhwc4(const float* x, float* y, int N, int C, int H, int W, cudaStream_t s);

// sums[0*ldsum + coff + c] += sum x, sums[1*ldsum + coff + c] += sum x^2 over P pixels, channels [coff, coff+cnt)
int launch_channel_stats(const float* x, size_t P, int ld, int coff, int cnt, double* sums, int ldsum, cudaStream_t s);

// stats [4][C] = mean, invstd, scale (=gamma*invstd), shift (=beta-mean*scale).
// train: batch statistics from sums (count pixels) and running-stat update; else running stats.
// fold > 1: sums are [2][fold*C] per (pixel slot, channel) of a packed view and are folded over the slots (ldsum =
// fold*C); stats_rep (nullable) [4][fold*C] receives the statistics replicated per slot.
int launch_bn_finalize(const double* sums, int ldsum, double count, int C, const float* gamma, const float* beta,
                       float* running_mean, float* running_var, float eps, float momentum, int train, float* stats,
                       cudaStream_t s, int fold = 1, float* stats_rep = nullptr, long long* nbt = nullptr);

// y = relu((u*scale+shift) * drop[n][c] + res)   (drop, res nullable)
int launch_bn_act(const float* u, const float* stats, const float* drop, const float* res, float* y, int N, size_t HW,
                  int C, cudaStream_t s);

// dz = dy * (y>0) * drop ; sums[0][c] += sum dz ; sums[1][c] += sum dz*uhat, uhat = (u-mean)*invstd
// (y, drop nullable: dz = dy)
int launch_bn_bwd_stats(const float* dy, const float* y, const float* drop, const float* u, const float* stats,
                        double* sums, int N, size_t HW, int C, cudaStream_t s);

// coef [3][C] = gamma*invstd, sum dz / n, sum dz*uhat / n ; dgamma/dbeta nullable
int launch_bn_bwd_finalize(const double* sums, double count, int C, const float* gamma, const float* stats, float* coef,
                           float* dgamma, float* dbeta, cudaStream_t s, int fold = 1);   // fold: sums are [2][fold*C]

// du = coef0 * (dz - coef1 - uhat*coef2)
// launch_bn_finalize + launch_bn_act in one launch (C <= 128; no replicated statistics)
int launch_bn_act_fused(const float* u, const double* sums, int ldsum, double count, const float* gamma, const float* beta,
                        float* rm, float* rv, float eps, float momentum, int train, int fold, float* stats, const float* drop,
                        const float* res, float* y, int N, size_t HW, int C, cudaStream_t s, long long* nbt = nullptr);
// launch_bn_bwd_finalize + launch_bn_bwd_apply in one launch (C <= 128)
int launch_bn_bwd_apply_fused(const float* dy, const float* y, const float* drop, const float* u, const float* stats,
                              const double* sums, double count, const float* gamma, int fold, float* dgamma, float* dbeta,
                              float* du, int N, size_t HW, int C, cudaStream_t s, int split = 0);
// split != 0: du is written in the "S16" format (per group of 4 channels = 16 bytes: 4 bf16 hi halves, 4 bf16 lo halves; x = hi + lo)
int launch_bn_bwd_apply(const float* dy, const float* y, const float* drop, const float* u, const float* stats,
                        const float* coef, float* du, int N, size_t HW, int C, cudaStream_t s, int split = 0);

// max-pool 2x2 s2 of x[N,H,W,ldin] channels [0,Cin) -> u[N,H/2,W/2,ldu] channels [coff, coff+Cin)
int launch_pool_fwd(const float* x, float* u, int N, int H, int W, int Cin, int ldin, int ldu, int coff, cudaStream_t s);
// dx[N,H,W,Cin] (+)= du routed to the first max of each window; accumulate != 0 adds to dx
int launch_pool_bwd(const float* x, const float* du, float* dx, int N, int H, int W, int Cin, int ldin, int ldu,
                    int coff, int accumulate, cudaStream_t s);

// rep[k][p*C + c] = stats[k][c] for k < 4, p < fold (BatchNorm statistics replicated per pixel slot of a packed view)
int launch_replicate_stats(const float* stats, float* rep, int C, int fold, cudaStream_t s);
int launch_scale(float* x, size_t n, const double* inv_den /*nullable: multiply by 1/(*inv_den)*/, const float* mul,
                 cudaStream_t s);
int launch_adam(float* param, const float* grad, float* m, float* v, size_t n, float lr, float b1, float b2, float eps,
                float wd, int step, float grad_scale, cudaStream_t s);
// state: device float[4] = {steps taken, lr, scratch, scratch}: graph-capturable (see elementwise.cu)
int launch_adam_dev(float* param, const float* grad, float* m, float* v, size_t n, float* state, float b1, float b2, float eps,
                    float wd, float grad_scale, cudaStream_t s);

// ---------------------------------------------------------------- conv_taps.cu
constexpr int kMaxTaps = 9;
constexpr int kMaxClasses = 4;

struct TapClass {
  int ntaps;
  int o_dy, o_dx;  // forward: output coordinate offset of this class
  int a_dy[kMaxTaps], a_dx[kMaxTaps];
  int widx[kMaxTaps];                  // which [CIN][COUT] weight slab
};

// Virtual grid (n, vy, vx). A coordinate = v*a_s + a_d[t]; G coordinate = v*g_s + o_d(class).
struct ConvGeom {
  int N, VH, VW;
  int AH, AW, lda, a_coff, a_sy, a_sx;
  int GH, GW, ldg, g_coff, g_sy, g_sx;
  int CIN, COUT, COUT_PAD;
  int CIN_VALID;  // wgrad: rows ci >= CIN_VALID are padding and not written
  int nclasses;
  TapClass cls[kMaxClasses];
};

// out[g pixel][g_coff+co] = bias[co] + sum_t sum_ci A[a pixel(t)][a_coff+ci] * Wp[widx[t]][ci][co]
int launch_conv_taps(const ConvGeom& g, const float* A, const float* Wp, const float* bias, float* out,
                     cudaStream_t s);

// dW[ci*s_ci + co*s_co + widx*s_t] += sum_pixels A'[a pixel(t)][ci] * G[g pixel][co]   (dW, db zeroed by the caller)
// A' = relu(A*a_scale+a_shift) when a_scale != NULL.  db[co] += sum of G over the tap-0 walk of every class.
int launch_wgrad_taps(const ConvGeom& g, const float* A, const float* a_scale, const float* a_shift, const float* G,
                      float* dW, long s_ci, long s_co, long s_t, float* db, cudaStream_t s);

// dst[(tmap(t)*Apad + a)*Bpad + b] = src[a*sa + b*sb + t*st] for a<A, b<B, else 0; flip: tmap(t)=T-1-t
int launch_pack(const float* src, float* dst, int T, int A, int Apad, int B, int Bpad, long sa, long sb, long st, int flip,
                cudaStream_t s);

// ---------------------------------------------------------------- conv_tc.cu
// The samplers' 3x3 stride-2 convolutions on tcgen05 (fp16 split operands forward, bf16 for gradients): the parity-class
// transposed form (UpsamplerBlock forward, DownsamplerBlock data gradient) and the strided form (DownsamplerBlock forward,
// UpsamplerBlock data gradient) of a ConvGeom.  wimg: 16-bit hi/lo chunk images built by launch_pack_conv_tc from the
// fp32 tap slabs Wp[widx][wp_rows][COUT_PAD] (conv_tc_image_floats(CIN, COUT_PAD) floats);
// sums (nullable, transposed form only): [2][COUT] fp64 sum / sum of squares of the output (zeroed by the caller)
bool conv_tc_ok(const ConvGeom& g, int grad);
size_t conv_tc_image_floats(int cin, int cout_pad);
int launch_pack_conv_tc(const ConvGeom& g, const float* Wp, int wp_rows, void* img, int grad, cudaStream_t s);
int launch_conv_tc(const ConvGeom& g, const float* A, const void* wimg, const float* bias, float* out, double* sums, int grad,
                   cudaStream_t s);

// ---------------------------------------------------------------- nb1d_pair.cu
// kEpiFwdBnRes (eval-mode forward of pair 2, tensor-core kernel only): out = relu((acc + b) * scale + shift + e0) with
// e_stats = the running-statistics BatchNorm [4][C] and e0 = the block input: BN2 + residual + ReLU without another pass
enum PairEpilogue { kEpiFwd = 0, kEpiBwdMaskStats = 1, kEpiBwdResidual = 2, kEpiFwdBnRes = 3 };

struct PairArgs {
  const float* in;         // [N,H,W,C]
  const float* in_scale;   // nullable: prologue v = relu(v*scale+shift)
  const float* in_shift;
  const float* wstream;    // packed [3][C][C] first conv, [3][C][C] second conv, [C][C] adapter (if has_adapter)
  const float* wstream_tc; // tensor-core path: hi/lo SWIZZLE_64B chunk images (16-bit for h3, TF32 for tc3) or NULL
  const float* b1;         // nullable
  const float* b2;         // nullable
  const float* bad;        // nullable (adapter bias)
  const float* mid_mask;   // nullable: mid = acc * (mid_mask > 0) instead of relu(acc + b1)
  float* mid_out;          // nullable: store mid (a / c / dc' / da')
  float* out;
  int epi;
  const float* e0;         // epi1: p            epi2: dy            epi3: block input x (residual)
  const float* e1;         // epi2: y
  const float* e_stats;    // epi1: [4][C] mean, invstd, scale, shift of BN1;  epi3: the same of BN2 (running statistics)
  double* sums;            // nullable: [2][C] (fwd: sum out, sum out^2; epi1: sum out, sum out*phat)
  int N, H, W, C, dil, has_adapter, vert_first;
  int trace;               // debug: CTA 0 prints its phase timestamps (MDIL_TC_TRACE=1)
  int view_c;              // 0, or the logical channel count when the launch runs on a packed view (profiling kinds only)
  int in_split;            // `in` is in the S16 format (per 4 channels: 4 bf16 hi, 4 bf16 lo halves): the loader copies (backward launches)
  int mid_out_split;       // write `mid_out` in the S16 format (backward launches: the operand halves are stored as they are)
};
int launch_pair(const PairArgs& a, cudaStream_t s);          // dispatch: tensor-core kernel when wstream_tc != NULL
int launch_pair_ffma(const PairArgs& a, cudaStream_t s);     // nb1d_pair.cu (FP32 FFMA; all C)
int launch_pair_h3(const PairArgs& a, cudaStream_t s);       // nb1d_pair_h3.cu (tcgen05 kind::f16, fp16/bf16 split operands; default)
int launch_pack_block_h3(const float* const* w6, void* packed, int C, int has_adapter, cudaStream_t s);
// packed-4 view of a C = 16 block without adapter ([N,H,W,16] read as [N,H,W/4,64]: four pixels of a row form one
// 64-"channel" row): the 16-bit hi/lo images of the equivalent 64 x 64 tap matrices (block-diagonal for the 3x1 convs,
// block-banded over the three GROUP taps for the 1x3 convs) + the four conv biases replicated per pixel slot (float [4][64])
int launch_pack_block_p4(const float* const* w4, const float* const* b4, void* packed16, float* bias_rep, cudaStream_t s);
// weight / bias gradients of the packed-4 view: acc [3][64][64] (tap, ci', co') of wgrad_tc<64> -> dW [16][16][3] (torch
// layout [co][ci][k]), dbacc [64] -> db [16].  horizontal: the conv runs along the packing axis (1x3).
struct UnpackP4Item { const float* acc; float* dW; const float* dbacc; float* db; int horizontal; };
struct UnpackP4List { int n; UnpackP4Item item[4]; };
int launch_wgrad_unpack_p4(const UnpackP4List& ul, cudaStream_t s);
int launch_pair_tc3(const PairArgs& a, cudaStream_t s);      // nb1d_pair_tc3.cu (persistent pipelined 3xTF32 tcgen05 kernel; A/B)
int launch_pack_tc3(const float* src_stream, float* dst_stream, int C, int has_adapter, cudaStream_t s);
// one launch: fp32 FFMA streams (write_fp32) and/or 3xTF32 tensor-core images (tc_order 0 = none, 3) of a whole block
int launch_pack_block(const float* const* w6, float* packed, int C, int has_adapter, int write_fp32, int tc_order,
                      cudaStream_t s);
int pair_impl_mode();   // MDIL_PAIR_IMPL: 0 = "ffma", 3 = "tc3" (3xTF32 tensor-core kernel), 4 = 16-bit split tensor-core kernel (default)
void pair_profile_record_begin(const PairArgs& a, cudaStream_t s, void** rec);
void pair_profile_record_end(cudaStream_t s, void* rec);
int pair_profile_begin();
int pair_profile_end(float* total_ms, int* counts, int nkinds);

// ---------------------------------------------------------------- wgrad_tc.cu
struct WgradTcArgs {
  const float* A;          // activation [N,H,W,C]
  const float* a_scale;    // nullable: A' = relu(A*scale+shift)
  const float* a_shift;
  const float* G;          // gradient [N,H,W,C]
  float* dWacc;            // [ntaps][C(ci)][C(co)] fp32, zeroed by the caller (accumulated with red.global.add)
  float* db;               // nullable [C], zeroed by the caller
  int N, H, W, C, dil, ntaps, vert;
  int trace;               // debug: CTA 0 prints its wait counters (MDIL_TC_TRACE=1)
  int g_split;             // G is in the S16 format (per 4 channels: 4 bf16 hi, 4 bf16 lo halves): the producers copy it
  // gathered job (the strided 3x3 / transposed 3x3 convolutions of the samplers; launch_wgrad_gather_tc): the walk is
  // over the virtual grid N x H x W of ConvGeom.  The 64 operand channels are `nslots` slots of `sw` channels; slot k of
  // the activation operand holds channels a_coff..a_coff+a_sw-1 of pixel (y*a_sy + dy_k, x*a_sx + dx_k) of an
  // [AH, AW, lda] tensor, slot k of the gradient operand channels g_coff.. of pixel (y*g_sy + dy_k, x*g_sx + dx_k) of
  // [GH, GW, ldg]; (dy_k + 1) | (dx_k + 1) << 2 is nibble k of a_slots / g_slots.  Several taps of a narrow convolution
  // thus share one job (stacked along M or N); pixels outside their tensor and unused slots contribute zeros.
  int AH, AW, lda, a_coff, a_sy, a_sx, a_sw, a_nslots;
  int GH, GW, ldg, g_coff, g_sy, g_sx, g_sw, g_nslots, g_cout;
  unsigned long long a_slots, g_slots;
};
int launch_wgrad_tc(const WgradTcArgs& a, cudaStream_t s);
// up to three independent jobs (same C) in ONE launch, CTAs split between them: one accumulator flush per CTA instead of three
int launch_wgrad_tc_multi(const WgradTcArgs* a, int n, cudaStream_t s);
// weight gradient of a tap-class convolution (ConvGeom: the samplers' strided / transposed 3x3) on the tensor-core kernel:
// gathered C = 64 jobs (one per tap and 64-channel block of CIN; the taps of a narrow operand stacked), all in ONE launch.  Same contract as
// launch_wgrad_taps, except that dW / db are overwritten (no zeroing by the caller); scratch: wgrad_gather_scratch_floats()
bool wgrad_gather_ok(const ConvGeom& g);
size_t wgrad_gather_scratch_floats();
int launch_wgrad_gather_tc(const ConvGeom& g, const float* A, const float* G, float* dW, long s_ci, long s_co, long s_t,
                           float* db, float* scratch, cudaStream_t s);
int launch_wgrad_unpack(const float* acc, float* dW, int C, int ntaps, long s_ci, long s_co, long s_t, cudaStream_t s);
struct UnpackItem { const float* acc; float* dW; int ntaps; long s_ci, s_co, s_t; const float* dbacc; float* db; };
struct UnpackList { int n; UnpackItem item[6]; };
int launch_wgrad_unpack_multi(const UnpackList& ul, int C, cudaStream_t s);   // one launch for a block's weight gradients

// ---------------------------------------------------------------- head_loss.cu
int launch_outconv_fwd(const float* x, const float* w, const float* bias, float* logits, int N, int H, int W, int Ccls,
                       cudaStream_t s);
int launch_outconv_bwd(const float* dlogits, const float* x, const float* w, float* dx, float* dw, float* db, int N,
                       int H, int W, int Ccls, cudaStream_t s);
int launch_ce2d(const float* logits, const int64_t* labels, const float* class_w, int N, int C, int H, int W,
                float* loss, double* acc, float* dlogits, cudaStream_t s);
// logit gradient of launch_ce2d recomputed from the logits: w[y] * (softmax - onehot) * (*grad_out or 1) / acc[1]
int launch_ce2d_bwd(const float* logits, const int64_t* labels, const float* class_w, int N, int C, int H, int W,
                    const double* acc, const float* grad_out, float* dlogits, cudaStream_t s);
int launch_kd(const float* student, const float* teacher, int N, int C, int H, int W, float* loss, double* acc,
              float* dstudent, cudaStream_t s);
int launch_argmax_confusion(const float* logits, const int64_t* labels, int N, int C, int H, int W, int64_t* pred,
                            long long* conf, cudaStream_t s);

// ---------------------------------------------------------------- cotransform.cu
int launch_cotransform(const unsigned char* img, const unsigned char* lab, int N, int Hs, int Ws, int H, int W, const int* xtab,
                       int KX, const int* ytab, int KY, const int* xnear, const int* ynear, const int* params, int num_classes,
                       float* out_img, long long* out_lab, cudaStream_t s);

}  // namespace mdil
