/*
 * mdil_b200.h — C ABI of libmdil_b200.so, the B200 (sm_100a) implementation of the
 * MDIL-SS hot path: ERFNet with parallel residual adapters (forward/backward) plus
 * the CrossEntropy2d and output-distillation losses.
 *
 * The reference (prachigarg23/MDIL-SS) is pure Python/PyTorch and has no FFI of its
 * own; the interface each entry point replaces is therefore the torch.nn call chain
 * of the reference module named beside it (paths relative to the reference root).
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's allocator);
 *     the library never allocates or frees device memory and keeps no global
 *     mutable device state (safe under nn.DataParallel's per-device threads);
 *   - activations are NHWC fp32 (channels innermost), except the network input
 *     (NCHW fp32, as the reference's ToTensor yields) and the logits / dlogits
 *     (NCHW fp32, what the reference's losses and argmax consume);
 *   - weights are passed in PyTorch's own layouts ([Cout,Cin,kh,kw] for Conv2d,
 *     [Cin,Cout,kh,kw] for ConvTranspose2d); kernel-friendly copies live in a
 *     caller-owned "packed" buffer refreshed by the *_pack entry points;
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it;
 *   - return value 0 = success, otherwise a cudaError_t or a negative library code;
 *     mdil_last_error_string() (thread-local) describes the last failure.
 */
#ifndef MDIL_B200_H
#define MDIL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDIL_ERR_BAD_ARG (-1)
#define MDIL_ERR_UNSUPPORTED (-2)
#define MDIL_ERR_WORKSPACE (-3)

const char* mdil_version(void);
const char* mdil_last_error_string(void);
/* Number of kernels this library has launched in this process (host-side counter). */
unsigned long long mdil_launch_count(void);
/* Opt-in timing of the fused nb1d pair kernel (bench.py's roofline leg): between begin and end every launch is
 * bracketed by CUDA events on its own stream; end() synchronises them and returns per-kind totals (HOST arrays).
 * kind = {C=16,64,128} * 4 + {forward pair 1, forward pair 2, backward pair 2, backward pair 1}. */
int mdil_profile_begin(void);
int mdil_profile_end(float* total_ms_host, int* counts_host, int nkinds);
/* 1 when the running device is compute capability 10.x (the only target built). */
int mdil_device_supported(int device);

/* ------------------------------------------------------------------ layout */
/* Net.forward entry (models/erfnet_RA_parallel.py:207-210): NCHW [N,C,H,W] (C<=4)
 * -> NHWC with 4 channels (zero padded). */
int mdil_nchw_to_nhwc4(const float* x_nchw, float* y_nhwc4, int N, int C, int H, int W, void* stream);

/* ------------------------------------------------- BatchNorm2d(eps=1e-3) state */
typedef struct {
  const float* weight;   /* gamma [C] */
  const float* bias;     /* beta  [C] */
  float* running_mean;   /* [C], updated in place when train != 0 */
  float* running_var;    /* [C] */
  long long* num_batches_tracked; /* nullable device scalar (nn.BatchNorm2d's buffer): += 1 by a train-mode forward launch */
} mdil_bn_params;

/* ------------------------------------------------------------- nb1d block */
/* non_bottleneck_1d_RAP.forward (models/erfnet_RA_parallel.py:90-113) and
 * non_bottleneck_1d.forward (:48-64, has_adapter = 0). */
typedef struct {
  int N, H, W, C;     /* activation [N,H,W,C]; C in {16,64,128} */
  int dil;            /* dilation of the second factorised pair */
  int has_adapter;    /* parallel_conv_1/2[task] present */
  int train;          /* batch-statistics BN + running-stat update; 0 = running stats */
  int save;           /* keep a,p,c,s for backward */
  float eps;          /* 1e-3 */
  float momentum;     /* 0.1 */
} mdil_nb1d_desc;

typedef struct {
  const float *w31_1, *b31_1;   /* conv3x1_1 [C,C,3,1] */
  const float *w13_1, *b13_1;   /* conv1x3_1 [C,C,1,3] */
  const float *w31_2, *b31_2;   /* conv3x1_2 (dilated) */
  const float *w13_2, *b13_2;   /* conv1x3_2 (dilated) */
  const float *wp1, *bp1;       /* parallel_conv_1[task] [C,C,1,1] or NULL */
  const float *wp2, *bp2;       /* parallel_conv_2[task] or NULL */
  mdil_bn_params bn1, bn2;      /* bns_1[task] / bn1, bns_2[task] / bn2 */
} mdil_nb1d_weights;

typedef struct {
  float *a, *p, *c, *s;   /* [N,H,W,C] each: relu(conv3x1_1), pre-BN1, relu(conv3x1_2), pre-BN2 (a,c may be NULL when !save) */
  float *stats;           /* [8,C]: mean1, invstd1, scale1, shift1, mean2, invstd2, scale2, shift2 */
} mdil_nb1d_saved;

typedef struct {
  float *w31_1, *b31_1, *w13_1, *b13_1, *w31_2, *b31_2, *w13_2, *b13_2; /* NULL = not needed */
  float *wp1, *bp1, *wp2, *bp2;
  float *bn1_w, *bn1_b, *bn2_w, *bn2_b;
} mdil_nb1d_grads;

size_t mdil_nb1d_packed_floats(int C);            /* size of the packed-weight buffer */
size_t mdil_nb1d_fwd_workspace_bytes(const mdil_nb1d_desc*);
size_t mdil_nb1d_bwd_workspace_bytes(const mdil_nb1d_desc*);
int mdil_nb1d_pack(const mdil_nb1d_desc*, const mdil_nb1d_weights*, float* packed, void* stream);
/* drop_mask: Dropout2d noise [N,C] already divided by (1-p), or NULL. */
int mdil_nb1d_fwd(const mdil_nb1d_desc*, const float* x, const mdil_nb1d_weights*, const float* packed,
                  const float* drop_mask, float* y, const mdil_nb1d_saved*, void* ws, size_t ws_bytes, void* stream);
int mdil_nb1d_bwd(const mdil_nb1d_desc*, const float* dy, const float* x, const float* y, const mdil_nb1d_weights*,
                  const float* packed, const float* drop_mask, const mdil_nb1d_saved*, float* dx,
                  const mdil_nb1d_grads*, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------- DownsamplerBlock */
/* models/erfnet_RA_parallel.py:13-25: relu(bn(cat[conv3x3 s2 p1, maxpool2x2])). */
typedef struct {
  int N, H, W;        /* INPUT spatial size */
  int Cin, Cout;      /* Cin real input channels (3,16,64); conv produces Cout-Cin */
  int ldin;           /* channel stride of x (4 for the padded image, else Cin) */
  int train, save;
  float eps, momentum;
} mdil_down_desc;

size_t mdil_down_packed_floats(const mdil_down_desc*);
size_t mdil_down_workspace_bytes(const mdil_down_desc*);
int mdil_down_pack(const mdil_down_desc*, const float* w /*[Cout-Cin,Cin,3,3]*/, float* packed, void* stream);
/* u: pre-BN concat [N,H/2,W/2,Cout] (saved), stats: [4,Cout] mean,invstd,scale,shift */
int mdil_down_fwd(const mdil_down_desc*, const float* x, const float* packed, const float* bias, const mdil_bn_params*,
                  float* u, float* stats, float* y, void* ws, size_t ws_bytes, void* stream);
/* dx may be NULL (initial block: the image needs no gradient). dw/db/dgamma/dbeta may be NULL. */
int mdil_down_bwd(const mdil_down_desc*, const float* dy, const float* x, const float* u, const float* y,
                  const float* stats, const float* packed, const mdil_bn_params*, float* dx, float* dw, float* db,
                  float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream);

/* --------------------------------------------------------- UpsamplerBlock */
/* models/erfnet_RA_parallel.py:152-162: relu(bn(ConvTranspose2d(3, s2, p1, op1))). */
typedef struct {
  int N, H, W;        /* INPUT spatial size; output is 2H x 2W */
  int Cin, Cout;
  int train, save;
  float eps, momentum;
} mdil_up_desc;

size_t mdil_up_packed_floats(const mdil_up_desc*);
size_t mdil_up_workspace_bytes(const mdil_up_desc*);
int mdil_up_pack(const mdil_up_desc*, const float* w /*[Cin,Cout,3,3]*/, float* packed, void* stream);
int mdil_up_fwd(const mdil_up_desc*, const float* x, const float* packed, const float* bias, const mdil_bn_params*,
                float* u, float* stats, float* y, void* ws, size_t ws_bytes, void* stream);
int mdil_up_bwd(const mdil_up_desc*, const float* dy, const float* x, const float* u, const float* y,
                const float* stats, const float* packed, const mdil_bn_params*, float* dx, float* dw, float* db,
                float* dgamma, float* dbeta, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------ output_conv */
/* Decoder.output_conv = ConvTranspose2d(16, Ccls, 2, stride 2) (:179-180,188).
 * x NHWC [N,H,W,16] -> logits NCHW [N,Ccls,2H,2W]. w is [16,Ccls,2,2]. */
int mdil_outconv_fwd(const float* x, const float* w, const float* bias, float* logits, int N, int H, int W, int Ccls,
                     void* stream);
/* dx NHWC [N,H,W,16] (may be NULL); dw [16,Ccls,2,2] and db [Ccls] (may be NULL) are OVERWRITTEN. */
int mdil_outconv_bwd(const float* dlogits, const float* x, const float* w, float* dx, float* dw, float* db, int N,
                     int H, int W, int Ccls, void* stream);

/* ----------------------------------------------------------------- losses */
/* CrossEntropyLoss2d (train_new_task_step2.py:84-92): class-weighted mean NLL of
 * log_softmax(dim=1).  One pass: reads logits NCHW + int64 labels [N,H,W], writes
 * acc[0]=sum w*nll, acc[1]=sum w (double, caller-zeroed is NOT required), loss[0]=
 * acc0/acc1 and dlogits = w[y]*(softmax-onehot)/acc1 * (*grad_scale if non-NULL at
 * mdil_ce2d_scale time). */
int mdil_ce2d_fwd_bwd(const float* logits, const int64_t* labels, const float* class_w, int N, int C, int H, int W,
                      float* loss, double* acc /*[2]*/, float* dlogits /*may be NULL*/, void* stream);
/* Two-phase form (what the autograd function of the host layer uses): mdil_ce2d_fwd_bwd with dlogits = NULL reads the
 * logits once and writes only loss / acc; mdil_ce2d_bwd then recomputes the softmax from the same logits and writes
 * dlogits = w[y] * (softmax - onehot) * (*grad_out) / acc[1] (grad_out: device scalar, may be NULL = 1; acc may have been
 * all-reduced over the ranks in between: SURVEY 8e).  Replaces the backward of CrossEntropyLoss2d,
 * train_new_task_step2.py:84-92. */
int mdil_ce2d_bwd(const float* logits, const int64_t* labels, const float* class_w, int N, int C, int H, int W,
                  const double* acc /*[2]*/, const float* grad_out, float* dlogits, void* stream);
/* dlogits *= (*grad_out) / acc[1]  (grad_out: device scalar, may be NULL = 1). */
int mdil_ce2d_scale(float* dlogits, size_t n, const double* acc, const float* grad_out, void* stream);

/* Output distillation (train_new_task_step2.py:241,296-297):
 * KLDivLoss(reduction='mean')(softmax(student), softmax(teacher)) with probabilities as
 * the input: mean(xlogy(T,T) - T*S). dstudent (may be NULL) receives d loss / d student logits. */
int mdil_kd_fwd_bwd(const float* student, const float* teacher, int N, int C, int H, int W, float* loss,
                    double* acc /*[1]*/, float* dstudent, void* stream);
/* x *= (*grad_out) (device scalar). */
int mdil_scale_by_device_scalar(float* x, size_t n, const float* grad_out, void* stream);

/* ------------------------------------------------------ validation (next row) */
/* argmax over C of NCHW logits (first max wins, as torch.max) -> int64 [N,H,W], and the
 * Ccls x Ccls confusion matrix conf[gt][pred] += 1 (int64, caller-zeroed) that
 * iouEval.addBatch (iouEval.py:21-70) reduces to tp/fp/fn. labels/conf may be NULL. */
int mdil_argmax_confusion(const float* logits, const int64_t* labels, int N, int C, int H, int W, int64_t* pred,
                          long long* conf, void* stream);

/* ----------------------------------------------- input co-transform (next row) */
/* MyCoTransform.__call__ (train_new_task_step2.py:48-81) + ToTensor / ToLabel / Relabel(255, C-1) (transform.py:63-79)
 * for a batch of uint8 source images [N,Hs,Ws,3] and labels [N,Hs,Ws]: Pillow's BILINEAR (image) / NEAREST (label)
 * resize to H x W, optional horizontal flip, translation by (tx, ty) in -2..2 with the reference's fill rules.
 * xtab [W][2+KX] / ytab [H][2+KY]: per output column / row the first source index, the tap count and Pillow's 22-bit
 * coefficients; xnear [W] / ynear [H]: source index of the NEAREST resize; params [N][3] = hflip, tx, ty or NULL (no
 * augmentation).  All tables are device int32 arrays built by the host (mdil_ss_b200/cotransform.py).
 * out_img float [N,3,H,W] in [0,1], out_lab int64 [N,1,H,W].  Bit-exact with the reference. */
int mdil_cotransform(const unsigned char* img, const unsigned char* lab, int N, int Hs, int Ws, int H, int W, const int* xtab,
                     int KX, const int* ytab, int KY, const int* xnear, const int* ynear, const int* params, int num_classes,
                     float* out_img, int64_t* out_lab, void* stream);

/* ------------------------------------------------------- optimiser (next row) */
/* Adam as the drivers configure it (train_new_task_step2.py:237-239): L2 weight decay
 * folded into the gradient, bias correction.  Flat multi-tensor form over n floats. */
int mdil_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

/* The same update with the step counter and the learning rate in DEVICE memory (state: float[4] = steps taken so far, lr,
 * two scratch words), so that an optimiser step captured in a CUDA graph advances its own bias corrections on replay. */
int mdil_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float* state, float beta1,
                       float beta2, float eps, float weight_decay, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDIL_B200_H */
