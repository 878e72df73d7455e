/*
 * ganslate_b200 -- C ABI of the B200 (sm_100a) hot path.
 *
 * The reference (ganslate-team/ganslate) is pure Python/PyTorch and has no FFI
 * of its own: every hot-path op is a torch.nn library call that dispatches to
 * cuDNN / ATen.  Each entry point below therefore cites the reference call
 * site whose library call it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C: raw device pointers + sizes, no torch types.
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as
 *     void*); no implicit synchronisation, no default-stream use.
 *   - the library owns no device memory: outputs, workspaces and saved
 *     statistics are allocated by the caller.
 *   - return 0 on success, non-zero on error; gb_last_error() returns a
 *     thread-local message.  Nothing throws across the ABI.
 *   - activations are channels-last (N, D, H, W, C) bf16 "views": pointer to
 *     the interior origin + element strides, so that a reflection-padded
 *     buffer and its interior are the same allocation.  C is the physical
 *     channel count (multiple of 8); channel stride is 1.
 */
#ifndef GANSLATE_B200_H
#define GANSLATE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB_VERSION 100
#define GB_MAX_TAPS 128
#define GB_MAX_CLASSES 8

/* channels-last bf16 (or fp32 where stated) tensor view */
typedef struct gb_view {
  void* ptr;                 /* element (n=0,z=0,y=0,x=0,c=0) of the interior */
  int64_t sn, sz, sy, sx;    /* element strides; channel stride is 1 */
  int32_t N, D, H, W, C;     /* interior extents; C = physical channels (%8==0) */
  int32_t pad;               /* reflection border (in y and x) materialised around the interior */
} gb_view;

enum { GB_ACT_NONE = 0, GB_ACT_RELU = 1, GB_ACT_LEAKY = 2, GB_ACT_TANH = 3, GB_ACT_PRELU = 4 };

/* One parity class of an implicit-GEMM "data" convolution.
 * Rows of the GEMM enumerate a q-grid (n,qz,qy,qx); the output pixel is
 * q*out_mul + off, the gathered input pixel for tap t is q*in_mul + d[t]. */
typedef struct gb_conv_class {
  int32_t off[3];            /* output offset (z,y,x) of this class */
  int32_t ntaps;             /* taps contributing to this class */
  int32_t tap_begin;         /* first entry in gb_conv_params.taps */
  int32_t kpad;              /* padded K (multiple of 64) of this class' weight matrix */
  int64_t w_offset;          /* element offset of the class matrix inside wpacked */
} gb_conv_class;

/* Implicit-GEMM convolution on tcgen05 tensor cores.
 *   out[q*out_mul+off][n] = act( sum_{t,c} in[q*in_mul + d_t][c] * W[n][t][c] + bias[n] )
 * Replaces: torch.nn.Conv2d/Conv3d forward (ganslate/nn/generators/resnet/resnet2d.py:24-35,81-86,
 * ganslate/nn/discriminators/patchgan/patchgan2d.py:29-62, patchgan3d.py:28-61,
 * ganslate/nn/generators/vnet/vnet3d.py:158-266), torch.nn.ConvTranspose2d/3d forward
 * (resnet2d.py:52-57, vnet3d.py:224-228) and the autograd data-gradient of each (cuDNN
 * bwd-data), which are the same implicit GEMM with the roles of Cin/Cout swapped. */
typedef struct gb_conv_params {
  gb_view in;                /* gathered operand (bf16) */
  gb_view out;               /* destination (bf16) */
  const void* wpacked;       /* bf16, per class [npad][kpad], K index = tap_local*in.C + c */
  const float* bias;         /* [ncols] fp32 or NULL */
  int32_t ncols;             /* real output channels (<= out.C) */
  int32_t npad;              /* rows of each packed class matrix (multiple of 16) */
  int32_t in_mul[3];         /* gather multiplier (z,y,x): conv stride, or 1 for class-decomposed */
  int32_t out_mul[3];        /* output multiplier: 1, or the stride for class-decomposed */
  int32_t nclass;
  gb_conv_class cls[GB_MAX_CLASSES];
  int8_t taps[GB_MAX_TAPS][4]; /* (dz,dy,dx,unused) */
  int32_t act;               /* GB_ACT_NONE / GB_ACT_TANH / GB_ACT_LEAKY applied in the epilogue */
  float act_slope;
  int32_t out_fp32;          /* 1: `out` is an fp32 view (activation gradients are carried in fp32) */
  int32_t accumulate;        /* 1: out += result (fp32 only; residual branches share one gradient buffer) */
  float* stats;              /* optional [N][out.C][2] fp32 (zeroed by the caller): the epilogue adds (sum, sum^2) of the
                                bf16-rounded outputs per (image, channel) -- the InstanceNorm statistics of the
                                layer that follows, so no separate pass reads the tensor (bf16 output only) */
  int32_t in_c_valid;        /* 0, or the number of leading channels of `in` that exist in memory (< in.C): the rest
                                reads as zero.  Used by "pixel-window" views (in.C = 64 spans 8 pixels of an
                                8-channel tensor with in.sx = 8, so that a k x 7 convolution over 3 channels is k
                                K-blocks of one contiguous 112-byte window each); TMA-fed kernels only */
} gb_conv_params;

int gb_conv_data(const gb_conv_params* p, void* stream);

/* Weight-gradient implicit GEMM (tcgen05, both operands MN-major, split over pixels).
 *   dw[r][t*gathered.C + c] += sum_q plain[q][r] * gathered[q*mul + d_t][c]
 * dw is fp32 [rows_pad][kpad], accumulated with red.global.add -- zero it first.
 * Replaces the autograd weight-gradient (cuDNN bwd-filter) of every conv above. */
typedef struct gb_wgrad_params {
  gb_view plain;             /* rows of dw come from its channels; its extents are the q-grid */
  gb_view gathered;
  float* dw;                 /* fp32 [rows_pad][kpad] */
  int32_t rows;              /* real rows (plain channels used) */
  int32_t kpad;              /* multiple of 64 */
  int32_t ntaps;
  int32_t mul[3];
  int8_t taps[GB_MAX_TAPS][4];
  int32_t splits;            /* 0 = choose */
  int32_t gathered_c_valid;  /* as gb_conv_params.in_c_valid, for `gathered` */
} gb_wgrad_params;

int gb_conv_wgrad(const gb_wgrad_params* p, void* stream);

/* Pack fp32 weights into the bf16 class matrices gb_conv_data reads.
 * dst[cls][n][tl][c] = src[n*sn + c*sc + tap_id[cls][tl]*st] (zero padded).
 * Replaces nothing in the reference (cuDNN consumes the fp32 weights directly). */
typedef struct gb_pack_params {
  const float* src;
  void* dst;                 /* bf16 */
  int64_t sn, sc, st;        /* strides of (row, k-channel, tap) in src */
  int32_t rows, rows_pad;    /* real / padded rows */
  int32_t chans, chans_pad;  /* real / padded k-channels (chans_pad = gathered view C) */
  int32_t nclass;
  int32_t ntaps[GB_MAX_CLASSES];
  int32_t kpad[GB_MAX_CLASSES];
  int64_t w_offset[GB_MAX_CLASSES];
  int32_t tap_begin[GB_MAX_CLASSES];
  int32_t tap_id[GB_MAX_TAPS]; /* source tap index of each (class-local) tap; < 0 = zero (padding tap) */
} gb_pack_params;

int gb_pack_weights(const gb_pack_params* p, void* stream);

/* Every convolution of a network in ONE launch: `table_dev` is an array of `count` gb_pack_params in DEVICE
 * memory (parameter and packed-buffer addresses are stable, so the host builds it once per network);
 * max_elems = largest rows_pad*kpad of any class (sizes the grid). */
int gb_pack_weights_multi(const gb_pack_params* table_dev, int count, int64_t max_elems, void* stream);

/* dst[r*dsr + c*dsc + t*dst_t] = scale * dw[r][t*chans_pad + c]   (fp32 -> fp32, PyTorch layout) */
int gb_unpack_wgrad(const float* dw, float* dst, int64_t dsr, int64_t dsc, int64_t dst_t, int rows,
                    int chans, int chans_pad, int ntaps, int kpad, void* stream);

/* The weight gradients of a whole backward pass in one launch (the batch is passed by value, so the call can be
 * captured in a CUDA graph).  accumulate=1: dst += (a parameter used twice in one pass). */
#define GB_UNPACK_BATCH 56
typedef struct gb_unpack_item {
  const float* dw;
  float* dst;
  int64_t dsr, dsc, dst_t;
  int32_t rows, chans, chans_pad, ntaps, kpad, accumulate;
} gb_unpack_item;
typedef struct gb_unpack_batch {
  int32_t count, pad_;
  gb_unpack_item item[GB_UNPACK_BATCH];
} gb_unpack_batch;
int gb_unpack_wgrad_multi(const gb_unpack_batch* b, void* stream);

/* per-channel sum over all pixels of a bf16 view -> fp32 out[C] (bias gradients). out is overwritten. */
int gb_colsum(const gb_view* x, float* out, void* stream);

/* ---- InstanceNorm (+activation, +residual), HBM-bound --------------------------------------
 * Replaces torch.nn.InstanceNorm2d/3d (affine=False, eps=1e-5; selected at ganslate/nn/utils.py:53-68)
 * and the following nn.ReLU / nn.LeakyReLU(0.2) / nn.PReLU / residual add
 * (resnet2d.py:26-27,36-37,83-87,93; patchgan2d.py:45-46,58-59; vnet3d.py:160-168,193-203). */

/* stats[n][c] = (sum x, sum x^2) in fp32; stats must be zeroed by the caller. */
int gb_in_stats(const gb_view* x, float* stats, void* stream);

typedef struct gb_in_fwd_params {
  gb_view x;                 /* raw conv output (bf16) */
  gb_view y;                 /* destination; if y.pad>0 the reflected border is written as well */
  gb_view res;               /* optional residual added AFTER the activation (ptr==NULL: none) */
  const float* stats;        /* [N][C][2] from gb_in_stats; NULL = no normalisation (activation only) */
  const float* prelu;        /* [C] slopes for GB_ACT_PRELU */
  float eps;
  int32_t act;
  float act_slope;
  int32_t res_before_act;    /* 1: y = act(norm(x) + res)  (V-Net), 0: y = out_scale * act(norm(x)) + res */
  float out_scale;           /* 1, or -1 for the inverse of an additive coupling (x2 = y2 - G(y1)); 0 is read as 1 */
} gb_in_fwd_params;

int gb_in_fwd(const gb_in_fwd_params* p, void* stream);

typedef struct gb_in_bwd_params {
  gb_view x;                 /* raw conv output saved by forward (bf16) */
  gb_view y;                 /* forward output (needed for tanh / no-norm activations), may be NULL ptr */
  gb_view dy_a;              /* FP32 gradient wrt y, plain view (ptr NULL: absent) */
  gb_view dy_b;              /* FP32 gradient wrt the reflection-PADDED y (pad>0: border folded in); ptr NULL: absent */
  gb_view dy_sum;            /* optional FP32: dy_a + fold(dy_b) is written here (residual chain); ptr NULL: skip */
  gb_view dx;                /* gradient wrt x (bf16) */
  const float* stats;        /* forward stats; NULL = no normalisation */
  float* bstats;             /* [N][C][2] workspace (sum g, sum g*xhat) followed by 4 spare words, all zeroed by the
                                caller (the spare words hold the grid-barrier counter of the single-launch path) */
  const float* prelu;
  float* dprelu;             /* [C] fp32 accumulated (zeroed by caller) or NULL */
  float* dbias;              /* [C] fp32 accumulated sum of dx over pixels = gradient of the conv bias that feeds x
                                (zeroed by caller) or NULL */
  float eps;
  int32_t act;
  float act_slope;
  gb_view res;               /* forward residual; needed when res_before_act (the activation mask depends on it) */
  int32_t res_before_act;    /* forward was act(norm(x) + res): dy_sum receives the MASKED gradient g */
  int32_t dy_sum_acc;        /* 1: dy_sum += (several consumers share the gradient buffer), 0: dy_sum = */
  int32_t dx_fp32_acc;       /* 1: dx is an FP32 view and is accumulated (x was an activation buffer, not a raw
                                convolution output) */
  float out_scale;           /* forward was out_scale * act(norm(x)) + res; 0 is read as 1 */
} gb_in_bwd_params;

/* two launches: reduction then apply */
int gb_in_bwd(const gb_in_bwd_params* p, void* stream);

/* ---- layout conversion at the network boundary (set_input / module outputs; cyclegan.py:89-90) ---- */
/* NC(D)HW fp32 -> channels-last bf16 view (zero-fills padded channels, writes reflected border if dst.pad>0).
 * If `pre` is non-NULL the value is multiplied by tanh'(pre) = 1 - tanh(pre)^2 (backward of the tanh export). */
int gb_nchw_to_cl(const float* src, int C, const gb_view* dst, const gb_view* pre, int dst_fp32, void* stream);
/* channels-last bf16 view (border folded in when src.pad>0 and fold!=0) -> NC(D)HW fp32.
 * act = GB_ACT_TANH applies the generator's output nn.Tanh (resnet2d.py:65) in fp32 on the way out. */
int gb_cl_to_nchw(const gb_view* src, float* dst, int C, int fold, int act, int src_fp32, void* stream);

/* ---- replicate padding (torch.nn.ReplicationPad3d in ganslate/nn/generators/resnet/resnet3d.py:24,64,80,84 and
 * piresnet3d.py:61,85,117) ----
 * fwd: dst[n,z,y,x,:] = src[n, clamp(z-pz), clamp(y-py), clamp(x-px), :]  (bf16 views, dst extents = src + 2*pad)
 * bwd: dsrc (FP32 view, accumulated) += sum of ddst (FP32 view on the padded domain) over the padded positions that
 *      read each source element (autograd of ReplicationPad3d). */
int gb_replicate_pad_fwd(const gb_view* src, const gb_view* dst, int pz, int py, int px, void* stream);
int gb_replicate_pad_bwd(const gb_view* ddst, const gb_view* dsrc, int pz, int py, int px, void* stream);

/* ---- losses --------------------------------------------------------------------------------
 * LSGAN: mean((p - t)^2) (ganslate/nn/losses/adversarial_loss.py:29,60-62); grad = 2(p-t)/n.
 * L1: mean(|a-b|) (ganslate/nn/losses/cyclegan_losses.py:64,73,97; pix2pix_losses.py:15); grad = sign(a-b)/n.
 * loss is a single fp32 (zeroed by the caller); grad may be NULL. */
int gb_mse_const(const float* pred, float target, int64_t n, float* loss, float* grad, void* stream);
int gb_l1(const float* a, const float* b, int64_t n, float* loss, float* grad_a, void* stream);

/* SSIM distance loss, ganslate/nn/losses/utils/ssim.py:51-99 (CycleLoss with proportion_ssim > 0,
 * ganslate/nn/losses/cyclegan_losses.py:77-101): x, y are [planes][H][W] fp32 (planes = N*C, or N*C*D for 5-D inputs,
 * whose depth slices the reference filters as channels), mapped by v = in * in_scale + in_shift before use;
 * 11-tap Gaussian (sigma 1.5), K = (0.01, 0.03).  loss (one fp32, zeroed by the caller) += mean over the valid
 * (H-10) x (W-10) map of sqrt(relu(2 - S1 - S2)).  gb_ssim_bwd writes grad_x = dloss[0] * d loss / d x (x is the
 * FIRST argument, the reconstructed image; dloss is a device scalar so the call can be captured in a CUDA graph). */
int gb_ssim_fwd(const float* x, const float* y, int planes, int H, int W, float in_scale, float in_shift,
                float data_range, float* loss, void* stream);
int gb_ssim_bwd(const float* x, const float* y, int planes, int H, int W, float in_scale, float in_shift,
                float data_range, const float* dloss, float* grad_x, void* stream);

/* PatchNCE (CUT): q, k are [B*P][D] fp32 (k is detached in the reference, ganslate/nn/losses/cut_losses.py:16);
 * loss[r] = CE(cat(q_r.k_r, q_r.K_b^T with the own patch masked to -10) / T, 0); probs [B*P][P+1] is saved for
 * gb_patchnce_bwd, which writes dq = d loss / d q scaled by dloss[r]. */
int gb_patchnce_fwd(const float* q, const float* k, int B, int P, int D, float T, float* loss, float* probs,
                    void* stream);
int gb_patchnce_bwd(const float* k, const float* probs, const float* dloss, int B, int P, int D, float T, float* dq,
                    void* stream);

/* FeaturePatchMLP (CUT, ganslate/nn/gans/unpaired/cut.py:229-294): for a feature map feat (N, C, F) fp32 (F = flattened
 * spatial positions) and P position ids shared by every image: x = feat[:, :, ids] as rows (N*P, C),
 * h = relu(x W1^T + b1), z = h W2^T + b2, y = z / (||z||_2 + 1e-7); W1 (nc, C), W2 (nc, nc) in torch.nn.Linear layout.
 * Forward saves x (xg), h, z for the backward; replaces feat[:, patch_id, :] + nn.Linear x 2 + LNorm (cut.py:262-276).
 * Backward: dy (N*P, nc) -> dfeat (N, C, F), zero-initialised by the caller (NULL: not needed), and dW1, db1, dW2,
 * db2 (NULL: not needed); dz, dh are (N*P, nc) scratch. fp32 CUDA-core arithmetic, deterministic. */
int gb_patch_mlp_fwd(const float* feat, const int64_t* ids, int N, int C, int64_t F, int P, const float* W1,
                     const float* b1, const float* W2, const float* b2, int nc, float* xg, float* h, float* z,
                     float* y, void* stream);
int gb_patch_mlp_bwd(const float* dy, const float* xg, const float* h, const float* z, const int64_t* ids, int N,
                     int C, int64_t F, int P, const float* W1, const float* W2, int nc, float* dz, float* dh,
                     float* dfeat, float* dW1, float* db1, float* dW2, float* db2, void* stream);

/* ---- optimizer -------------------------------------------------------------------------------
 * Multi-tensor Adam step with torch.optim.Adam semantics (betas, eps; no weight decay / amsgrad), replacing the
 * optimizer.step() calls at ganslate/nn/gans/unpaired/cyclegan.py:108,121 (optimizers built at :76-82).
 * `lr` and `step` are DEVICE pointers (one float each; step already holds this step's count) so the launch can be
 * replayed in a CUDA graph.  The batch travels by value. */
#define GB_ADAM_BATCH 96
typedef struct gb_adam_item {
  float* p;                  /* parameter (fp32 master) */
  const float* g;            /* gradient */
  float* m;                  /* exp_avg */
  float* v;                  /* exp_avg_sq */
  int32_t n;                 /* elements */
  int32_t vec4;              /* 1: all four pointers are 16-byte aligned */
} gb_adam_item;
typedef struct gb_adam_batch {
  const float* lr;
  const float* step;
  float beta1, beta2, eps;
  int32_t count;
  gb_adam_item item[GB_ADAM_BATCH];
} gb_adam_batch;
int gb_adam_multi(const gb_adam_batch* b, void* stream);

/* ---- misc ---- */
int gb_version(void);
const char* gb_last_error(void);
/* number of kernels launched through this library since load (bench.py's gpu_launches) */
unsigned long long gb_launch_count(void);
/* 1 when the driver accepts the overlapping-stride tensor map that pixel-window views need (gb_conv_params.in_c_valid) */
int gb_tma_window_supported(void);
/* Size of the caller-allocated, zero-initialised FP32 workspace of an operator (the library owns no device memory):
 *   GB_WS_WGRAD  (params = const gb_wgrad_params*): the `dw` matrix gb_conv_wgrad accumulates into,
 *                rows rounded up to 128 x kpad floats;
 *   GB_WS_IN_BWD (params = const gb_view* x): gb_in_bwd_params.bstats, N x C x 2 floats + 4 words for the grid barrier
 *                of the single-launch path.
 * Host arithmetic only. */
#define GB_WS_WGRAD 0
#define GB_WS_IN_BWD 1
int gb_workspace_bytes(int op, const void* params, int64_t* bytes);

/* Host-only replay of the work decomposition of the persistent convolution kernels (csrc/igemm_cg2.cu; mode 1 = CTA
 * pair, 2 = single CTA): see the definition.  Used by the CPU tests; touches no device. */
int gb_debug_cg2_plan(const gb_conv_params* p, int mode, int32_t* info, int32_t* out, int64_t out_ints);

/* Report of the persistent kernels' bring-up watchdog (gb_debug_knob(21, limit in millions of clocks)): see the
 * definition in csrc/igemm_cg2.cu.  out: 8 ints. */
int gb_debug_cg2_watchdog(int32_t* out);

/* debug knobs for bring-up (e.g. descriptor variants); returns previous value */
int gb_debug_knob(int knob, int value);
/* Bring-up: per-CTA time stamps of the TMA-fed data kernel. buf = device memory of 8 x uint64 per CTA (SM id,
   globaltimer, clock64 at: start, setup done, first operands landed, last MMA issued, accumulator complete, epilogue
   done); launches with more CTAs than max_ctas are not recorded; NULL switches it off. tools/conv_timeline.py. */
int gb_debug_timeline(void* buf, long long max_ctas);

#ifdef __cplusplus
}
#endif
#endif
