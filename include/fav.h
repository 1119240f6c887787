/*
 * fav.h — C-ABI of the B200-native flickering-attack engine (libfav.so).
 *
 * The reference (roiponytch/Flickering_Adversarial_Video) has no FFI: its hot path is a Python
 * attribute bag of TF tensors (`utils/kinetics_i3d_utils.py:76-307`, class kinetics_i3d) driven by
 * `sess.run(train_op …)` (`i3d_adversarial_main_single_video_npy.py:211-215`).  This header is the
 * boundary a maintainer binds with ctypes; every entry point names the
 * reference code it replaces.
 *
 * Conventions
 *   - every `void*`/`float*` marked DEVICE is a caller-owned device pointer on the handle's GPU
 *     (e.g. a torch tensor's data_ptr()); HOST pointers are plain host memory;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all hot-path calls are
 *     asynchronous on it and are CUDA-graph capturable (no allocation, no sync inside);
 *   - return value: 0 (FAV_OK) or a negative fav_status; never throws across the boundary;
 *     `fav_last_error` returns a thread-local message for the last failing call;
 *   - a handle is bound to one device and is not thread-safe (one handle per rank);
 *   - the library owns only its activation/gradient arena and packed weights (allocated in
 *     fav_create / fav_load_weights).
 *
 * Activation layout inside the library: NDHWC, 16 bits per element, channel stride padded to a multiple of 16:
 * FORWARD activations (and the forward weights) are IEEE fp16, GRADIENTS (and the data-gradient weights) bf16;
 * every contraction accumulates in fp32 (DESIGN.md section 4 says why the forward is not bf16).
 */
#ifndef FAV_H_
#define FAV_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fav_handle fav_handle;

typedef enum {
  FAV_OK = 0,
  FAV_ERR_ARG = -1,      /* bad argument / unsupported shape */
  FAV_ERR_CUDA = -2,     /* a CUDA runtime/driver call failed */
  FAV_ERR_STATE = -3,    /* call order violated (e.g. forward before load_weights) */
  FAV_ERR_NOGPU = -4,    /* no CUDA device / not an sm_100 device */
  FAV_ERR_MISSING = -5   /* a required named tensor was not supplied */
} fav_status;

/* I3D (i3d.py) and the torchvision video ResNets the torch stack attacks
 * (utils_cv/action_recognition/model.py:403-441: r3d_18 / mc3_18 / r2plus1d_18). */
typedef enum { FAV_NET_I3D = 0, FAV_NET_R3D_18 = 1, FAV_NET_MC3_18 = 2, FAV_NET_R2PLUS1D_18 = 3 } fav_arch;
typedef enum { FAV_U8 = 0, FAV_F32 = 1 } fav_dtype;

/* Which framework's semantics to follow where the two reference stacks differ
 * (SURVEY App. B.5/B.6): regulariser weighting, Adam epsilon placement. */
typedef enum { FAV_STACK_TF = 0, FAV_STACK_TORCH = 1 } fav_stack;

typedef struct {
  int32_t arch;         /* fav_arch */
  int32_t batch;        /* clips resident per step on this GPU */
  int32_t frames;       /* T  (reference constant _SAMPLE_VIDEO_FRAMES=90, kinetics_i3d_utils.py:12) */
  int32_t height;       /* 224 */
  int32_t width;        /* 224 */
  int32_t num_classes;  /* 400 (kinetics_i3d_utils.py:19) */
} fav_net_desc;

/* A named fp32 HOST tensor in the reference's native layout (TF ckpt variable names,
 * `RGB/inception_i3d/<unit>/conv_3d/w` [kt,kh,kw,Cin,Cout], `…/batch_norm/{beta,moving_mean,
 * moving_variance}` [1,1,1,1,C] or [C], `…/Logits/Conv3d_0c_1x1/conv_3d/{w,b}`), cf.
 * utils/kinetics_i3d_utils.py:41-62. */
typedef struct {
  const char* name;
  const float* data;   /* HOST */
  int32_t ndim;
  int64_t dims[5];
} fav_tensor;

/* Input normalisation and range clamp of the torch stack: x = (u8/255 - mean)/std
 * (references/functional_video.py:65-97, dataset.py:28-29), adv = clamp(x + delta/std, lo, hi)
 * with the scalar bounds of Perturbation (model.py:72-75). */
typedef struct {
  float mean[3];
  float std[3];
  float lo, hi;
} fav_norm_params;

/* Adversarial-loss selection: improve_adversarial_loss (kinetics_i3d_utils.py:253-288) or
 * ce_adversarial_loss (:290-307). */
typedef struct {
  int32_t improve_loss;   /* IMPROVE_ADV_LOSS */
  int32_t targeted;       /* TARGETED_ATTACK  */
  int32_t use_logits;     /* USE_LOGITS       */
  float   margin;         /* PROB_MARGIN      */
  float   grad_scale;     /* multiplies dloss/dlogits (1/n_ranks for the mean-reduced CE loss when sharded) */
  int32_t global_batch;   /* batch over all ranks (CE mean divisor); 0 => local batch */
  int32_t stack;          /* fav_stack: which reference stack's selection rules */
} fav_loss_params;

/* Regulariser weights: loss = adv + beta0*(beta1*thick + beta2*diff + beta3*lap)
 * (i3d_adversarial_main_single_video_npy.py:56-59). */
typedef struct {
  float beta0, beta1, beta2, beta3;
  float delta_clip;       /* 0.4  (kinetics_i3d_utils.py:104-105) */
} fav_reg_params;

typedef struct {
  float lr, b1, b2, eps;  /* tf.train.AdamOptimizer defaults 1e-3, .9, .999, 1e-8 */
  int32_t stack;          /* fav_stack: TF => eps outside the bias correction */
} fav_adam_params;

/* Layout of the per-step device scalar block written by fav_loss / fav_delta_update
 * (floats; indices below).  Reference fetch list: single_video_npy.py:213-215. */
enum {
  FAV_S_ADV_LOSS = 0,     /* adversarial loss (sum or mean over the local batch) */
  FAV_S_FOOLED = 1,       /* # local clips with argmax != label (or == target) */
  FAV_S_SUM_P_MIN = 2,    /* sum of to_min_prob */
  FAV_S_SUM_P_MAX = 3,    /* sum of to_max_prob */
  FAV_S_NORM_REG = 4,     /* norm_reg        (+1e-12) */
  FAV_S_DIFF_REG = 5,     /* diff_norm_reg   (+1e-12) */
  FAV_S_LAP_REG = 6,      /* laplacian_norm_reg (+1e-12) */
  FAV_S_THICKNESS = 7,    /* mean|delta|  (before the update, as the reference fetches) */
  FAV_S_ROUGHNESS = 8,    /* mean|delta - roll(delta,1)| */
  FAV_S_TOTAL_LOSS = 9,   /* adv + beta0*reg (local adv term) */
  FAV_S_SAT_COUNT = 10,   /* # (pixel,channel) entries where the [-1,1] clip fired */
  FAV_S_COUNT = 16
};

/* ---- lifecycle ------------------------------------------------------------------------- */
int fav_create(fav_handle** out, int device, const fav_net_desc* desc);
int fav_destroy(fav_handle* h);
const char* fav_last_error(void);
/* bytes of device memory owned by the handle (arena + weights) */
int64_t fav_device_bytes(const fav_handle* h);

/* Fold BN into conv weights, cast to fp16 (forward operands) / bf16 (data-gradient operands), pack for the
 * tensor-core kernels.
 * Replaces tf.train.Saver.restore (kinetics_i3d_utils.py:41-62) + snt.BatchNorm inference
 * (i3d.py:66-68). */
int fav_load_weights(fav_handle* h, const fav_tensor* tensors, int n);

/* ---- hot path -------------------------------------------------------------------------- */
/* K1: adv = clip(x + adv_flag*clip(delta,+-delta_clip), -1, 1) with x = u8/128-1
 * (pre_process_rgb_flow.py:234; kinetics_i3d_utils.py:104-105,139-142).
 *   clip     DEVICE [B,T,H,W,3] u8 or f32 (already normalised, cf. single_video_npy.py:121)
 *   delta    DEVICE [T,3] f32 (the reference's eps_rgb [T,1,1,3])
 *   adv_u8   DEVICE [B,T,H,W,3] or NULL — ((adv+1.0)*127.5).astype(uint8)  (stats_plots.py:57), bit-exact
 *   adv_f32  DEVICE [B,T,H,W,3] or NULL — the reference's `adversarial_inputs_rgb`
 * Also fills the engine's stem input and the pass bitmap of the range clip used by the backward pass. */
int fav_apply_flicker(fav_handle* h, const void* clip, int in_dtype, const float* delta,
                      float adv_flag, float delta_clip, uint8_t* adv_u8, float* adv_f32, void* stream);

/* Forward of the frozen network on the last applied input; logits DEVICE [B,num_classes] f32.
 * Replaces rgb_model(adversarial_inputs_rgb) (kinetics_i3d_utils.py:150; i3d.py:144-474). */
int fav_forward(fav_handle* h, float* logits, void* stream);

/* softmax + adversarial loss + dloss/dlogits (kept inside the handle).
 *   labels DEVICE [B] int64;  probs DEVICE [B,num_classes] f32 or NULL;
 *   scalars DEVICE [FAV_S_COUNT] f32. */
int fav_loss(fav_handle* h, const int64_t* labels, const fav_loss_params* p, float* probs,
             float* scalars, void* stream);

/* Backward-to-input and the H*W*B collapse: grad DEVICE [T,3] f32 receives
 * sum_{b,h,w} mask * dL/dx  (the data term of compute_gradients(loss, var_list=perturbation),
 * single_video_npy.py:82), *before* the |delta|<=clip mask and regulariser terms. */
int fav_backward_delta(fav_handle* h, float* grad, void* stream);

/* (c) regulariser gradients + clip mask + Adam + metrics on delta [T,3]
 * (kinetics_i3d_utils.py:177-200; single_video_npy.py:56-59,79-84).
 *   grad is the (all-reduced) data gradient from fav_backward_delta;
 *   step DEVICE int64 counter (incremented), m/v DEVICE [T,3] Adam slots. */
int fav_delta_update(fav_handle* h, float* delta, const float* grad, float* m, float* v,
                     int64_t* step, const fav_reg_params* reg, const fav_adam_params* adam,
                     float* scalars, void* stream);

/* ---- sparse per-pixel attack (FLICKERING_ATTACK = False) -------------------------------------
 * kinetics_i3d_L12 (utils/kinetics_i3d_utils.py:308-521: eps [T,224,224,3], no +-0.4 clip, regulariser
 * beta_1 * loss_L12, i3d_adversarial_main_universal.py:133) and the torch stack with attack_type "L12"
 * (pert_size [3,T,112,112], model.py:383-384; Losses.L12_regularization_loss :211-214).
 * delta_px / grad_px / m / v are DEVICE f32 [T,H,W,3] (the torch stack's [3,T,H,W] permuted by the caller). */
/* allocate the dense stem data-gradient buffer and plan its kernels (once, after fav_load_weights) */
int fav_pixels_enable(fav_handle* h);
/* adv = clamp(x + adv_flag*delta) (TF: [-1,1]; torch: clamp(delta,+-delta_clip)/std and the scalar bounds);
 * delta_clip <= 0 disables the delta clamp (kinetics_i3d_utils.py:336).  adv_f32 (or NULL): NTHWC (I3D) /
 * NCTHW (torch stack).  clip_u8 DEVICE [B,T,H,W,3]. */
int fav_apply_pixels(fav_handle* h, const void* clip_u8, const float* delta_px, float adv_flag, float delta_clip,
                     float* adv_f32, void* stream);
/* grad_px[t,h,w,c] = sum_b mask * dL/d(adv) (times 1/std_c on the torch stack): the full backward-to-input,
 * with the stem data gradient as parity-class tensor-core GEMMs. */
int fav_backward_pixels(fav_handle* h, float* grad_px, void* stream);
/* L1,2 regulariser gradient (reg_weight * d/d delta sum_t sqrt(mean delta_t^2)) + clamp mask + Adam on the
 * per-pixel delta; scalars: FAV_S_NORM_REG <- L12, thickness, roughness, total loss. */
int fav_pixels_update(fav_handle* h, float* delta_px, const float* grad_px, float* m, float* v, int64_t* step,
                      float reg_weight, float delta_clip, const fav_adam_params* adam, float* scalars, void* stream);

/* ---- fused evaluation pass ---------------------------------------------
 * kinetics_i3d.evaluate (utils/kinetics_i3d_utils.py:217-250: two sess.run per validation batch, adv_flag 1 and 0), the
 * EVAL branch of model_fn (i3d_adversarial_main_universal.py:139-161) and the torch stack's validation phase
 * (utils_cv/action_recognition/model.py:697-713 + Adversarial_metrics.accuracy_for_eval :293-323) forward the clean and
 * the perturbed version of every validation batch and count
 *     valid = argmax(clean) == label          (all clips when exclude_misclassify == 0)
 *     miss  = valid && (targeted ? argmax(adv) == target_class : argmax(adv) != label)
 * An evaluation handle is forward-only: desc->batch = Bu clips per validation batch; its plan runs 2*Bu clips (clean
 * rows [0,Bu), perturbed rows [Bu,2Bu)) through one set of packed forward weights in ONE pass and owns no gradient
 * buffers, pool codes or data-gradient weights (about 40 % of a training handle's arena per resident clip).  The training
 * entry points (fav_apply_flicker, fav_loss, fav_backward_*, fav_pixels_*) return FAV_ERR_STATE on it. */
int fav_create_eval(fav_handle** out, int device, const fav_net_desc* desc);
/* One validation batch.
 *   clip_clean  DEVICE [Bu,T,H,W,3] u8 (or f32 for I3D);  clip_adv: the clips the perturbation is added to, or NULL for
 *               the same ones (the reference's cyclic evaluation rolls only the perturbed input, kinetics_i3d_utils.py:228)
 *   delta       DEVICE [T,3] f32;  labels DEVICE [Bu] i64;  n_clips <= Bu: clips of a ragged last batch that count
 *   loss/scalars: both NULL, or the adversarial-loss selection and a DEVICE [FAV_S_COUNT] block that receives the loss
 *               scalars of the perturbed rows (what the validation phase logs, model.py:706); labels must hold Bu entries
 *   counts      DEVICE int64[2]: counts[0] += miss, counts[1] += valid (accumulates over batches; zero it once)
 *   probs       DEVICE [2*Bu,num_classes] f32 or NULL: softmax of the clean rows, then of the perturbed rows */
int fav_eval_batch(fav_handle* h, const void* clip_clean, const void* clip_adv, int in_dtype, const float* delta,
                   float delta_clip, const int64_t* labels, int n_clips, int targeted, int64_t target_class,
                   int exclude_misclassify, const fav_loss_params* loss, int64_t* counts, float* probs, float* scalars,
                   void* stream);

/* ---- op-level entry points (layer-wise parity tests; same kernels the engine runs) ------- */
/* stride-1 SAME Conv3d (+bias, +ReLU) on NDHWC 16-bit tensors via the tcgen05 implicit-GEMM kernel.
 *   Formats follow the engine's: dgrad == 0 reads fp16 x and writes fp16 y; dgrad != 0 reads bf16 dY and writes bf16 dX.
 *   x [B,T,H,W,x_cs] (channels x_coff..x_coff+cin), w HOST f32 [kt,kh,kw,cin,cout] (TF layout),
 *   bias HOST f32 [cout] or NULL, y [B,T,H,W,y_cs] (channels y_coff..y_coff+cout).
 *   dgrad != 0: computes the data gradient instead (x is dY with cout channels, y is dX with cin);
 *   relu_src (DEVICE fp16 activation, same geometry as y, stride relu_cs/offset relu_coff) != NULL multiplies
 *   the result by (relu_src > 0). */
int fav_op_conv3d(int device, const void* x, int64_t x_cs, int64_t x_coff,
                  const float* w, const float* bias, int kt, int kh, int kw, int cin, int cout,
                  void* y, int64_t y_cs, int64_t y_coff, int B, int T, int H, int W,
                  int relu, int dgrad, const void* relu_src, int64_t relu_cs, int64_t relu_coff,
                  void* stream);

/* tf.nn.max_pool3d SAME on NDHWC fp16 (i3d.py:174 etc.); idx receives the arg-max tap (u8). */
int fav_op_maxpool3d(int device, const void* x, void* y, uint8_t* idx, int B, int T, int H, int W,
                     int C, int kt, int kh, int kw, int st, int sh, int sw, void* stream);
/* its backward: dx = (add ? add : 0) + scatter(dy) ; then * (relu_src>0) if relu_src != NULL
 * (dy, add, dx bf16 gradients; relu_src an fp16 activation) */
int fav_op_maxpool3d_bwd(int device, const void* dy, const uint8_t* idx, const void* add,
                         const void* relu_src, void* dx, int B, int T, int H, int W, int C,
                         int kt, int kh, int kw, int st, int sh, int sw, void* stream);

/* softmax + adversarial loss + dloss/dlogits on caller buffers (same kernel as fav_loss):
 * logits DEVICE [B,K] f32, labels DEVICE [B] i64, probs DEVICE [B,K] or NULL, dlogits DEVICE [B,K],
 * scalars DEVICE [FAV_S_COUNT].  Torch-stack rules (model.py:177-250) when p->stack == FAV_STACK_TORCH. */
int fav_op_loss(int device, const float* logits, const int64_t* labels, const fav_loss_params* p, int B, int K,
                float* probs, float* dlogits, float* scalars, void* stream);
/* kernel (c) on caller buffers (same kernel as fav_delta_update); delta/grad/m/v DEVICE [T,3]. */
int fav_op_delta_update(int device, float* delta, const float* grad, float* m, float* v, int64_t* step,
                        const fav_reg_params* reg, const fav_adam_params* adam, float adv_flag, float* scalars,
                        int T, void* stream);

/* Clip loader of the torch stack: ToTensorVideo -> ResizeVideo(128, keep_ratio) -> CenterCropVideo(112)
 * [-> NormalizeVideo] (utils_cv/action_recognition/dataset.py:84-123; references/transforms_video.py:23-53,182-200;
 * references/functional_video.py:52-97) in one pass over decoded frames, the crop folded into the bilinear resize
 * (torch interpolate, align_corners=False; arithmetic order of the CPU kernel, no FMA contraction).
 *   frames_u8 DEVICE [n_frames,H,W,3];  resized_h/w: size after ResizeVideo;  ratio_h/w: float(1/scale_factor) as
 *   interpolate uses it (input/output size when no scale factor is given);  crop_i/j, out_h/w: crop window in the
 *   resized frame;  out_u8 DEVICE [n_frames,out_h,out_w,3] or NULL: nearest uint8 of the resized [0,1] clip (what
 *   fav_apply_flicker consumes);  out_f32 DEVICE [n_frames/frames_per_clip,3,frames_per_clip,out_h,out_w] or NULL:
 *   the reference's transform output (v - mean)/std, needs norm. */
int fav_op_resize_crop(int device, const uint8_t* frames_u8, int n_frames, int H, int W, int resized_h,
                       int resized_w, float ratio_h, float ratio_w, int crop_i, int crop_j, int out_h, int out_w,
                       int frames_per_clip, const fav_norm_params* norm, uint8_t* out_u8, float* out_f32,
                       void* stream);

/* debug/introspection: copy a named internal activation (fp16, or bf16 for "grad:" -> f32, NDHWC, unpadded channels)
 * to a DEVICE f32 buffer; returns element count or <0.  Names follow i3d.py end points
 * ("Conv3d_1a_7x7", "Mixed_3b", ...), prefix "grad:" for the gradient buffer. */
int64_t fav_debug_read(fav_handle* h, const char* name, float* out, int64_t capacity, void* stream);

/* Per-kernel-family timing with CUDA events on the launching stream (bench.py's live roofline numbers).
 * Between _begin and _end every launch is bracketed by an event pair; _end synchronises and writes
 * out[kind*4 + {0: ms, 1: launches, 2: algorithmic FLOPs, 3: algorithmic bytes}] (HOST doubles). */
enum { FAV_PK_APPLY = 0, FAV_PK_STEM, FAV_PK_CONV_HALO, FAV_PK_CONV_TAP, FAV_PK_POOL_FWD, FAV_PK_POOL_BWD,
       FAV_PK_HEAD_LOSS, FAV_PK_STEM_BWD, FAV_PK_DELTA_UPDATE, FAV_PK_OTHER, FAV_PK_COUNT };
int fav_profile_begin(void);
int fav_profile_end(double* out, int capacity);

/* number of CUDA kernels libfav has launched in this process (bench.py `gpu_launches`) */
int64_t fav_launch_count(void);

/* host-side fp32 -> fp16 bit pattern used when packing the forward weights (round to nearest even, subnormals kept,
 * saturating at +-65504); exported so that the conversion can be checked without a GPU */
uint16_t fav_debug_f32_to_f16(float f);

/* host-only: the frame schedule the temporal-sharing stem kernel runs for a KT-tap stem of temporal stride st and bn output
 * channels (csrc/conv_stem.cu).  Per class c < st (frame index mod st = tap index mod st): nfr[c] input frames per tile,
 * nslot[c] taps, ktmax[c] the largest tap; tab[c*8 + f] for frame f of the class packs the first / top output frame it
 * feeds, the accumulator column and the weight-row offset.  Returns the output frames per tile (4) or < 0.  Exported so
 * that the schedule's invariants can be checked without a GPU (tests/test_cpu_lib.py). */
int fav_debug_stem_ts_schedule(int KT, int st, int bn, int* nfr, int* nslot, int* ktmax, uint32_t* tab);

/* library build info: "sm_100a;<compile date>" */
const char* fav_build_info(void);

#ifdef __cplusplus
}
#endif
#endif  /* FAV_H_ */
