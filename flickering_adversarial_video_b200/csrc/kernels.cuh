// kernels.cuh — HBM-bound kernels of the attack loop (everything that is not a GEMM).
#pragma once
#include "fav_common.cuh"

namespace fav {

struct PoolGeom {
  int B, T, H, W, C;       // input
  int To, Ho, Wo;          // output (TF SAME: ceil(in/stride))
  int kt, kh, kw, st, sh, sw;
  int pt, ph, pw;          // pad_before (TF SAME: floor(pad_total/2))
};
PoolGeom make_pool_geom(int B, int T, int H, int W, int C, int kt, int kh, int kw, int st, int sh, int sw);

// (a) flicker apply. Writes the stem input x' (fp16 RGBX, W padded) and optionally the uint8 /
// fp32 adversarial video; with pass_bits, the pass nibbles of stem_grad.cu (bit c of pixel nibble = entry (pixel, c)
// was not range-clipped; rows of round_up((W + 16) / 8, 4) words, H + 7 rows per frame, zero borders).
int launch_apply(const void* clip, int in_dtype, const float* delta, float adv_flag, float delta_clip,
                 __half* xpad, int Wp, int padl, uint8_t* adv_u8, float* adv_f32,
                 uint32_t* pass_bits, int B, int T, int H, int W, cudaStream_t s);

// delta-dependent stem bias table [To][4][4][64]
int launch_stem_bias(const float* delta, float adv_flag, float delta_clip, const float* wc /*[7][16][3][64]*/,
                     const float* bnbias /*[64]*/, float* table, int T, int To, int pt, cudaStream_t s);

int launch_stem_bias_ex(const float* delta, float adv_flag, float delta_clip, const float* wc, const float* bnbias,
                        float* table, int T, int To, int pt, int KT, int st, int C1, const float* cst3,
                        const float* dscale3, cudaStream_t s);

// torch-stack apply (uint8 NTHWC clip -> stem input in uint8 units; adv_f32 optional, NCTHW)
int launch_apply_torch(const uint8_t* clip, const float* delta, float adv_flag, float delta_clip,
                       const fav_norm_params& nrm, __half* xpad, int Wp, int padl, float* adv_f32,
                       uint32_t* pass_bits, int B, int T, int H, int W, cudaStream_t s);

// (c) dense reduce of the stem data gradient dX [B,T,H,W,16] bf16 into grad [T,3] with the recomputed clip mask
int launch_stem_dx_reduce(const __nv_bfloat16* dx, const uint8_t* clip, const float* delta, float adv_flag,
                          float delta_clip, const fav_norm_params& nrm, int torch_mode, float* partial, float* grad,
                          int B, int T, int H, int W, cudaStream_t s);
int stem_dx_reduce_chunks(int H);

// sparse per-pixel attack: apply with delta [T,H,W,3], per-pixel gradient, L1,2 regulariser + Adam
int launch_apply_pixels(const uint8_t* clip, const float* delta_px, float adv_flag, float delta_clip,
                        const fav_norm_params& nrm, int torch_mode, __half* xpad, int Wp, int padl,
                        float* adv_f32, int B, int T, int H, int W, cudaStream_t s);
int launch_stem_dx_pixels(const __nv_bfloat16* dx, const uint8_t* clip, const float* delta_px, float adv_flag,
                          float delta_clip, const fav_norm_params& nrm, int torch_mode, float* grad, int B, int T, int H,
                          int W, cudaStream_t s);
int launch_pixels_update(float* delta_px, const float* grad_px, float* m, float* v, int64_t* step, float* partial,
                         float reg_weight, float delta_clip, const fav_adam_params& adam, float* scalars, int T, int H,
                         int W, cudaStream_t s);
int pixels_partial_floats(int T, int H, int W);

// pools: forward on fp16 activations; backward on bf16 gradients, ReLU mask from the producer's fp16 output
int launch_maxpool_fwd(const __half* x, __half* y, uint8_t* idx, const PoolGeom& g,
                       cudaStream_t s);
// `pooled` (optional): the pool's forward OUTPUT.  Without an addend the strided kernels then take the ReLU mask of the
// input from it (an element that receives gradient is the arg-max of that window, so input > 0 <=> pooled > 0) and never
// read the full-resolution `relu_src`; results are bit-identical.
int launch_maxpool_bwd(const __nv_bfloat16* dy, const uint8_t* idx, const __nv_bfloat16* addend,
                       const __half* relu_src, __nv_bfloat16* dx, const PoolGeom& g,
                       cudaStream_t s, const __half* pooled = nullptr);

// 3x3x3 / stride 1 / SAME pools: separable streaming kernels (pool3.cu).  idx holds three 2-bit stage
// codes per element instead of a 27-tap index; launch_maxpool_fwd/bwd dispatch to them when applicable.
bool pool3s1_applicable(const PoolGeom& g);
int launch_pool3s1_fwd(const __half* x, __half* y, uint8_t* idx, const PoolGeom& g, cudaStream_t s);
int launch_pool3s1_bwd(const __nv_bfloat16* dy, const uint8_t* idx, const __nv_bfloat16* addend,
                       const __half* relu_src, __nv_bfloat16* dx, const PoolGeom& g, cudaStream_t s);

// stride-2 pools ([1,3,3]/[1,2,2], 3x3x3/2x2x2): patch-per-thread backward (pool3.cu), standard 27-tap idx
bool pool_s2_applicable(const PoolGeom& g);
int launch_pool_s2_bwd(const __nv_bfloat16* dy, const uint8_t* idx, const __nv_bfloat16* addend,
                       const __half* relu_src, __nv_bfloat16* dx, const PoolGeom& g, cudaStream_t s,
                       const __half* pooled = nullptr);

// head: feat[b,c] = sum_t coef[t]*sum_hw Y / (HW*2*(T5-1)); logits = feat @ Wl + bl
int launch_head_fwd(const __half* y, int B, int T5, int HW, int C, float* feat,
                    const float* wl /*[C][K]*/, const float* bl, int K, float* logits, cudaStream_t s);
int launch_head_bwd(const float* dlogits, const float* wl, int K, const __half* y,
                    __nv_bfloat16* gy, float* dfeat, int B, int T5, int HW, int C, cudaStream_t s);

int launch_loss(const float* logits, const int64_t* labels, const fav_loss_params& p, int B, int K,
                float* probs, float* dlogits, float* scalars, cudaStream_t s);

// fooling-ratio counts of one validation batch (clean rows [0,Bu), perturbed rows [Bu,2Bu) of `logits`)
int launch_eval_counts(const float* logits, const int64_t* labels, int Bu, int n, int K, int targeted, long long target,
                       int exclude, int64_t* counts, float* probs, cudaStream_t s);

int launch_delta_update(float* delta, const float* grad, float* m, float* v, int64_t* step,
                        const fav_reg_params& reg, const fav_adam_params& adam, float adv_flag,
                        float* scalars, int T, cudaStream_t s);

int launch_f16_to_f32(const __half* src, long long cs, int coff, int C, long long npos, float* dst, cudaStream_t s);
int launch_bf16_to_f32(const __nv_bfloat16* src, long long cs, int coff, int C, long long npos, float* dst,
                       cudaStream_t s);
int launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long cs, int coff, int C, long long npos,
                       cudaStream_t s);

}  // namespace fav
