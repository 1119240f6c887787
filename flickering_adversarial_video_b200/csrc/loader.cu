// loader.cu — GPU side of the torch stack's clip loader: the test-time transform chain
// ToTensorVideo -> ResizeVideo(128, keep_ratio) -> CenterCropVideo(112) [-> NormalizeVideo] of the reference
// (utils_cv/action_recognition/dataset.py:84-123, references/transforms_video.py:23-53,182-200,
// references/functional_video.py:52-97) as ONE pass over the decoded uint8 frames: the crop is folded into the
// resize (only the pixels that survive the crop are interpolated), nothing is materialised in between.
//
// Arithmetic follows torch.nn.functional.interpolate(mode="bilinear", align_corners=False, scale_factor=s) on the
// float clip u8/255 operation by operation (IEEE single, no FMA contraction), so that the optional normalised fp32
// output is the reference's transform output and the uint8 output is its nearest uint8 (round half to even of 255·v):
//   src = ratio·(dst + 0.5) − 0.5, clamped at 0;  i0 = min(⌊src⌋, n−1);  i1 = i0 + (i0 < n−1);  λ1 = clamp(src − i0, 0, 1)
//   v = λ0h·(λ0w·p00 + λ1w·p01) + λ1h·(λ0w·p10 + λ1w·p11),   a dimension with n_out == n_in is copied.
// HBM-bound and tiny next to the attack step (a 16×256×340 clip is 4.2 MB in, 0.6 MB out).
#include "fav_common.cuh"

namespace fav {

struct AxisTap { int i0, i1; float l0, l1; };

__device__ __forceinline__ AxisTap axis_tap(int dst, int n_in, int n_out, float ratio) {
  AxisTap a;
  if (n_in == n_out) { a.i0 = a.i1 = dst; a.l0 = 1.f; a.l1 = 0.f; return a; }
  float s = __fsub_rn(__fmul_rn(ratio, __fadd_rn(static_cast<float>(dst), 0.5f)), 0.5f);
  s = s < 0.f ? 0.f : s;
  int i0 = min(static_cast<int>(floorf(s)), n_in - 1);
  float l1 = fminf(fmaxf(__fsub_rn(s, static_cast<float>(i0)), 0.f), 1.f);
  a.i0 = i0;
  a.i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  a.l1 = l1;
  a.l0 = __fsub_rn(1.f, l1);
  return a;
}

struct LoaderNorm { float mean[3], stdv[3]; };

// grid: (ceil(out_h*out_w / 256), frame groups); one thread per output pixel, frames strided over grid.y (the taps of
// a pixel are the same in every frame, so a thread that owns several frames computes them once)
__global__ void __launch_bounds__(256)
resize_crop_kernel(const uint8_t* __restrict__ src, int n_frames, int H, int W, int rh, int rw, float ratio_h,
                   float ratio_w, int crop_i, int crop_j, int oh, int ow, int frames_per_clip,
                   uint8_t* __restrict__ dst_u8, float* __restrict__ dst_f32, LoaderNorm nrm) {
  // u8 / 255 as a 256-entry table (the IEEE quotient, computed once per CTA instead of 12 divisions per pixel)
  __shared__ float unit[256];
  unit[threadIdx.x] = __fdiv_rn(static_cast<float>(threadIdx.x), 255.f);
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= oh * ow) return;
  const int oy = p / ow, ox = p - oy * ow;
  const AxisTap ty = axis_tap(oy + crop_i, H, rh, ratio_h);
  const AxisTap tx = axis_tap(ox + crop_j, W, rw, ratio_w);
  const size_t plane = static_cast<size_t>(oh) * ow;
  for (int f = blockIdx.y; f < n_frames; f += gridDim.y) {
    const uint8_t* fr = src + static_cast<size_t>(f) * H * W * 3;
    const uint8_t* r0 = fr + static_cast<size_t>(ty.i0) * W * 3;
    const uint8_t* r1 = fr + static_cast<size_t>(ty.i1) * W * 3;
    const int b = f / frames_per_clip, t = f - b * frames_per_clip;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float p00 = unit[r0[tx.i0 * 3 + c]];
      const float p01 = unit[r0[tx.i1 * 3 + c]];
      const float p10 = unit[r1[tx.i0 * 3 + c]];
      const float p11 = unit[r1[tx.i1 * 3 + c]];
      const float top = __fadd_rn(__fmul_rn(p00, tx.l0), __fmul_rn(p01, tx.l1));
      const float bot = __fadd_rn(__fmul_rn(p10, tx.l0), __fmul_rn(p11, tx.l1));
      const float v = __fadd_rn(__fmul_rn(top, ty.l0), __fmul_rn(bot, ty.l1));
      if (dst_u8) {
        int q = __float2int_rn(__fmul_rn(v, 255.f));
        q = q < 0 ? 0 : (q > 255 ? 255 : q);
        dst_u8[(static_cast<size_t>(f) * plane + p) * 3 + c] = static_cast<uint8_t>(q);
      }
      if (dst_f32) {
        const float z = __fdiv_rn(__fsub_rn(v, nrm.mean[c]), nrm.stdv[c]);
        dst_f32[((static_cast<size_t>(b) * 3 + c) * frames_per_clip + t) * plane + p] = z;
      }
    }
  }
}

}  // namespace fav

extern "C" int fav_op_resize_crop(int device, const uint8_t* frames_u8, int n_frames, int H, int W, int resized_h,
                                  int resized_w, float ratio_h, float ratio_w, int crop_i, int crop_j, int out_h,
                                  int out_w, int frames_per_clip, const fav_norm_params* norm, uint8_t* out_u8,
                                  float* out_f32, void* stream) {
  using namespace fav;
  FAV_CHECK_ARG(frames_u8 && (out_u8 || out_f32), "fav_op_resize_crop: null argument");
  FAV_CHECK_ARG(n_frames > 0 && H > 0 && W > 0 && resized_h > 0 && resized_w > 0 && out_h > 0 && out_w > 0,
                "fav_op_resize_crop: non-positive size");
  FAV_CHECK_ARG(crop_i >= 0 && crop_j >= 0 && crop_i + out_h <= resized_h && crop_j + out_w <= resized_w,
                "fav_op_resize_crop: crop [%d:%d, %d:%d] outside the resized frame %dx%d", crop_i, crop_i + out_h,
                crop_j, crop_j + out_w, resized_h, resized_w);
  FAV_CHECK_ARG(frames_per_clip > 0 && n_frames % frames_per_clip == 0,
                "fav_op_resize_crop: n_frames %d is not a multiple of frames_per_clip %d", n_frames, frames_per_clip);
  FAV_CHECK_ARG(!out_f32 || norm, "fav_op_resize_crop: the normalised fp32 output needs norm (mean / std)");
  FAV_CHECK_ARG(static_cast<int64_t>(out_h) * out_w < (1ll << 30) && ratio_h > 0.f && ratio_w > 0.f,
                "fav_op_resize_crop: bad output plane / ratio");
  FAV_CUDA(cudaSetDevice(device));
  LoaderNorm nrm{};
  for (int c = 0; c < 3; ++c) {
    nrm.mean[c] = norm ? norm->mean[c] : 0.f;
    nrm.stdv[c] = norm ? norm->std[c] : 1.f;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int groups = n_frames >= 64 ? ceil_div(n_frames, 4) : n_frames;
  dim3 grid(ceil_div(out_h * out_w, 256), groups < 65535 ? groups : 65535);
  ProfScope prof(PK_OTHER, s, 0.0, static_cast<double>(n_frames) * (3.0 * H * W + 3.0 * out_h * out_w));
  resize_crop_kernel<<<grid, 256, 0, s>>>(frames_u8, n_frames, H, W, resized_h, resized_w, ratio_h, ratio_w, crop_i,
                                          crop_j, out_h, out_w, frames_per_clip, out_u8, out_f32, nrm);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}
