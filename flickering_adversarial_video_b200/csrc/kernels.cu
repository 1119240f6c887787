// kernels.cu — HBM-bound kernels of the attack loop: flicker apply (a), max-pool fwd/bwd, head,
// loss, stem backward collapse, and the fused regulariser-gradient + Adam update on delta (c).
#include "kernels.cuh"

#include <math.h>
#include <algorithm>
#include <stdlib.h>

namespace fav {

// =============================================================================================
// (a) flicker apply
//   reference: x = u8/128 - 1 (utils/pre_process_rgb_flow.py:234);
//              eps_rgb_clip = clip(eps, +-0.4) (utils/kinetics_i3d_utils.py:104-105);
//              adv = clip(x + adv_flag*eps_clip, -1, 1) (:139-142);
//              uint8 view ((adv+1.0)*127.5).astype(uint8) (utils/stats_and_plot/stats_plots.py:57).
//   Each thread owns 16 consecutive pixels of one row (48 bytes of uint8 = three 128-bit loads).
//   The stem input written here is x' = adv - delta' (== x wherever the range clip did not fire);
//   delta' reaches the network as an fp32 per-frame bias of the stem (launch_stem_bias).
// =============================================================================================
// ---- per-lane arithmetic on 16 consecutive pixels (48 uint8 in wds[12]) ----------------------------------------------
// TF stack.  a[48]: adversarial values; xq[32]: 16 x (RG, B0) fp16 pairs of x'; pw0 / pw1: pass nibbles of pixels 0-7 / 8-15
__device__ __forceinline__ void apply_math_tf(const float (&x)[48], const float (&d)[3], float (&a)[48], uint32_t (&xq)[32],
                                              uint32_t& pw0, uint32_t& pw1) {
  pw0 = 0; pw1 = 0;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    float q[3];
    uint32_t m = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float s = __fadd_rn(x[3 * p + c], d[c]);
      const float av = fminf(fmaxf(s, -1.0f), 1.0f);
      a[3 * p + c] = av;
      const bool sat = (s < -1.0f) || (s > 1.0f);
      q[c] = sat ? __fsub_rn(av, d[c]) : x[3 * p + c];
      m |= sat ? (1u << c) : 0u;
    }
    if (p < 8) pw0 |= (~m & 7u) << (4 * p); else pw1 |= (~m & 7u) << (4 * (p - 8));
    xq[2 * p] = pack_f16x2(q[0], q[1]);
    xq[2 * p + 1] = pack_f16x2(q[2], 0.0f);
  }
}

// The I3D clip given as fp32 (the reference's .npy clips are stored normalised, single_video_npy.py:121): the original
// one-thread-per-16-pixels form (a rare path: drivers upload uint8 whenever the floats are on the u8/128 - 1 grid).
__global__ void __launch_bounds__(256)
apply_f32in_kernel(const float* __restrict__ clip, const float* __restrict__ delta, float adv_flag,
                   float dclip, __half* __restrict__ xpad, int Wp, int padl,
                   uint8_t* __restrict__ adv_u8, float* __restrict__ adv_f32,
                   uint32_t* __restrict__ pass_bits, int T, int H, int W, long long groups) {
  pdl_sync();
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= groups) return;
  const int gpr = W >> 4;
  const int wg = static_cast<int>(gid % gpr);
  const long long row = gid / gpr;  // (b*T + t)*H + h
  const int t = static_cast<int>((row / H) % T);
  float d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    d[c] = __fmul_rn(adv_flag, fminf(fmaxf(__ldg(delta + t * 3 + c), -dclip), dclip));
  float x[48];
  const long long e0 = (row * W + wg * 16) * 3;  // first element of this thread
  const float4* src = reinterpret_cast<const float4*>(clip + e0);
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float4 v = __ldg(src + i);
    x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
  }
  float a[48];
  uint32_t xq[32], pw0, pw1;
  apply_math_tf(x, d, a, xq, pw0, pw1);
  uint4* dst = reinterpret_cast<uint4*>(xpad + (row * Wp + padl + wg * 16) * 4);
#pragma unroll
  for (int i = 0; i < 8; ++i) dst[i] = make_uint4(xq[4 * i], xq[4 * i + 1], xq[4 * i + 2], xq[4 * i + 3]);
  if (pass_bits) {
    const long long bt = row / H;
    const int hh = static_cast<int>(row - bt * H);
    uint32_t* dstb = pass_bits + (bt * (H + 7) + hh + 3) * ((((W + 16) >> 3) + 3) & ~3) + 1 + 2 * wg;
    dstb[0] = pw0;
    dstb[1] = pw1;
  }
  if (adv_u8) {
    uint4* du = reinterpret_cast<uint4*>(adv_u8 + e0);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t wv = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          wv |= (__float2uint_rz(__fmul_rn(__fadd_rn(a[16 * i + 4 * j + k], 1.0f), 127.5f)) & 0xffu) << (8 * k);
        o[j] = wv;
      }
      du[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  if (adv_f32) {
    float4* df = reinterpret_cast<float4*>(adv_f32 + e0);
#pragma unroll
    for (int i = 0; i < 12; ++i) df[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
  }
}

// ---- table-driven variant of the hot path (uint8 clip, no adversarial-video output) -------------------------------
// ncu on apply_u8_kernel (profiles/r02_final_ncu_table.md): 60 % of the HBM peak with 68 % of the issue slots busy —
// ~50 instructions per pixel of float arithmetic on values that only depend on (frame, channel, uint8 level).  One
// CTA works inside ONE frame, so the whole map level -> (fp16 stem operand, pass bit) is a 3 x 256-entry table it
// builds once with exactly the arithmetic of the direct kernel (the two formulas below are the ones above, entry by
// entry); a pixel then costs three shared-memory lookups and a handful of byte permutes.  Entry: fp16 bits of x' in
// the low half, pass bit of channel c at bit 16 + c.
template <bool kTorch>
__device__ __forceinline__ uint32_t apply_lut_entry(int u8, int c, float dc, const fav_norm_params& nrm) {
  const float u = static_cast<float>(u8);
  float q;
  bool sat;
  if (!kTorch) {
    const float x = __fsub_rn(__fmul_rn(u, 0.0078125f), 1.0f);
    const float s = __fadd_rn(x, dc);
    const float av = fminf(fmaxf(s, -1.0f), 1.0f);
    sat = (s < -1.0f) || (s > 1.0f);
    q = sat ? __fsub_rn(av, dc) : x;
  } else {
    const float x = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), nrm.mean[c]), nrm.std[c]);
    const float sv = __fadd_rn(x, dc);
    const float av = fminf(fmaxf(sv, nrm.lo), nrm.hi);
    sat = (sv < nrm.lo) || (sv > nrm.hi);
    q = (sat ? ((av - dc) * nrm.std[c] + nrm.mean[c]) * 255.0f : u) - 128.0f;   // centred uint8 units
  }
  return (pack_f16x2(q, 0.0f) & 0xffffu) | (sat ? 0u : (1u << (16 + c)));
}

constexpr int kLutWarps = 6;
template <bool kTorch>
__global__ void __launch_bounds__(kLutWarps * 32, 4)
apply_u8_lut_kernel(const uint8_t* __restrict__ clip, const float* __restrict__ delta, float adv_flag, float dclip,
                    const fav_norm_params nrm, __half* __restrict__ xpad, int Wp, int padl,
                    uint32_t* __restrict__ pass_bits, int T, int H, int W, int chunks_per_frame) {
  __shared__ uint32_t lut[3 * 256];
  __shared__ uint4 s_in[kLutWarps][96];
  __shared__ uint4 s_out[kLutWarps][32 * 9];
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long bt = blockIdx.y;                      // frame index b*T + t
  const int t = static_cast<int>(bt % T);
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
    const int c = i >> 8;
    const float dcl = fminf(fmaxf(__ldg(delta + t * 3 + c), -dclip), dclip);
    const float dc = kTorch ? __fdiv_rn(__fmul_rn(adv_flag, dcl), nrm.std[c]) : __fmul_rn(adv_flag, dcl);
    lut[i] = apply_lut_entry<kTorch>(i & 255, c, dc, nrm);
  }
  __syncthreads();
  const int HW = H * W;
  uint4* in = s_in[warp];
  uint4* out = s_out[warp];
  const uint32_t lut_base = static_cast<uint32_t>(__cvta_generic_to_shared(lut));
  // chunks of 512 pixels of this frame, dealt round-robin to the warps of the CTAs that share the frame
  for (int chunk = blockIdx.x * kLutWarps + warp; chunk < chunks_per_frame; chunk += gridDim.x * kLutWarps) {
    const int q0 = chunk * 512;                          // first pixel of the chunk inside the frame
    const int left = HW - q0;
    const int nv = left < 512 ? left : 512;              // a multiple of 16 (W % 16 == 0)
    {
      const uint4* src = reinterpret_cast<const uint4*>(clip + (bt * HW + q0) * 3);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int i = j * 32 + lane;
        if (i * 16 < nv * 3) in[i] = __ldg(src + i);
      }
    }
    __syncwarp();
    const bool active = lane * 16 < nv;
    const int qp = q0 + lane * 16;
    const int h = qp / W;
    const int w0 = qp - h * W;
    const long long dst_off = ((bt * H + h) * Wp + padl + w0) * 4;
    uint32_t xq[32], pw0 = 0, pw1 = 0;
    if (active) {
      uint32_t wds[12];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const uint4 v = in[lane * 3 + i];
        wds[4 * i] = v.x; wds[4 * i + 1] = v.y; wds[4 * i + 2] = v.z; wds[4 * i + 3] = v.w;
      }
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        uint32_t ent[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int e = 3 * p + c;
          const uint32_t byte4 = ((wds[e >> 2] >> (8 * (e & 3))) & 0xffu) << 2;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ent[c]) : "r"(lut_base + c * 1024u + byte4));
        }
        xq[2 * p] = __byte_perm(ent[0], ent[1], 0x5410);
        xq[2 * p + 1] = ent[2] & 0xffffu;
        const uint32_t nib = ((ent[0] | ent[1] | ent[2]) >> 16) & 7u;
        if (p < 8) pw0 |= nib << (4 * p); else pw1 |= nib << (4 * (p - 8));
      }
    }
    if (!kTorch) {
      if (active) {
#pragma unroll
        for (int i = 0; i < 8; ++i) out[lane * 9 + i] = make_uint4(xq[4 * i], xq[4 * i + 1], xq[4 * i + 2], xq[4 * i + 3]);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = j * 32 + lane;
        const int L = i >> 3, part = i & 7;
        const long long off = __shfl_sync(0xffffffffu, dst_off, L);
        if (L * 16 < nv) *reinterpret_cast<uint4*>(xpad + off + part * 8) = out[L * 9 + part];
      }
    } else {
      uint2* out2 = reinterpret_cast<uint2*>(out);
      if (active) {
#pragma unroll
        for (int i = 0; i < 16; ++i) out2[lane * 17 + i] = make_uint2(xq[2 * i], xq[2 * i + 1]);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int i = j * 32 + lane;
        const int L = i >> 4, part = i & 15;
        const long long off = __shfl_sync(0xffffffffu, dst_off, L);
        if (L * 16 < nv) *reinterpret_cast<uint2*>(xpad + off + part * 4) = out2[L * 17 + part];
      }
    }
    if (pass_bits && active) {
      uint32_t* dstb = pass_bits + (bt * (H + 7) + h + 3) * ((((W + 16) >> 3) + 3) & ~3) + 1 + (w0 >> 3);
      dstb[0] = pw0;
      dstb[1] = pw1;
    }
    __syncwarp();
  }
}

// grid of the table-driven kernel: enough CTAs per frame that the table (768 entries) costs < 10 % of the pixel work
static dim3 apply_lut_grid(int B, int T, int H, int W, int* chunks_per_frame) {
  const int chunks = ceil_div(H * W, 512);
  *chunks_per_frame = chunks;
  const int per_cta = 2 * kLutWarps;                    // two 512-pixel chunks per warp
  return dim3(static_cast<unsigned>(std::max(1, ceil_div(chunks, per_cta))), static_cast<unsigned>(B * T), 1);
}
// Measured (B200, bench.py kernels.apply): torch stack 16 x 16 x 112^2: 36 -> 25 us (its normalisation has three
// divisions per entry); TF stack 8 x 64 x 224^2: 69 -> 75 us (the direct arithmetic is cheap there and the lookups pay
// ~3-way bank conflicts) — so the table is the default for the torch stack only.  FAV_APPLY_LUT=0 / 1 forces either
// kernel (bit-identity test); read per call.
static bool apply_use_lut(bool torch_stack) {
  const char* ev = getenv("FAV_APPLY_LUT");
  return ev ? atoi(ev) != 0 : torch_stack;
}

// uint8 clips, both stacks.  One warp owns 512 consecutive pixels: the 1536 input bytes arrive as three fully coalesced
// 512-byte loads and are re-dealt through shared memory so that every lane computes 16 whole pixels (conflict-free: a
// lane's 48 bytes start 12 banks after its neighbour's); the 4 KB of fp16 RGBX output (and the uint8 adversarial video)
// go back through shared memory the other way, so that every store instruction of the warp covers 512 (torch stack:
// 256) contiguous bytes.  Round 1 had each lane store its own 128-byte run 16 bytes at a time: 32 half-used sectors
// per instruction and 43 % of the HBM peak.
// kAdv = false (the attack loop: neither adversarial-video output is requested) drops the 48 adversarial values per
// lane from the register budget, which is what sets the occupancy of this latency-bound kernel.
constexpr int kApplyWarps = 4;
template <bool kTorch, bool kAdv>
__global__ void __launch_bounds__(kApplyWarps * 32, kAdv ? 3 : 6)
apply_u8_kernel(const uint8_t* __restrict__ clip, const float* __restrict__ delta, float adv_flag, float dclip,
                const fav_norm_params nrm, __half* __restrict__ xpad, int Wp, int padl, uint8_t* __restrict__ adv_u8,
                float* __restrict__ adv_f32, uint32_t* __restrict__ pass_bits, int T, int H, int W, long long npix) {
  __shared__ uint4 s_in[kApplyWarps][96];
  __shared__ uint4 s_out[kApplyWarps][32 * 9];
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long p0 = (static_cast<long long>(blockIdx.x) * kApplyWarps + warp) * 512;
  if (p0 >= npix) return;                                             // warp-uniform
  const long long left = npix - p0;
  const int nv = left < 512 ? static_cast<int>(left) : 512;           // valid pixels of the chunk (a multiple of 16)
  uint4* in = s_in[warp];
  uint4* out = s_out[warp];
  {
    const uint4* src = reinterpret_cast<const uint4*>(clip + p0 * 3);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int i = j * 32 + lane;
      if (i * 16 < nv * 3) in[i] = __ldg(src + i);
    }
  }
  __syncwarp();
  const bool active = lane * 16 < nv;
  const long long px = p0 + lane * 16;
  const long long row = px / W;                                       // (b*T + t)*H + h
  const int w0 = static_cast<int>(px - row * W);
  const long long bt = row / H;
  const int h = static_cast<int>(row - bt * H);
  const int t = static_cast<int>(bt % T);
  const long long b = bt / T;
  const long long dst_off = (row * Wp + padl + w0) * 4;               // element offset of this lane's run in xpad
  float a[48];
  uint32_t xq[32], pw0 = 0, pw1 = 0;
  if (active) {
    uint32_t wds[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint4 v = in[lane * 3 + i];
      wds[4 * i] = v.x; wds[4 * i + 1] = v.y; wds[4 * i + 2] = v.z; wds[4 * i + 3] = v.w;
    }
    float d[3];
    if (!kTorch) {
#pragma unroll
      for (int c = 0; c < 3; ++c) d[c] = __fmul_rn(adv_flag, fminf(fmaxf(__ldg(delta + t * 3 + c), -dclip), dclip));
      float x[48];
#pragma unroll
      for (int e = 0; e < 48; ++e)
        x[e] = __fsub_rn(__fmul_rn(static_cast<float>((wds[e >> 2] >> (8 * (e & 3))) & 0xffu), 0.0078125f), 1.0f);
      apply_math_tf(x, d, a, xq, pw0, pw1);
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c)
        d[c] = __fdiv_rn(__fmul_rn(adv_flag, fminf(fmaxf(__ldg(delta + t * 3 + c), -dclip), dclip)), nrm.std[c]);
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        float q[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int e = 3 * p + c;
          const float u = static_cast<float>((wds[e >> 2] >> (8 * (e & 3))) & 0xffu);
          const float x = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), nrm.mean[c]), nrm.std[c]);
          const float sv = __fadd_rn(x, d[c]);
          const float av = fminf(fmaxf(sv, nrm.lo), nrm.hi);
          a[e] = av;
          const bool sat = (sv < nrm.lo) || (sv > nrm.hi);
          q[c] = (sat ? ((av - d[c]) * nrm.std[c] + nrm.mean[c]) * 255.0f : u) - 128.0f;   // centred uint8 units
          if (!sat) { if (p < 8) pw0 |= 1u << (4 * p + c); else pw1 |= 1u << (4 * (p - 8) + c); }
        }
        xq[2 * p] = pack_f16x2(q[0], q[1]);
        xq[2 * p + 1] = pack_f16x2(q[2], 0.0f);
      }
    }
  }
  // ---- x' (fp16 RGBX): lane-major into shared memory, position-major out of it ----
  if (!kTorch) {
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) out[lane * 9 + i] = make_uint4(xq[4 * i], xq[4 * i + 1], xq[4 * i + 2], xq[4 * i + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = j * 32 + lane;
      const int L = i >> 3, part = i & 7;
      const long long off = __shfl_sync(0xffffffffu, dst_off, L);
      if (L * 16 < nv) *reinterpret_cast<uint4*>(xpad + off + part * 8) = out[L * 9 + part];
    }
  } else {
    // torch pads 3 columns on the left: positions are 8 bytes, so the runs are 8-byte (not 16-byte) aligned
    uint2* out2 = reinterpret_cast<uint2*>(out);
    if (active) {
#pragma unroll
      for (int i = 0; i < 16; ++i) out2[lane * 17 + i] = make_uint2(xq[2 * i], xq[2 * i + 1]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int i = j * 32 + lane;
      const int L = i >> 4, part = i & 15;
      const long long off = __shfl_sync(0xffffffffu, dst_off, L);
      if (L * 16 < nv) *reinterpret_cast<uint2*>(xpad + off + part * 4) = out2[L * 17 + part];
    }
  }
  if (pass_bits && active) {
    // one nibble per pixel, bit c = entry (pixel, c) passes the gradient; 8 zero nibbles left of w = 0, 3 zero rows above h = 0
    uint32_t* dstb = pass_bits + (bt * (H + 7) + h + 3) * ((((W + 16) >> 3) + 3) & ~3) + 1 + (w0 >> 3);
    dstb[0] = pw0;
    dstb[1] = pw1;
  }
  if (kAdv && !kTorch && adv_u8) {
    // uint8 view ((adv + 1.0) * 127.5).astype(uint8) (stats_plots.py:57): back through the input staging buffer
    if (active) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t wv = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            wv |= (__float2uint_rz(__fmul_rn(__fadd_rn(a[16 * i + 4 * j + k], 1.0f), 127.5f)) & 0xffu) << (8 * k);
          o[j] = wv;
        }
        in[lane * 3 + i] = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    __syncwarp();
    uint4* du = reinterpret_cast<uint4*>(adv_u8 + p0 * 3);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int i = j * 32 + lane;
      if (i * 16 < nv * 3) du[i] = in[i];
    }
  }
  if (kAdv && adv_f32 && active) {
    if (!kTorch) {           // NTHWC like the clip
      float4* df = reinterpret_cast<float4*>(adv_f32 + px * 3);
#pragma unroll
      for (int i = 0; i < 12; ++i) df[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
    } else {                 // NCTHW like the torch tensors of the reference
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float* o = adv_f32 + (((b * 3 + c) * T + t) * H + h) * static_cast<long long>(W) + w0;
#pragma unroll
        for (int p = 0; p < 16; p += 4)
          *reinterpret_cast<float4*>(o + p) = make_float4(a[3 * p + c], a[3 * (p + 1) + c], a[3 * (p + 2) + c], a[3 * (p + 3) + c]);
      }
    }
  }
}

int launch_apply(const void* clip, int in_dtype, const float* delta, float adv_flag, float delta_clip,
                 __half* xpad, int Wp, int padl, uint8_t* adv_u8, float* adv_f32,
                 uint32_t* pass_bits, int B, int T, int H, int W, cudaStream_t s) {
  ProfScope ps(PK_APPLY, s, 0.0, static_cast<double>(B) * T * H * W * (3.0 * (in_dtype == FAV_F32 ? 4 : 1) + 8.0 + (pass_bits ? 0.5 : 0.0) + (adv_u8 ? 3.0 : 0.0) + (adv_f32 ? 12.0 : 0.0)));
  FAV_CHECK_ARG(W % 16 == 0, "apply: W=%d must be a multiple of 16", W);
  FAV_CHECK_ARG(static_cast<long long>(B) * T * H * W < (1ll << 28), "apply: too many pixels per call");
  const long long npix = static_cast<long long>(B) * T * H * W;
  if (in_dtype == FAV_F32) {
    const long long groups = npix / 16;
    FAV_CUDA(launch_pdl(apply_f32in_kernel, static_cast<int>(ceil_div64(groups, 256)), 256, 0, s, static_cast<const float*>(clip),
                        delta, adv_flag, delta_clip, xpad, Wp, padl, adv_u8, adv_f32, pass_bits, T, H, W, groups));
  } else {
    fav_norm_params none{};
    const int grid = static_cast<int>(ceil_div64(npix, 512 * kApplyWarps));
    if (adv_u8 || adv_f32)
      FAV_CUDA(launch_pdl(apply_u8_kernel<false, true>, grid, kApplyWarps * 32, 0, s, static_cast<const uint8_t*>(clip), delta,
                          adv_flag, delta_clip, none, xpad, Wp, padl, adv_u8, adv_f32, pass_bits, T, H, W, npix));
    else if (apply_use_lut(false) && B * T <= 65535) {
      int cpf = 0;
      const dim3 g2 = apply_lut_grid(B, T, H, W, &cpf);
      FAV_CUDA(launch_pdl(apply_u8_lut_kernel<false>, g2, kLutWarps * 32, 0, s, static_cast<const uint8_t*>(clip), delta, adv_flag,
                          delta_clip, none, xpad, Wp, padl, pass_bits, T, H, W, cpf));
    } else
      FAV_CUDA(launch_pdl(apply_u8_kernel<false, false>, grid, kApplyWarps * 32, 0, s, static_cast<const uint8_t*>(clip), delta,
                          adv_flag, delta_clip, none, xpad, Wp, padl, adv_u8, adv_f32, pass_bits, T, H, W, npix));
  }
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

// ---------------------------------------------------------------------------------------------
// delta-dependent stem bias: conv(x' + delta') = conv(x') + bias[t_o, hclass, wclass, co]
// ---------------------------------------------------------------------------------------------
// table[to][cls][co] = bnbias[co] + sum_{kt valid} sum_c (cst[c] + dscale[c] * delta'[t,c]) * wc[kt][cls][c][co],
// t = st*to + kt - pt.  I3D: cst = 0, dscale = 1.  Torch stack: cst = -mean/std (the stem input is kept in
// uint8 units), dscale = 1/std (F.normalize(delta, 0, std), model.py:89).
__global__ void stem_bias_kernel(const float* __restrict__ delta, float adv_flag, float dclip,
                                 const float* __restrict__ wc, const float* __restrict__ bnbias,
                                 float* __restrict__ table, int T, int To, int pt, int KT, int st, int C1,
                                 float cst0, float cst1, float cst2, float ds0, float ds1, float ds2) {
  const int to = blockIdx.x >> 4;
  const int cls = blockIdx.x & 15;
  const int co = threadIdx.x;
  if (co >= C1) return;
  const float cst[3] = {cst0, cst1, cst2}, ds[3] = {ds0, ds1, ds2};
  float acc = bnbias[co];
  for (int kt = 0; kt < KT; ++kt) {
    const int t = st * to + kt - pt;
    if (t < 0 || t >= T) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = cst[c] + ds[c] * (adv_flag * fminf(fmaxf(delta[t * 3 + c], -dclip), dclip));
      acc = fmaf(d, wc[((kt * 16 + cls) * 3 + c) * C1 + co], acc);
    }
  }
  table[(to * 16 + cls) * C1 + co] = acc;
  (void)To;
}

int launch_stem_bias(const float* delta, float adv_flag, float delta_clip, const float* wc,
                     const float* bnbias, float* table, int T, int To, int pt, cudaStream_t s) {
  ProfScope ps(PK_OTHER, s);
  stem_bias_kernel<<<To * 16, 64, 0, s>>>(delta, adv_flag, delta_clip, wc, bnbias, table, T, To, pt, 7, 2, 64,
                                          0.f, 0.f, 0.f, 1.f, 1.f, 1.f);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

int launch_stem_bias_ex(const float* delta, float adv_flag, float delta_clip, const float* wc, const float* bnbias,
                        float* table, int T, int To, int pt, int KT, int st, int C1, const float* cst3,
                        const float* dscale3, cudaStream_t s) {
  ProfScope ps(PK_OTHER, s);
  FAV_CHECK_ARG(C1 <= 64, "stem bias: at most 64 stem channels");
  stem_bias_kernel<<<To * 16, 64, 0, s>>>(delta, adv_flag, delta_clip, wc, bnbias, table, T, To, pt, KT, st, C1,
                                          cst3[0], cst3[1], cst3[2], dscale3[0], dscale3[1], dscale3[2]);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

// =============================================================================================
// torch-stack flicker apply (Perturbation.forward, utils_cv/action_recognition/model.py:80-96):
//   x = (u8/255 - mean_c)/std_c  (functional_video.py:65-97, dataset.py:28-29)
//   adv = clamp(x + adv_flag*clamp(delta, +-max_norm)/std_c, min_value, max_value)   (scalar bounds :72-75)
// The stem input is written in centred uint8 units (u - 128, exact in fp16): x' = u - 128 where the clamp did not
// fire, else the value that reproduces adv under the folded normalisation; delta and the -mean/std constant reach the
// network through the fp32 stem bias table.
// =============================================================================================
int launch_apply_torch(const uint8_t* clip, const float* delta, float adv_flag, float delta_clip,
                       const fav_norm_params& nrm, __half* xpad, int Wp, int padl, float* adv_f32,
                       uint32_t* pass_bits, int B, int T, int H, int W, cudaStream_t s) {
  ProfScope ps(PK_APPLY, s, 0.0, static_cast<double>(B) * T * H * W * (3.0 + 8.0 + (pass_bits ? 0.5 : 0.0) + (adv_f32 ? 12.0 : 0.0)));
  FAV_CHECK_ARG(W % 16 == 0, "apply: W=%d must be a multiple of 16", W);
  const long long npix = static_cast<long long>(B) * T * H * W;
  // the same warp-per-512-pixels kernel as the TF stack (apply_u8_kernel above), with the torch normalisation
  const int grid = static_cast<int>(ceil_div64(npix, 512 * kApplyWarps));
  if (adv_f32)
    apply_u8_kernel<true, true><<<grid, kApplyWarps * 32, 0, s>>>(clip, delta, adv_flag, delta_clip, nrm, xpad, Wp, padl,
                                                                  nullptr, adv_f32, pass_bits, T, H, W, npix);
  else if (apply_use_lut(true) && B * T <= 65535) {
    int cpf = 0;
    const dim3 g2 = apply_lut_grid(B, T, H, W, &cpf);
    apply_u8_lut_kernel<true><<<g2, kLutWarps * 32, 0, s>>>(clip, delta, adv_flag, delta_clip, nrm, xpad, Wp, padl, pass_bits, T,
                                                            H, W, cpf);
  } else
    apply_u8_kernel<true, false><<<grid, kApplyWarps * 32, 0, s>>>(clip, delta, adv_flag, delta_clip, nrm, xpad, Wp, padl,
                                                                   nullptr, nullptr, pass_bits, T, H, W, npix);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

// =============================================================================================
// (c) dense reduce: g[t,c] = dscale_c * sum_{b,h,w} mask(b,t,h,w,c) * dX[b,t,h,w,c]
//   dX is the stem data gradient (bf16, NDHWC with 16-channel pixels, channels 0-2 used); the range-clip
//   mask (lo <= x + delta' <= hi, inclusive like tf.clip_by_value / torch.clamp) is recomputed from the
//   uint8 clip.  Warp-shuffle + shared-memory tree per block, then a fixed-order second pass: deterministic.
//   reference: compute_gradients(loss, var_list=perturbation) single_video_npy.py:82 / loss.backward() model.py:732
// =============================================================================================
constexpr int kRedRows = 8;
__global__ void __launch_bounds__(256)
stem_dx_reduce_kernel(const __nv_bfloat16* __restrict__ dx, const uint8_t* __restrict__ clip,
                      const float* __restrict__ delta, float adv_flag, float dclip, const fav_norm_params nrm,
                      int torch_mode, float* __restrict__ partial, int T, int H, int W, int chunks) {
  __shared__ float red[8][3];
  const int chunk = blockIdx.x % chunks;
  const int t = blockIdx.x / chunks;
  const int b = blockIdx.y;
  float d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float dc = __fmul_rn(adv_flag, fminf(fmaxf(__ldg(delta + t * 3 + c), -dclip), dclip));
    d[c] = torch_mode ? __fdiv_rn(dc, nrm.std[c]) : dc;
  }
  float acc[3] = {0.f, 0.f, 0.f};
  const int h0 = chunk * kRedRows, h1 = min(H, h0 + kRedRows);
  const long long base = ((static_cast<long long>(b) * T + t) * H) * W;
  const int npx = (h1 - h0) * W;
  for (int i = threadIdx.x; i < npx; i += blockDim.x) {
    const long long px = base + static_cast<long long>(h0) * W + i;
    const uint2 g = __ldg(reinterpret_cast<const uint2*>(dx + px * 16));   // channels 0..3
    const float gv[3] = {bf16_lo(g.x), bf16_hi(g.x), bf16_lo(g.y)};
    const uint8_t* up = clip + px * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float u = static_cast<float>(__ldg(up + c));
      const float x = torch_mode ? __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), nrm.mean[c]), nrm.std[c])
                                 : __fsub_rn(__fmul_rn(u, 0.0078125f), 1.0f);
      const float sv = __fadd_rn(x, d[c]);
      if (sv >= nrm.lo && sv <= nrm.hi) acc[c] += gv[c];
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
    if (lane == 0) red[wid][c] = acc[c];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
    partial[((static_cast<long long>(b) * T + t) * chunks + chunk) * 3 + threadIdx.x] = sum;
  }
}

__global__ void stem_dx_final_kernel(const float* __restrict__ partial, float* __restrict__ grad, int B, int T,
                                     int chunks, float s0, float s1, float s2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * 3) return;
  const int t = i / 3, c = i - t * 3;
  float sum = 0.f;
  for (int b = 0; b < B; ++b)
    for (int k = 0; k < chunks; ++k) sum += partial[((static_cast<long long>(b) * T + t) * chunks + k) * 3 + c];
  grad[i] = sum * (c == 0 ? s0 : (c == 1 ? s1 : s2));
}

int launch_stem_dx_reduce(const __nv_bfloat16* dx, const uint8_t* clip, const float* delta, float adv_flag,
                          float delta_clip, const fav_norm_params& nrm, int torch_mode, float* partial, float* grad,
                          int B, int T, int H, int W, cudaStream_t s) {
  ProfScope ps(PK_STEM_BWD, s, 0.0, static_cast<double>(B) * T * H * W * 9.0);
  const int chunks = ceil_div(H, kRedRows);
  stem_dx_reduce_kernel<<<dim3(T * chunks, B), 256, 0, s>>>(dx, clip, delta, adv_flag, delta_clip, nrm, torch_mode,
                                                            partial, T, H, W, chunks);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  const float s0 = torch_mode ? 1.0f / nrm.std[0] : 1.0f, s1 = torch_mode ? 1.0f / nrm.std[1] : 1.0f,
              s2 = torch_mode ? 1.0f / nrm.std[2] : 1.0f;
  stem_dx_final_kernel<<<ceil_div(T * 3, 128), 128, 0, s>>>(partial, grad, B, T, chunks, s0, s1, s2);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}
int stem_dx_reduce_chunks(int H) { return ceil_div(H, kRedRows); }

// =============================================================================================
// MaxPool3d, TF SAME semantics (i3d.py:174,189,212,252,398): -inf padding, first arg-max wins.
// =============================================================================================
PoolGeom make_pool_geom(int B, int T, int H, int W, int C, int kt, int kh, int kw, int st, int sh, int sw) {
  PoolGeom g;
  g.B = B; g.T = T; g.H = H; g.W = W; g.C = C;
  g.kt = kt; g.kh = kh; g.kw = kw; g.st = st; g.sh = sh; g.sw = sw;
  g.To = ceil_div(T, st); g.Ho = ceil_div(H, sh); g.Wo = ceil_div(W, sw);
  auto padb = [](int in, int out, int k, int s) {
    int total = (out - 1) * s + k - in;
    if (total < 0) total = 0;
    return total / 2;
  };
  g.pt = padb(T, g.To, kt, st); g.ph = padb(H, g.Ho, kh, sh); g.pw = padb(W, g.Wo, kw, sw);
  return g;
}

// forward: one thread = one output position x 8 channels, packed fp16x2 compare/select
// (3 ALU ops per tap per channel pair); 2-D grid so the only runtime division is by C/8.
template <int KT, int KH, int KW, int ST, int SH, int SW>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                   uint8_t* __restrict__ idx, const PoolGeom gg) {
  pdl_sync();
  PoolGeom g = gg;
  if (KT > 0) { g.kt = KT; g.kh = KH; g.kw = KW; g.st = ST; g.sh = SH; g.sw = SW; }
  const int cg = g.C >> 3;
  const int i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= g.Wo * cg) return;
  const int wo = i / cg;
  const int c8 = i - wo * cg;
  int row = blockIdx.x;
  const int ho = row % g.Ho; row /= g.Ho;
  const int to = row % g.To;
  const int b = row / g.To;
  __half2 best[4];
  uint32_t bidx[4];
  const __half2 ninf = __halves2half2(__ushort_as_half(0xFC00), __ushort_as_half(0xFC00));
#pragma unroll
  for (int j = 0; j < 4; ++j) { best[j] = ninf; bidx[j] = 0; }
  const int t0 = to * g.st - g.pt, h0 = ho * g.sh - g.ph, w0 = wo * g.sw - g.pw;
#pragma unroll
  for (int dt = 0; dt < g.kt; ++dt) {
    const int t = t0 + dt;
    if (t < 0 || t >= g.T) continue;
#pragma unroll
    for (int dh = 0; dh < g.kh; ++dh) {
      const int h = h0 + dh;
      if (h < 0 || h >= g.H) continue;
      const __half* rowp = x + ((static_cast<long long>(b) * g.T + t) * g.H + h) * g.W * g.C + c8 * 8;
      const uint32_t tap0 = static_cast<uint32_t>((dt * g.kh + dh) * g.kw);
#pragma unroll
      for (int dw = 0; dw < g.kw; ++dw) {
        const int w = w0 + dw;
        if (w < 0 || w >= g.W) continue;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(rowp + w * g.C));
        const uint32_t tap2 = (tap0 + dw) * 0x00010001u;
        const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __half2 vv = *reinterpret_cast<const __half2*>(&wv[j]);
          const uint32_t m = __hgt2_mask(vv, best[j]);   // strict >: the first arg-max wins
          best[j] = __hmax2(best[j], vv);
          bidx[j] = (bidx[j] & ~m) | (tap2 & m);
        }
      }
    }
  }
  const long long ooff = ((static_cast<long long>(blockIdx.x) * g.Wo + wo) * cg + c8) * 8;
  uint4 o;
  o.x = *reinterpret_cast<uint32_t*>(&best[0]); o.y = *reinterpret_cast<uint32_t*>(&best[1]);
  o.z = *reinterpret_cast<uint32_t*>(&best[2]); o.w = *reinterpret_cast<uint32_t*>(&best[3]);
  *reinterpret_cast<uint4*>(y + ooff) = o;
  if (idx) {
    uint2 iv;   // 2 x u16 -> 2 bytes per word
    iv.x = __byte_perm(bidx[0], bidx[1], 0x6420);
    iv.y = __byte_perm(bidx[2], bidx[3], 0x6420);
    *reinterpret_cast<uint2*>(idx + ooff) = iv;
  }
}

// Loads-first variant for the pools of the I3D trunk (compile-time window).  ncu on the generic kernel above
// (profiles/r02_c1_ncu_full_details.txt, MaxPool3d_2a): 346 instructions per thread, 70 % of the warp cycles stalled on
// L1TEX with one load outstanding at a time — the bounds tests around each tap keep the compiler from hoisting the
// loads, so the kernel ran at the memory latency (46 % of HBM).  Here every tap of one temporal slice is loaded
// unconditionally from a clamped (always valid) address before the first compare; taps outside the input become -inf,
// which a strict > never selects, so the first-arg-max rule and the tap numbering are unchanged.
template <int KT, int KH, int KW, int ST, int SH, int SW>
__global__ void __launch_bounds__(256)
maxpool_fwd_win_kernel(const __half* __restrict__ x, __half* __restrict__ y, uint8_t* __restrict__ idx, const PoolGeom g) {
  pdl_sync();
  const int cg = g.C >> 3;
  const int i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= g.Wo * cg) return;
  const int wo = i / cg;
  const int c8 = i - wo * cg;
  int row = blockIdx.x;
  const int ho = row % g.Ho; row /= g.Ho;
  const int to = row % g.To;
  const int b = row / g.To;
  const int t0 = to * ST - g.pt, h0 = ho * SH - g.ph, w0 = wo * SW - g.pw;
  uint32_t best[4] = {kF16NegInf2, kF16NegInf2, kF16NegInf2, kF16NegInf2};
  uint32_t bidx[4] = {0u, 0u, 0u, 0u};
  int woff[KW];
  bool wok[KW];
#pragma unroll
  for (int dw = 0; dw < KW; ++dw) {
    const int w = w0 + dw;
    wok[dw] = w >= 0 && w < g.W;
    woff[dw] = min(max(w, 0), g.W - 1) * g.C;
  }
#pragma unroll
  for (int dt = 0; dt < KT; ++dt) {
    const int t = t0 + dt;
    const bool tok = t >= 0 && t < g.T;
    const int tc = min(max(t, 0), g.T - 1);
    uint4 v[KH][KW];
    bool hok[KH];
#pragma unroll
    for (int dh = 0; dh < KH; ++dh) {
      const int h = h0 + dh;
      hok[dh] = tok && h >= 0 && h < g.H;
      const int hc = min(max(h, 0), g.H - 1);
      const __half* rowp = x + ((static_cast<long long>(b) * g.T + tc) * g.H + hc) * g.W * g.C + c8 * 8;
#pragma unroll
      for (int dw = 0; dw < KW; ++dw) v[dh][dw] = __ldg(reinterpret_cast<const uint4*>(rowp + woff[dw]));
    }
#pragma unroll
    for (int dh = 0; dh < KH; ++dh)
#pragma unroll
      for (int dw = 0; dw < KW; ++dw) {
        const bool ok = hok[dh] && wok[dw];
        const uint32_t tap2 = static_cast<uint32_t>((dt * KH + dh) * KW + dw) * 0x00010001u;
        const uint32_t wv[4] = {v[dh][dw].x, v[dh][dw].y, v[dh][dw].z, v[dh][dw].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t u = ok ? wv[j] : kF16NegInf2;
          const __half2 vv = *reinterpret_cast<const __half2*>(&u);
          const __half2 bb = *reinterpret_cast<const __half2*>(&best[j]);
          const uint32_t m = __hgt2_mask(vv, bb);   // strict >: the first arg-max wins
          const __half2 mx = __hmax2(bb, vv);
          best[j] = *reinterpret_cast<const uint32_t*>(&mx);
          bidx[j] = (bidx[j] & ~m) | (tap2 & m);
        }
      }
  }
  const long long ooff = ((static_cast<long long>(blockIdx.x) * g.Wo + wo) * cg + c8) * 8;
  *reinterpret_cast<uint4*>(y + ooff) = make_uint4(best[0], best[1], best[2], best[3]);
  if (idx) {
    uint2 iv;
    iv.x = __byte_perm(bidx[0], bidx[1], 0x6420);
    iv.y = __byte_perm(bidx[2], bidx[3], 0x6420);
    *reinterpret_cast<uint2*>(idx + ooff) = iv;
  }
}

int launch_maxpool_fwd(const __half* x, __half* y, uint8_t* idx, const PoolGeom& g,
                       cudaStream_t s) {
  ProfScope ps(PK_POOL_FWD, s, 0.0, static_cast<double>(g.B) * g.C * (2.0 * g.T * g.H * g.W + 3.0 * g.To * g.Ho * g.Wo));
  FAV_CHECK_ARG(g.C % 8 == 0, "maxpool: C=%d must be a multiple of 8", g.C);
  FAV_CHECK_ARG(g.kt * g.kh * g.kw <= 255, "maxpool: window too large");
  if (pool3s1_applicable(g)) return launch_pool3s1_fwd(x, y, idx, g, s);   // idx == nullptr: forward-only plan, no codes
  dim3 grid(g.B * g.To * g.Ho, ceil_div(g.Wo * (g.C / 8), 256));
  const int key = ((g.kt * 10 + g.kh) * 10 + g.kw) * 1000 + (g.st * 10 + g.sh) * 10 + g.sw;
  static int win = -1;   // FAV_POOL_FWD_WIN=0: the generic tap-by-tap kernel (A/B)
  if (win < 0) {
    const char* ev = getenv("FAV_POOL_FWD_WIN");
    win = (ev && atoi(ev) == 0) ? 0 : 1;
  }
  switch (win ? key : -key) {
    case 133122: FAV_CUDA(launch_pdl(maxpool_fwd_win_kernel<1, 3, 3, 1, 2, 2>, grid, 256, 0, s, x, y, idx, g)); break;
    case 333222: FAV_CUDA(launch_pdl(maxpool_fwd_win_kernel<3, 3, 3, 2, 2, 2>, grid, 256, 0, s, x, y, idx, g)); break;
    case 222222: FAV_CUDA(launch_pdl(maxpool_fwd_win_kernel<2, 2, 2, 2, 2, 2>, grid, 256, 0, s, x, y, idx, g)); break;
    case -133122: FAV_CUDA(launch_pdl(maxpool_fwd_kernel<1, 3, 3, 1, 2, 2>, grid, 256, 0, s, x, y, idx, g)); break;
    case 333111: case -333111: FAV_CUDA(launch_pdl(maxpool_fwd_kernel<3, 3, 3, 1, 1, 1>, grid, 256, 0, s, x, y, idx, g)); break;
    case -333222: FAV_CUDA(launch_pdl(maxpool_fwd_kernel<3, 3, 3, 2, 2, 2>, grid, 256, 0, s, x, y, idx, g)); break;
    case -222222: FAV_CUDA(launch_pdl(maxpool_fwd_kernel<2, 2, 2, 2, 2, 2>, grid, 256, 0, s, x, y, idx, g)); break;
    default: FAV_CUDA(launch_pdl(maxpool_fwd_kernel<0, 0, 0, 0, 0, 0>, grid, 256, 0, s, x, y, idx, g)); break;
  }
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

// backward in gather form: every input element sums the windows whose recorded arg-max is itself
template <int KT, int KH, int KW, int ST, int SH, int SW>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                   const __nv_bfloat16* __restrict__ addend, const __half* __restrict__ relu_src,
                   __nv_bfloat16* __restrict__ dx, const PoolGeom gg) {
  pdl_sync();
  PoolGeom g = gg;
  if (KT > 0) { g.kt = KT; g.kh = KH; g.kw = KW; g.st = ST; g.sh = SH; g.sw = SW; }
  const int cg = g.C >> 3;
  const int i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= g.W * cg) return;
  const int w = i / cg;
  const int c8 = i - w * cg;
  int row = blockIdx.x;
  const int h = row % g.H; row /= g.H;
  const int t = row % g.T;
  const int b = row / g.T;
  const long long eoff = ((static_cast<long long>(blockIdx.x) * g.W + w) * cg + c8) * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  if (addend) {
    const uint4 a = *reinterpret_cast<const uint4*>(addend + eoff);
    acc[0] = bf16_lo(a.x); acc[1] = bf16_hi(a.x); acc[2] = bf16_lo(a.y); acc[3] = bf16_hi(a.y);
    acc[4] = bf16_lo(a.z); acc[5] = bf16_hi(a.z); acc[6] = bf16_lo(a.w); acc[7] = bf16_hi(a.w);
  }
  // windows (to) containing t: to*st - pt <= t <= to*st - pt + kt - 1
  const int tp = t + g.pt, hp = h + g.ph, wp = w + g.pw;
  int to_lo = tp - g.kt + g.st; to_lo = to_lo < 0 ? 0 : to_lo / g.st;
  int ho_lo = hp - g.kh + g.sh; ho_lo = ho_lo < 0 ? 0 : ho_lo / g.sh;
  int wo_lo = wp - g.kw + g.sw; wo_lo = wo_lo < 0 ? 0 : wo_lo / g.sw;
  const int to_hi = min(tp / g.st, g.To - 1);
  const int ho_hi = min(hp / g.sh, g.Ho - 1);
  const int wo_hi = min(wp / g.sw, g.Wo - 1);
  for (int to = to_lo; to <= to_hi; ++to) {
    const int dt = tp - to * g.st;
    for (int ho = ho_lo; ho <= ho_hi; ++ho) {
      const int dh = hp - ho * g.sh;
      const long long rbase = (((static_cast<long long>(b) * g.To + to) * g.Ho + ho) * g.Wo) * g.C + c8 * 8;
#pragma unroll 3
      for (int wo = wo_lo; wo <= wo_hi; ++wo) {
        const int dw = wp - wo * g.sw;
        const uint32_t tap4 = static_cast<uint32_t>((dt * g.kh + dh) * g.kw + dw) * 0x01010101u;
        const long long off = rbase + wo * g.C;
        const uint2 iv = __ldg(reinterpret_cast<const uint2*>(idx + off));
        const uint32_t ex = __vcmpeq4(iv.x, tap4);   // 0xff per channel whose arg-max is this element
        const uint32_t ey = __vcmpeq4(iv.y, tap4);
        if ((ex | ey) == 0u) continue;
        const uint4 d = __ldg(reinterpret_cast<const uint4*>(dy + off));
        const uint32_t m0 = __byte_perm(ex, 0, 0x1100), m1 = __byte_perm(ex, 0, 0x3322);
        const uint32_t m2 = __byte_perm(ey, 0, 0x1100), m3 = __byte_perm(ey, 0, 0x3322);
        const uint32_t d0 = d.x & m0, d1 = d.y & m1, d2 = d.z & m2, d3 = d.w & m3;
        acc[0] += bf16_lo(d0); acc[1] += bf16_hi(d0); acc[2] += bf16_lo(d1); acc[3] += bf16_hi(d1);
        acc[4] += bf16_lo(d2); acc[5] += bf16_hi(d2); acc[6] += bf16_lo(d3); acc[7] += bf16_hi(d3);
      }
    }
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
  if (relu_src) {   // the producer's ReLU mask, from its fp16 output
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(relu_src + eoff));
    o.x &= relu_mask2(r.x); o.y &= relu_mask2(r.y); o.z &= relu_mask2(r.z); o.w &= relu_mask2(r.w);
  }
  *reinterpret_cast<uint4*>(dx + eoff) = o;
}

int launch_maxpool_bwd(const __nv_bfloat16* dy, const uint8_t* idx, const __nv_bfloat16* addend,
                       const __half* relu_src, __nv_bfloat16* dx, const PoolGeom& g,
                       cudaStream_t s, const __half* pooled) {
  // FAV_POOL_BWD_POOLED=0: always mask with the full-resolution producer output (A/B and the bit-identity test; read at
  // every call — 13 per step, once per graph capture)
  const char* ev = getenv("FAV_POOL_BWD_POOLED");
  if ((ev && atoi(ev) == 0) || addend || !relu_src) pooled = nullptr;
  const bool s2 = !pool3s1_applicable(g) && pool_s2_applicable(g);
  const double in_bytes = (addend ? 6.0 : (pooled && s2 ? 2.0 : 4.0));
  ProfScope ps(PK_POOL_BWD, s, 0.0, static_cast<double>(g.B) * g.C * (in_bytes * g.T * g.H * g.W + (pooled && s2 ? 5.0 : 3.0) * g.To * g.Ho * g.Wo));
  FAV_CHECK_ARG(g.C % 8 == 0, "maxpool_bwd: C=%d must be a multiple of 8", g.C);
  if (pool3s1_applicable(g)) return launch_pool3s1_bwd(dy, idx, addend, relu_src, dx, g, s);
  if (pool_s2_applicable(g)) return launch_pool_s2_bwd(dy, idx, addend, relu_src, dx, g, s, pooled);
  dim3 grid(g.B * g.T * g.H, ceil_div(g.W * (g.C / 8), 256));
  const int key = ((g.kt * 10 + g.kh) * 10 + g.kw) * 1000 + (g.st * 10 + g.sh) * 10 + g.sw;
  switch (key) {
    case 133122: FAV_CUDA(launch_pdl(maxpool_bwd_kernel<1, 3, 3, 1, 2, 2>, grid, 256, 0, s, dy, idx, addend, relu_src, dx, g)); break;
    case 333111: FAV_CUDA(launch_pdl(maxpool_bwd_kernel<3, 3, 3, 1, 1, 1>, grid, 256, 0, s, dy, idx, addend, relu_src, dx, g)); break;
    case 333222: FAV_CUDA(launch_pdl(maxpool_bwd_kernel<3, 3, 3, 2, 2, 2>, grid, 256, 0, s, dy, idx, addend, relu_src, dx, g)); break;
    case 222222: FAV_CUDA(launch_pdl(maxpool_bwd_kernel<2, 2, 2, 2, 2, 2>, grid, 256, 0, s, dy, idx, addend, relu_src, dx, g)); break;
    default: FAV_CUDA(launch_pdl(maxpool_bwd_kernel<0, 0, 0, 0, 0, 0>, grid, 256, 0, s, dy, idx, addend, relu_src, dx, g)); break;
  }
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

// =============================================================================================
// head: avg_pool3d [2,7,7] VALID s1 -> 1x1x1 conv + bias -> mean over T' (i3d.py:459-472).
// All linear, so logits = bl + Wl^T * feat with feat[b,c] = sum_t coef[t] sum_hw Y / (HW*2*(T5-1)),
// coef[t] = number of length-2 windows containing frame t.
// =============================================================================================
// T5 < 0 encodes the uniform mean over all -T5 frames (AdaptiveAvgPool3d(1) of the torchvision video ResNets)
__device__ __forceinline__ float head_coef(int t, int T5) {
  if (T5 <= 1) return 1.0f;
  return (t == 0 || t == T5 - 1) ? 1.0f : 2.0f;
}
__device__ __forceinline__ float head_scale(int T5, int HW) {
  if (T5 < 0) return 1.0f / (static_cast<float>(HW) * static_cast<float>(-T5));
  return T5 == 1 ? 1.0f / HW : 1.0f / (static_cast<float>(HW) * 2.0f * (T5 - 1));
}

__global__ void __launch_bounds__(256)
head_feat_kernel(const __half* __restrict__ y, float* __restrict__ feat, int T5, int HW, int C) {
  __shared__ float red[8][32][8];
  const int b = blockIdx.y;
  const int cgp = blockIdx.x * 32 + (threadIdx.x & 31);  // 8-channel group
  const int pl = threadIdx.x >> 5;                        // position lane 0..7
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
  const int npos = (T5 < 0 ? -T5 : T5) * HW;
  if (cgp * 8 < C) {
    for (int p = pl; p < npos; p += 8) {
      const float cf = head_coef(p / HW, T5);
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(y + (static_cast<long long>(b) * npos + p) * C + cgp * 8));
      acc[0] += cf * f16_lo(v.x); acc[1] += cf * f16_hi(v.x);
      acc[2] += cf * f16_lo(v.y); acc[3] += cf * f16_hi(v.y);
      acc[4] += cf * f16_lo(v.z); acc[5] += cf * f16_hi(v.z);
      acc[6] += cf * f16_lo(v.w); acc[7] += cf * f16_hi(v.w);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[pl][threadIdx.x & 31][j] = acc[j];
  __syncthreads();
  if (pl == 0 && cgp * 8 < C) {
    const float sc = head_scale(T5, HW);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float sum = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) sum += red[k][threadIdx.x & 31][j];
      feat[static_cast<long long>(b) * C + cgp * 8 + j] = sum * sc;
    }
  }
}

// logits[b,k] = bl[k] + sum_c feat[b,c] * wl[c,k]; one block = 64 classes x 8 channel slices of one clip
__global__ void __launch_bounds__(512)
head_logits_kernel(const float* __restrict__ feat, const float* __restrict__ wl, const float* __restrict__ bl,
                   float* __restrict__ logits, int C, int K) {
  __shared__ float red[8][64];
  const int b = blockIdx.y;
  const int kl = threadIdx.x & 63, sl = threadIdx.x >> 6;
  const int k = blockIdx.x * 64 + kl;
  float acc = 0.0f;
  if (k < K) {
    const int c0 = sl * ((C + 7) / 8), c1 = min(C, c0 + (C + 7) / 8);
    const float* f = feat + static_cast<long long>(b) * C;
#pragma unroll 8
    for (int c = c0; c < c1; ++c) acc = fmaf(__ldg(f + c), __ldg(wl + static_cast<long long>(c) * K + k), acc);
  }
  red[sl][kl] = acc;
  __syncthreads();
  if (sl == 0 && k < K) {
    float sum = bl[k];
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += red[i][kl];
    logits[static_cast<long long>(b) * K + k] = sum;
  }
}

int launch_head_fwd(const __half* y, int B, int T5, int HW, int C, float* feat, const float* wl,
                    const float* bl, int K, float* logits, cudaStream_t s) {
  ProfScope ps(PK_HEAD_LOSS, s);
  FAV_CHECK_ARG(C % 8 == 0, "head: C must be a multiple of 8");
  dim3 grid(ceil_div(C / 8, 32), B);
  head_feat_kernel<<<grid, 256, 0, s>>>(y, feat, T5, HW, C);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  head_logits_kernel<<<dim3(ceil_div(K, 64), B), 512, 0, s>>>(feat, wl, bl, logits, C, K);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

__global__ void head_dfeat_kernel(const float* __restrict__ dlogits, const float* __restrict__ wl,
                                  float* __restrict__ dfeat, int C, int K) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  float acc = 0.0f;
  for (int k = lane; k < K; k += 32)
    acc = fmaf(dlogits[static_cast<long long>(b) * K + k], wl[static_cast<long long>(c) * K + k], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) dfeat[static_cast<long long>(b) * C + c] = acc;
}

__global__ void __launch_bounds__(256)
head_gy_kernel(const float* __restrict__ dfeat, const __half* __restrict__ y,
               __nv_bfloat16* __restrict__ gy, int T5, int HW, int C, long long total) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cg = C >> 3;
  const int c8 = static_cast<int>(gid % cg);
  const long long p = gid / cg;
  const int npos = (T5 < 0 ? -T5 : T5) * HW;
  const int b = static_cast<int>(p / npos);
  const int t = static_cast<int>((p % npos) / HW);
  const float sc = head_scale(T5, HW) * head_coef(t, T5);
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(y + gid * 8));
  const float* df = dfeat + static_cast<long long>(b) * C + c8 * 8;
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = sc * df[j];
  uint4 ov;   // masked by the final block's ReLU (its fp16 output)
  ov.x = pack_bf16x2(o[0], o[1]) & relu_mask2(v.x); ov.y = pack_bf16x2(o[2], o[3]) & relu_mask2(v.y);
  ov.z = pack_bf16x2(o[4], o[5]) & relu_mask2(v.z); ov.w = pack_bf16x2(o[6], o[7]) & relu_mask2(v.w);
  *reinterpret_cast<uint4*>(gy + gid * 8) = ov;
}

int launch_head_bwd(const float* dlogits, const float* wl, int K, const __half* y,
                    __nv_bfloat16* gy, float* dfeat, int B, int T5, int HW, int C, cudaStream_t s) {
  ProfScope ps(PK_HEAD_LOSS, s);
  dim3 grid(ceil_div(C, 8), B);
  head_dfeat_kernel<<<grid, 256, 0, s>>>(dlogits, wl, dfeat, C, K);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  const long long total = static_cast<long long>(B) * (T5 < 0 ? -T5 : T5) * HW * (C / 8);
  head_gy_kernel<<<static_cast<int>(ceil_div64(total, 256)), 256, 0, s>>>(dfeat, y, gy, T5, HW, C, total);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

// =============================================================================================
// softmax + adversarial loss + dloss/dlogits
//   improve_adversarial_loss: utils/kinetics_i3d_utils.py:253-288 (TF) / model.py:216-250 (torch)
//   ce_adversarial_loss:      utils/kinetics_i3d_utils.py:290-307 (TF) / model.py:177-196 (torch)
// One block, one warp per clip; the batch sums are added in clip order, so they are deterministic.
// =============================================================================================
template <typename T>
__device__ __forceinline__ T block_reduce(T v, T* sh, bool is_max) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? (u > v ? u : v) : v + u;
  }
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  T r = sh[0];
  for (int i = 1; i < nw; ++i) r = is_max ? (sh[i] > r ? sh[i] : r) : r + sh[i];
  return r;
}

// arg-max with lowest-index tie-break over values val(k), k in [0,K)
struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax block_argmax(float v, int i, ArgMax* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float uv = __shfl_xor_sync(0xffffffffu, v, o);
    const int ui = __shfl_xor_sync(0xffffffffu, i, o);
    if (uv > v || (uv == v && ui < i)) { v = uv; i = ui; }
  }
  __syncthreads();
  if (lane == 0) { sh[wid].v = v; sh[wid].i = i; }
  __syncthreads();
  ArgMax r = sh[0];
  const int nw = blockDim.x >> 5;
  for (int k = 1; k < nw; ++k)
    if (sh[k].v > r.v || (sh[k].v == r.v && sh[k].i < r.i)) r = sh[k];
  return r;
}

// warp-level versions (one warp owns one clip)
__device__ __forceinline__ float warp_reduce(float v, bool is_max) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float u = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(u, v) : v + u;
  }
  return v;
}
__device__ __forceinline__ ArgMax warp_argmax(float v, int i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float uv = __shfl_xor_sync(0xffffffffu, v, o);
    const int ui = __shfl_xor_sync(0xffffffffu, i, o);
    if (uv > v || (uv == v && ui < i)) { v = uv; i = ui; }
  }
  ArgMax r;
  r.v = v; r.i = i;
  return r;
}

constexpr int kLossWarps = 8;
// One warp per clip (8 clips in flight, no block barriers inside a clip); the batch sums are added in clip order by one
// thread at the end, so they stay deterministic.
__global__ void __launch_bounds__(kLossWarps * 32)
loss_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const fav_loss_params p,
            int B, int K, float* __restrict__ probs, float* __restrict__ dlogits,
            float* __restrict__ scalars) {
  extern __shared__ float sp[];  // per warp: [K] probabilities, [K] dL/dp; then per clip: 4 partial sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* prob = sp + warp * 2 * K;
  float* dp = prob + K;
  float* part = sp + kLossWarps * 2 * K;   // [B][4]
  const float bdiv = static_cast<float>(p.global_batch > 0 ? p.global_batch : B);
  const bool torch_stack = p.stack == FAV_STACK_TORCH;
  for (int b = warp; b < B; b += kLossWarps) {
    const float* z = logits + static_cast<long long>(b) * K;
    const int y = static_cast<int>(labels[b]);
    float adv_sum = 0.0f, fooled = 0.0f, sum_min = 0.0f, sum_max = 0.0f;
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, z[k]);
    mx = warp_reduce(mx, true);
    float se = 0.0f;
    for (int k = lane; k < K; k += 32) {
      const float e = expf(z[k] - mx);
      prob[k] = e;
      se += e;
    }
    se = warp_reduce(se, false);
    const float inv = 1.0f / se;
    for (int k = lane; k < K; k += 32) {
      prob[k] *= inv;
      dp[k] = 0.0f;
      if (probs) probs[static_cast<long long>(b) * K + k] = prob[k];
    }
    __syncwarp();
    // selections
    float bv = -INFINITY; int bi = K;       // arg-max prob (prediction)
    float nv = -INFINITY; int ni = K;       // max non-label prob
    float lv = -INFINITY; int li = K;       // max "non-label" logit
    for (int k = lane; k < K; k += 32) {
      const float pk = prob[k];
      if (pk > bv) { bv = pk; bi = k; }
      const float pn = torch_stack ? (k == y ? -INFINITY : pk) : pk - (k == y ? 1.0f : 0.0f);
      if (pn > nv) { nv = pn; ni = k; }
      const float zn = torch_stack ? (k == y ? -INFINITY : z[k]) : z[k] - (k == y ? 1.0f : 0.0f);
      if (zn > lv) { lv = zn; li = k; }
    }
    const ArgMax pred = warp_argmax(bv, bi);
    const ArgMax nonl = warp_argmax(nv, ni);
    const ArgMax nonz = warp_argmax(lv, li);
    const float py = prob[y];
    const float pmaxnl = nonl.v;  // TF: value of (p - onehot) at its arg-max == p[k*] when k* != y
    float dd_logit = 0.0f;   // direct logit gradient: +dd at ia, -dd at ib (logits mode only)
    int ia = -1, ib = -1;
    if (lane == 0) {
      float loss = 0.0f;
      if (p.improve_loss) {
        float a, bb, mm;
        float dmm_dp = 0.0f; int mm_idx = -1;  // in logits mode the margin depends on one probability
        if (!p.targeted) {
          if (p.use_logits) {
            a = z[y]; bb = nonz.v; ia = y; ib = nonz.i;
            const float pm = torch_stack ? py : pmaxnl;
            mm = logf(1.0f + p.margin * (1.0f / (0.00001f + pm)));
            dmm_dp = -p.margin / ((0.00001f + pm) * (0.00001f + pm) * (1.0f + p.margin / (0.00001f + pm)));
            mm_idx = torch_stack ? y : nonl.i;
          } else {
            a = py; bb = pmaxnl; mm = p.margin; ia = y; ib = nonl.i;
          }
          sum_min += py; sum_max += pmaxnl;
        } else {
          if (p.use_logits) {
            a = nonz.v; bb = z[y]; ia = nonz.i; ib = y;
            mm = logf(1.0f + p.margin * (1.0f / py));
            dmm_dp = -p.margin / (py * py * (1.0f + p.margin / py));
            mm_idx = y;
          } else {
            a = pmaxnl; bb = py; mm = p.margin; ia = nonl.i; ib = y;
          }
          sum_min += pmaxnl; sum_max += py;
        }
        const float d = a - (bb - mm);
        const float l2 = d * d / mm, l3 = d;
        const float mn = fminf(l2, l3);
        float dd = 0.0f, dmm = 0.0f;
        if (mn > 0.0f) {   // tf.maximum(0.0, x) passes the gradient to x only when x > 0
          loss = mn;
          if (l2 <= l3) { dd = 2.0f * d / mm; dmm = dd - d * d / (mm * mm); }
          else { dd = 1.0f; dmm = 1.0f; }
        }
        if (p.use_logits) {
          dd_logit = dd;
          if (mm_idx >= 0) dp[mm_idx] += dmm * dmm_dp;
        } else {
          dp[ia] += dd;
          dp[ib] -= dd;
        }
        adv_sum += loss;
      } else {
        if (!p.targeted) {
          loss = -logf(1.0f - py + 1e-6f);
          dp[y] += 1.0f / (1.0f - py + 1e-6f) / bdiv;
          sum_min += py; sum_max += pmaxnl;
        } else if (torch_stack) {
          loss = -logf(py + 1e-6f);
          dp[y] += -1.0f / (py + 1e-6f) / bdiv;
          sum_min += pmaxnl; sum_max += py;
        } else {
          loss = -logf(fmaxf(py, 1e-38f));  // sparse softmax cross entropy
          dp[y] += -1.0f / fmaxf(py, 1e-38f) / bdiv;
          sum_min += pmaxnl; sum_max += py;
        }
        adv_sum += loss / bdiv;
      }
      const bool is_fooled = p.targeted ? (pred.i == y) : (pred.i != y);
      fooled += is_fooled ? 1.0f : 0.0f;
      part[b * 4 + 0] = adv_sum;
      part[b * 4 + 1] = fooled;
      part[b * 4 + 2] = sum_min;
      part[b * 4 + 3] = sum_max;
    }
    __syncwarp();
    const float ddl = __shfl_sync(0xffffffffu, dd_logit, 0);
    ia = __shfl_sync(0xffffffffu, ia, 0);
    ib = __shfl_sync(0xffffffffu, ib, 0);
    // softmax backward: dz_k = p_k * (dp_k - sum_j dp_j p_j)  (+ direct logit terms)
    float dot = 0.0f;
    for (int k = lane; k < K; k += 32) dot += dp[k] * prob[k];
    dot = warp_reduce(dot, false);
    for (int k = lane; k < K; k += 32) {
      float g = prob[k] * (dp[k] - dot);
      if (p.use_logits && p.improve_loss) g += (k == ia ? ddl : 0.0f) - (k == ib ? ddl : 0.0f);
      dlogits[static_cast<long long>(b) * K + k] = g * p.grad_scale;
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float adv_sum = 0.0f, fooled = 0.0f, sum_min = 0.0f, sum_max = 0.0f;
    for (int b = 0; b < B; ++b) {   // clip order: deterministic batch sums
      adv_sum += part[b * 4 + 0]; fooled += part[b * 4 + 1]; sum_min += part[b * 4 + 2]; sum_max += part[b * 4 + 3];
    }
    scalars[FAV_S_ADV_LOSS] = adv_sum;
    scalars[FAV_S_FOOLED] = fooled;
    scalars[FAV_S_SUM_P_MIN] = sum_min;
    scalars[FAV_S_SUM_P_MAX] = sum_max;
  }
}

int launch_loss(const float* logits, const int64_t* labels, const fav_loss_params& p, int B, int K,
                float* probs, float* dlogits, float* scalars, cudaStream_t s) {
  ProfScope ps(PK_HEAD_LOSS, s);
  const size_t smem = (static_cast<size_t>(kLossWarps) * 2 * K + static_cast<size_t>(B) * 4) * sizeof(float);
  FAV_CHECK_ARG(smem <= 200 * 1024, "loss: K=%d / B=%d need %zu bytes of shared memory", K, B, smem);
  if (smem > 48 * 1024) FAV_CUDA(cudaFuncSetAttribute(loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  loss_kernel<<<1, kLossWarps * 32, smem, s>>>(logits, labels, p, B, K, probs, dlogits, scalars);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

// =============================================================================================
// Fused evaluation pass: the fooling-ratio counts of one validation batch from the logits of
// its clean rows [0, Bu) and perturbed rows [Bu, 2 Bu).
//   kinetics_i3d.evaluate (utils/kinetics_i3d_utils.py:226-243): miss_cond = argmax(prob_adv) != label (targeted: ==
//   target class); exclude_misclassify: valid = argmax(prob_clean) == label, miss += (miss_cond & valid).sum(), total +=
//   valid.sum(); otherwise miss += miss_cond.sum(), total += batch.  Adversarial_metrics.accuracy_for_eval
//   (model.py:293-323, untargeted) counts the same two numbers.  argmax keeps the first maximum, like numpy / topk.
// One warp per clip; counts[0] += miss, counts[1] += total (int64, atomics: the counters run over all batches).
// =============================================================================================
__device__ __forceinline__ int warp_argmax_first(const float* __restrict__ row, int K, int lane, float* vmax) {
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int k = lane; k < K; k += 32) {
    const float v = row[k];
    if (v > best || bi == 0x7fffffff) { best = v; bi = k; }   // strict >: the earlier index keeps ties
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  *vmax = best;
  return bi;
}

__global__ void __launch_bounds__(256)
eval_count_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int Bu, int n, int K,
                  int targeted, long long target, int exclude, unsigned long long* __restrict__ counts,
                  float* __restrict__ probs) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= Bu) return;
  int am[2];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const float* row = logits + static_cast<size_t>(half * Bu + b) * K;
    float mx;
    am[half] = warp_argmax_first(row, K, lane, &mx);
    if (probs) {   // softmax of both halves (the reference fetches `softmax` / `scores_no_adv` next to the counts)
      float sum = 0.0f;
      for (int k = lane; k < K; k += 32) sum += __expf(row[k] - mx);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.0f / sum;
      float* prow = probs + static_cast<size_t>(half * Bu + b) * K;
      for (int k = lane; k < K; k += 32) prow[k] = __expf(row[k] - mx) * inv;
    }
  }
  if (lane == 0 && b < n) {
    const long long lab = labels[b];
    const bool valid = !exclude || am[0] == lab;
    const bool miss = targeted ? am[1] == target : am[1] != lab;
    if (valid) atomicAdd(counts + 1, 1ull);
    if (valid && miss) atomicAdd(counts + 0, 1ull);
  }
}

int launch_eval_counts(const float* logits, const int64_t* labels, int Bu, int n, int K, int targeted, long long target,
                       int exclude, int64_t* counts, float* probs, cudaStream_t s) {
  ProfScope ps(PK_HEAD_LOSS, s);
  FAV_CHECK_ARG(n >= 0 && n <= Bu, "eval counts: n_clips %d outside [0, %d]", n, Bu);
  eval_count_kernel<<<ceil_div(Bu, 8), 256, 0, s>>>(logits, labels, Bu, n, K, targeted, target, exclude,
                                                    reinterpret_cast<unsigned long long*>(counts), probs);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

// =============================================================================================
// (c) regulariser gradients + clip mask + Adam + metrics on delta [T,3]
//   norm_reg / diff_norm_reg / laplacian_norm_reg, thickness, roughness:
//     utils/kinetics_i3d_utils.py:177-200 (TF, on the raw delta);  model.py:198-209 (torch, on the
//     clamped delta, weights beta1 / (1-beta1) passed in as beta1/beta2/beta3)
//   loss = adv + beta0*(beta1*thick + beta2*diff + beta3*lap): single_video_npy.py:56-59
//   Adam: tf.train.AdamOptimizer (single_video_npy.py:79-84) or torch.optim.Adam (model.py:542)
// =============================================================================================
__global__ void __launch_bounds__(256)
delta_update_kernel(float* __restrict__ delta, const float* __restrict__ grad, float* __restrict__ m,
                    float* __restrict__ v, int64_t* __restrict__ step, const fav_reg_params reg,
                    const fav_adam_params adam, float adv_flag, float* __restrict__ scalars, int T) {
  extern __shared__ float sd[];  // [T*3] regularised copy of delta
  __shared__ float sh[8];
  const int N = T * 3;
  const bool torch_stack = adam.stack == FAV_STACK_TORCH;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float d = delta[i];
    sd[i] = torch_stack ? fminf(fmaxf(d, -reg.delta_clip), reg.delta_clip) : d;
  }
  __syncthreads();
  float s_norm = 0.0f, s_diff = 0.0f, s_lap = 0.0f, s_abs = 0.0f, s_rough = 0.0f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int t = i / 3, c = i % 3;
    const float d0 = sd[i];
    const float dm = sd[((t + T - 1) % T) * 3 + c];
    const float dp = sd[((t + 1) % T) * 3 + c];
    const float df = d0 - dm;
    const float lp = -2.0f * d0 + dm + dp;
    s_norm += d0 * d0;
    s_diff += df * df;
    s_lap += lp * lp;
    // metrics: TF on the raw delta, torch on the clamped one (model.py:1114-1116)
    s_abs += fabsf(d0);
    s_rough += fabsf(df);
  }
  s_norm = block_reduce<float>(s_norm, sh, false);
  s_diff = block_reduce<float>(s_diff, sh, false);
  s_lap = block_reduce<float>(s_lap, sh, false);
  s_abs = block_reduce<float>(s_abs, sh, false);
  s_rough = block_reduce<float>(s_rough, sh, false);
  const float invN = 1.0f / static_cast<float>(N);
  const int64_t tstep = *step + 1;
  const float b1t = powf(adam.b1, static_cast<float>(tstep));
  const float b2t = powf(adam.b2, static_cast<float>(tstep));
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int t = i / 3, c = i % 3;
    const int tm = (t + T - 1) % T, tp = (t + 1) % T;
    const int tmm = (t + T - 2) % T, tpp = (t + 2) % T;
    const float d0 = sd[i], dm = sd[tm * 3 + c], dp = sd[tp * 3 + c];
    const float dmm = sd[tmm * 3 + c], dpp = sd[tpp * 3 + c];
    // d/d delta_t of sum_s (d_s - d_{s-1})^2 = 2 (D_t - D_{t+1}),  D_t = d_t - d_{t-1}
    const float g_diff = 2.0f * ((d0 - dm) - (dp - d0));
    // d/d delta_t of sum_s L_s^2 = 2 (-2 L_t + L_{t-1} + L_{t+1}),  L_t = -2 d_t + d_{t-1} + d_{t+1}
    const float l0 = -2.0f * d0 + dm + dp;
    const float lm = -2.0f * dm + dmm + d0;
    const float lp = -2.0f * dp + d0 + dpp;
    const float g_lap = 2.0f * (-2.0f * l0 + lm + lp);
    const float g_norm = 2.0f * d0;
    float g_reg = reg.beta0 * invN * (reg.beta1 * g_norm + reg.beta2 * g_diff + reg.beta3 * g_lap);
    const float raw = delta[i];
    const bool inside = fabsf(raw) <= reg.delta_clip;  // clip_by_value / clamp pass the gradient inclusively
    if (torch_stack && !inside) g_reg = 0.0f;
    const float g_data = inside ? adv_flag * grad[i] : 0.0f;
    const float g = g_data + g_reg;
    const float mi = adam.b1 * m[i] + (1.0f - adam.b1) * g;
    const float vi = adam.b2 * v[i] + (1.0f - adam.b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    float upd;
    if (torch_stack) {
      upd = (adam.lr / (1.0f - b1t)) * mi / (sqrtf(vi) / sqrtf(1.0f - b2t) + adam.eps);
    } else {
      const float lr_t = adam.lr * sqrtf(1.0f - b2t) / (1.0f - b1t);
      upd = lr_t * mi / (sqrtf(vi) + adam.eps);
    }
    delta[i] = raw - upd;
  }
  if (threadIdx.x == 0) {
    *step = tstep;
    const float nr = s_norm * invN + 1e-12f, dr = s_diff * invN + 1e-12f, lr_ = s_lap * invN + 1e-12f;
    scalars[FAV_S_NORM_REG] = nr;
    scalars[FAV_S_DIFF_REG] = dr;
    scalars[FAV_S_LAP_REG] = lr_;
    scalars[FAV_S_THICKNESS] = s_abs * invN;
    scalars[FAV_S_ROUGHNESS] = s_rough * invN;
    scalars[FAV_S_TOTAL_LOSS] =
        scalars[FAV_S_ADV_LOSS] + reg.beta0 * (reg.beta1 * nr + reg.beta2 * dr + reg.beta3 * lr_);
  }
}

int launch_delta_update(float* delta, const float* grad, float* m, float* v, int64_t* step,
                        const fav_reg_params& reg, const fav_adam_params& adam, float adv_flag,
                        float* scalars, int T, cudaStream_t s) {
  ProfScope ps(PK_DELTA_UPDATE, s);
  delta_update_kernel<<<1, 256, T * 3 * sizeof(float), s>>>(delta, grad, m, v, step, reg, adam, adv_flag,
                                                            scalars, T);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

// =============================================================================================
// sparse per-pixel attack (FLICKERING_ATTACK = False): kinetics_i3d_L12 (utils/kinetics_i3d_utils.py:308-521,
// delta [T,H,W,3], no +-0.4 clip) and the torch stack with attack_type "L12" (pert_size [3,T,112,112],
// model.py:383-384).  delta is per pixel, so it is added in fp32 and the sum is what the stem reads.
// =============================================================================================
__global__ void __launch_bounds__(256)
apply_pixels_kernel(const uint8_t* __restrict__ clip, const float* __restrict__ dpx, float adv_flag, float dclip,
                    const fav_norm_params nrm, int torch_mode, __half* __restrict__ xpad, int Wp, int padl,
                    float* __restrict__ adv_f32, int T, int H, int W, long long groups) {
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= groups) return;
  const int gpr = W >> 4;
  const int wg = static_cast<int>(gid % gpr);
  const long long row = gid / gpr;  // (b*T + t)*H + h
  const int h = static_cast<int>(row % H);
  const int t = static_cast<int>((row / H) % T);
  const long long b = row / (static_cast<long long>(H) * T);
  const long long e0 = (row * W + wg * 16) * 3;
  const float* dp = dpx + ((static_cast<long long>(t) * H + h) * W + wg * 16) * 3;
  uint2* dst = reinterpret_cast<uint2*>(xpad + (row * Wp + padl + wg * 16) * 4);
#pragma unroll 4
  for (int p = 0; p < 16; ++p) {
    float q[3], a[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float u = static_cast<float>(__ldg(clip + e0 + 3 * p + c));
      float dv = __ldg(dp + 3 * p + c);
      if (dclip > 0.0f) dv = fminf(fmaxf(dv, -dclip), dclip);
      dv = __fmul_rn(adv_flag, dv);
      float x;
      if (torch_mode) {
        x = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), nrm.mean[c]), nrm.std[c]);
        dv = __fdiv_rn(dv, nrm.std[c]);
      } else {
        x = __fsub_rn(__fmul_rn(u, 0.0078125f), 1.0f);
      }
      const float av = fminf(fmaxf(__fadd_rn(x, dv), nrm.lo), nrm.hi);
      a[c] = av;
      q[c] = torch_mode ? (av * nrm.std[c] + nrm.mean[c]) * 255.0f - 128.0f : av;
    }
    dst[p] = make_uint2(pack_f16x2(q[0], q[1]), pack_f16x2(q[2], 0.0f));
    if (adv_f32) {
      if (torch_mode) {   // NCTHW
#pragma unroll
        for (int c = 0; c < 3; ++c)
          adv_f32[(((b * 3 + c) * T + t) * H + h) * static_cast<long long>(W) + wg * 16 + p] = a[c];
      } else {            // NTHWC
#pragma unroll
        for (int c = 0; c < 3; ++c) adv_f32[e0 + 3 * p + c] = a[c];
      }
    }
  }
}

int launch_apply_pixels(const uint8_t* clip, const float* delta_px, float adv_flag, float delta_clip,
                        const fav_norm_params& nrm, int torch_mode, __half* xpad, int Wp, int padl,
                        float* adv_f32, int B, int T, int H, int W, cudaStream_t s) {
  ProfScope ps(PK_APPLY, s, 0.0, static_cast<double>(B) * T * H * W * (3.0 + 8.0 + (adv_f32 ? 12.0 : 0.0)) + static_cast<double>(T) * H * W * 12.0);
  FAV_CHECK_ARG(W % 16 == 0, "apply: W=%d must be a multiple of 16", W);
  const long long groups = static_cast<long long>(B) * T * H * (W / 16);
  apply_pixels_kernel<<<static_cast<int>(ceil_div64(groups, 256)), 256, 0, s>>>(clip, delta_px, adv_flag, delta_clip, nrm,
                                                                                 torch_mode, xpad, Wp, padl, adv_f32, T, H,
                                                                                 W, groups);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

// grad[t,h,w,c] = scale_c * sum_b mask(b,t,h,w,c) * dX[b,t,h,w,c]   (one thread per pixel)
__global__ void __launch_bounds__(256)
stem_dx_pixels_kernel(const __nv_bfloat16* __restrict__ dx, const uint8_t* __restrict__ clip,
                      const float* __restrict__ dpx, float adv_flag, float dclip, const fav_norm_params nrm,
                      int torch_mode, float* __restrict__ grad, int B, long long npix) {
  const long long px = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (px >= npix) return;
  float d[3], acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float dv = __ldg(dpx + px * 3 + c);
    if (dclip > 0.0f) dv = fminf(fmaxf(dv, -dclip), dclip);
    dv = __fmul_rn(adv_flag, dv);
    d[c] = torch_mode ? __fdiv_rn(dv, nrm.std[c]) : dv;
  }
  for (int b = 0; b < B; ++b) {
    const long long q = b * npix + px;
    const uint2 g = __ldg(reinterpret_cast<const uint2*>(dx + q * 16));
    const float gv[3] = {bf16_lo(g.x), bf16_hi(g.x), bf16_lo(g.y)};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float u = static_cast<float>(__ldg(clip + q * 3 + c));
      const float x = torch_mode ? __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), nrm.mean[c]), nrm.std[c])
                                 : __fsub_rn(__fmul_rn(u, 0.0078125f), 1.0f);
      const float sv = __fadd_rn(x, d[c]);
      if (sv >= nrm.lo && sv <= nrm.hi) acc[c] += gv[c];
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) grad[px * 3 + c] = acc[c] * (torch_mode ? 1.0f / nrm.std[c] : 1.0f);
}

int launch_stem_dx_pixels(const __nv_bfloat16* dx, const uint8_t* clip, const float* delta_px, float adv_flag,
                          float delta_clip, const fav_norm_params& nrm, int torch_mode, float* grad, int B, int T, int H,
                          int W, cudaStream_t s) {
  ProfScope ps(PK_STEM_BWD, s, 0.0, static_cast<double>(B) * T * H * W * 9.0 + static_cast<double>(T) * H * W * 24.0);
  const long long npix = static_cast<long long>(T) * H * W;
  stem_dx_pixels_kernel<<<static_cast<int>(ceil_div64(npix, 256)), 256, 0, s>>>(dx, clip, delta_px, adv_flag, delta_clip,
                                                                                 nrm, torch_mode, grad, B, npix);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

// L1,2 regulariser + Adam on the per-pixel delta.
//   L12 = sum_t sqrt(mean_{h,w,c} delta_t^2) + 1e-12  (kinetics_i3d_utils.py:409; Losses.L12_regularization_loss
//   model.py:211-214, on the clamped delta);  d L12 / d delta = delta / (n * sqrt(mean_t)).
// pass 1: per-frame partial sums (sum d^2, sum |d|, sum |d - d_prev|); pass 2: fixed-order totals + update.
constexpr int kPixChunk = 4096;
__global__ void __launch_bounds__(256)
pixels_stats_kernel(const float* __restrict__ dpx, float dclip, float* __restrict__ partial, int T, int n_frame,
                    int chunks) {
  __shared__ float red[8][3];
  const int t = blockIdx.y, chunk = blockIdx.x;
  const int tp = (t + T - 1) % T;
  float ssq = 0.f, sab = 0.f, sro = 0.f;
  const int i1 = min(n_frame, (chunk + 1) * kPixChunk);
  for (int i = chunk * kPixChunk + threadIdx.x; i < i1; i += blockDim.x) {
    float d = dpx[static_cast<long long>(t) * n_frame + i];
    float dprev = dpx[static_cast<long long>(tp) * n_frame + i];
    if (dclip > 0.0f) {
      d = fminf(fmaxf(d, -dclip), dclip);
      dprev = fminf(fmaxf(dprev, -dclip), dclip);
    }
    ssq += d * d;
    sab += fabsf(d);
    sro += fabsf(d - dprev);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
    sab += __shfl_xor_sync(0xffffffffu, sab, o);
    sro += __shfl_xor_sync(0xffffffffu, sro, o);
  }
  if (lane == 0) { red[wid][0] = ssq; red[wid][1] = sab; red[wid][2] = sro; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sum += red[w][threadIdx.x];
    partial[(static_cast<long long>(t) * chunks + chunk) * 3 + threadIdx.x] = sum;
  }
}

__global__ void __launch_bounds__(256)
pixels_update_kernel(float* __restrict__ dpx, const float* __restrict__ grad, float* __restrict__ m,
                     float* __restrict__ v, const int64_t* __restrict__ step, const float* __restrict__ partial,
                     float reg_weight, float dclip, const fav_adam_params adam, float* __restrict__ scalars, int T,
                     int n_frame, int chunks) {
  __shared__ float s_mean;
  const int t = blockIdx.y, chunk = blockIdx.x;
  if (threadIdx.x == 0) {
    float ssq = 0.f;
    for (int k = 0; k < chunks; ++k) ssq += partial[(static_cast<long long>(t) * chunks + k) * 3];
    s_mean = ssq / static_cast<float>(n_frame);
  }
  __syncthreads();
  const float rms = sqrtf(s_mean);
  const float gcoef = rms > 0.0f ? reg_weight / (static_cast<float>(n_frame) * rms) : 0.0f;
  const int64_t tstep = *step + 1;
  const float b1t = powf(adam.b1, static_cast<float>(tstep)), b2t = powf(adam.b2, static_cast<float>(tstep));
  const bool torch_stack = adam.stack == FAV_STACK_TORCH;
  const int i1 = min(n_frame, (chunk + 1) * kPixChunk);
  for (int i = chunk * kPixChunk + threadIdx.x; i < i1; i += blockDim.x) {
    const long long e = static_cast<long long>(t) * n_frame + i;
    const float raw = dpx[e];
    const bool inside = dclip <= 0.0f || fabsf(raw) <= dclip;
    const float dc = dclip > 0.0f ? fminf(fmaxf(raw, -dclip), dclip) : raw;
    const float g = inside ? grad[e] + gcoef * dc : 0.0f;
    const float mi = adam.b1 * m[e] + (1.0f - adam.b1) * g;
    const float vi = adam.b2 * v[e] + (1.0f - adam.b2) * g * g;
    m[e] = mi;
    v[e] = vi;
    float upd;
    if (torch_stack) {
      upd = (adam.lr / (1.0f - b1t)) * mi / (sqrtf(vi) / sqrtf(1.0f - b2t) + adam.eps);
    } else {
      upd = adam.lr * sqrtf(1.0f - b2t) / (1.0f - b1t) * mi / (sqrtf(vi) + adam.eps);
    }
    dpx[e] = raw - upd;
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    // metrics of the delta this step was computed with (fixed summation order)
    float l12 = 0.f, sab = 0.f, sro = 0.f;
    for (int tt = 0; tt < T; ++tt) {
      float ssq = 0.f;
      for (int k = 0; k < chunks; ++k) {
        const float* p = partial + (static_cast<long long>(tt) * chunks + k) * 3;
        ssq += p[0]; sab += p[1]; sro += p[2];
      }
      l12 += sqrtf(ssq / static_cast<float>(n_frame));
    }
    const float inv = 1.0f / (static_cast<float>(T) * static_cast<float>(n_frame));
    scalars[FAV_S_NORM_REG] = l12 + 1e-12f;
    scalars[FAV_S_DIFF_REG] = 0.0f;
    scalars[FAV_S_LAP_REG] = 0.0f;
    scalars[FAV_S_THICKNESS] = sab * inv;
    scalars[FAV_S_ROUGHNESS] = sro * inv;
    scalars[FAV_S_TOTAL_LOSS] = scalars[FAV_S_ADV_LOSS] + reg_weight * (l12 + 1e-12f);
  }
}

__global__ void step_increment_kernel(int64_t* step) { *step += 1; }

int launch_pixels_update(float* delta_px, const float* grad_px, float* m, float* v, int64_t* step, float* partial,
                         float reg_weight, float delta_clip, const fav_adam_params& adam, float* scalars, int T, int H,
                         int W, cudaStream_t s) {
  ProfScope ps(PK_DELTA_UPDATE, s, 0.0, static_cast<double>(T) * H * W * 3.0 * 28.0);
  const int n_frame = H * W * 3;
  const int chunks = ceil_div(n_frame, kPixChunk);
  pixels_stats_kernel<<<dim3(chunks, T), 256, 0, s>>>(delta_px, delta_clip, partial, T, n_frame, chunks);
  FAV_COUNT_LAUNCH();
  pixels_update_kernel<<<dim3(chunks, T), 256, 0, s>>>(delta_px, grad_px, m, v, step, partial, reg_weight, delta_clip, adam,
                                                       scalars, T, n_frame, chunks);
  FAV_COUNT_LAUNCH();
  step_increment_kernel<<<1, 1, 0, s>>>(step);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}
int pixels_partial_floats(int T, int H, int W) { return T * ceil_div(H * W * 3, kPixChunk) * 3; }

// ---------------------------------------------------------------------------------------------
// layout helpers for tests / debug reads
// ---------------------------------------------------------------------------------------------
__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ src, long long cs, int coff, int C,
                                   long long total, float* __restrict__ dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long p = i / C;
  const int c = static_cast<int>(i % C);
  dst[i] = __bfloat162float(src[p * cs + coff + c]);
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                   long long cs, int coff, int C, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long p = i / C;
  const int c = static_cast<int>(i % C);
  dst[p * cs + coff + c] = __float2bfloat16_rn(src[i]);
}
__global__ void f16_to_f32_kernel(const __half* __restrict__ src, long long cs, int coff, int C, long long total,
                                  float* __restrict__ dst) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long p = i / C;
  const int c = static_cast<int>(i % C);
  dst[i] = __half2float(src[p * cs + coff + c]);
}
int launch_f16_to_f32(const __half* src, long long cs, int coff, int C, long long npos, float* dst, cudaStream_t s) {
  const long long total = npos * C;
  f16_to_f32_kernel<<<static_cast<int>(ceil_div64(total, 256)), 256, 0, s>>>(src, cs, coff, C, total, dst);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}
int launch_bf16_to_f32(const __nv_bfloat16* src, long long cs, int coff, int C, long long npos, float* dst,
                       cudaStream_t s) {
  const long long total = npos * C;
  bf16_to_f32_kernel<<<static_cast<int>(ceil_div64(total, 256)), 256, 0, s>>>(src, cs, coff, C, total, dst);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}
int launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long cs, int coff, int C, long long npos,
                       cudaStream_t s) {
  const long long total = npos * C;
  f32_to_bf16_kernel<<<static_cast<int>(ceil_div64(total, 256)), 256, 0, s>>>(src, dst, cs, coff, C, total);
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

}  // namespace fav
