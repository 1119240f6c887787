// pool3.cu — MaxPool3d 3x3x3 / stride 1 / SAME (the Branch_3 pool of every Inception block,
// i3d.py:212 and its eight siblings), forward and backward, in separable streaming form.
//
// max over a 3x3x3 window = max_t(max_h(max_w x)), and "first arg-max in (t,h,w) scan order" (where TF
// and torch route the gradient) is exactly what three first-wins 1-D stages select.  The forward
// therefore keeps one code byte per element describing the three stages *centred on this element*:
// bits 0-2 one-hot W-stage winner, bits 3-5 one-hot H-stage winner, bits 6-7 the T-stage winner (0..2);
// the backward pulls through three 3-tap stages: 9 masked accumulates per element instead of 27.
//
// One CTA owns a whole HxW plane of a channel group and streams over T: the W and H stages go through
// shared memory once per frame, the T stage is a register ring, so nothing is re-read from global
// memory except the two frames at the ends of a T segment.
//
// ncu (profiles/r02_c1_ncu_full_details.txt): round 1's kernels were ISSUE bound, not memory bound — the backward ran
// 413 instructions per 8-channel item and frame (emulated __vcmpeq4 code tests, fp32 unpack + predicated adds) at 65 %
// issue-slot utilisation and 29 % of DRAM.  Now a code test is one shift + one PRMT in sign-replication mode per channel
// pair (the one-hot bit moved to the byte's MSB becomes a halfword mask), and the stage sums are packed bf16x2 adds
// (at most three terms per stage; an element that wins a single window — the usual case — is exact).
#include "kernels.cuh"

#include <algorithm>
#include <stdlib.h>

namespace fav {
namespace {

constexpr uint32_t kNegInf2 = kF16NegInf2;   // fp16x2 (-inf, -inf): the forward runs on fp16 activations
constexpr int kPoolThreads = 800;

__device__ __forceinline__ void first_max(uint32_t (&best)[4], uint32_t (&code)[4], const uint4 v, const uint32_t d2) {
  const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 a = *reinterpret_cast<const __half2*>(&vv[j]);
    const __half2 b = *reinterpret_cast<const __half2*>(&best[j]);
    const uint32_t m = __hgt2_mask(a, b);          // strict >: the earlier candidate keeps ties
    const __half2 mx = __hmax2(b, a);
    best[j] = *reinterpret_cast<const uint32_t*>(&mx);
    code[j] = (code[j] & ~m) | (d2 & m);
  }
}
// the first candidate of a stage needs no compare: it is the running maximum and its code until a later one beats it
__device__ __forceinline__ void first_set(uint32_t (&best)[4], uint32_t (&code)[4], const uint4 v, const uint32_t d2) {
  best[0] = v.x; best[1] = v.y; best[2] = v.z; best[3] = v.w;
  code[0] = code[1] = code[2] = code[3] = d2;
}

// per-halfword code constants (two channels per 32-bit word; the low byte of each halfword is the channel's code byte)
constexpr uint32_t kCW0 = 0x00010001u, kCW1 = 0x00020002u, kCW2 = 0x00040004u;   // W stage, one-hot bits 0-2
constexpr uint32_t kCH0 = 0x00080008u, kCH1 = 0x00100010u, kCH2 = 0x00200020u;   // H stage, one-hot bits 3-5
constexpr uint32_t kCT0 = 0x00000000u, kCT1 = 0x00400040u, kCT2 = 0x00800080u;   // T stage, value in bits 6-7

// forward.  grid (channel blocks x row tiles, T segments, B); one thread = one (row, w, 8-channel group) item.
// Small planes are owned whole (halo = 0, R = H); large ones are cut into tiles of R rows that also load one halo
// row above and below (W stage only), so that cgn consecutive lanes cover 16*cgn contiguous bytes per position.
__global__ void __launch_bounds__(1024)
pool3s1_fwd_kernel(const __half* __restrict__ x, __half* __restrict__ y, uint8_t* __restrict__ idx,
                   const int T, const int H, const int W, const int C, const int cgn, const int tseg, const int R,
                   const int nth, const int halo) {
  extern __shared__ uint4 smem4[];
  const int Rt = R + 2 * halo;
  uint4* X = smem4;                                   // [Rt][W+2][cgn]
  uint4* M1 = smem4 + Rt * (W + 2) * cgn;             // [R+2][W][cgn], row index = local row + 1
  const int tile = blockIdx.x % nth, cblk = blockIdx.x / nth;
  const int r0 = tile * R;
  const int it = threadIdx.x;
  const int cgi = it % cgn;
  const int pos = it / cgn;
  const int lr = pos / W, w = pos - lr * W;
  const int lrow = lr - halo;
  const int h = r0 + lrow;
  const bool live = lr < Rt && h >= 0 && h < H;
  const bool core = live && lrow >= 0 && lrow < R;
  const uint4 ninf = make_uint4(kNegInf2, kNegInf2, kNegInf2, kNegInf2);
  // -inf borders / rows outside the image (never overwritten afterwards)
  for (int i = threadIdx.x; i < Rt * 2 * cgn; i += blockDim.x) {
    const int hh = i / (2 * cgn), r = i - hh * 2 * cgn;
    X[(hh * (W + 2) + (r < cgn ? 0 : W + 1)) * cgn + (r % cgn)] = ninf;
  }
  for (int i = threadIdx.x; i < (R + 2) * W * cgn; i += blockDim.x) M1[i] = ninf;
  __syncthreads();
  pdl_sync();
  const int b = blockIdx.z;
  const int t_begin = blockIdx.y * tseg;
  const int t_end = min(t_begin + tseg, T);
  const long long plane = static_cast<long long>(H) * W * C;
  const long long eoff = live ? (static_cast<long long>(h) * W + w) * C + (cblk * cgn + cgi) * 8 : 0;
  const __half* xb = x + static_cast<long long>(b) * T * plane + eoff;
  __half* yb = y + static_cast<long long>(b) * T * plane + eoff;
  uint8_t* ib = idx + static_cast<long long>(b) * T * plane + eoff;
  const int xs = (lr * (W + 2) + w + 1) * cgn + cgi;       // own cell in X
  const int ms = ((lrow + 1) * W + w) * cgn + cgi;         // own cell in M1

  uint32_t p2[4] = {kNegInf2, kNegInf2, kNegInf2, kNegInf2};   // in-plane max of frame to-1
  uint32_t p1[4] = {kNegInf2, kNegInf2, kNegInf2, kNegInf2};   // in-plane max of frame to
  uint2 pcode = make_uint2(0u, 0u);                             // W | H code bytes of frame to
  uint4 nxt = ninf;
  if (live && t_begin - 1 >= 0) nxt = __ldg(reinterpret_cast<const uint4*>(xb + (t_begin - 1) * plane));
  for (int tt = t_begin - 1; tt <= t_end; ++tt) {
    uint32_t m2[4] = {kNegInf2, kNegInf2, kNegInf2, kNegInf2};
    uint2 ccode = make_uint2(0u, 0u);
    const uint4 cur = nxt;
    if (live && tt + 1 <= t_end && tt + 1 < T) nxt = __ldg(reinterpret_cast<const uint4*>(xb + (tt + 1) * plane));
    if (tt >= 0 && tt < T) {   // CTA-uniform
      if (live) X[xs] = cur;
      __syncthreads();
      uint32_t cw[4] = {0u, 0u, 0u, 0u}, ch[4] = {0u, 0u, 0u, 0u};
      if (live) {
        uint32_t m1[4];
        first_set(m1, cw, X[xs - cgn], kCW0);
        first_max(m1, cw, cur, kCW1);
        first_max(m1, cw, X[xs + cgn], kCW2);
        M1[ms] = make_uint4(m1[0], m1[1], m1[2], m1[3]);
      }
      __syncthreads();
      if (core) {
        first_set(m2, ch, M1[ms - W * cgn], kCH0);
        first_max(m2, ch, M1[ms], kCH1);
        first_max(m2, ch, M1[ms + W * cgn], kCH2);
      }
      ccode.x = __byte_perm(cw[0] | ch[0], cw[1] | ch[1], 0x6420);
      ccode.y = __byte_perm(cw[2] | ch[2], cw[3] | ch[3], 0x6420);
    }
    const int to = tt - 1;
    if (core && to >= t_begin && to < t_end) {
      uint32_t best[4], c3[4];
      first_set(best, c3, make_uint4(p2[0], p2[1], p2[2], p2[3]), kCT0);
      first_max(best, c3, make_uint4(p1[0], p1[1], p1[2], p1[3]), kCT1);
      first_max(best, c3, make_uint4(m2[0], m2[1], m2[2], m2[3]), kCT2);
      *reinterpret_cast<uint4*>(yb + to * plane) = make_uint4(best[0], best[1], best[2], best[3]);
      if (idx) {   // nullptr: forward-only (evaluation) plan, no backward will read the stage codes
        uint2 o;
        o.x = pcode.x | __byte_perm(c3[0], c3[1], 0x6420);
        o.y = pcode.y | __byte_perm(c3[2], c3[3], 0x6420);
        *reinterpret_cast<uint2*>(ib + to * plane) = o;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { p2[j] = p1[j]; p1[j] = m2[j]; }
    pcode = ccode;
  }
}

// ---- backward helpers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t sel) {   // generic PRMT: selector bit 3 = replicate the byte's MSB
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
  return d;
}
__device__ __forceinline__ uint32_t bf2_add(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
// acc (8 channels, packed bf16x2) += v where the code bytes' MSB is set: `sx`, `sy` hold the tested bit of every code
// byte in bit 7 (channels 0-3 / 4-7)
__device__ __forceinline__ void acc_msb(uint32_t (&acc)[4], const uint4 v, const uint32_t sx, const uint32_t sy) {
  acc[0] = bf2_add(acc[0], v.x & prmt_b32(sx, 0x9988u));
  acc[1] = bf2_add(acc[1], v.y & prmt_b32(sx, 0xbbaau));
  acc[2] = bf2_add(acc[2], v.z & prmt_b32(sy, 0x9988u));
  acc[3] = bf2_add(acc[3], v.w & prmt_b32(sy, 0xbbaau));
}
template <int BIT>   // one-hot / single-bit test: bit BIT of every code byte
__device__ __forceinline__ void acc_bit(uint32_t (&acc)[4], const uint4 v, const uint2 code) {
  acc_msb(acc, v, code.x << (7 - BIT), code.y << (7 - BIT));
}
// T-stage value 0: neither bit 6 nor bit 7
__device__ __forceinline__ void acc_t0(uint32_t (&acc)[4], const uint4 v, const uint2 code) {
  acc_msb(acc, v, ~(code.x | (code.x << 1)), ~(code.y | (code.y << 1)));
}

// backward.  dx = relu_mask(addend + pool^T(dy)); same grid / item mapping as the forward (halo rows run the T stage only).
__global__ void __launch_bounds__(kPoolThreads)
pool3s1_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                   const __nv_bfloat16* __restrict__ addend, const __half* __restrict__ relu_src,
                   __nv_bfloat16* __restrict__ dx, const int T, const int H, const int W, const int C,
                   const int cgn, const int tseg, const int R, const int nth, const int halo) {
  extern __shared__ uint4 smem4[];
  const int Rt = R + 2 * halo;
  const int n2 = (R + 2) * W * cgn, n1 = R * (W + 2) * cgn, nc = (R + 2) * (W + 2) * cgn;
  uint4* G2 = smem4;                                  // [R+2][W][cgn]   T-stage sums (bf16 x 8), row index = local row + 1
  uint4* G1 = G2 + n2;                                // [R][W+2][cgn]   H-stage sums
  uint2* CD = reinterpret_cast<uint2*>(G1 + n1);      // [2][R+2][W+2][cgn] code bytes
  const int tile = blockIdx.x % nth, cblk = blockIdx.x / nth;
  const int r0 = tile * R;
  const int it = threadIdx.x;
  const int cgi = it % cgn;
  const int pos = it / cgn;
  const int lr = pos / W, w = pos - lr * W;
  const int lrow = lr - halo;
  const int h = r0 + lrow;
  const bool live = lr < Rt && h >= 0 && h < H;
  const bool core = live && lrow >= 0 && lrow < R;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < n2; i += blockDim.x) G2[i] = z;                       // rows outside the image stay 0
  for (int i = threadIdx.x; i < n1; i += blockDim.x) G1[i] = z;                       // incl. the border columns
  for (int i = threadIdx.x; i < 2 * nc; i += blockDim.x) CD[i] = make_uint2(0u, 0u);  // code 0: no one-hot bit set
  __syncthreads();
  pdl_sync();

  const int b = blockIdx.z;
  const int t_begin = blockIdx.y * tseg;
  const int t_end = min(t_begin + tseg, T);
  const long long plane = static_cast<long long>(H) * W * C;
  const long long eoff = live ? (static_cast<long long>(h) * W + w) * C + (cblk * cgn + cgi) * 8 : 0;
  const long long boff = static_cast<long long>(b) * T * plane + eoff;
  const __nv_bfloat16* dyb = dy + boff;
  const uint8_t* ib = idx + boff;
  const int g2s = ((lrow + 1) * W + w) * cgn + cgi;
  const int g1s = (lrow * (W + 2) + w + 1) * cgn + cgi;
  const int cds = ((lrow + 1) * (W + 2) + w + 1) * cgn + cgi;
  const int crow = (W + 2) * cgn;

  auto ld_dy = [&](int t) { return (live && t >= 0 && t < T) ? __ldg(reinterpret_cast<const uint4*>(dyb + t * plane)) : z; };
  auto ld_cd = [&](int t) {
    return (live && t >= 0 && t < T) ? __ldg(reinterpret_cast<const uint2*>(ib + t * plane)) : make_uint2(0u, 0u);
  };
  uint4 dA = ld_dy(t_begin - 1), dB = ld_dy(t_begin), dC = ld_dy(t_begin + 1);
  uint2 cA = ld_cd(t_begin - 1), cB = ld_cd(t_begin), cC = ld_cd(t_begin + 1);
  for (int t = t_begin; t < t_end; ++t) {
    // prefetch: next frame of the ring, and this frame's epilogue operands
    const uint4 dN = ld_dy(t + 2);
    const uint2 cN = ld_cd(t + 2);
    uint4 av = z, rv = z;
    if (core) {
      if (addend) av = __ldg(reinterpret_cast<const uint4*>(addend + boff + t * plane));
      if (relu_src) rv = __ldg(reinterpret_cast<const uint4*>(relu_src + boff + t * plane));
    }
    uint2* cd = CD + (t & 1) * nc;
    // T stage: the window centred on frame t+1-d selected frame t iff its T code == d
    uint32_t g2[4] = {0u, 0u, 0u, 0u};
    acc_t0(g2, dC, cC);
    acc_bit<6>(g2, dB, cB);
    acc_bit<7>(g2, dA, cA);
    if (live) {
      G2[g2s] = make_uint4(g2[0], g2[1], g2[2], g2[3]);
      cd[cds] = cB;
    }
    __syncthreads();
    // H stage: the window centred on row h+1-d selected row h iff its H one-hot bit d is set
    uint32_t g1[4] = {0u, 0u, 0u, 0u};
    if (core) {
      acc_bit<3>(g1, G2[g2s + W * cgn], cd[cds + crow]);
      acc_bit<4>(g1, make_uint4(g2[0], g2[1], g2[2], g2[3]), cB);     // own row: still in registers
      acc_bit<5>(g1, G2[g2s - W * cgn], cd[cds - crow]);
      G1[g1s] = make_uint4(g1[0], g1[1], g1[2], g1[3]);
    }
    __syncthreads();
    // W stage: the window centred on column w+1-d selected column w iff its W one-hot bit d is set
    if (core) {
      uint32_t g0[4] = {av.x, av.y, av.z, av.w};                      // the sibling branches' gradient (zeros without one)
      acc_bit<0>(g0, G1[g1s + cgn], cd[cds + cgn]);
      acc_bit<1>(g0, make_uint4(g1[0], g1[1], g1[2], g1[3]), cB);
      acc_bit<2>(g0, G1[g1s - cgn], cd[cds - cgn]);
      uint4 o = make_uint4(g0[0], g0[1], g0[2], g0[3]);
      if (relu_src) { o.x &= relu_mask2(rv.x); o.y &= relu_mask2(rv.y); o.z &= relu_mask2(rv.z); o.w &= relu_mask2(rv.w); }
      *reinterpret_cast<uint4*>(dx + boff + t * plane) = o;
    }
    dA = dB; dB = dC; dC = dN;
    cA = cB; cB = cC; cC = cN;
  }
}


// ---------------------------------------------------------------------------------------------
// Backward of the stride-2 pools ([1,3,3]/[1,2,2] at i3d.py:174,189 and 3x3x3/2x2x2 at :252).
// One thread owns a 2x2(x2) patch of input elements x 8 channels: the patch is covered by at most
// 2x2(x2) windows, so every arg-max byte and every dy value is loaded once per patch (all loads are
// issued before the first use) instead of once per input element.
// padded coordinate hp = h + pad_before; patch q holds hp = 2q, 2q+1; window ho = q contains both
// (tap dh = 0 / 1), window ho = q-1 contains only hp = 2q (tap dh = 2).
// ---------------------------------------------------------------------------------------------
// bit 7 of every byte of the result is set iff that arg-max byte equals the tap (exact per byte: no borrow between
// bytes); acc_msb turns those bits into halfword masks with PRMT
__device__ __forceinline__ uint32_t eq_msb(uint32_t iv, uint32_t tap4) {
  const uint32_t x = iv ^ tap4;
  return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x);
}

// ReLU mask of the pool INPUT from the pool OUTPUT (`pooled`, used when there is no addend): an input element only
// receives gradient from windows it won, and for those the pooled value IS its own value, so (input > 0) == (pooled > 0)
// on every element whose gradient is non-zero.  Masking dy by (pooled > 0) window by window gives bit-identical results
// and replaces the read of the full-resolution producer output (4x / 8x the pooled tensor: 411 MB at MaxPool3d_2a)
// by a read of the pooled tensor — a third of the kernel's DRAM traffic (ncu r02_c3: 5.3 TB/s, already HBM-bound).
template <int KT, bool POOLED>   // temporal kernel: 1 (stride 1) or 3 (stride 2); POOLED: ReLU mask from the pool output
__global__ void __launch_bounds__(256)
pool_s2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                   const __nv_bfloat16* __restrict__ addend, const __half* __restrict__ relu_src,
                   const __half* __restrict__ pooled,
                   __nv_bfloat16* __restrict__ dx, const PoolGeom g, const int Qt, const int Qh, const int Qw) {
  pdl_sync();
  constexpr int NA = KT == 3 ? 2 : 1;
  const int cg = g.C >> 3;
  const int i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i >= Qw * cg) return;
  const int qw = i / cg;
  const int c8 = i - qw * cg;
  int row = blockIdx.x;
  const int qh = row % Qh; row /= Qh;
  const int qt = row % Qt;
  const int b = row / Qt;
  // ---- all window loads first ----
  uint2 iv[NA][2][2];
  uint4 dv[NA][2][2];
  uint4 pv[POOLED ? NA : 1][2][2];
  // one 64-bit offset per tensor, the windows / patch elements are 32-bit strides away from it (the SASS of the first
  // version spent 40 % of its 650 instructions on 64-bit index arithmetic; the kernel is issue bound)
  const long long woff0 = (((static_cast<long long>(b) * g.To + qt) * g.Ho + qh) * g.Wo + qw) * g.C + c8 * 8;
  const int wsW = g.C, wsH = g.Wo * g.C;
  const long long wsT = static_cast<long long>(g.Ho) * wsH;
#pragma unroll
  for (int a = 0; a < NA; ++a)
#pragma unroll
    for (int bb = 0; bb < 2; ++bb)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int to = qt - a, ho = qh - bb, wo = qw - c;
        const bool ok = to >= 0 && to < g.To && ho >= 0 && ho < g.Ho && wo >= 0 && wo < g.Wo;
        const long long off = woff0 - (a ? wsT : 0) - (bb ? wsH : 0) - (c ? wsW : 0);
        iv[a][bb][c] = ok ? __ldg(reinterpret_cast<const uint2*>(idx + off)) : make_uint2(0xffffffffu, 0xffffffffu);
        dv[a][bb][c] = ok ? __ldg(reinterpret_cast<const uint4*>(dy + off)) : make_uint4(0u, 0u, 0u, 0u);
        if (POOLED) pv[a][bb][c] = ok ? __ldg(reinterpret_cast<const uint4*>(pooled + off)) : make_uint4(0u, 0u, 0u, 0u);
      }
  if (POOLED) {
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int bb = 0; bb < 2; ++bb)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          dv[a][bb][c].x &= relu_mask2(pv[a][bb][c].x); dv[a][bb][c].y &= relu_mask2(pv[a][bb][c].y);
          dv[a][bb][c].z &= relu_mask2(pv[a][bb][c].z); dv[a][bb][c].w &= relu_mask2(pv[a][bb][c].w);
        }
  }
#pragma unroll
  for (int et = 0; et < NA; ++et) {
    const int t = KT == 3 ? 2 * qt - g.pt + et : qt;
    if (t < 0 || t >= g.T) continue;
    // epilogue operands of the four in-plane patch elements
    uint4 rv[2][2], av[2][2];
    long long eo[2][2];
    bool live[2][2];
    const long long eo00 = (((static_cast<long long>(b) * g.T + t) * g.H + (2 * qh - g.ph)) * g.W + (2 * qw - g.pw)) * g.C + c8 * 8;
#pragma unroll
    for (int eh = 0; eh < 2; ++eh)
#pragma unroll
      for (int ew = 0; ew < 2; ++ew) {
        const int h = 2 * qh - g.ph + eh, w = 2 * qw - g.pw + ew;
        live[eh][ew] = h >= 0 && h < g.H && w >= 0 && w < g.W;
        eo[eh][ew] = eo00 + (eh ? g.W * g.C : 0) + (ew ? g.C : 0);
        rv[eh][ew] = (!POOLED && live[eh][ew] && relu_src) ? __ldg(reinterpret_cast<const uint4*>(relu_src + eo[eh][ew]))
                                                 : make_uint4(kF16One2, kF16One2, kF16One2, kF16One2);
        av[eh][ew] = (live[eh][ew] && addend) ? __ldg(reinterpret_cast<const uint4*>(addend + eo[eh][ew]))
                                               : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
    for (int eh = 0; eh < 2; ++eh)
#pragma unroll
      for (int ew = 0; ew < 2; ++ew) {
        if (!live[eh][ew]) continue;
        const uint4 a4 = av[eh][ew];
        uint32_t acc[4] = {a4.x, a4.y, a4.z, a4.w};   // packed bf16x2 sums (at most 4 / 8 windows cover an element)
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          if (a == 1 && et != 0) continue;
          const int dt = KT == 3 ? (a ? 2 : et) : 0;
#pragma unroll
          for (int bb = 0; bb < 2; ++bb) {
            if (bb == 1 && eh != 0) continue;
            const int dh = bb ? 2 : eh;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c == 1 && ew != 0) continue;
              const int dw = c ? 2 : ew;
              const uint32_t tap4 = static_cast<uint32_t>((dt * 3 + dh) * 3 + dw) * 0x01010101u;
              acc_msb(acc, dv[a][bb][c], eq_msb(iv[a][bb][c].x, tap4), eq_msb(iv[a][bb][c].y, tap4));
            }
          }
        }
        const uint4 r = rv[eh][ew];
        uint4 o = make_uint4(acc[0], acc[1], acc[2], acc[3]);
        if (!POOLED) { o.x &= relu_mask2(r.x); o.y &= relu_mask2(r.y); o.z &= relu_mask2(r.z); o.w &= relu_mask2(r.w); }
        *reinterpret_cast<uint4*>(dx + eo[eh][ew]) = o;
      }
  }
}

// tiling of a plane: whole plane per CTA when it fits with >= 4 channel groups side by side (64 contiguous bytes per
// position), else row tiles of R rows (+1 halo row each side) with cgn = 4 (or the largest divisor of C/8 below)
struct PoolTiling { int cgn, R, nth, halo, threads; };
PoolTiling pick_tiling(int H, int W, int C, int max_threads, int max_tiled = 0) {
  if (max_tiled <= 0) max_tiled = max_threads;
  PoolTiling t{0, 0, 0, 0, 0};
  if (C % 8) return t;
  const int cg = C / 8, hw = H * W;
  int full = 0;
  for (int c = 1; c <= 16 && c * hw <= max_threads; ++c)
    if (cg % c == 0) full = c;
  int tc = 0;
  static int cg_pref = -1;
  if (cg_pref < 0) {
    const char* ev = getenv("FAV_POOL_CGN");
    cg_pref = ev ? atoi(ev) : 4;
  }
  for (int c = cg_pref; c >= 1; --c)
    if (cg % c == 0) { tc = c; break; }
  const int Rtile = tc > 0 ? max_tiled / (W * tc) - 2 : 0;
  if (full >= 4 || (full > 0 && (Rtile < 2 || tc <= full))) {
    t.cgn = full; t.R = H; t.nth = 1; t.halo = 0; t.threads = round_up(hw * full, 32);
  } else if (Rtile >= 2) {
    t.cgn = tc; t.R = std::min(Rtile, H); t.nth = ceil_div(H, t.R); t.halo = 1;
    t.threads = round_up((t.R + 2) * W * tc, 32);
  } else if (full > 0) {
    t.cgn = full; t.R = H; t.nth = 1; t.halo = 0; t.threads = round_up(hw * full, 32);
  }
  return t;
}

}  // namespace

// T segment length: one CTA streams `seg` output frames and reads seg + 2 input frames, and the grid runs in waves of
// one CTA per SM, so the time is ~ waves x (seg + 2 + fixed cost).  Fewer, longer segments win whenever they remove a
// mostly empty last wave (Mixed_3b backward: 192 CTAs = 2 waves of 10 frames -> 144 CTAs = 1 wave of 13).
static int pool_threads();
static int pick_tseg(int T, long long units_per_segment) {
  int dev = 0;
  cudaGetDevice(&dev);
  const int sms = sm_count(dev) * (pool_threads() <= kPoolThreads / 2 ? 2 : 1);   // CTAs resident at once
  double best = 1e30;
  int best_seg = T;
  for (int seg = 2; seg <= T; ++seg) {
    const int nseg = ceil_div(T, seg);
    const long long waves = ceil_div64(units_per_segment * nseg, sms);
    const double cost = static_cast<double>(waves) * (seg + 2 + 1.5);
    if (cost < best - 1e-9) { best = cost; best_seg = seg; }
  }
  if (const char* ev = getenv("FAV_POOL_TSEG")) best_seg = std::max(1, std::min(T, atoi(ev)));
  return best_seg;
}

// threads per CTA of the streaming kernels (<= kPoolThreads).  FAV_POOL_THREADS=400 lets two CTAs share an SM (each
// waits at its own barriers) at the price of narrower tiles.
static int pool_threads() {
  static int v = -1;
  if (v < 0) {
    const char* ev = getenv("FAV_POOL_THREADS");
    v = ev ? std::max(128, std::min(kPoolThreads, atoi(ev))) : kPoolThreads;
  }
  return v;
}

bool pool3s1_applicable(const PoolGeom& g) {
  return g.kt == 3 && g.kh == 3 && g.kw == 3 && g.st == 1 && g.sh == 1 && g.sw == 1 &&
         pick_tiling(g.H, g.W, g.C, kPoolThreads).cgn > 0;
}

int launch_pool3s1_fwd(const __half* x, __half* y, uint8_t* idx, const PoolGeom& g, cudaStream_t s) {
  const PoolTiling t = pick_tiling(g.H, g.W, g.C, pool_threads(), pool_threads() < kPoolThreads ? pool_threads() : 1024);
  FAV_CHECK_ARG(t.cgn > 0, "pool3s1: plane %dx%d with C=%d not supported", g.H, g.W, g.C);
  const int tseg = pick_tseg(g.T, static_cast<long long>(g.C / (8 * t.cgn)) * t.nth * g.B);
  const size_t smem = (static_cast<size_t>(t.R + 2 * t.halo) * (g.W + 2) + static_cast<size_t>(t.R + 2) * g.W) * t.cgn * 16;
  static bool attr = false;
  if (!attr) {
    FAV_CUDA(cudaFuncSetAttribute(pool3s1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  dim3 grid(g.C / (8 * t.cgn) * t.nth, ceil_div(g.T, tseg), g.B);
  FAV_CUDA(launch_pdl(pool3s1_fwd_kernel, grid, t.threads, smem, s, x, y, idx, g.T, g.H, g.W, g.C, t.cgn, tseg, t.R, t.nth, t.halo));
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

int launch_pool3s1_bwd(const __nv_bfloat16* dy, const uint8_t* idx, const __nv_bfloat16* addend,
                       const __half* relu_src, __nv_bfloat16* dx, const PoolGeom& g, cudaStream_t s) {
  const PoolTiling t = pick_tiling(g.H, g.W, g.C, pool_threads());
  FAV_CHECK_ARG(t.cgn > 0, "pool3s1: plane %dx%d with C=%d not supported", g.H, g.W, g.C);
  const int tseg = pick_tseg(g.T, static_cast<long long>(g.C / (8 * t.cgn)) * t.nth * g.B);
  const size_t smem = (static_cast<size_t>(t.R + 2) * g.W + static_cast<size_t>(t.R) * (g.W + 2)) * t.cgn * 16 +
                      static_cast<size_t>(2) * (t.R + 2) * (g.W + 2) * t.cgn * 8;
  static bool attr = false;
  if (!attr) {
    FAV_CUDA(cudaFuncSetAttribute(pool3s1_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  dim3 grid(g.C / (8 * t.cgn) * t.nth, ceil_div(g.T, tseg), g.B);
  FAV_CUDA(launch_pdl(pool3s1_bwd_kernel, grid, t.threads, smem, s, dy, idx, addend, relu_src, dx, g.T, g.H, g.W, g.C, t.cgn,
                      tseg, t.R, t.nth, t.halo));
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

bool pool_s2_applicable(const PoolGeom& g) {
  const bool hw = g.kh == 3 && g.kw == 3 && g.sh == 2 && g.sw == 2;
  return hw && ((g.kt == 1 && g.st == 1) || (g.kt == 3 && g.st == 2)) && g.C % 8 == 0;
}

int launch_pool_s2_bwd(const __nv_bfloat16* dy, const uint8_t* idx, const __nv_bfloat16* addend,
                       const __half* relu_src, __nv_bfloat16* dx, const PoolGeom& g, cudaStream_t s, const __half* pooled) {
  FAV_CHECK_ARG(pool_s2_applicable(g), "pool_s2_bwd: unsupported geometry");
  if (pooled && !addend) relu_src = nullptr;   // the window masks carry the ReLU
  else pooled = nullptr;
  const int Qt = g.kt == 3 ? (g.T - 1 + g.pt) / 2 + 1 : g.T;
  const int Qh = (g.H - 1 + g.ph) / 2 + 1, Qw = (g.W - 1 + g.pw) / 2 + 1;
  dim3 grid(g.B * Qt * Qh, ceil_div(Qw * (g.C / 8), 256));
  if (g.kt == 3 && pooled) FAV_CUDA(launch_pdl(pool_s2_bwd_kernel<3, true>, grid, 256, 0, s, dy, idx, addend, relu_src, pooled, dx, g, Qt, Qh, Qw));
  else if (g.kt == 3) FAV_CUDA(launch_pdl(pool_s2_bwd_kernel<3, false>, grid, 256, 0, s, dy, idx, addend, relu_src, pooled, dx, g, Qt, Qh, Qw));
  else if (pooled) FAV_CUDA(launch_pdl(pool_s2_bwd_kernel<1, true>, grid, 256, 0, s, dy, idx, addend, relu_src, pooled, dx, g, Qt, Qh, Qw));
  else FAV_CUDA(launch_pdl(pool_s2_bwd_kernel<1, false>, grid, 256, 0, s, dy, idx, addend, relu_src, pooled, dx, g, Qt, Qh, Qw));
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

}  // namespace fav
