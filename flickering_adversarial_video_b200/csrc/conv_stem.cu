// conv_stem.cu — the strided RGB stem as a tcgen05 implicit GEMM with shared-memory halo reuse.
//
// Replaces Conv3d_1a_7x7 (7x7x7, stride 2, SAME; i3d.py:168-171) and, with KT/st/pads as parameters,
// the torchvision video stems ((3,7,7) and (1,7,7), stride (1,2,2), padding 3).
//
// Input is the RGBX buffer the apply kernel writes: [B,T,H,Wp,4] fp16 with W physically padded, so the
// 7 W taps x RGB of output column wo are the 8 positions x 4 channels = 32 contiguous fp16 starting at
// column 2*wo: one 64-byte K row per output position per (kt,kh) tap.  A tensor map whose W stride is
// 16 bytes (two positions) exposes those overlapping rows directly; four maps cover the (T,H)
// parities, so that "input row 2*ho + kh - pad" is row ho + qh of the parity-(kh-pad)&1 map.
//
// A CTA tile is 16 output columns x 8*mt output rows of one output frame.  Per temporal tap kt (one
// pipeline stage) the producer brings, for each H parity, ONE slab of 8*mt + span rows x 16 columns
// x 64 B; the 7 kh taps are 1024-byte-aligned row windows of those two slabs (2*8*mt + 5 slab rows
// instead of 7*8*mt per stage), and the stage's 7 weight sub-tiles are shared by the mt M tiles.
#include "conv_umma.cuh"

#include <string.h>
#include <stdlib.h>
#include <algorithm>

namespace fav {
extern __device__ unsigned long long g_halo_prof[8];
namespace {

constexpr int kThreads = 192;
constexpr int kMaxStages = 6;

struct StemTile {
  int b, t, h0, w0;
};
__device__ __forceinline__ StemTile decode_stem_tile(const StemGeom& g, int tile) {
  StemTile c;
  const int wi = tile % g.tw;
  int m = tile / g.tw;
  const int hi = m % g.th;
  m /= g.th;
  c.t = m % g.To;
  c.b = m / g.To;
  c.h0 = hi * 8 * g.mt;
  c.w0 = wi * 16;
  return c;
}
__device__ __forceinline__ int border_cls(int o, int n, int nlo, int nhi) {
  return o < nlo ? o : (o >= n - nhi ? nlo + 1 + (o - (n - nhi)) : nlo);
}

__global__ void __launch_bounds__(kThreads, 1)
conv_stem_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
                 const __grid_constant__ CUtensorMap tmB, const StemGeom g, const ConvEpilogue e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(g.stages) * g.stage_bytes);
  uint64_t* full_bar = bars;                    // [stages]
  uint64_t* empty_bar = bars + kMaxStages;      // [stages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint4* stage_all = reinterpret_cast<uint4*>(bars + 2 * kMaxStages + 6);   // 4 epilogue warps x 32 rows x 5 uint4

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int acc_cols = g.mt * g.bn;             // TMEM columns per accumulator stage

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmA3);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < g.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // prologue done; global memory only after the previous kernels of the stream have completed

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
        const StemTile tc = decode_stem_tile(g, tile);
        for (int kt = 0; kt < g.KT; ++kt) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + static_cast<size_t>(stage) * g.stage_bytes;
          mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(g.a_bytes + g.b_bytes));
          // input frame = st*t + kt - pt = st*(t + qt) + parity
          int par_t = 0, tcoord;
          if (g.st == 2) {
            const int offt = kt - g.pt;
            par_t = offt & 1;
            tcoord = tc.t + ((offt - par_t) >> 1);
          } else {
            tcoord = tc.t + kt - g.pt;
          }
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const int rows_p = p ? g.rows[1] : g.rows[0];
            if (rows_p == 0) continue;
            const int mi = par_t * 2 + p;
            const CUtensorMap* tm = mi == 0 ? &tmA0 : (mi == 1 ? &tmA1 : (mi == 2 ? &tmA2 : &tmA3));
            tma_load_5d(sa + (p ? g.slab_off[1] : g.slab_off[0]), tm, &full_bar[stage], 0, tc.w0,
                        tc.h0 + (p ? g.qmin[1] : g.qmin[0]), tcoord, tc.b);
          }
          tma_load_3d(sa + g.a_bytes, &tmB, &full_bar[stage], 0, 0, kt * g.KH);
          if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    const uint32_t idesc = umma_idesc(128, g.bn, true);   // fp16 clip operand x fp16 weights
    const uint32_t desc_hi = umma_desc_hi(64);
    const uint32_t b_sub = static_cast<uint32_t>(g.bn) * 64u;   // bytes per kh weight sub-tile
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * acc_cols);
      for (int kt = 0; kt < g.KT; ++kt) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * g.stage_bytes);
        const uint32_t sb = sa + static_cast<uint32_t>(g.a_bytes);
        if (elect_one()) {
          for (int kh = 0; kh < g.KH; ++kh) {
            const int offh = kh - g.ph;
            const int p = offh & 1;
            const int qh = (offh - p) >> 1;
            const int so = p ? g.slab_off[1] : g.slab_off[0];
            const int q0 = p ? g.qmin[1] : g.qmin[0];
            const uint32_t a_lo = umma_desc_lo(sa + static_cast<uint32_t>(so + (qh - q0) * 1024));
            const uint32_t b_lo = umma_desc_lo(sb + kh * b_sub);
            const uint32_t accum = (kt | kh) ? 1u : 0u;
            for (int i = 0; i < g.mt; ++i) {
              const uint32_t a_i = a_lo + static_cast<uint32_t>(i) * 512u;   // 8 rows x 1024 B, in 16-byte units
              const uint32_t d_i = d_tmem + static_cast<uint32_t>(i * g.bn);
              umma_bf16(d_i, make_desc(desc_hi, a_i), make_desc(desc_hi, b_lo), idesc, accum);
              umma_bf16(d_i, make_desc(desc_hi, a_i + 2), make_desc(desc_hi, b_lo + 2), idesc, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (kt == g.KT - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == g.stages) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rw = row & 15;
    const int rh = row >> 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
      const StemTile tc = decode_stem_tile(g, tile);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int w = tc.w0 + rw;
      for (int i = 0; i < g.mt; ++i) {
        const int h = tc.h0 + i * 8 + rh;
        const bool valid = (w < g.Wo) && (h < g.Ho);
        const long long pos = valid ? ((static_cast<long long>(tc.b) * g.To + tc.t) * g.Ho + h) * g.Wo + w : 0;
        h16* out_row = e.out + pos * e.out_cs + e.out_coff;
        const float* bias_row = nullptr;
        if (e.bias) {
          int br = 0;
          if (e.bias_stem) {
            const int hc = border_cls(min(h, g.Ho - 1), g.Ho, g.nlo_h, g.nhi_h);
            const int wc = border_cls(min(w, g.Wo - 1), g.Wo, g.nlo_w, g.nhi_w);
            br = (tc.t * 4 + hc) * 4 + wc;
          }
          bias_row = e.bias + static_cast<long long>(br) * e.bias_ld;
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc * acc_cols + i * g.bn);
        // coalesced stores through shared memory (the bias row differs per lane: border classes of the delta table)
        epilogue_columns_staged(e, g.bn, 0, taddr, valid, out_row, bias_row, e.cout_store, stage_all + (warp - 2) * 160, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Raw-row variant.  ncu on the kernel above (profiles/r02_c1_ncu_full_details.txt): tensor pipe busy 74 % of the time
// but at 66 cycles per 128x64x16 MMA instead of the 48 the operand reads need — the TMA writes of the im2col'd A tiles
// (every input pixel 4x, 66 KB per stage, 7 GB of L2 -> SM traffic per step) share the 128 B/clk shared-memory port
// with the MMA's operand reads.  Here the producer brings RAW input rows (40 RGBX pixels = 320 B per row) and the MMA
// reads the im2col windows straight out of them: a no-swizzle descriptor with LBO = 16 B, SBO = 320 B makes row r of
// the M tile the 32-byte window at 16*(r%8) + 320*(r/8), i.e. M tile = 8 output columns x 16 output rows (the K = 32
// of a (kt,kh) tap = pixels 2w..2w+7 x RGBX; second K = 16 half at +32 B).  A stage holds, per H parity, the rows of
// nf = 2 output frames (24 KB instead of 76) next to the 28 KB of weights, and the 2*nf = 4 M tiles
// ((8-column half, frame)) share every weight sub-tile: the TMA write share per MMA drops from 2.4 KB to 0.9 KB.
// ---------------------------------------------------------------------------------------------------------------------
struct RawTile {
  int b, t0, h0, w0;
};
__device__ __forceinline__ RawTile decode_raw_tile(const StemGeom& g, int tile) {
  RawTile c;
  const int wi = tile % g.tw;
  int m = tile / g.tw;
  const int hi = m % g.th;
  m /= g.th;
  c.t0 = (m % g.tp) * g.nf;
  c.b = m / g.tp;
  c.h0 = hi * 16;
  c.w0 = wi * 16;
  return c;
}

__global__ void __launch_bounds__(kThreads, 1)
conv_stem_raw_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                     const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
                     const __grid_constant__ CUtensorMap tmB, const StemGeom g, const ConvEpilogue e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(g.stages) * g.stage_bytes);
  uint64_t* full_bar = bars;                    // [stages]
  uint64_t* empty_bar = bars + kMaxStages;      // [stages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint4* stage_all = reinterpret_cast<uint4*>(bars + 2 * kMaxStages + 6);   // 4 epilogue warps x 32 rows x 5 uint4

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int acc_cols = g.mt * g.bn;             // TMEM columns per accumulator stage
  // stage layout: [weights KH x bn x 64 B (SW64)] [parity-0 slab: nf x rows[0] x pitch] [parity-1 slab]

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmA3);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < g.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t a_tx = static_cast<uint32_t>((g.rows[0] + g.rows[1]) * g.nf * g.pitch);
      for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
        const RawTile tc = decode_raw_tile(g, tile);
        for (int kt = 0; kt < g.KT; ++kt) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sb = smem + static_cast<size_t>(stage) * g.stage_bytes;
          mbar_expect_tx(&full_bar[stage], a_tx + static_cast<uint32_t>(g.b_bytes));
          // input frame of output frame t: st*t + kt - pt = st*(t + qt) + parity; the nf frames are consecutive in the map
          int par_t = 0, tcoord;
          if (g.st == 2) {
            const int offt = kt - g.pt;
            par_t = offt & 1;
            tcoord = tc.t0 + ((offt - par_t) >> 1);
          } else {
            tcoord = tc.t0 + kt - g.pt;
          }
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            if ((p ? g.rows[1] : g.rows[0]) == 0) continue;
            const int mi = par_t * 2 + p;
            const CUtensorMap* tm = mi == 0 ? &tmA0 : (mi == 1 ? &tmA1 : (mi == 2 ? &tmA2 : &tmA3));
            tma_load_5d(sb + (p ? g.slab_off[1] : g.slab_off[0]), tm, &full_bar[stage], 8 * tc.w0,
                        tc.h0 + (p ? g.qmin[1] : g.qmin[0]), tcoord, tc.b, 0);
          }
          tma_load_3d(sb, &tmB, &full_bar[stage], 0, 0, kt * g.KH);
          if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc(128, g.bn, true);
    const uint32_t hi_a = umma_desc_hi_nosw(static_cast<uint32_t>(g.pitch));   // SBO = one raw row; LBO = 16 B (in the low word)
    const uint32_t hi_b = umma_desc_hi(64);
    const uint32_t b_sub = static_cast<uint32_t>(g.bn) * 64u;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * acc_cols);
      for (int kt = 0; kt < g.KT; ++kt) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sb = smem_u32(smem + static_cast<size_t>(stage) * g.stage_bytes);
        if (elect_one()) {
          for (int kh = 0; kh < g.KH; ++kh) {
            const int offh = kh - g.ph;
            const int p = offh & 1;
            const int qh = (offh - p) >> 1;
            const uint32_t slab = sb + static_cast<uint32_t>((p ? g.slab_off[1] : g.slab_off[0]) +
                                                             (qh - (p ? g.qmin[1] : g.qmin[0])) * g.pitch);
            const uint32_t fstep = static_cast<uint32_t>((p ? g.rows[1] : g.rows[0]) * g.pitch);
            const uint32_t b_lo = umma_desc_lo(sb + kh * b_sub);
            const uint32_t accum = (kt | kh) ? 1u : 0u;
            for (int i = 0; i < g.mt; ++i) {
              // M tile i: frame i >> 1, 8-column half i & 1 (128 B = 8 output columns x 16 B)
              const uint32_t a_lo = umma_desc_lo(slab + static_cast<uint32_t>(i >> 1) * fstep + static_cast<uint32_t>(i & 1) * 128u);
              const uint32_t d_i = d_tmem + static_cast<uint32_t>(i * g.bn);
              umma_bf16(d_i, make_desc(hi_a, a_lo), make_desc(hi_b, b_lo), idesc, accum);
              umma_bf16(d_i, make_desc(hi_a, a_lo + 2), make_desc(hi_b, b_lo + 2), idesc, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (kt == g.KT - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == g.stages) { stage = 0; phase ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rw = row & 7;
    const int rh = row >> 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
      const RawTile tc = decode_raw_tile(g, tile);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int h = tc.h0 + rh;
      for (int i = 0; i < g.mt; ++i) {
        const int t = tc.t0 + (i >> 1);
        const int w = tc.w0 + (i & 1) * 8 + rw;
        const bool valid = (w < g.Wo) && (h < g.Ho) && (t < g.To);
        const long long pos = valid ? ((static_cast<long long>(tc.b) * g.To + t) * g.Ho + h) * g.Wo + w : 0;
        h16* out_row = e.out + pos * e.out_cs + e.out_coff;
        const float* bias_row = nullptr;
        if (e.bias) {
          int br = 0;
          if (e.bias_stem) {
            const int hc = border_cls(min(h, g.Ho - 1), g.Ho, g.nlo_h, g.nhi_h);
            const int wc = border_cls(min(w, g.Wo - 1), g.Wo, g.nlo_w, g.nhi_w);
            br = (min(t, g.To - 1) * 4 + hc) * 4 + wc;
          }
          bias_row = e.bias + static_cast<long long>(br) * e.bias_ld;
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc * acc_cols + i * g.bn);
        epilogue_columns_staged(e, g.bn, 0, taddr, valid, out_row, bias_row, e.cout_store, stage_all + (warp - 2) * 160, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Temporal-sharing variant.  The raw-row kernel above is bound by its A-operand reads: a 128x64x16 MMA needs 32 cycles of
// the tensor pipe but reads 4 KB of A + 2 KB of B through the 128 B/clk shared-memory port (48 cycles; ncu: tensor pipe
// 49 % busy).  An input frame feeds up to ceil(KT/st) output frames (t_in = st*t_o + kt - pt), each with its own kt tap —
// so here ONE MMA multiplies a frame's row windows by the stacked weight sub-tiles of all those taps (N = up to 4 x 64)
// and writes the TMEM accumulators of all those output frames at once: 4 KB of A per 128x256x16 MMA instead of per
// 128x64x16, tensor-bound (128 cycles) instead of port-bound (4 x 48).
//   tile       = 8 output columns x 16 output rows x G = 4 consecutive output frames (G accumulators of bn columns, x2 stages)
//   A set      = the raw rows of the tile's input frames of one class c (frame index mod st; st*(G-1)+KT frames in all),
//                loaded once per tile; two sets in flight
//   weight ring= blocks of ts_khg (3) kh taps of one class: per kh the class's kt sub-tiles stacked by DESCENDING kt, so
//                the sub-tiles of consecutive output frames j, j+1, ... (kt = dd - st*j) are consecutive rows of the B
//                operand; 36-48 KB per block from L2, one barrier hand-off per block
//   loop       = class -> weight block -> kh -> frame -> K half; the frames of a kh are issued narrowest first, widest last
// The first MMA that touches accumulator j is (class 0, kh 0, frame dd = st*j, K half 0) with j the TOP of the frame's
// range: that MMA is split into an accumulating part and a fresh (accumulate = 0) part of N = bn.
// Measured (FAV_STEM_PROF, profiles/r02b_stem_ts_prof.txt; I3D 8 x 64: 0.58 -> 0.44 ms): the kernel now sits on the
// shared-memory / L1 port — per tile 1.5 MB of operand reads + 0.28 MB of TMA writes + 0.3 MB of epilogue traffic = 16.5
// kcycles at 128 B/clk against 13.4 k of tensor time; 17.2 k measured.  The tensor pipe queues only ~2 MMAs behind the
// running one, so what is queued when the issuing thread goes through a barrier hand-off decides how long the pipe idles:
// hence the widest-last order (894 -> 798 kcycles per CTA) and 3 kh taps per block (826 / 749 / 730 at 1 / 2 / 3).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kThreadsTs = 352;   // warp 0: A producer, 1: MMA issuer, 6: weight producer, 2..5 and 7..10: two epilogue sets
constexpr int kMaxW = 8;
constexpr int kTsMaxFr = 7;   // input frames of one class per tile (st*(G-1)+KT frames over st classes)

struct TsTile {
  int b, t0, h0, w0;
};
// 64-bit descriptor from its two words without 64-bit arithmetic in the issue loop
__device__ __forceinline__ uint64_t pack_desc(uint32_t hi, uint32_t lo) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ TsTile decode_ts_tile(const StemGeom& g, int tile) {
  TsTile c;
  const int wi = tile % g.tw;
  int m = tile / g.tw;
  const int hi = m % g.th;
  m /= g.th;
  c.t0 = (m % g.tp) * g.tsG;
  c.b = m / g.tp;
  c.h0 = hi * 16;
  c.w0 = wi * 8;
  return c;
}

__global__ void __launch_bounds__(kThreadsTs, 1)
conv_stem_ts_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
                    const __grid_constant__ CUtensorMap tmB1, const StemGeom g, const ConvEpilogue e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  // [weight ring: ts_nw x ts_wblk_bytes][A sets: 2 x ts_set_bytes][barriers][staging]
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + static_cast<size_t>(g.ts_nw) * g.ts_wblk_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + 2 * static_cast<size_t>(g.ts_set_bytes));
  uint64_t* a_full = bars;                 // [2]
  uint64_t* a_empty = bars + 2;            // [2]
  uint64_t* w_full = bars + 4;             // [kMaxW]
  uint64_t* w_empty = bars + 4 + kMaxW;    // [kMaxW]
  uint64_t* tfull_bar = bars + 4 + 2 * kMaxW;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint4* stage_all = reinterpret_cast<uint4*>(bars + 4 + 2 * kMaxW + 6);   // 8 epilogue warps x 32 rows x 5 uint4

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int acc_cols = g.tsG * g.bn;
  const int nclass = g.st;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmA3);
    tma_prefetch_desc(&tmB1);
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < g.ts_nw; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 8); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  if (warp == 0) {
    // ===================== TMA producer: the input frames of one class = one A set =====================
    if (lane == 0) {
      int sa = 0;
      uint32_t pa = 0;
      const uint32_t frame_tx = static_cast<uint32_t>((g.rows[0] + g.rows[1]) * g.pitch);
      for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
        const TsTile tc = decode_ts_tile(g, tile);
        for (int c = 0; c < nclass; ++c) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          mbar_expect_tx(&a_full[sa], frame_tx * static_cast<uint32_t>(g.ts_nfr[c]));
          uint8_t* set = smem_a + static_cast<size_t>(sa) * g.ts_set_bytes;
          for (int f = 0; f < g.ts_nfr[c]; ++f) {
            const int t_in = g.st * tc.t0 - g.pt + c + f * g.st;   // may be < 0 or >= T: TMA zero fill
            int par_t = 0, tcoord = t_in;
            if (g.st == 2) {
              par_t = t_in & 1;
              tcoord = (t_in - par_t) >> 1;
            }
#pragma unroll
            for (int p = 0; p < 2; ++p) {
              if ((p ? g.rows[1] : g.rows[0]) == 0) continue;
              const int mi = par_t * 2 + p;
              const CUtensorMap* tm = mi == 0 ? &tmA0 : (mi == 1 ? &tmA1 : (mi == 2 ? &tmA2 : &tmA3));
              tma_load_5d(set + static_cast<size_t>(f) * g.ts_slot_bytes + (p ? g.slab_off[1] : g.slab_off[0]), tm,
                          &a_full[sa], 8 * tc.w0, tc.h0 + (p ? g.qmin[1] : g.qmin[0]), tcoord, tc.b, 0);
            }
          }
          sa ^= 1;
          if (sa == 0) pa ^= 1;
        }
      }
    }
  } else if (warp == 6) {
    // ===================== TMA producer: weight blocks of ts_khg kh taps of one class =====================
    int sb = 0;
    uint32_t pb = 0;
    const uint32_t sub = static_cast<uint32_t>(g.bn) * 64u;
    for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
      for (int c = 0; c < nclass; ++c) {
        const int nslot = g.ts_nslot[c];
        for (int kh0 = 0; kh0 < g.KH; kh0 += g.ts_khg) {
          const int nsub = min(g.ts_khg, g.KH - kh0) * nslot;   // sub-tiles of this block: [kh][slot], <= 32
          if (lane == 0) {
            mbar_wait(&w_empty[sb], pb ^ 1);
            mbar_expect_tx(&w_full[sb], sub * static_cast<uint32_t>(nsub));
          }
          __syncwarp();
          if (lane < nsub) {   // slot s holds tap kt = ktmax - s*st
            const int khi = lane / nslot;
            const int kt = g.ts_ktmax[c] - (lane - khi * nslot) * g.st;
            tma_load_3d(smem_w + static_cast<size_t>(sb) * g.ts_wblk_bytes + static_cast<size_t>(lane) * sub, &tmB1,
                        &w_full[sb], 0, 0, kt * g.KH + kh0 + khi);
          }
          if (++sb == g.ts_nw) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The issuing thread is the critical resource (~45 cycles per tcgen05.mma, ~110 per barrier wait): the per-frame
    // schedule (first output frame, top output frame, weight-row and accumulator-column offsets) comes from a table
    // built by stem_plan, one packed word per frame, held in registers for the whole class.
    const uint32_t hi_a = umma_desc_hi_nosw(static_cast<uint32_t>(g.pitch));
    const uint32_t hi_b = umma_desc_hi(64);
    const uint32_t sub16 = (static_cast<uint32_t>(g.bn) * 64u) >> 4;   // weight sub-tile in descriptor units
    const uint32_t slot16 = static_cast<uint32_t>(g.ts_slot_bytes) >> 4;
    const uint32_t idesc0 = umma_idesc(128, 0, true);
    const uint32_t bn_n = static_cast<uint32_t>(g.bn >> 3) << 17;      // N field of the instruction descriptor per output frame
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long w_te = 0, w_a = 0, w_w = 0, c0 = 0;   // FAV_STEM_PROF: barrier wait cycles of this warp
    const long long t_start = clock64();
    for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
      const int t0 = ((tile / (g.tw * g.th)) % g.tp) * g.tsG;
      const int geff = min(g.tsG, g.To - t0);   // output frames of this tile that exist
      c0 = g.prof ? clock64() : 0;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      if (g.prof) w_te += clock64() - c0;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * acc_cols);
      for (int c = 0; c < nclass; ++c) {
        // per-frame operands of this (tile, class), in registers: N (output frames fed), weight-row offset, accumulator
        // address and instruction descriptor — the block loop below only adds the block's base addresses
        int f_n[kTsMaxFr];
        uint32_t f_b[kTsMaxFr], f_d[kTsMaxFr], f_i[kTsMaxFr];
        bool f_fresh[kTsMaxFr];
        const int nfr = g.ts_nfr[c];
#pragma unroll
        for (int f = 0; f < kTsMaxFr; ++f) {
          // tab: jlo [0,4) | jtop [4,8) | accumulator column of jlo [8,18) | weight-row offset of jlo, 16-B units [18,32)
          const uint32_t tab = g.ts_tab[c][f];
          const int jlo = static_cast<int>(tab & 15u);
          const int jtop = static_cast<int>((tab >> 4) & 15u);
          const int n = f < nfr ? min(jtop, geff - 1) - jlo + 1 : 0;
          f_n[f] = n;
          f_b[f] = tab >> 18;
          f_d[f] = d_tmem + ((tab >> 8) & 1023u);
          f_i[f] = idesc0 + static_cast<uint32_t>(max(n, 1)) * bn_n;
          f_fresh[f] = jtop <= geff - 1;   // at (class 0, kh 0) the top output frame's tap is kt = 0: fresh accumulator
        }
        c0 = g.prof ? clock64() : 0;
        mbar_wait(&a_full[sa], pa);
        if (g.prof) w_a += clock64() - c0;
        const uint32_t set_lo = umma_desc_lo(smem_u32(smem_a + static_cast<size_t>(sa) * g.ts_set_bytes));
        const uint32_t kh16 = static_cast<uint32_t>(g.ts_nslot[c]) * sub16;   // one kh's sub-tiles inside a weight block
        for (int kh0 = 0; kh0 < g.KH; kh0 += g.ts_khg) {
          const int kh1 = min(g.KH, kh0 + g.ts_khg);
          c0 = g.prof ? clock64() : 0;
          mbar_wait(&w_full[sb], pb);
          if (g.prof) w_w += clock64() - c0;
          tc_fence_after();
          uint32_t wblk = umma_desc_lo(smem_u32(smem_w + static_cast<size_t>(sb) * g.ts_wblk_bytes));
          if (elect_one()) {
            for (int kh = kh0; kh < kh1; ++kh, wblk += kh16) {
              const int offh = kh - g.ph;
              const int p = offh & 1;
              const int qh = (offh - p) >> 1;
              const uint32_t a_kh = set_lo + (static_cast<uint32_t>((p ? g.slab_off[1] : g.slab_off[0]) +
                                                                    (qh - (p ? g.qmin[1] : g.qmin[0])) * g.pitch) >> 4);
              if (g.prof & 4) {
                // prof bit 2: timing without the MMAs
              } else if (c == 0 && kh == 0) {
                uint32_t a_lo = a_kh;
#pragma unroll
                for (int f = 0; f < kTsMaxFr; ++f) {
                  const int n = f_n[f];
                  if (n > 0) {
                    const uint32_t b_lo = wblk + f_b[f];
                    if (f_fresh[f]) {
                      if (n > 1) umma_bf16(f_d[f], pack_desc(hi_a, a_lo), pack_desc(hi_b, b_lo), f_i[f] - bn_n, 1u);
                      umma_bf16(f_d[f] + static_cast<uint32_t>((n - 1) * g.bn), pack_desc(hi_a, a_lo),
                                pack_desc(hi_b, b_lo + static_cast<uint32_t>(n - 1) * sub16), idesc0 + bn_n, 0u);
                    } else {
                      umma_bf16(f_d[f], pack_desc(hi_a, a_lo), pack_desc(hi_b, b_lo), f_i[f], 1u);
                    }
                    umma_bf16(f_d[f], pack_desc(hi_a, a_lo + 2), pack_desc(hi_b, b_lo + 2), f_i[f], 1u);
                  }
                  a_lo += slot16;
                }
              } else {
                // Narrow frames first, the widest last: between two blocks this thread spends ~400 cycles on the barrier
                // hand-offs, and the MMAs still queued must keep the tensor pipe busy meanwhile (N = 256: 128 cycles
                // each, N = 64: 51).  Any order is legal: every accumulator was initialised by the (class 0, kh 0) pass.
                constexpr int order[kTsMaxFr] = {0, 6, 1, 5, 2, 4, 3};
#pragma unroll
                for (int u = 0; u < kTsMaxFr; ++u) {
                  const int f = order[u];
                  if (f_n[f] > 0) {
                    const uint32_t a_lo = a_kh + static_cast<uint32_t>(f) * slot16;
                    const uint32_t b_lo = wblk + f_b[f];
                    umma_bf16(f_d[f], pack_desc(hi_a, a_lo), pack_desc(hi_b, b_lo), f_i[f], 1u);
                    umma_bf16(f_d[f], pack_desc(hi_a, a_lo + 2), pack_desc(hi_b, b_lo + 2), f_i[f], 1u);
                  }
                }
              }
            }
            umma_commit(&w_empty[sb]);
            if (kh1 == g.KH) {
              umma_commit(&a_empty[sa]);
              if (c == nclass - 1) umma_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++sb == g.ts_nw) { sb = 0; pb ^= 1; }
        }
        sa ^= 1;
        if (sa == 0) pa ^= 1;
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (g.prof && lane == 0) {
      atomicAdd(&g_halo_prof[0], static_cast<unsigned long long>(w_te));
      atomicAdd(&g_halo_prof[1], static_cast<unsigned long long>(w_a));
      atomicAdd(&g_halo_prof[2], static_cast<unsigned long long>(w_w));
      atomicAdd(&g_halo_prof[3], static_cast<unsigned long long>(clock64() - t_start));
    }
  } else {
    // ===================== epilogue: G M tiles = G output frames of the 8 x 16 patch =====================
    // Two sets of four warps (one warp per TMEM lane quarter each): set 0 drains the even output frames, set 1 the odd
    // ones (FAV_STEM_PROF: one set needs ~13 kcycles per tile, as long as the tile's MMAs).
    const int eset = warp >= 7 ? 1 : 0;
    const int ew = eset ? warp - 3 : warp - 2;   // staging slot 0..7
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rw = row & 7;
    const int rh = row >> 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long w_tf = 0;
    for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
      const TsTile tc = decode_ts_tile(g, tile);
      const long long c0 = g.prof ? clock64() : 0;
      mbar_wait(&tfull_bar[acc], acc_phase);
      if (g.prof) w_tf += clock64() - c0;
      tc_fence_after();
      const int h = tc.h0 + rh;
      const int w = tc.w0 + rw;
      for (int j = eset; j < g.tsG; j += 2) {
        const int t = tc.t0 + j;
        if (t >= g.To || (g.prof & 2)) break;   // warp-uniform (prof bit 1: timing without the epilogue's work)
        const bool valid = (w < g.Wo) && (h < g.Ho);
        const long long pos = valid ? ((static_cast<long long>(tc.b) * g.To + t) * g.Ho + h) * g.Wo + w : 0;
        h16* out_row = e.out + pos * e.out_cs + e.out_coff;
        const float* bias_row = nullptr;
        if (e.bias) {
          int br = 0;
          if (e.bias_stem) {
            const int hc = border_cls(min(h, g.Ho - 1), g.Ho, g.nlo_h, g.nhi_h);
            const int wc = border_cls(min(w, g.Wo - 1), g.Wo, g.nlo_w, g.nhi_w);
            br = (t * 4 + hc) * 4 + wc;
          }
          bias_row = e.bias + static_cast<long long>(br) * e.bias_ld;
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc * acc_cols + j * g.bn);
        epilogue_columns_staged(e, g.bn, 0, taddr, valid, out_row, bias_row, e.cout_store, stage_all + ew * 160, lane);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (g.prof && warp == 2 && lane == 0) atomicAdd(&g_halo_prof[4], static_cast<unsigned long long>(w_tf));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Frame schedule of conv_stem_ts_kernel: a tile of G = tsG output frames reads st*(G-1)+KT input frames; frame dd
// (relative to the tile's first input frame st*t0 - pt) feeds output frame j through tap kt = dd - st*j, so it belongs to
// class dd % st (= kt % st) and serves the consecutive output frames jlo .. jtop.  The class's kt sub-tiles are stacked by
// descending kt (slot s holds kt = ktmax - s*st): output frame jlo reads slot (ktmax - dd)/st + jlo, jlo+1 the next one.
int stem_ts_schedule(StemGeom* gp, int KT, int st, int bn, int* max_nfr_out, int* max_slot_out) {
  StemGeom& g = *gp;
  FAV_CHECK_ARG(st >= 1 && st <= 2 && KT >= 1 && KT <= 7 && g.tsG >= 1 && g.tsG <= 8, "stem schedule: KT=%d st=%d G=%d", KT,
                st, g.tsG);
  const int nd = st * (g.tsG - 1) + KT;        // input frames per tile
  int max_nfr = 0, max_slot = 0;
  for (int c = 0; c < st; ++c) {
    g.ts_nfr[c] = (nd - c + st - 1) / st;
    g.ts_ktmax[c] = c + ((KT - 1 - c) / st) * st;
    g.ts_nslot[c] = (g.ts_ktmax[c] - c) / st + 1;
    max_nfr = std::max(max_nfr, g.ts_nfr[c]);
    max_slot = std::max(max_slot, g.ts_nslot[c]);
  }
  FAV_CHECK_ARG(max_nfr <= kTsMaxFr, "stem: %d input frames per class", max_nfr);
  for (int c = 0; c < st; ++c)
    for (int f = 0; f < g.ts_nfr[c]; ++f) {
      const int dd = c + f * st;                       // frame index inside the tile: kt of output frame j = dd - st*j
      const int jtop = dd / st;
      const int num = dd - KT + 1;
      const int jlo = num > 0 ? (num + st - 1) / st : 0;
      const int slot0 = (g.ts_ktmax[c] - dd) / st + jlo;   // exact: dd and ktmax are in the same class
      // jlo [0,4) | jtop [4,8) | accumulator column of jlo [8,18) | weight-row offset of jlo in 16-byte units [18,32)
      g.ts_tab[c][f] = static_cast<uint32_t>(jlo) | (static_cast<uint32_t>(jtop) << 4) |
                       (static_cast<uint32_t>(jlo * bn) << 8) | (static_cast<uint32_t>(slot0 * bn * 4) << 18);
    }
  *max_nfr_out = max_nfr;
  *max_slot_out = max_slot;
  return FAV_OK;
}

int stem_plan(StemLaunch* L, int device, const void* xpad, int B, int T, int H, int Wp, const void* wpk, int bn,
              int To, int Ho, int Wo, int KT, int KH, int st, int pt, int ph) {
  FAV_CHECK_ARG(bn % 16 == 0 && bn >= 16 && bn <= 256, "stem: cout=%d must be a multiple of 16 <= 256", bn);
  FAV_CHECK_ARG(st == 1 || st == 2, "stem: temporal stride %d", st);
  FAV_CHECK_ARG(KT >= 1 && KT <= 7 && KH >= 1 && KH <= 7, "stem: taps %dx%d", KT, KH);
  memset(L, 0, sizeof(*L));
  StemGeom& g = L->g;
  g.B = B; g.To = To; g.Ho = Ho; g.Wo = Wo;
  g.KT = KT; g.KH = KH; g.st = st; g.pt = pt; g.ph = ph; g.bn = bn;
  {
    const char* ev = getenv("FAV_STEM_RAW");   // 0: the im2col-tile kernel (A/B)
    g.raw = (ev && atoi(ev) == 0) ? 0 : 1;
    if (bn > 128) g.raw = 0;                    // four accumulators of bn columns, double-buffered
  }
  {
    const char* ev = getenv("FAV_STEM_TS");    // 0: keep the raw-row kernel (A/B)
    g.ts = (g.raw && !(ev && atoi(ev) == 0) && 4 * bn <= 256 && KT >= 2 * st && To >= 2) ? 1 : 0;
  }
  if (g.ts) {
    g.tsG = 4;
    g.tp = ceil_div(To, g.tsG);
    g.th = ceil_div(Ho, 16);
    g.tw = ceil_div(Wo, 8);
    g.m_tiles = B * g.tp * g.th * g.tw;
    g.mt = g.tsG;
    {
      const char* ev = getenv("FAV_STEM_TS_PITCH");
      g.pitch = ev ? atoi(ev) : 176;            // 2*7 + 8 = 22 pixels of 8 B
      FAV_CHECK_ARG(g.pitch >= 176 && g.pitch % 16 == 0 && g.pitch <= 512, "stem: bad FAV_STEM_TS_PITCH");
    }
    int qlo[2] = {1 << 20, 1 << 20}, qhi[2] = {-(1 << 20), -(1 << 20)};
    for (int kh = 0; kh < KH; ++kh) {
      const int offh = kh - ph;
      const int p = offh & 1;
      const int qh = (offh - p) / 2;
      qlo[p] = std::min(qlo[p], qh);
      qhi[p] = std::max(qhi[p], qh);
    }
    int off = 0;
    for (int p = 0; p < 2; ++p) {
      if (qhi[p] < qlo[p]) { g.qmin[p] = 0; g.rows[p] = 0; g.slab_off[p] = off; continue; }
      g.qmin[p] = qlo[p];
      g.rows[p] = 16 + (qhi[p] - qlo[p]);
      g.slab_off[p] = off;
      off += round_up(g.rows[p] * g.pitch, 128);
    }
    g.ts_slot_bytes = off;
    int max_nfr = 0, max_slot = 0;
    FAV_TRY(stem_ts_schedule(&g, KT, st, bn, &max_nfr, &max_slot));
    g.ts_set_bytes = round_up(max_nfr * g.ts_slot_bytes, 1024);
    {
      const char* ev = getenv("FAV_STEM_TS_KHG");   // kh taps per weight block (one barrier hand-off per block)
      g.ts_khg = ev ? atoi(ev) : 3;      // measured (I3D 8 x 64): 1: 826, 2: 749, 3: 730 kcycles per CTA
      FAV_CHECK_ARG(g.ts_khg >= 1 && g.ts_khg <= KH && g.ts_khg * max_slot <= 32, "stem: bad FAV_STEM_TS_KHG");
    }
    g.b_bytes = bn * 64;
    const int fixed = 1024 + 2 * g.ts_set_bytes + 512 + 8 * 32 * 5 * 16;   // alignment, A sets, barriers, epilogue staging
    for (;; --g.ts_khg) {   // at least a double-buffered ring
      g.ts_wblk_bytes = round_up(g.ts_khg * max_slot * bn * 64, 1024);
      g.ts_nw = std::min(kMaxW, (227 * 1024 - fixed) / g.ts_wblk_bytes);
      if (g.ts_nw >= 2 || g.ts_khg == 1) break;
    }
    FAV_CHECK_ARG(g.ts_nw >= 2, "stem: temporal-sharing stages do not fit (%d + %d x n)", fixed, g.ts_wblk_bytes);
    L->smem_bytes = static_cast<size_t>(fixed) + static_cast<size_t>(g.ts_nw) * g.ts_wblk_bytes;
    auto classes = [](int in, int out, int k, int s, int pad, int* nlo, int* nhi) {
      *nlo = ceil_div(pad, s);
      const int pad_after = std::max(0, (out - 1) * s + k - pad - in);
      *nhi = ceil_div(pad_after, s);
    };
    classes(H, Ho, KH, 2, ph, &g.nlo_h, &g.nhi_h);
    classes(2 * Wo, Wo, 7, 2, ph, &g.nlo_w, &g.nhi_w);
    FAV_CHECK_ARG(g.nlo_h + g.nhi_h <= 3 && g.nlo_w + g.nhi_w <= 3, "stem: more than 4 border classes");
    const uint64_t pos_bytes = 8;
    const uint64_t row_pitch = static_cast<uint64_t>(Wp) * pos_bytes;
    const uint64_t frame_pitch = row_pitch * H;
    const uint64_t clip_pitch = frame_pitch * T;
    for (int p_t = 0; p_t < 2; ++p_t)
      for (int p_h = 0; p_h < 2; ++p_h) {
        if (g.rows[p_h] == 0 || (st == 1 && p_t == 1)) { L->tmA[p_t * 2 + p_h] = L->tmA[0]; continue; }
        // raw rows of one (T, H) parity: [Wp*4 elements][rows of this parity][frames of this parity][B][1]; one frame per box
        uint64_t dims[5] = {static_cast<uint64_t>(Wp) * 4, static_cast<uint64_t>((H - p_h + 1) / 2),
                            static_cast<uint64_t>(st == 2 ? (T - p_t + 1) / 2 : T), static_cast<uint64_t>(B), 1};
        uint64_t strides[4] = {2 * row_pitch, st * frame_pitch, clip_pitch, clip_pitch * B};
        uint32_t box[5] = {static_cast<uint32_t>(g.pitch / 2), static_cast<uint32_t>(g.rows[p_h]), 1, 1, 1};
        const char* base = static_cast<const char*>(xpad) + (st == 2 ? p_t : 0) * frame_pitch + p_h * row_pitch;
        FAV_TRY(make_tmap_bf16(&L->tmA[p_t * 2 + p_h], base, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE));
      }
    if (g.rows[0] == 0) L->tmA[0] = L->tmA[1];
    uint64_t bd[3] = {32, static_cast<uint64_t>(bn), static_cast<uint64_t>(KT) * KH};
    uint64_t bs[2] = {64, static_cast<uint64_t>(bn) * 64};
    uint32_t bb[3] = {32, static_cast<uint32_t>(bn), 1};
    FAV_TRY(make_tmap_bf16(&L->tmB1, wpk, 3, bd, bs, bb, CU_TENSOR_MAP_SWIZZLE_64B));
    L->tmB = L->tmB1;
    L->grid = std::max(1, std::min(g.m_tiles, sm_count(device)));
    return FAV_OK;
  }
  if (g.raw) {
    g.nf = To >= 2 ? 2 : 1;
    if (2 * 2 * g.nf * bn > 512) g.nf = 1;
    g.mt = 2 * g.nf;
    g.tp = ceil_div(To, g.nf);
    g.th = ceil_div(Ho, 16);
    g.tw = ceil_div(Wo, 16);
    g.m_tiles = B * g.tp * g.th * g.tw;
    g.pitch = 320;                              // 2*15 + 8 = 38 pixels of 8 B, padded to a multiple of 64 B
    int qlo[2] = {1 << 20, 1 << 20}, qhi[2] = {-(1 << 20), -(1 << 20)};
    for (int kh = 0; kh < KH; ++kh) {
      const int offh = kh - ph;
      const int p = offh & 1;
      const int qh = (offh - p) / 2;
      qlo[p] = std::min(qlo[p], qh);
      qhi[p] = std::max(qhi[p], qh);
    }
    g.b_bytes = KH * bn * 64;
    int off = round_up(g.b_bytes, 1024);
    for (int p = 0; p < 2; ++p) {
      if (qhi[p] < qlo[p]) { g.qmin[p] = 0; g.rows[p] = 0; g.slab_off[p] = off; continue; }
      g.qmin[p] = qlo[p];
      g.rows[p] = 16 + (qhi[p] - qlo[p]);
      g.slab_off[p] = off;
      off += round_up(g.nf * g.rows[p] * g.pitch, 128);
    }
    g.a_bytes = off - round_up(g.b_bytes, 1024);
    g.stage_bytes = round_up(off, 1024);
    g.stages = std::max(2, std::min(kMaxStages, (215 * 1024) / g.stage_bytes));
    FAV_CHECK_ARG(g.stages * g.stage_bytes <= 215 * 1024, "stem: stage of %d bytes does not fit", g.stage_bytes);
    L->smem_bytes = static_cast<size_t>(g.stages) * g.stage_bytes + 1024 + 512 + 4 * 32 * 5 * 16;
    auto classes = [](int in, int out, int k, int s, int pad, int* nlo, int* nhi) {
      *nlo = ceil_div(pad, s);
      const int pad_after = std::max(0, (out - 1) * s + k - pad - in);
      *nhi = ceil_div(pad_after, s);
    };
    classes(H, Ho, KH, 2, ph, &g.nlo_h, &g.nhi_h);
    classes(2 * Wo, Wo, 7, 2, ph, &g.nlo_w, &g.nhi_w);
    FAV_CHECK_ARG(g.nlo_h + g.nhi_h <= 3 && g.nlo_w + g.nhi_w <= 3, "stem: more than 4 border classes");
    const uint64_t pos_bytes = 8;
    const uint64_t row_pitch = static_cast<uint64_t>(Wp) * pos_bytes;
    const uint64_t frame_pitch = row_pitch * H;
    const uint64_t clip_pitch = frame_pitch * T;
    for (int p_t = 0; p_t < 2; ++p_t)
      for (int p_h = 0; p_h < 2; ++p_h) {
        if (g.rows[p_h] == 0 || (st == 1 && p_t == 1)) { L->tmA[p_t * 2 + p_h] = L->tmA[0]; continue; }
        // raw rows of one (T, H) parity: [Wp*4 elements][rows of this parity][frames of this parity][B][1]
        uint64_t dims[5] = {static_cast<uint64_t>(Wp) * 4, static_cast<uint64_t>((H - p_h + 1) / 2),
                            static_cast<uint64_t>(st == 2 ? (T - p_t + 1) / 2 : T), static_cast<uint64_t>(B), 1};
        uint64_t strides[4] = {2 * row_pitch, st * frame_pitch, clip_pitch, clip_pitch * B};
        uint32_t box[5] = {static_cast<uint32_t>(g.pitch / 2), static_cast<uint32_t>(g.rows[p_h]),
                           static_cast<uint32_t>(g.nf), 1, 1};
        const char* base = static_cast<const char*>(xpad) + (st == 2 ? p_t : 0) * frame_pitch + p_h * row_pitch;
        FAV_TRY(make_tmap_bf16(&L->tmA[p_t * 2 + p_h], base, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE));
      }
    if (g.rows[0] == 0) L->tmA[0] = L->tmA[1];
    uint64_t bd[3] = {32, static_cast<uint64_t>(bn), static_cast<uint64_t>(KT) * KH};
    uint64_t bs[2] = {64, static_cast<uint64_t>(bn) * 64};
    uint32_t bb[3] = {32, static_cast<uint32_t>(bn), static_cast<uint32_t>(KH)};
    FAV_TRY(make_tmap_bf16(&L->tmB, wpk, 3, bd, bs, bb, CU_TENSOR_MAP_SWIZZLE_64B));
    L->grid = std::max(1, std::min(g.m_tiles, sm_count(device)));
    return FAV_OK;
  }
  // M tiles per CTA tile: two accumulator stages of mt*bn columns must fit the 512 TMEM columns
  g.mt = std::max(1, std::min(2, 256 / bn));
  if (Ho <= 8) g.mt = 1;
  g.th = ceil_div(Ho, 8 * g.mt);
  g.tw = ceil_div(Wo, 16);
  g.m_tiles = B * To * g.th * g.tw;
  // per-parity slabs
  int qlo[2] = {1 << 20, 1 << 20}, qhi[2] = {-(1 << 20), -(1 << 20)};
  for (int kh = 0; kh < KH; ++kh) {
    const int offh = kh - ph;
    const int p = offh & 1;
    const int qh = (offh - p) / 2;
    qlo[p] = std::min(qlo[p], qh);
    qhi[p] = std::max(qhi[p], qh);
  }
  int off = 0;
  for (int p = 0; p < 2; ++p) {
    if (qhi[p] < qlo[p]) { g.qmin[p] = 0; g.rows[p] = 0; g.slab_off[p] = off; continue; }
    g.qmin[p] = qlo[p];
    g.rows[p] = 8 * g.mt + (qhi[p] - qlo[p]);
    g.slab_off[p] = off;
    off += g.rows[p] * 1024;
  }
  g.a_bytes = off;
  g.b_bytes = KH * bn * 64;
  g.stage_bytes = round_up(g.a_bytes + g.b_bytes, 1024);
  g.stages = std::max(2, std::min(kMaxStages, (215 * 1024) / g.stage_bytes));
  FAV_CHECK_ARG(g.stages * g.stage_bytes <= 215 * 1024, "stem: stage of %d bytes does not fit", g.stage_bytes);
  L->smem_bytes = static_cast<size_t>(g.stages) * g.stage_bytes + 1024 + 512 + 4 * 32 * 5 * 16;   // + epilogue staging
  // border classes of the delta-bias table: output rows whose window touches the zero padding
  auto classes = [](int in, int out, int k, int s, int pad, int* nlo, int* nhi) {
    *nlo = ceil_div(pad, s);
    const int pad_after = std::max(0, (out - 1) * s + k - pad - in);
    *nhi = ceil_div(pad_after, s);
  };
  classes(H, Ho, KH, 2, ph, &g.nlo_h, &g.nhi_h);
  {
    const int W = 2 * Wo;   // the engine requires even W; the W pad equals the H pad for square kernels
    classes(W, Wo, 7, 2, ph, &g.nlo_w, &g.nhi_w);
  }
  FAV_CHECK_ARG(g.nlo_h + g.nhi_h <= 3 && g.nlo_w + g.nhi_w <= 3, "stem: more than 4 border classes");

  const uint64_t pos_bytes = 8;                       // 4 channels bf16
  const uint64_t row_pitch = static_cast<uint64_t>(Wp) * pos_bytes;
  const uint64_t frame_pitch = row_pitch * H;
  const uint64_t clip_pitch = frame_pitch * T;
  for (int p_t = 0; p_t < 2; ++p_t) {
    for (int p_h = 0; p_h < 2; ++p_h) {
      if (g.rows[p_h] == 0 || (st == 1 && p_t == 1)) { L->tmA[p_t * 2 + p_h] = L->tmA[0]; continue; }
      uint64_t dims[5], strides[4];
      uint32_t box[5];
      dims[0] = 32;                                    // 8 W-positions x 4 channels, contiguous
      dims[1] = static_cast<uint64_t>(Wo);             // output column; window start moves 2 positions
      dims[2] = static_cast<uint64_t>((H - p_h + 1) / 2);
      dims[3] = static_cast<uint64_t>(st == 2 ? (T - p_t + 1) / 2 : T);
      dims[4] = static_cast<uint64_t>(B);
      strides[0] = 2 * pos_bytes;                      // 16 B: overlapping windows
      strides[1] = 2 * row_pitch;
      strides[2] = st * frame_pitch;
      strides[3] = clip_pitch;
      box[0] = 32; box[1] = 16; box[2] = static_cast<uint32_t>(g.rows[p_h]); box[3] = 1; box[4] = 1;
      const char* base = static_cast<const char*>(xpad) + (st == 2 ? p_t : 0) * frame_pitch + p_h * row_pitch;
      FAV_TRY(make_tmap_bf16(&L->tmA[p_t * 2 + p_h], base, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B));
    }
  }
  if (g.rows[0] == 0) L->tmA[0] = L->tmA[1];
  // weights [KT*KH taps][bn][32]: one box = the KH sub-tiles of one kt
  uint64_t bd[3] = {32, static_cast<uint64_t>(bn), static_cast<uint64_t>(KT) * KH};
  uint64_t bs[2] = {64, static_cast<uint64_t>(bn) * 64};
  uint32_t bb[3] = {32, static_cast<uint32_t>(bn), static_cast<uint32_t>(KH)};
  FAV_TRY(make_tmap_bf16(&L->tmB, wpk, 3, bd, bs, bb, CU_TENSOR_MAP_SWIZZLE_64B));
  L->grid = std::max(1, std::min(g.m_tiles, sm_count(device)));
  return FAV_OK;
}

int stem_launch(const StemLaunch& L, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    FAV_CUDA(cudaFuncSetAttribute(conv_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  ProfScope ps(PK_STEM, stream, L.flops);
  if (L.g.ts) {
    static bool attr_ts = false;
    if (!attr_ts) {
      FAV_CUDA(cudaFuncSetAttribute(conv_stem_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_ts = true;
    }
    static int prof = -1;
    if (prof < 0) prof = getenv("FAV_STEM_PROF") ? atoi(getenv("FAV_STEM_PROF")) : 0;
    if (prof) {   // debug: wait-cycle breakdown of the MMA warp (not capturable: synchronises the stream)
      StemGeom gp = L.g;
      gp.prof = prof | 1;   // 1: breakdown; +2: no epilogue work; +4: no MMAs (results are then wrong: timing only)
      unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}, r[8];
      cudaMemcpyToSymbol(g_halo_prof, z, sizeof(z));
      conv_stem_ts_kernel<<<L.grid, kThreadsTs, L.smem_bytes, stream>>>(L.tmA[0], L.tmA[1], L.tmA[2], L.tmA[3], L.tmB1, gp, L.e);
      cudaStreamSynchronize(stream);
      cudaMemcpyFromSymbol(r, g_halo_prof, sizeof(r));
      const double n = L.grid;
      fprintf(stderr, "[fav] stem ts prof=%d tiles=%d grid=%d nw=%d pitch=%d: per-CTA kclk total %.0f, MMA warp waits: tempty %.0f, a_full %.0f, "
              "w_full %.0f; epilogue warp waits tfull %.0f\n", gp.prof, L.g.m_tiles, L.grid, L.g.ts_nw, L.g.pitch, r[3] / n / 1e3, r[0] / n / 1e3,
              r[1] / n / 1e3, r[2] / n / 1e3, r[4] / n / 1e3);
      FAV_COUNT_LAUNCH();
      return FAV_OK;
    }
    FAV_CUDA(launch_pdl(conv_stem_ts_kernel, L.grid, kThreadsTs, L.smem_bytes, stream, L.tmA[0], L.tmA[1], L.tmA[2], L.tmA[3],
                        L.tmB1, L.g, L.e));
    FAV_COUNT_LAUNCH();
    FAV_CUDA(cudaGetLastError());
    return FAV_OK;
  }
  if (L.g.raw) {
    static bool attr_raw = false;
    if (!attr_raw) {
      FAV_CUDA(cudaFuncSetAttribute(conv_stem_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_raw = true;
    }
    FAV_CUDA(launch_pdl(conv_stem_raw_kernel, L.grid, kThreads, L.smem_bytes, stream, L.tmA[0], L.tmA[1], L.tmA[2], L.tmA[3],
                        L.tmB, L.g, L.e));
    FAV_COUNT_LAUNCH();
    FAV_CUDA(cudaGetLastError());
    return FAV_OK;
  }
  FAV_CUDA(launch_pdl(conv_stem_kernel, L.grid, kThreads, L.smem_bytes, stream, L.tmA[0], L.tmA[1], L.tmA[2], L.tmA[3], L.tmB,
                      L.g, L.e));
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

}  // namespace fav

// host-only view of the schedule for the CPU tests (include/fav.h)
extern "C" int fav_debug_stem_ts_schedule(int KT, int st, int bn, int* nfr, int* nslot, int* ktmax, uint32_t* tab) {
  if (!nfr || !nslot || !ktmax || !tab) {
    fav::set_error("fav_debug_stem_ts_schedule: null argument");
    return FAV_ERR_ARG;
  }
  fav::StemGeom g;
  memset(&g, 0, sizeof(g));
  g.tsG = 4;
  int a = 0, b = 0;
  const int s = fav::stem_ts_schedule(&g, KT, st, bn, &a, &b);
  if (s != FAV_OK) return s;
  for (int c = 0; c < 2; ++c) {
    nfr[c] = c < st ? g.ts_nfr[c] : 0;
    nslot[c] = c < st ? g.ts_nslot[c] : 0;
    ktmax[c] = c < st ? g.ts_ktmax[c] : 0;
    for (int f = 0; f < 8; ++f) tab[c * 8 + f] = c < st ? g.ts_tab[c][f] : 0u;
  }
  return g.tsG;
}
