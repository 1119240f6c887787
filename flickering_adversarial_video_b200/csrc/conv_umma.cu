// conv_umma.cu — tcgen05/TMEM/TMA implicit-GEMM Conv3d for sm_100a.
//
// Replaces the reference's snt.Conv3D + snt.BatchNorm + ReLU (i3d.py:61-70, cuDNN fp32 in the
// reference) and the backward-data convolutions autograd builds for
// compute_gradients(loss, var_list=perturbation) (i3d_adversarial_main_single_video_npy.py:82).
//
// One persistent CTA per SM, 6 warps:
//   warp 0    TMA producer  (one elected lane; A box + B box per k-block into a smem ring)
//   warp 1    TMEM allocator + MMA issuer (one lane issues tcgen05.mma, commits to mbarriers)
//   warps 2-5 epilogue (tcgen05.ld 32 lanes x 16 cols, bias/ReLU/mask, fp16 / bf16 NDHWC stores)
// Two 256-column fp32 accumulators in TMEM let the epilogue of tile i overlap the MMAs of tile i+1.
#include "conv_umma.cuh"

#include <string.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

namespace fav {

__device__ int g_fav_timeout_flag = 0;
// debug: cycles the halo MMA warp spent waiting [tempty, a_full, b_full, total, tiles] summed over CTAs (FAV_HALO_PROF=1)
__device__ unsigned long long g_halo_prof[8];

namespace {

constexpr int kThreads = 192;
constexpr int kTapThreads = 320;   // generic kernel: TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr int kAccCols = 256;   // TMEM columns per accumulator stage
constexpr int kMaxStages = 8;

struct TileCoord {
  int b, t0, h0, w0, n0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvGeom& g, int tile) {
  TileCoord c;
  int nt = tile % g.n_tiles;
  int m = tile / g.n_tiles;
  if (g.m_tiles == g.tw) {   // flat 1x1x1 tiling (B = T = H = 1): no further divisions (they cost the epilogue warps ~60 instructions a tile)
    c.b = 0; c.t0 = 0; c.h0 = 0;
    c.w0 = m * g.bw;
    c.n0 = nt * g.bn;
    return c;
  }
  int wi = m % g.tw;
  m /= g.tw;
  int hi = m % g.th;
  m /= g.th;
  int ti = m % g.tt;
  c.b = m / g.tt;
  c.w0 = wi * g.bw;
  c.h0 = hi * g.bh;
  c.t0 = ti * g.bt;
  c.n0 = nt * g.bn;
  return c;
}

__global__ void __launch_bounds__(kTapThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB, const ConvGeom g,
                 const ConvEpilogue e,
                 const int stages, const int a_bytes, const int b_bytes, const int stage_bytes, const int kg) {
  // `stages` ring entries of `kg` k-blocks each: ONE barrier hand-off per kg k-blocks.  FAV_TAP_PROF on the (3,1,1) and
  // strided convs of r2plus1d_18 (N = 64: 204 cycles of MMA per k-block) showed 870 cycles per k-block with one hand-off
  // each: ~450 on the producer lane (empty wait + expect_tx + two cp.async.bulk.tensor at ~150 cycles per issue) and
  // ~490 on the MMA-issuing thread (full wait + commit), and the tensor pipe only queues ~2 instructions behind the
  // running one, so it idles through every hand-off.  Now the 2*kg loads of a group are issued by 2*kg lanes in one
  // instruction slot, and the MMA warp waits / commits once per group.
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(stages) * kg * stage_bytes);
  uint64_t* full_bar = bars;                    // [stages]
  uint64_t* empty_bar = bars + kMaxStages;      // [stages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;  // [nacc <= 8]
  uint64_t* tempty_bar = tfull_bar + 8;         // [nacc <= 8]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 8);
  uint4* stage_all = reinterpret_cast<uint4*>(bars + 2 * kMaxStages + 24);   // 8 epilogue warps x 32 rows x 5 uint4
  // Accumulator ring: 512 TMEM columns hold nacc = 8 / 4 / 2 accumulators of 64 / 128 / 256 columns.  FAV_TAP_PROF showed
  // the narrow 1x1x1 launches bound by the LATENCY of one epilogue pass (Conv3d_2b: 55 of 95 kcycles waiting for a free
  // accumulator, ~2 kcycles per pass with all 8 warps on one tile), so with nacc >= 4 the two warp groups drain
  // ALTERNATE tiles (each warp all columns of its 32 rows) and two passes are in flight; with nacc == 2 both groups share
  // every tile as before (each drains half of the columns).
  const int nacc = g.acc_stages;
  const int acc_stride = 512 / nacc;
  const bool alt_groups = nacc >= 4;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = g.m_tiles * g.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB);
    if (g.nsrc > 1) { tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmA2); }
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < nacc; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], alt_groups ? 4 : 8);  // one arrival per epilogue warp that drains the stage
    }
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
    if (kg == 1) {
      // one k-block per hand-off (1x1x1 convs: memory-bound, few k-blocks per tile): a single lane issues both loads
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          const TileCoord tc = decode_tile(g, tile);
          for (int kb = 0; kb < g.nkb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
            uint8_t* sb = sa + a_bytes;
            mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(a_bytes + b_bytes));
            if (g.nsrc > 1) {
              // 1x1x1 with K concatenated from several tensors: k-block -> (source, local block)
              int src = 0, cb = kb;
              if (cb >= g.src_blocks[0]) { cb -= g.src_blocks[0]; src = 1; }
              if (src == 1 && cb >= g.src_blocks[1]) { cb -= g.src_blocks[1]; src = 2; }
              const CUtensorMap* tm = src == 0 ? &tmA0 : (src == 1 ? &tmA1 : &tmA2);
              tma_load_5d(sa, tm, &full_bar[stage], cb * 64, tc.w0, tc.h0, tc.t0, tc.b);
              tma_load_2d(sb, &tmB, &full_bar[stage], kb * 64, tc.n0);
            } else {
              const int tap = kb / g.cblocks;
              const int cb = kb - tap * g.cblocks;
              const int dw = tap % g.kw;
              const int dh = (tap / g.kw) % g.kh;
              const int dt = tap / (g.kw * g.kh);
              tma_load_5d(sa, &tmA0, &full_bar[stage], cb * 64, tc.w0 * g.sw + dw + g.ow, tc.h0 * g.sh + dh + g.oh,
                          tc.t0 * g.st + dt + g.ot, tc.b);
              tma_load_2d(sb, &tmB, &full_bar[stage], kb * 64, tc.n0);
            }
            if (++stage == stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    } else {
      // the whole warp walks the ring; lane 0 polls, lanes 0 .. 2*nk-1 issue the group's loads (even: A box, odd: B tile)
      {
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          const TileCoord tc = decode_tile(g, tile);
          for (int kb0 = 0; kb0 < g.nkb; kb0 += kg) {
            const int nk = min(kg, g.nkb - kb0);
            if (lane == 0) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(nk * (a_bytes + b_bytes)));
            }
            __syncwarp();
            if (lane < 2 * nk) {
              const int i = lane >> 1;
              const int kb = kb0 + i;
              uint8_t* sa = smem + static_cast<size_t>(stage * kg + i) * stage_bytes;
              if (lane & 1) {
                tma_load_2d(sa + a_bytes, &tmB, &full_bar[stage], kb * 64, tc.n0);
              } else if (g.nsrc > 1) {
                // 1x1x1 with K concatenated from several tensors: k-block -> (source, local block)
                int src = 0, cb = kb;
                if (cb >= g.src_blocks[0]) { cb -= g.src_blocks[0]; src = 1; }
                if (src == 1 && cb >= g.src_blocks[1]) { cb -= g.src_blocks[1]; src = 2; }
                const CUtensorMap* tm = src == 0 ? &tmA0 : (src == 1 ? &tmA1 : &tmA2);
                tma_load_5d(sa, tm, &full_bar[stage], cb * 64, tc.w0, tc.h0, tc.t0, tc.b);
              } else {
                const int tap = kb / g.cblocks;
                const int cb = kb - tap * g.cblocks;
                const int dw = tap % g.kw;
                const int dh = (tap / g.kw) % g.kh;
                const int dt = tap / (g.kw * g.kh);
                tma_load_5d(sa, &tmA0, &full_bar[stage], cb * 64, tc.w0 * g.sw + dw + g.ow, tc.h0 * g.sh + dh + g.oh,
                            tc.t0 * g.st + dt + g.ot, tc.b);
              }
            }
            if (++stage == stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the pipeline in lock-step (warp-uniform state stays in uniform registers);
    // one elected lane issues the tcgen05 instructions.  The issue loop is the critical resource for
    // small tiles: every instruction in it costs tensor-pipe idle time.
    const uint32_t idesc = umma_idesc(128, g.bn, g.f16 != 0);
    const uint32_t desc_hi = umma_desc_hi(128);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long w_te = 0, w_f = 0, c0 = 0, ntile = 0;
    const long long t_start = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      c0 = g.prof ? clock64() : 0;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      if (g.prof) { w_te += clock64() - c0; ++ntile; }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * acc_stride);
      uint32_t accum = 0;
      int cb = 0;
      if (kg == 1) {
        for (int kb = 0; kb < g.nkb; ++kb) {
          c0 = g.prof ? clock64() : 0;
          mbar_wait(&full_bar[stage], phase);
          if (g.prof) w_f += clock64() - c0;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t sb = sa + static_cast<uint32_t>(a_bytes);
          int ksteps = min(4, (g.cin - cb * 64) >> 4);
          if (g.nsrc > 1) {
            int src = 0, lb = kb;
            if (lb >= g.src_blocks[0]) { lb -= g.src_blocks[0]; src = 1; }
            if (src == 1 && lb >= g.src_blocks[1]) { lb -= g.src_blocks[1]; src = 2; }
            const int cs_ = src == 0 ? g.src_cin[0] : (src == 1 ? g.src_cin[1] : g.src_cin[2]);
            ksteps = min(4, (cs_ - lb * 64) >> 4);
          }
          if (elect_one()) {
            const uint32_t a_lo = umma_desc_lo(sa);
            const uint32_t b_lo = umma_desc_lo(sb);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k < ksteps) {
                umma_bf16(d_tmem, make_desc(desc_hi, a_lo + 2 * k), make_desc(desc_hi, b_lo + 2 * k), idesc, accum);
                accum = 1;
              }
            }
            umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            if (kb == g.nkb - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++cb == g.cblocks) cb = 0;
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      } else {
        for (int kb0 = 0; kb0 < g.nkb; kb0 += kg) {
          const int nk = min(kg, g.nkb - kb0);
          c0 = g.prof ? clock64() : 0;
          mbar_wait(&full_bar[stage], phase);
          if (g.prof) w_f += clock64() - c0;
          tc_fence_after();
          const uint32_t sa0 = smem_u32(smem + static_cast<size_t>(stage * kg) * stage_bytes);
          if (elect_one()) {
            int cbl = cb;
            for (int i = 0; i < nk; ++i) {
              const int kb = kb0 + i;
              int ksteps = min(4, (g.cin - cbl * 64) >> 4);
              if (g.nsrc > 1) {
                int src = 0, lb = kb;
                if (lb >= g.src_blocks[0]) { lb -= g.src_blocks[0]; src = 1; }
                if (src == 1 && lb >= g.src_blocks[1]) { lb -= g.src_blocks[1]; src = 2; }
                const int cs_ = src == 0 ? g.src_cin[0] : (src == 1 ? g.src_cin[1] : g.src_cin[2]);
                ksteps = min(4, (cs_ - lb * 64) >> 4);
              }
              const uint32_t sa = sa0 + static_cast<uint32_t>(i * stage_bytes);
              const uint32_t a_lo = umma_desc_lo(sa);
              const uint32_t b_lo = umma_desc_lo(sa + static_cast<uint32_t>(a_bytes));
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < ksteps) {
                  umma_bf16(d_tmem, make_desc(desc_hi, a_lo + 2 * k), make_desc(desc_hi, b_lo + 2 * k), idesc, accum);
                  accum = 1;
                }
              }
              if (++cbl == g.cblocks) cbl = 0;
            }
            umma_commit(&empty_bar[stage]);  // frees the group's smem slots when these MMAs retire
            if (kb0 + nk == g.nkb) umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
          }
          __syncwarp();
          cb += nk;
          while (cb >= g.cblocks) cb -= g.cblocks;
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (++acc == nacc) { acc = 0; acc_phase ^= 1; }
    }
    if (g.prof && lane == 0) {
      atomicAdd(&g_halo_prof[0], static_cast<unsigned long long>(w_te));
      atomicAdd(&g_halo_prof[1], static_cast<unsigned long long>(w_f));
      atomicAdd(&g_halo_prof[3], static_cast<unsigned long long>(clock64() - t_start));
      atomicAdd(&g_halo_prof[4], static_cast<unsigned long long>(ntile));
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // two warps per TMEM lane quarter, each draining half of the tile's columns: the epilogue is a chain of TMEM /
    // global-memory latencies and bounds the 1x1x1 launches (FAV_TAP_PROF), so it gets the spare warps
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int rw = row % g.bw;
    const int rh = (row / g.bw) % g.bh;
    const int rt = row / (g.bw * g.bh);
    int acc = -1;
    uint32_t acc_phase = 1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      if (++acc == nacc) acc = 0;
      if (acc == 0) acc_phase ^= 1;
      if (alt_groups && (acc & 1) != chalf) continue;   // the other warp group's tile (nacc is even: a stage keeps its group)
      const TileCoord tc = decode_tile(g, tile);
      const int w = tc.w0 + rw, h = tc.h0 + rh, t = tc.t0 + rt;
      const bool valid = (w < g.W) && (h < g.H) && (t < g.T);
      // output position (identity for plain convs; strided scatter for the parity classes of a strided dgrad)
      const long long pos = ((static_cast<long long>(tc.b) * g.oT + (t * g.est + g.eot)) * g.oH + (h * g.esh + g.eoh)) * g.oW +
                            (w * g.esw + g.eow);
      h16* out_row = e.out + pos * e.out_cs + e.out_coff;
      const h16* mask_row = e.mask ? e.mask + pos * e.mask_cs + e.mask_coff : nullptr;
      const h16* add_row = e.addend ? e.addend + pos * e.add_cs + e.add_coff : nullptr;
      const float* bias_row = e.bias;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(acc * acc_stride);
      // this warp's column range of the tile (multiples of 16): everything, or its half when both groups share a tile
      const int half_cols = ((g.bn >> 4) + 1) / 2 * 16;
      const int c_lo = alt_groups ? 0 : chalf * half_cols, c_hi = alt_groups ? g.bn : min(g.bn, c_lo + half_cols);
      if (e.nseg > 1) {
        // column segments of this N tile go to different tensors (fused same-input 1x1x1 convs)
        const int n_lo = tc.n0 + c_lo, n_hi = tc.n0 + c_hi;
#pragma unroll
        for (int sgi = 0; sgi < 3; ++sgi) {
          if (sgi >= e.nseg) break;
          const int a = max(n_lo, e.seg_n0[sgi]), b = min(n_hi, e.seg_n0[sgi + 1]);
          if (a >= b) continue;   // warp-uniform
          h16* srow = e.seg_out[sgi] + pos * e.seg_cs[sgi] + e.seg_coff[sgi];
          epilogue_columns_staged(e, b - a, a - e.seg_n0[sgi], taddr + static_cast<uint32_t>(a - tc.n0), valid, srow,
                                  bias_row ? bias_row + e.seg_n0[sgi] : nullptr, e.seg_n0[sgi + 1] - e.seg_n0[sgi],
                                  stage_all + (warp - 2) * 160, lane);
        }
      } else if (c_lo < c_hi) {
        if (add_row == nullptr)
          epilogue_columns_staged(e, c_hi - c_lo, tc.n0 + c_lo, taddr + static_cast<uint32_t>(c_lo), valid, out_row, bias_row,
                                  e.cout_store, stage_all + (warp - 2) * 160, lane, mask_row);
        else
          epilogue_columns<2>(e, c_hi - c_lo, tc.n0 + c_lo, taddr + static_cast<uint32_t>(c_lo), valid, out_row, mask_row,
                              add_row, bias_row);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ---------------------------------------------------------------------------------------------
// Halo kernel: 3x3x3, stride 1, SAME.  The per-tap kernel above streams a fresh 16 KB A box through
// L2 for each of the 27 taps; here one TMA box per (64-channel block, dt) brings (nrows+2) x (W+2)
// zero-padded positions into shared memory once and the 9 in-plane taps are row windows of it.
// ---------------------------------------------------------------------------------------------
struct HaloTile {
  int b, t, h0, n0;
};
__device__ __forceinline__ HaloTile decode_halo_tile(const ConvGeom& g, int tile) {
  HaloTile c;
  const int nt = tile % g.n_tiles;
  int m = tile / g.n_tiles;
  const int hi = m % g.th;
  m /= g.th;
  c.t = m % g.T;
  c.b = m / g.T;
  c.h0 = hi * g.nrows * g.mt;
  c.n0 = nt * g.bn;
  return c;
}

constexpr int kHaloStageBytes = 4 * 32 * 5 * 16;   // epilogue staging: 4 warps x 32 rows x 5 uint4
constexpr int kHaloThreads = 224;   // + warp 6: weight-tile producer (A slabs and B tiles must not block each other)

__global__ void __launch_bounds__(kHaloThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvGeom g, const ConvEpilogue e, const int b_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + static_cast<size_t>(g.na) * g.slab_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + static_cast<size_t>(g.nb) * g.bgroup * b_bytes);
  uint64_t* a_full = bars;            // [4]
  uint64_t* a_empty = bars + 4;       // [4]
  uint64_t* b_full = bars + 8;        // [8]
  uint64_t* b_empty = bars + 16;      // [8]
  uint64_t* tfull_bar = bars + 24;    // [2]
  uint64_t* tempty_bar = bars + 26;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  uint4* stage_all = reinterpret_cast<uint4*>(bars + 32);   // 4 epilogue warps x 32 rows x 5 uint4 (coalesced stores)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = g.m_tiles * g.n_tiles;
  const int slabs_per_tile = g.cblocks * g.kt;   // kt = 3 (3x3x3) or 1 ((1,3,3) convs)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < g.na; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < g.nb; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
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
    // ===================== TMA producer: activation slabs =====================
    if (lane == 0) {
      int sa = 0;
      uint32_t pa = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const HaloTile tc = decode_halo_tile(g, tile);
        for (int sidx = 0; sidx < slabs_per_tile; ++sidx) {
          const int cb = sidx / g.kt;
          const int dt = sidx - cb * g.kt;
          mbar_wait(&a_empty[sa], pa ^ 1);
          mbar_expect_tx(&a_full[sa], static_cast<uint32_t>(g.slab_tx));
          tma_load_5d(smem_a + static_cast<size_t>(sa) * g.slab_bytes, &tmA, &a_full[sa], cb * 64, -1, tc.h0 - 1,
                      tc.t + dt + g.ot, tc.b);
          if (++sa == g.na) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    // ===================== TMA producer: weight tiles =====================
    // the whole warp walks the ring; lanes 0..bgroup-1 issue the group's tile loads in one instruction slot (a single
    // lane pays ~100-150 cycles per cp.async.bulk.tensor it issues: tools/ubench/tma_rate.cu)
    {
      int sb = 0;
      uint32_t pb = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = (tile % g.n_tiles) * g.bn;
        for (int sidx = 0; sidx < slabs_per_tile; ++sidx) {
          const int cb = sidx / g.kt;
          const int dt = sidx - cb * g.kt;
          for (int j = 0; j < 9; j += g.bgroup) {
            if (lane == 0) {   // one lane polls (32 spinning lanes would compete with the MMA warp for issue slots)
              mbar_wait(&b_empty[sb], pb ^ 1);
              mbar_expect_tx(&b_full[sb], static_cast<uint32_t>(b_bytes * g.bgroup));
            }
            __syncwarp();
            if (lane < g.bgroup) {
              const int tap = dt * 9 + j + lane;
              tma_load_2d(smem_b + (static_cast<size_t>(sb) * g.bgroup + lane) * b_bytes, &tmB, &b_full[sb],
                          (tap * g.cblocks + cb) * 64, n0);
            }
            if (++sb == g.nb) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, elected lane issues) =====================
    const uint32_t idesc = umma_idesc(128, g.bn, g.f16 != 0);
    const uint32_t desc_hi = umma_desc_hi(128);
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t wp16 = static_cast<uint32_t>(g.Wp) * 8u;   // one padded row in 16-byte units
    const uint32_t tile16 = wp16 * static_cast<uint32_t>(g.nrows);
    const int acc_cols = g.acc_stages == 2 ? kAccCols : 0;
    long long w_te = 0, w_a = 0, w_b = 0, n_tiles_done = 0;
    const long long t_start = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      long long c0 = g.prof ? clock64() : 0;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      if (g.prof) { w_te += clock64() - c0; ++n_tiles_done; }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * acc_cols);
      uint32_t accum = 0;
      int cb = 0, dt = 0;
      for (int sidx = 0; sidx < slabs_per_tile; ++sidx) {
        const int ksteps = min(4, (g.cin - cb * 64) >> 4);
        c0 = g.prof ? clock64() : 0;
        mbar_wait(&a_full[sa], pa);
        if (g.prof) w_a += clock64() - c0;
        const uint32_t slab_lo = umma_desc_lo(smem_u32(smem_a + static_cast<size_t>(sa) * g.slab_bytes));
        for (int j0 = 0; j0 < 9; j0 += g.bgroup) {   // weight groups of 1, 3 (one dh row) or 9 (the whole slab) taps
          c0 = g.prof ? clock64() : 0;
          mbar_wait(&b_full[sb], pb);
          if (g.prof) w_b += clock64() - c0;
          tc_fence_after();
          const uint32_t b_grp = umma_desc_lo(smem_u32(smem_b + static_cast<size_t>(sb) * g.bgroup * b_bytes));
          if (elect_one()) {
            for (int u = 0; u < g.bgroup; ++u) {
              const int tap = j0 + u;
              const int dh = (tap >= 3) + (tap >= 6);
              const int dw = tap - 3 * dh;
              const uint32_t b_lo = b_grp + static_cast<uint32_t>(u) * static_cast<uint32_t>(b_bytes >> 4);
              // row window of tap (dh, dw): dh padded rows down, dw positions right (one position = 128 B = 8 x 16 B)
              uint32_t a_i = slab_lo + static_cast<uint32_t>(dh) * wp16 + 8u * static_cast<uint32_t>(dw);
              uint32_t d_i = d_tmem;
              for (int i = 0; i < g.mt; ++i) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (k < ksteps) umma_bf16(d_i, make_desc(desc_hi, a_i + 2 * k), make_desc(desc_hi, b_lo + 2 * k), idesc, accum | (k > 0 ? 1u : 0u));
                }
                a_i += tile16;                       // next M tile: nrows padded rows further down the slab
                d_i += static_cast<uint32_t>(g.bn);
              }
              accum = 1;
            }
            umma_commit(&b_empty[sb]);
            if (j0 + g.bgroup == 9) {
              umma_commit(&a_empty[sa]);
              if (sidx == slabs_per_tile - 1) umma_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++sb == g.nb) { sb = 0; pb ^= 1; }
        }
        if (++sa == g.na) { sa = 0; pa ^= 1; }
        if (++dt == g.kt) { dt = 0; ++cb; }
      }
      if (g.acc_stages == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
    }
    if (g.prof && lane == 0) {
      atomicAdd(&g_halo_prof[0], static_cast<unsigned long long>(w_te));
      atomicAdd(&g_halo_prof[1], static_cast<unsigned long long>(w_a));
      atomicAdd(&g_halo_prof[2], static_cast<unsigned long long>(w_b));
      atomicAdd(&g_halo_prof[3], static_cast<unsigned long long>(clock64() - t_start));
      atomicAdd(&g_halo_prof[4], static_cast<unsigned long long>(n_tiles_done));
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int hm = row / g.Wp;
    const int wm = row - hm * g.Wp;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const HaloTile tc = decode_halo_tile(g, tile);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int acc_cols = g.acc_stages == 2 ? kAccCols : 0;
      for (int i = 0; i < g.mt; ++i) {
        const int h = tc.h0 + i * g.nrows + hm;
        const bool valid = (wm < g.W) && (hm < g.nrows) && (h < g.H);
        const long long pos = valid ? ((static_cast<long long>(tc.b) * g.T + tc.t) * g.H + h) * g.W + wm : 0;
        h16* out_row = e.out + pos * e.out_cs + e.out_coff;
        const h16* mask_row = e.mask ? e.mask + pos * e.mask_cs + e.mask_coff : nullptr;
        const h16* add_row = e.addend ? e.addend + pos * e.add_cs + e.add_coff : nullptr;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc * acc_cols + i * g.bn);
        // ncu (profiles/r01_ncu_full_conv_halo_dgrad_2c.txt): per-lane 16-byte row stores used 16 of 32 bytes per
        // sector and kept L1/TEX 78 % busy; the staged form stores 8 rows x 64 contiguous bytes per instruction
        if (add_row == nullptr)
          epilogue_columns_staged(e, g.bn, tc.n0, taddr, valid, out_row, e.bias, e.cout_store, stage_all + (warp - 2) * 160,
                                  lane, mask_row);
        else
          epilogue_columns(e, g.bn, tc.n0, taddr, valid, out_row, mask_row, add_row, e.bias);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (g.acc_stages == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
uint16_t f32_to_bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);  // NaN
  const uint32_t rounding = 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>((u + rounding) >> 16);
}
float bf16_bits_to_f32(uint16_t b) {
  uint32_t u = static_cast<uint32_t>(b) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

uint16_t f32_to_f16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint16_t sign = static_cast<uint16_t>((u >> 16) & 0x8000u);
  const uint32_t a = u & 0x7fffffffu;
  if (a > 0x7f800000u) return static_cast<uint16_t>(sign | 0x7e00u);          // NaN
  if (a >= 0x477ff000u) return static_cast<uint16_t>(sign | 0x7bffu);         // >= 65520 rounds past the largest finite: saturate
  if (a < 0x33000000u) return sign;                                            // < 2^-25: rounds to zero
  const int e = static_cast<int>(a >> 23) - 127;                               // unbiased exponent
  uint32_t mant = (a & 0x7fffffu) | 0x800000u;                                 // 24-bit significand
  int shift;                                                                    // bits dropped from the significand
  uint32_t base;
  if (e >= -14) { shift = 13; base = static_cast<uint32_t>(e + 15) << 10; mant &= 0x7fffffu; }
  else { shift = 13 + (-14 - e); base = 0; }                                   // subnormal: keep the leading 1
  const uint32_t kept = mant >> shift;
  const uint32_t rem = mant & ((1u << shift) - 1u);
  const uint32_t half = 1u << (shift - 1);
  uint32_t r = base + kept;                                                    // a carry out of the mantissa bumps the exponent
  if (rem > half || (rem == half && (kept & 1u))) r += 1;
  return static_cast<uint16_t>(sign | r);
}

void choose_box(int T, int H, int W, int kt, int kh, int kw, int* bw, int* bh, int* bt) {
  (void)kt; (void)kh; (void)kw;
  double best = -1.0;
  int best_bw = 128, best_bh = 1, best_bt = 1;
  for (int a = 1; a <= 128; a *= 2) {
    for (int b = 1; a * b <= 128; b *= 2) {
      const int c = 128 / (a * b);
      if (a > 256 || b > 256 || c > 256) continue;
      const double eff = (static_cast<double>(W) / (ceil_div(W, a) * a)) *
                         (static_cast<double>(H) / (ceil_div(H, b) * b)) *
                         (static_cast<double>(T) / (ceil_div(T, c) * c));
      // prefer wide-in-W boxes on ties (longer contiguous runs per TMA box row group)
      const double score = eff + 1e-6 * a + 1e-7 * b;
      if (score > best) {
        best = score;
        best_bw = a;
        best_bh = b;
        best_bt = c;
      }
    }
  }
  *bw = best_bw;
  *bh = best_bh;
  *bt = best_bt;
}

static void finish_plan(ConvLaunch* L, int device) {
  ConvGeom& g = L->g;
  g.tw = ceil_div(g.W, g.bw);
  g.th = ceil_div(g.H, g.bh);
  g.tt = ceil_div(g.T, g.bt);
  g.m_tiles = g.B * g.tt * g.th * g.tw;
  L->a_bytes = 128 * 128;
  L->b_bytes = g.bn * 128;
  L->stage_bytes = round_up(L->a_bytes + L->b_bytes, 1024);
  // ring: `stages` groups of `kg` k-blocks (one barrier hand-off per group, see conv_umma_kernel)
  const int budget = 200 * 1024;
  const int slots = std::max(2, std::min(16, budget / L->stage_bytes));
  {
    // ~800 cycles of MMA per group hide a hand-off; at least three groups in the ring keep the loads ahead of the MMAs
    const int mma_clk = 4 * std::max(g.bn / 2, 32 + g.bn / 4);
    int kg = std::max(1, std::min(4, ceil_div(800, mma_clk)));
    if (const char* ev = getenv("FAV_TAP_KG")) kg = std::max(1, std::min(8, atoi(ev)));
    if (g.kt * g.kh * g.kw == 1 && !getenv("FAV_TAP_KG")) kg = 1;   // 1x1x1: memory-bound, measured slightly slower grouped (I3D conv_tap 0.98 -> 1.02 ms)
    kg = std::min(kg, std::max(1, g.nkb));
    while (kg > 1 && slots / kg < 3) --kg;
    L->kg = kg;
  }
  L->stages = std::max(2, std::min(kMaxStages, slots / L->kg));
  L->smem_bytes = static_cast<size_t>(L->stages) * L->kg * L->stage_bytes + 1024 + 512 + 8 * 32 * 5 * 16;   // + epilogue staging
  const int tiles = g.m_tiles * g.n_tiles;
  L->grid = std::max(1, std::min(tiles, sm_count(device)));
  // accumulator ring of the per-tap kernel (conv_umma_kernel): 8 x 64, 4 x 128 or 2 x 256 TMEM columns
  g.acc_stages = g.bn <= 64 ? 8 : (g.bn <= 128 ? 4 : 2);
  if (const char* ev = getenv("FAV_TAP_ACC")) {   // A/B: FAV_TAP_ACC=2 restores the double-buffered pair
    const int v = atoi(ev);
    if ((v == 2 || v == 4 || v == 8) && g.bn * v <= 512) g.acc_stages = v;
  }
}

int conv_plan_ex(ConvLaunch* L, int device, const ConvSpec& sp) {
  FAV_CHECK_ARG(sp.cin % 16 == 0 && sp.cin > 0, "conv: cin=%d must be a positive multiple of 16", sp.cin);
  FAV_CHECK_ARG(sp.x_cs % 8 == 0 && sp.x_coff % 8 == 0, "conv: channel stride/offset must be multiples of 8");
  FAV_CHECK_ARG(sp.cout_pad % 16 == 0 && sp.cout_pad > 0, "conv: padded cout=%d must be a multiple of 16", sp.cout_pad);
  FAV_CHECK_ARG(sp.st >= 1 && sp.sh >= 1 && sp.sw >= 1 && sp.st <= 8 && sp.sh <= 8 && sp.sw <= 8, "conv: bad strides");
  memset(L, 0, sizeof(*L));
  ConvGeom& g = L->g;
  g.kt = sp.kt; g.kh = sp.kh; g.kw = sp.kw;
  g.ot = sp.ot; g.oh = sp.oh; g.ow = sp.ow;
  g.st = sp.st; g.sh = sp.sh; g.sw = sp.sw;
  g.cin = sp.cin;
  g.cblocks = ceil_div(sp.cin, 64);
  g.nkb = sp.kt * sp.kh * sp.kw * g.cblocks;
  g.n_tiles = ceil_div(sp.cout_pad, 256);
  g.bn = round_up(ceil_div(sp.cout_pad, g.n_tiles), 16);
  {
    // few M tiles (7x7 planes): split N further so that the persistent grid covers the machine; every CTA then
    // streams 1/n of the weights and the per-SM L2 ingest (the bound of these launches) drops with it
    const long long m_tiles_est = ceil_div64(static_cast<long long>(sp.B) * sp.T * sp.H * sp.W, 128);
    const int sms = sm_count(device);
    while (m_tiles_est * g.n_tiles * 2 <= sms && g.bn > 64) {
      g.n_tiles += 1;
      g.bn = round_up(ceil_div(sp.cout_pad, g.n_tiles), 16);
    }
    while (g.bn * (g.n_tiles - 1) >= sp.cout_pad) g.n_tiles -= 1;
  }
  // the last N tile may be partial: weight rows past cout_pad are TMA out-of-bounds zeros and the epilogue
  // stores only channels below cout_store
  g.B = sp.B; g.T = sp.T; g.H = sp.H; g.W = sp.W;
  g.oT = sp.oT; g.oH = sp.oH; g.oW = sp.oW;
  g.est = sp.est; g.esh = sp.esh; g.esw = sp.esw;
  g.eot = sp.eot; g.eoh = sp.eoh; g.eow = sp.eow;
  choose_box(sp.T, sp.H, sp.W, sp.kt, sp.kh, sp.kw, &g.bw, &g.bh, &g.bt);
  // element-strided boxes traverse bw*sw elements (<= 256)
  while (g.bw * sp.sw > 256) { g.bw /= 2; g.bh *= 2; }
  while (g.bh * sp.sh > 256) { g.bh /= 2; g.bt *= 2; }
  FAV_CHECK_ARG(g.bt * sp.st <= 256, "conv: strided box too large");
  uint64_t dims[5], strides[4];
  uint32_t box[5], estr[5];
  const char* base = static_cast<const char*>(sp.x) + static_cast<long long>(sp.x_coff) * 2;
  dims[0] = static_cast<uint64_t>(sp.cin);
  dims[1] = static_cast<uint64_t>(sp.aW);
  dims[2] = static_cast<uint64_t>(sp.aH);
  dims[3] = static_cast<uint64_t>(sp.aT);
  dims[4] = static_cast<uint64_t>(sp.B);
  strides[0] = static_cast<uint64_t>(sp.x_cs) * 2;
  strides[1] = strides[0] * sp.aW;
  strides[2] = strides[1] * sp.aH;
  strides[3] = strides[2] * sp.aT;
  box[0] = 64; box[1] = g.bw * sp.sw; box[2] = g.bh * sp.sh; box[3] = g.bt * sp.st; box[4] = 1;
  estr[0] = 1; estr[1] = sp.sw; estr[2] = sp.sh; estr[3] = sp.st; estr[4] = 1;
  FAV_TRY(make_tmap_bf16(&L->tmA[0], base, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, estr));
  L->tmA[1] = L->tmA[0]; L->tmA[2] = L->tmA[0]; L->tmA[3] = L->tmA[0];
  uint64_t bd[2] = {static_cast<uint64_t>(g.nkb) * 64, static_cast<uint64_t>(sp.cout_pad)};
  uint64_t bs[1] = {bd[0] * 2};
  uint32_t bb[2] = {64, static_cast<uint32_t>(g.bn)};
  FAV_TRY(make_tmap_bf16(&L->tmB, sp.wpk, 2, bd, bs, bb, CU_TENSOR_MAP_SWIZZLE_128B));
  finish_plan(L, device);
  return FAV_OK;
}

int conv_plan_generic(ConvLaunch* L, int device, const void* x, long long x_cs, int x_coff, int cin,
                      const void* wpk, int cout_pad, int B, int T, int H, int W, int kt, int kh,
                      int kw, int flat) {
  ConvSpec sp{};
  sp.x = x; sp.x_cs = x_cs; sp.x_coff = x_coff; sp.cin = cin;
  sp.wpk = wpk; sp.cout_pad = cout_pad;
  sp.kt = kt; sp.kh = kh; sp.kw = kw;
  sp.ot = -((kt - 1) / 2); sp.oh = -((kh - 1) / 2); sp.ow = -((kw - 1) / 2);
  sp.st = sp.sh = sp.sw = 1;
  sp.est = sp.esh = sp.esw = 1;
  sp.eot = sp.eoh = sp.eow = 0;
  if (flat) {
    FAV_CHECK_ARG(kt == 1 && kh == 1 && kw == 1, "conv: flat tiling needs a 1x1x1 kernel");
    const long long M = static_cast<long long>(B) * T * H * W;
    FAV_CHECK_ARG(M < (1ll << 31), "conv: too many positions");
    sp.B = 1; sp.T = 1; sp.H = 1; sp.W = static_cast<int>(M);
  } else {
    sp.B = B; sp.T = T; sp.H = H; sp.W = W;
  }
  sp.aT = sp.T; sp.aH = sp.H; sp.aW = sp.W;
  sp.oT = sp.T; sp.oH = sp.H; sp.oW = sp.W;
  return conv_plan_ex(L, device, sp);
}


bool conv_halo_applicable(int T, int H, int W, int kt, int kh, int kw) {
  (void)T;
  if ((kt != 3 && kt != 1) || kh != 3 || kw != 3) return false;
  // (1,3,3): only 9 taps share a slab, which pays on the large planes only (measured on mc3_18 / r2plus1d_18)
  if (kt == 1 && H * W < 28 * 28) return false;
  const int Wp = W + 2;
  if (Wp > 64) return false;                 // at least two output rows per 128-row tile
  const int nrows = std::min(128 / Wp, H);
  // tiny planes (7x7): the launch is a chain of TMA / barrier latencies, not throughput: one slab + three 3-tap
  // weight groups per (channel block, dt) is 3x fewer round trips than 27 per-tap stages (FAV_HALO_SMALL=0 disables)
  static int small_ok = -1;
  if (small_ok < 0) {
    const char* ev = getenv("FAV_HALO_SMALL");
    small_ok = (ev && !atoi(ev)) ? 0 : 1;
  }
  if (small_ok && H * W <= 64) return true;
  // useful rows per tile vs the per-tap path's box efficiency (~0.9): keep halo when >= 60 %
  return nrows * W * 10 >= 128 * 6;
}

int conv_plan_halo(ConvLaunch* L, int device, const void* x, long long x_cs, int x_coff, int cin,
                   const void* wpk, int cout_pad, int B, int T, int H, int W, int kt) {
  FAV_CHECK_ARG(kt == 1 || kt == 3, "conv halo: temporal taps %d", kt);
  FAV_CHECK_ARG(cin % 16 == 0 && cin > 0, "conv: cin=%d must be a positive multiple of 16", cin);
  FAV_CHECK_ARG(cout_pad % 16 == 0 && cout_pad > 0, "conv: padded cout=%d must be a multiple of 16", cout_pad);
  memset(L, 0, sizeof(*L));
  ConvGeom& g = L->g;
  g.halo = 1;
  g.kt = kt; g.kh = g.kw = 3;
  g.ot = -(kt / 2); g.oh = g.ow = -1;
  g.cin = cin;
  g.cblocks = ceil_div(cin, 64);
  g.nkb = 9 * kt * g.cblocks;
  g.B = B; g.T = T; g.H = H; g.W = W;
  g.Wp = W + 2;
  g.nrows = std::min(128 / g.Wp, H);
  FAV_CHECK_ARG(g.nrows >= 1, "conv halo: W=%d too wide", W);
  g.swz_base_offset = 0;   // tcgen05 applies the 128-byte swizzle on absolute smem addresses (measured)
  // ---- choose the N split and the number of M tiles that share every weight tile ----
  // Cost model from measurements on B200 (tools/ubench/, FAV_HALO_PROF, profiles/): shared-memory bandwidth
  // (operand reads + TMA refills) and L2 -> SM bandwidth (41 B/cycle/SM in a pure-load microbenchmark; the planner uses 70, which fits whole-step timings) bound these
  // kernels, not the tensor pipe.
  {
    const int groups = ceil_div(H, g.nrows);
    const int budget = 212 * 1024;   // + 10 KB of epilogue staging + barriers = the 227 KB a CTA may have
    const int sms = sm_count(device);
    static int force_mt = -1, force_nt = -1;
    static double l2_rate = 70.0;   // B per cycle per SM; measured on whole steps (41 -> 70: 7.44 -> 7.35 ms/step)
    if (force_mt < 0) {
      if (const char* lr = getenv("FAV_HALO_L2RATE")) l2_rate = atof(lr);
      const char* ev = getenv("FAV_HALO_MT");
      force_mt = ev ? atoi(ev) : 0;
      ev = getenv("FAV_HALO_NT");
      force_nt = ev ? atoi(ev) : 0;
    }
    double best = 1e30;
    int best_nt = 0, best_mt = 0, best_na = 0, best_nb = 0, best_bg = 1;
    const char* ev9 = getenv("FAV_HALO_BG9");   // 0: at most 3 taps per weight group (A/B)
    const bool bg9 = !(ev9 && atoi(ev9) == 0);
    for (int nt = 1; nt <= 6; ++nt) {
      const int bn = round_up(ceil_div(cout_pad, nt), 16);   // the last N tile may be partial (TMA zero fill)
      if (bn > 256 || bn * (nt - 1) >= cout_pad) continue;
      if (force_nt > 0 && nt != force_nt) continue;
      for (int mt = 1; mt <= 4 && mt <= groups; ++mt) {
        if (mt * bn > 512) break;
        if (force_mt > 0 && mt != std::min(force_mt, groups)) continue;
        const int slab_tx = (mt * g.nrows + 2) * g.Wp * 128;
        const int slab_bytes = round_up(std::max(((mt - 1) * g.nrows * g.Wp + 130 + 2 * g.Wp) * 128, slab_tx), 1024);
        const int b_bytes = bn * 128;
        // weight tiles travel in groups of 3 taps (one dh row) per barrier when two groups fit: the issuing
        // thread pays ~100 cycles per mbarrier wait and ~70 per commit, so fewer handshakes per MMA matter
        int na = 3, bgroup = 3;
        int nb = std::min(4, (budget - na * slab_bytes) / (3 * b_bytes));
        if (nb < 2) { na = 2; nb = std::min(4, (budget - na * slab_bytes) / (3 * b_bytes)); }
        // narrow weight tiles: all 9 taps of a slab per barrier (FAV_HALO_PROF on the Branch_2 convs: ~250 cycles per
        // (tap, M tile) against ~50 of MMA — the issuing thread's hand-offs, not the tensor pipe)
        if (bg9 && b_bytes <= 8 * 1024 && (budget - 2 * slab_bytes) / (9 * b_bytes) >= 2) {
          bgroup = 9;
          na = (budget - 3 * slab_bytes) / (9 * b_bytes) >= 2 ? 3 : 2;
          nb = std::min(3, (budget - na * slab_bytes) / (9 * b_bytes));
        }
        if (nb < 2) {
          bgroup = 1; na = 3;
          nb = std::min(8, (budget - na * slab_bytes) / b_bytes);
          if (nb < 4) { na = 2; nb = std::min(8, (budget - na * slab_bytes) / b_bytes); }
          if (nb < 3) continue;
        }
        const int acc_stages = (2 * mt * bn <= 512) ? 2 : 1;
        const double ksteps = std::min(4.0, cin / 16.0 / g.cblocks);   // average k-steps per 64-channel block
        // cycles per 128 x bn x 16 MMA (measured, tools/ubench/umma_rate.cu + FAV_HALO_PROF): the tensor pipe needs
        // bn/2; shared memory moves 128 B/cycle and must serve the operand reads (4 KB of A + 32*bn of B) AND the
        // TMA writes that refill them (weight tile shared by mt M tiles, slab shared by 9 taps); the issuing
        // thread needs ~45 cycles per tcgen05.mma plus ~200 per mbarrier wait + commit pair.
        const double per_mma = std::max(std::max(bn / 2.0, (4096.0 + 32.0 * bn * (1.0 + 1.0 / mt) +
                                                            slab_tx / (9.0 * ksteps * mt)) / 128.0),
                                        45.0 + 200.0 / (bgroup * ksteps * mt));
        const double mma = 9.0 * kt * g.cblocks * ksteps * mt * per_mma;
        const double l2 = (9.0 * kt * g.cblocks * b_bytes + 1.0 * kt * g.cblocks * slab_tx) / l2_rate;
        const double epi = mt * (bn / 16.0) * 400.0;   // epilogue of one tile; exposed (and slower: nothing to overlap) when single-buffered
        const double per_tile = acc_stages == 2 ? std::max(std::max(mma, l2), epi) : std::max(mma, l2) + 1.5 * epi;
        const long long tiles = static_cast<long long>(B) * T * ceil_div(H, g.nrows * mt) * nt;
        const double waves = static_cast<double>(ceil_div64(tiles, sms));
        const double cost = waves * per_tile;
        if (cost < best) { best = cost; best_nt = nt; best_mt = mt; best_na = na; best_nb = nb; best_bg = bgroup; }
      }
    }
    FAV_CHECK_ARG(best_nt > 0, "conv halo: no tiling fits (cout_pad=%d, W=%d)", cout_pad, W);
    g.n_tiles = best_nt;
    g.bn = round_up(ceil_div(cout_pad, best_nt), 16);
    g.mt = best_mt;
    g.na = best_na;
    g.nb = best_nb;
    g.bgroup = best_bg;
    g.acc_stages = (2 * g.mt * g.bn <= 512) ? 2 : 1;
    if (getenv("FAV_DEBUG_PLAN"))
      fprintf(stderr, "[fav] halo %dx%dx%d cin=%d cout=%d: bn=%d x%d, mt=%d, na=%d nb=%dx%d acc_stages=%d (model %.0f kclk)\n",
              T, H, W, cin, cout_pad, g.bn, g.n_tiles, g.mt, g.na, g.nb, g.bgroup, g.acc_stages, best / 1e3);
  }
  g.th = ceil_div(H, g.nrows * g.mt);
  g.tw = 1; g.tt = T;
  g.m_tiles = B * T * g.th;
  g.slab_tx = (g.mt * g.nrows + 2) * g.Wp * 128;
  g.slab_bytes = round_up(std::max(((g.mt - 1) * g.nrows * g.Wp + 130 + 2 * g.Wp) * 128, g.slab_tx), 1024);
  const int b_bytes = g.bn * 128;
  L->b_bytes = b_bytes;
  L->a_bytes = g.slab_bytes;
  L->stages = g.nb;
  L->stage_bytes = b_bytes;
  L->smem_bytes = static_cast<size_t>(g.na) * g.slab_bytes + static_cast<size_t>(g.nb) * g.bgroup * b_bytes + 1024 + 512 +
                  kHaloStageBytes;

  uint64_t dims[5], strides[4];
  uint32_t box[5];
  const char* base = static_cast<const char*>(x) + static_cast<long long>(x_coff) * 2;
  dims[0] = static_cast<uint64_t>(cin); dims[1] = W; dims[2] = H; dims[3] = T; dims[4] = B;
  strides[0] = static_cast<uint64_t>(x_cs) * 2;
  strides[1] = strides[0] * W;
  strides[2] = strides[1] * H;
  strides[3] = strides[2] * T;
  box[0] = 64; box[1] = g.Wp; box[2] = g.mt * g.nrows + 2; box[3] = 1; box[4] = 1;
  FAV_CHECK_ARG(box[2] <= 256, "conv halo: box too tall");
  FAV_TRY(make_tmap_bf16(&L->tmA[0], base, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  L->tmA[1] = L->tmA[0]; L->tmA[2] = L->tmA[0]; L->tmA[3] = L->tmA[0];
  // CTA pairs (tcgen05 cta_group::2, conv_halo2.cu): each CTA loads bn/2 rows of every weight tile
  {
    static int pair_mode = -1;
    if (pair_mode < 0) {
      const char* ev = getenv("FAV_HALO_2CTA");
      pair_mode = ev ? atoi(ev) : 1;
    }
    // Pairs pay off where the tile is shared-memory bound and the weight tiles are small enough for deep rings
    // (measured with FAV_HALO_PROF: Conv3d_2c data gradient, N = 64: 1136 -> 717 kclk; N = 192 tiles are bound by the
    // weight supply either way).  FAV_HALO_2CTA=0 disables, =2 forces pairs on small problems too (tests).
    g.pair = (pair_mode && g.bn % 16 == 0 &&
              (pair_mode == 2 ? g.m_tiles >= 2 : (g.bn <= 128 && cin >= 64 && g.m_tiles >= 2 * sm_count(device)))) ? 1 : 0;
    if (getenv("FAV_DEBUG_PLAN")) fprintf(stderr, "[fav]   -> %s\n", g.pair ? "CTA pairs (cta_group::2)" : "single CTA");
  }
  uint64_t bd[2] = {static_cast<uint64_t>(g.nkb) * 64, static_cast<uint64_t>(cout_pad)};
  uint64_t bs[1] = {bd[0] * 2};
  uint32_t bb[2] = {64, static_cast<uint32_t>(g.pair ? g.bn / 2 : g.bn)};
  FAV_TRY(make_tmap_bf16(&L->tmB, wpk, 2, bd, bs, bb, CU_TENSOR_MAP_SWIZZLE_128B));
  if (g.pair) {
    // half-size weight tiles: re-derive the ring depths (deeper rings hide the longer cross-CTA signalling latency)
    const int hb = (g.bn / 2) * 128;
    L->b_bytes = hb;
    const int budget = 212 * 1024;
    int na = 3, bg = 3;
    int nb = std::min(6, (budget - na * g.slab_bytes) / (3 * hb));
    const char* ev9 = getenv("FAV_HALO_BG9");
    const bool bg9 = !(ev9 && atoi(ev9) == 0) && hb <= 4 * 1024 && (budget - 2 * g.slab_bytes) / (9 * hb) >= 2;
    // FAV_HALO_PROF: pairs spend 20-40 % of their time waiting for weight groups (every group crosses the pair: remote
    // expect_tx, remote TMA completion, multicast commit), while a slab lives for 9 taps x k-steps x mt MMAs — so a
    // third slab buys less than one or two more weight groups in flight (FAV_HALO_PAIR_NA=3 restores the old choice)
    static int pair_na = -1;
    if (pair_na < 0) { const char* ev = getenv("FAV_HALO_PAIR_NA"); pair_na = ev ? atoi(ev) : 2; }
    if (nb < 5 && pair_na == 2) {
      const int nb2 = std::min(6, (budget - 2 * g.slab_bytes) / (3 * hb));
      if (nb2 > nb) { na = 2; nb = nb2; }
    }
    if (nb < 3) { na = 2; nb = std::min(6, (budget - na * g.slab_bytes) / (3 * hb)); }
    if (bg9) {   // narrow tiles: the 9 taps of a slab per (cross-CTA) weight hand-off
      bg = 9;
      na = (budget - 3 * g.slab_bytes) / (9 * hb) >= 2 ? 3 : 2;
      nb = std::min(3, (budget - na * g.slab_bytes) / (9 * hb));
    }
    if (nb >= 2) {
      g.na = na; g.nb = nb; g.bgroup = bg;
      L->smem_bytes = static_cast<size_t>(g.na) * g.slab_bytes + static_cast<size_t>(g.nb) * g.bgroup * hb + 1024 + 512 +
                      kHaloStageBytes;
    }
    if (getenv("FAV_DEBUG_PLAN")) fprintf(stderr, "[fav]      pair rings: na=%d nb=%dx%d\n", g.na, g.nb, g.bgroup);
  }
  const int tiles = g.m_tiles * g.n_tiles;
  L->grid = std::max(1, std::min(tiles, sm_count(device)));
  return FAV_OK;
}

int conv_launch(const ConvLaunch& L, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    FAV_CUDA(cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  227 * 1024));
    attr_set = true;
  }
  if (L.use_t3) return conv_t3_launch(L.t3, L.e, L.flops, stream);
  ProfScope ps(L.g.halo ? PK_CONV_HALO : PK_CONV_TAP, stream, L.flops);
  if (L.g.halo && L.g.pair) return conv_launch_halo_pair(L, stream);
  if (L.g.halo) {
    static bool attr2 = false;
    if (!attr2) {
      FAV_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
      attr2 = true;
    }
    static int prof = -1;
    if (prof < 0) prof = getenv("FAV_HALO_PROF") ? 1 : 0;
    if (prof) {
      ConvGeom gp = L.g;
      gp.prof = 1;
      unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}, r[8];
      cudaMemcpyToSymbol(g_halo_prof, z, sizeof(z));
      conv_halo_kernel<<<L.grid, kHaloThreads, L.smem_bytes, stream>>>(L.tmA[0], L.tmB, gp, L.e, L.b_bytes);
      cudaStreamSynchronize(stream);
      cudaMemcpyFromSymbol(r, g_halo_prof, sizeof(r));
      const double n = L.grid;
      fprintf(stderr, "[fav] halo prof T%d H%d W%d cin=%d bn=%dx%d mt=%d: per-CTA kclk total %.0f, wait tempty %.0f, a_full %.0f, b_full %.0f, tiles %.1f\n",
              L.g.T, L.g.H, L.g.W, L.g.cin, L.g.bn, L.g.n_tiles, L.g.mt, r[3] / n / 1e3, r[0] / n / 1e3, r[1] / n / 1e3,
              r[2] / n / 1e3, r[4] / n);
      FAV_COUNT_LAUNCH();
      return FAV_OK;
    }
    FAV_CUDA(launch_pdl(conv_halo_kernel, L.grid, kHaloThreads, L.smem_bytes, stream, L.tmA[0], L.tmB, L.g, L.e, L.b_bytes));
    FAV_COUNT_LAUNCH();
    FAV_CUDA(cudaGetLastError());
    return FAV_OK;
  }
  {
    static int prof = -1;
    if (prof < 0) prof = getenv("FAV_TAP_PROF") ? 1 : 0;
    if (prof) {
      ConvGeom gp = L.g;
      gp.prof = 1;
      unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}, r[8];
      cudaMemcpyToSymbol(g_halo_prof, z, sizeof(z));
      conv_umma_kernel<<<L.grid, kTapThreads, L.smem_bytes, stream>>>(L.tmA[0], L.tmA[1], L.tmA[2], L.tmB, gp, L.e, L.stages,
                                                                   L.a_bytes, L.b_bytes, L.stage_bytes, L.kg);
      cudaStreamSynchronize(stream);
      cudaMemcpyFromSymbol(r, g_halo_prof, sizeof(r));
      const double n = L.grid;
      fprintf(stderr, "[fav] tap prof M %dx%dx%dx%d k%dx%dx%d cin=%d nkb=%d bn=%dx%d stages=%dx%d grid=%d: per-CTA kclk total %.1f, wait tempty %.1f, full %.1f, tiles %.1f\n",
              L.g.B, L.g.T, L.g.H, L.g.W, L.g.kt, L.g.kh, L.g.kw, L.g.cin, L.g.nkb, L.g.bn, L.g.n_tiles, L.stages, L.kg, L.grid,
              r[3] / n / 1e3, r[0] / n / 1e3, r[1] / n / 1e3, r[4] / n);
      FAV_COUNT_LAUNCH();
      return FAV_OK;
    }
  }
  FAV_CUDA(launch_pdl(conv_umma_kernel, L.grid, kTapThreads, L.smem_bytes, stream, L.tmA[0], L.tmA[1], L.tmA[2], L.tmB, L.g,
                      L.e, L.stages, L.a_bytes, L.b_bytes, L.stage_bytes, L.kg));
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

void pack_weights_taps(uint16_t* dst, const float* w, const float* scale, const int* src, int ntaps, int cin_real,
                       int cout_real, int kch, int n_pad, bool dgrad) {
  const int cblocks = ceil_div(kch, 64);
  const size_t K = static_cast<size_t>(ntaps) * cblocks * 64;
  memset(dst, 0, K * n_pad * sizeof(uint16_t));
  for (int j = 0; j < ntaps; ++j) {
    const float* wt = w + static_cast<size_t>(src[j]) * cin_real * cout_real;
    for (int ci = 0; ci < cin_real; ++ci)
      for (int co = 0; co < cout_real; ++co) {
        const float v = wt[static_cast<size_t>(ci) * cout_real + co] * (scale ? scale[co] : 1.0f);
        const int kc = dgrad ? co : ci, n = dgrad ? ci : co;
        const size_t kidx = (static_cast<size_t>(j) * cblocks + kc / 64) * 64 + kc % 64;
        dst[static_cast<size_t>(n) * K + kidx] = dgrad ? f32_to_bf16_bits(v) : f32_to_f16_bits(v);
      }
  }
}

void pack_weights_fwd(uint16_t* dst, const float* w, const float* scale, int taps, int cin_real,
                      int cin_k, int cout_real, int n_pad) {
  const int cblocks = ceil_div(cin_k, 64);
  const size_t K = static_cast<size_t>(taps) * cblocks * 64;
  memset(dst, 0, K * n_pad * sizeof(uint16_t));
  for (int tap = 0; tap < taps; ++tap)
    for (int ci = 0; ci < cin_real; ++ci) {
      const size_t kidx = (static_cast<size_t>(tap) * cblocks + ci / 64) * 64 + ci % 64;
      const float* src = w + (static_cast<size_t>(tap) * cin_real + ci) * cout_real;
      for (int co = 0; co < cout_real; ++co) {
        const float s = scale ? scale[co] : 1.0f;
        dst[static_cast<size_t>(co) * K + kidx] = f32_to_f16_bits(src[co] * s);
      }
    }
}

void pack_weights_dgrad(uint16_t* dst, const float* w, const float* scale, int taps, int cin_real,
                        int cout_real, int cout_k, int n_pad) {
  const int cblocks = ceil_div(cout_k, 64);
  const size_t K = static_cast<size_t>(taps) * cblocks * 64;
  memset(dst, 0, K * n_pad * sizeof(uint16_t));
  for (int tap = 0; tap < taps; ++tap) {
    const int ftap = taps - 1 - tap;  // flipping all three axes == reversing the flattened tap index
    for (int ci = 0; ci < cin_real; ++ci) {
      const float* src = w + (static_cast<size_t>(ftap) * cin_real + ci) * cout_real;
      for (int co = 0; co < cout_real; ++co) {
        const float s = scale ? scale[co] : 1.0f;
        const size_t kidx = (static_cast<size_t>(tap) * cblocks + co / 64) * 64 + co % 64;
        dst[static_cast<size_t>(ci) * K + kidx] = f32_to_bf16_bits(src[co] * s);
      }
    }
  }
}

}  // namespace fav
