// conv_halo2.cu — the halo-reuse 3x3x3 conv of conv_umma.cu on CTA pairs (tcgen05 cta_group::2).
//
// Measured (tools/ubench, FAV_HALO_PROF): a cta_group::1 tile is bound by the 128 B/cycle shared-memory port —
// every 128xNx16 MMA reads 4 KB of A and 32*N bytes of B, and the TMA refills of both go through the same port.
// A CTA pair halves the B side: each CTA holds N/2 rows of every weight tile, the pair's tensor cores exchange
// them, and one tcgen05.mma (M = 256) issued by the leader drives both SMs.
//
// Protocol (per pair; rank 0 = leader):
//   * both CTAs run the two producer warps: their own A slab (their own M tiles) and their half of each weight
//     group; every TMA signals the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2), which counts one
//     arrive.expect_tx per CTA;
//   * the leader's MMA warp waits on its full barriers and issues tcgen05.mma.cta_group::2; tcgen05.commit with
//     multicast mask 0b11 releases the smem slots and publishes the accumulators in BOTH CTAs;
//   * each CTA's epilogue warps drain their own TMEM; the peer's warps arrive remotely on the leader's tempty
//     barrier (count 8).
#include "conv_umma.cuh"

namespace fav {
extern __device__ unsigned long long g_halo_prof[8];
namespace {

constexpr int kThreads2 = 224;
constexpr int kAccCols = 256;

struct HaloTile {
  int b, t, h0, n0;
};
// m runs over the M tiles of the layer; n over the N tiles
__device__ __forceinline__ HaloTile decode_pair_tile(const ConvGeom& g, int m, int nt) {
  HaloTile c;
  const int hi = m % g.th;
  m /= g.th;
  c.t = m % g.T;
  c.b = m / g.T;          // >= B for the padding tile of an odd tile count: TMA reads zeros, nothing is stored
  c.h0 = hi * g.nrows * g.mt;
  c.n0 = nt * g.bn;
  return c;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t raddr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(raddr), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(void* smem_dst, const CUtensorMap* m, uint32_t mbar_raddr, int c0, int c1,
                                                 int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_raddr), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t mbar_raddr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar_raddr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair when the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
conv_halo2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvGeom g,
                  const ConvEpilogue e, const int b_bytes /* per CTA: bn/2 rows x 128 B */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + static_cast<size_t>(g.na) * g.slab_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + static_cast<size_t>(g.nb) * g.bgroup * b_bytes);
  uint64_t* a_full = bars;            // [4]  (used in the leader)
  uint64_t* a_empty = bars + 4;       // [4]
  uint64_t* b_full = bars + 8;        // [8]  (used in the leader)
  uint64_t* b_empty = bars + 16;      // [8]
  uint64_t* tfull_bar = bars + 24;    // [2]
  uint64_t* tempty_bar = bars + 26;   // [2]  (used in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
  uint4* stage_all = reinterpret_cast<uint4*>(bars + 32);   // 4 epilogue warps x 32 rows x 5 uint4 (coalesced stores)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int m_pairs = (g.m_tiles + 1) >> 1;
  const int total = m_pairs * g.n_tiles;         // pair tiles
  const int slabs_per_tile = g.cblocks * g.kt;   // kt = 3 (3x3x3) or 1 ((1,3,3) convs)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < g.na; ++s) { mbar_init(&a_full[s], 2); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < g.nb; ++s) { mbar_init(&b_full[s], 2); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 8); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();            // barriers of both CTAs are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // prologue done; global memory only after the previous kernels of the stream have completed

  if (warp == 0) {
    // ===================== TMA producer: this CTA's activation slabs =====================
    if (lane == 0) {
      int sa = 0;
      uint32_t pa = 0;
      for (int pt = pair; pt < total; pt += npairs) {
        const HaloTile tc = decode_pair_tile(g, (pt / g.n_tiles) * 2 + static_cast<int>(rank), pt % g.n_tiles);
        for (int sidx = 0; sidx < slabs_per_tile; ++sidx) {
          const int cb = sidx / g.kt;
          const int dt = sidx - cb * g.kt;
          mbar_wait(&a_empty[sa], pa ^ 1);
          const uint32_t lead = mapa_u32(smem_u32(&a_full[sa]), 0);
          mbar_expect_tx_cluster(lead, static_cast<uint32_t>(g.slab_tx));
          tma_load_5d_pair(smem_a + static_cast<size_t>(sa) * g.slab_bytes, &tmA, lead, cb * 64, -1, tc.h0 - 1, tc.t + dt + g.ot,
                           tc.b);
          if (++sa == g.na) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    // ===================== TMA producer: this CTA's half (bn/2 rows) of every weight tile =====================
    // The whole warp walks the ring; lanes 0..bgroup-1 issue the group's weight-tile loads in ONE instruction slot (a
    // single lane pays ~100-150 cycles per cp.async.bulk.tensor it issues — tools/ubench/tma_rate.cu — and the pair
    // kernel's MMA warp waited up to 40 % of its time for weight groups).
    {
      int sb = 0;
      uint32_t pb = 0;
      for (int pt = pair; pt < total; pt += npairs) {
        const int n0 = (pt % g.n_tiles) * g.bn + static_cast<int>(rank) * (g.bn >> 1);
        for (int sidx = 0; sidx < slabs_per_tile; ++sidx) {
          const int cb = sidx / g.kt;
          const int dt = sidx - cb * g.kt;
          for (int j = 0; j < 9; j += g.bgroup) {
            const uint32_t lead = mapa_u32(smem_u32(&b_full[sb]), 0);
            if (lane == 0) {   // one lane polls (32 spinning lanes would compete with the MMA warp for issue slots)
              mbar_wait(&b_empty[sb], pb ^ 1);
              mbar_expect_tx_cluster(lead, static_cast<uint32_t>(b_bytes * g.bgroup));
            }
            __syncwarp();
            if (lane < g.bgroup) {
              const int tap = dt * 9 + j + lane;
              tma_load_2d_pair(smem_b + (static_cast<size_t>(sb) * g.bgroup + lane) * b_bytes, &tmB, lead,
                               (tap * g.cblocks + cb) * 64, n0);
            }
            if (++sb == g.nb) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: leader CTA only =====================
    if (rank == 0) {
      const uint32_t idesc = umma_idesc(256, g.bn, g.f16 != 0);
      const uint32_t desc_hi = umma_desc_hi(128);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t wp16 = static_cast<uint32_t>(g.Wp) * 8u;
      const uint32_t tile16 = wp16 * static_cast<uint32_t>(g.nrows);
      const int acc_cols = g.acc_stages == 2 ? kAccCols : 0;
      long long w_te = 0, w_a = 0, w_b = 0, c0 = 0;
      const long long t_start = clock64();
      for (int pt = pair; pt < total; pt += npairs) {
        c0 = g.prof ? clock64() : 0;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        if (g.prof) w_te += clock64() - c0;
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
                uint32_t a_i = slab_lo + static_cast<uint32_t>(dh) * wp16 + 8u * static_cast<uint32_t>(dw);
                uint32_t d_i = d_tmem;
                for (int i = 0; i < g.mt; ++i) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    if (k < ksteps)
                      umma_bf16_pair(d_i, make_desc(desc_hi, a_i + 2 * k), make_desc(desc_hi, b_lo + 2 * k), idesc,
                                     accum | (k > 0 ? 1u : 0u));
                  }
                  a_i += tile16;
                  d_i += static_cast<uint32_t>(g.bn);
                }
                accum = 1;
              }
              umma_commit_pair(&b_empty[sb]);
              if (j0 + g.bgroup == 9) {
                umma_commit_pair(&a_empty[sa]);
                if (sidx == slabs_per_tile - 1) umma_commit_pair(&tfull_bar[acc]);
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
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) of this CTA's M tiles =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int hm = row / g.Wp;
    const int wm = row - hm * g.Wp;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int pt = pair; pt < total; pt += npairs) {
      const int m = (pt / g.n_tiles) * 2 + static_cast<int>(rank);
      const HaloTile tc = decode_pair_tile(g, m, pt % g.n_tiles);
      const bool tile_valid = m < g.m_tiles;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int acc_cols = g.acc_stages == 2 ? kAccCols : 0;
      for (int i = 0; i < g.mt; ++i) {
        const int h = tc.h0 + i * g.nrows + hm;
        const bool valid = tile_valid && (wm < g.W) && (hm < g.nrows) && (h < g.H);
        const long long pos = valid ? ((static_cast<long long>(tc.b) * g.T + tc.t) * g.H + h) * g.W + wm : 0;
        h16* out_row = e.out + pos * e.out_cs + e.out_coff;
        const h16* mask_row = e.mask ? e.mask + pos * e.mask_cs + e.mask_coff : nullptr;
        const h16* add_row = e.addend ? e.addend + pos * e.add_cs + e.add_coff : nullptr;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc * acc_cols + i * g.bn);
        if (add_row == nullptr)   // coalesced stores through shared memory (see conv_halo_kernel)
          epilogue_columns_staged(e, g.bn, tc.n0, taddr, valid, out_row, e.bias, e.cout_store, stage_all + (warp - 2) * 160,
                                  lane, mask_row);
        else
          epilogue_columns(e, g.bn, tc.n0, taddr, valid, out_row, mask_row, add_row, e.bias);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
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
  cluster_sync_all();            // no CTA of the pair leaves while the other may still signal / read it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace

int conv_launch_halo_pair(const ConvLaunch& L, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    FAV_CUDA(cudaFuncSetAttribute(conv_halo2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr = true;
  }
  const int grid = L.grid & ~1;
  static int prof = -1;
  if (prof < 0) prof = getenv("FAV_HALO_PROF") ? 1 : 0;
  if (prof) {
    ConvGeom gp = L.g;
    gp.prof = 1;
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}, r[8];
    cudaMemcpyToSymbol(g_halo_prof, z, sizeof(z));
    conv_halo2_kernel<<<grid, kThreads2, L.smem_bytes, stream>>>(L.tmA[0], L.tmB, gp, L.e, L.b_bytes);
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(r, g_halo_prof, sizeof(r));
    const double n = grid / 2;
    fprintf(stderr, "[fav] halo2 prof T%d H%d W%d cin=%d bn=%dx%d mt=%d na=%d nb=%dx%d: per-pair kclk total %.0f, wait tempty %.0f, a_full %.0f, b_full %.0f\n",
            L.g.T, L.g.H, L.g.W, L.g.cin, L.g.bn, L.g.n_tiles, L.g.mt, L.g.na, L.g.nb, L.g.bgroup, r[3] / n / 1e3, r[0] / n / 1e3,
            r[1] / n / 1e3, r[2] / n / 1e3);
    FAV_COUNT_LAUNCH();
    return FAV_OK;
  }
  FAV_CUDA(launch_pdl(conv_halo2_kernel, grid, kThreads2, L.smem_bytes, stream, L.tmA[0], L.tmB, L.g, L.e, L.b_bytes));
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

}  // namespace fav
