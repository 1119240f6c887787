// conv_t3.cu — (3,1,1) stride-1 temporal convolutions with few output channels on tcgen05, input frames shared over taps.
//
// Replaces the temporal half of torchvision's Conv2Plus1D where cout <= 64 (r2plus1d_18: the stem's 45 -> 64 and layer1's
// 144 -> 64 convolutions; model.py:403-441 builds them through torchvision.models.video.resnet).  The per-tap kernel
// (conv_umma.cu) streams a fresh 16 KB activation tile from L2 for each of the 3 taps of every 128 x 64 output tile:
// FAV_TAP_PROF showed those launches L2-bound (28 B/clk per SM, 870 cycles per k-block against 204 of MMA).  Here, as in
// conv_stem_ts_kernel, an input frame is loaded ONCE and multiplied by the stacked weight tiles of all taps it serves:
// frame tau feeds output frames tau+1, tau, tau-1 with taps kt = 0, 1, 2, so one MMA of N = 3 x cout writes the TMEM
// accumulators of three output frames.
//   tile   = 128 flat positions of a frame plane x G = 4 consecutive output frames (G accumulators of bn columns, x2 stages)
//   A ring = groups of one input frame: its ceil(cin/64) channel blocks [128 positions x 64 channels], SW128 (a 16- or
//            32-channel tail block is loaded as a 32 / 64 byte-row tile: SW32 / SW64 descriptor); G + 2 frames per tile
//   B      = all weights resident: per channel block the three taps stacked by DESCENDING kt, so the taps of
//            consecutive output frames j, j+1, ... (kt = dd - j) are consecutive rows of the B operand
// The first MMA that touches accumulator j is (frame dd = j, block 0, k-step 0) with j the top of the frame's range:
// that MMA is split into an accumulating part and a fresh part of N = bn.
#include "conv_umma.cuh"

#include <stdlib.h>
#include <string.h>
#include <algorithm>

namespace fav {
extern __device__ unsigned long long g_halo_prof[8];
namespace {

constexpr int kT3Threads = 352;   // warp 0: A producer, 1: MMA issuer, 6: weight loader, 2..5 and 7..10: two epilogue sets
constexpr int kT3G = 4;
constexpr int kT3MaxSt = 6;

__device__ __forceinline__ uint64_t pack_desc2(uint32_t hi, uint32_t lo) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

struct T3Tile {
  int b, t0, p0;
};
__device__ __forceinline__ T3Tile decode_t3(const T3Geom& g, int tile) {
  T3Tile c;
  const int pi = tile % g.ptiles;
  const int m = tile / g.ptiles;
  c.t0 = (m % g.tp) * kT3G;
  c.b = m / g.tp;
  c.p0 = pi * 128;
  return c;
}

__global__ void __launch_bounds__(kT3Threads, 1)
conv_t3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAt,
               const __grid_constant__ CUtensorMap tmB, const T3Geom g, const ConvEpilogue e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  // [weights: cblocks x 3 x bn x 128 B][A ring: nst x grp_bytes][barriers][staging]
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + g.w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + static_cast<size_t>(g.nst) * g.grp_bytes);
  uint64_t* a_full = bars;                   // [kT3MaxSt]
  uint64_t* a_empty = bars + kT3MaxSt;       // [kT3MaxSt]
  uint64_t* w_full = bars + 2 * kT3MaxSt;    // [1]
  uint64_t* tfull_bar = w_full + 1;          // [2]
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint4* stage_all = reinterpret_cast<uint4*>(bars + 2 * kT3MaxSt + 6);   // 8 epilogue warps x 32 rows x 5 uint4

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int acc_cols = kT3G * g.bn;
  const int nfr = kT3G + 2;                  // input frames of a tile
  const uint32_t sub = static_cast<uint32_t>(g.bn) * 128u;   // one tap's weight tile of one channel block

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmAt); tma_prefetch_desc(&tmB);
    for (int s = 0; s < g.nst; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    mbar_init(w_full, 1);
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
    // ===================== TMA producer: one input frame (all channel blocks) per ring entry =====================
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
      const T3Tile tc = decode_t3(g, tile);
      for (int dd = 0; dd < nfr; ++dd) {
        if (lane == 0) {
          mbar_wait(&a_empty[st], ph ^ 1);
          mbar_expect_tx(&a_full[st], static_cast<uint32_t>(g.grp_tx));
        }
        __syncwarp();
        if (lane < g.cblocks) {
          uint8_t* dst = smem_a + static_cast<size_t>(st) * g.grp_bytes + static_cast<size_t>(lane) * (128 * 128);
          const int t_in = tc.t0 - 1 + dd;   // < 0 or >= T: TMA zero fill (the conv's zero padding)
          if (lane == g.cblocks - 1 && g.tail_bytes != 128)
            tma_load_4d(dst, &tmAt, &a_full[st], lane * 64, tc.p0, t_in, tc.b);
          else
            tma_load_4d(dst, &tmA, &a_full[st], lane * 64, tc.p0, t_in, tc.b);
        }
        if (++st == g.nst) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 6) {
    // ===================== weights: loaded once, resident =====================
    // block cb, slot s (tap kt = 2 - s) at smem_w + (cb * 3 + s) * sub; packed k-block of (kt, cb) is kt * cblocks + cb
    if (lane == 0) mbar_expect_tx(w_full, sub * 3u * static_cast<uint32_t>(g.cblocks));
    __syncwarp();
    for (int i = lane; i < 3 * g.cblocks; i += 32) {
      const int cb = i / 3, s = i - cb * 3;
      tma_load_2d(smem_w + static_cast<size_t>(i) * sub, &tmB, w_full, ((2 - s) * g.cblocks + cb) * 64, 0);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t hi128 = umma_desc_hi(128);
    const uint32_t hi_tail = umma_desc_hi(static_cast<uint32_t>(g.tail_bytes));
    const uint32_t idesc0 = umma_idesc(128, 0, g.f16 != 0);
    const uint32_t bn_n = static_cast<uint32_t>(g.bn >> 3) << 17;
    const uint32_t sub16 = sub >> 4;
    const uint32_t w_lo = umma_desc_lo(smem_u32(smem_w));
    int st = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long w_te = 0, w_a = 0, c0 = 0;
    const long long t_start = clock64();
    mbar_wait(w_full, 0);
    for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
      const int t0 = ((tile / g.ptiles) % g.tp) * kT3G;
      const int geff = min(kT3G, g.T - t0);   // output frames of this tile that exist
      c0 = g.prof ? clock64() : 0;
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      if (g.prof) w_te += clock64() - c0;
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * acc_cols);
      for (int dd = 0; dd < nfr; ++dd) {
        // output frames fed by input frame dd: j = dd - kt, kt in [0, 2]
        const int jtop = dd;
        const int jhi = min(jtop, geff - 1);
        const int jlo = max(0, dd - 2);
        const int n = jhi - jlo + 1;
        c0 = g.prof ? clock64() : 0;
        mbar_wait(&a_full[st], ph);
        if (g.prof) w_a += clock64() - c0;
        tc_fence_after();
        const uint32_t grp_lo = umma_desc_lo(smem_u32(smem_a + static_cast<size_t>(st) * g.grp_bytes));
        if (elect_one()) {
          if (n > 0) {
            const uint32_t slot0 = static_cast<uint32_t>(2 - dd + jlo);       // kt of jlo is dd - jlo
            const uint32_t d0 = d_tmem + static_cast<uint32_t>(jlo * g.bn);
            const uint32_t idn = idesc0 + static_cast<uint32_t>(n) * bn_n;
            const bool fresh = jtop <= geff - 1;                               // the top output frame's tap is kt = 0
            for (int cb = 0; cb < g.cblocks; ++cb) {
              const bool tail = cb == g.cblocks - 1;
              const uint32_t hi_a = tail ? hi_tail : hi128;
              const int ksteps = tail ? g.ktail : 4;
              const uint32_t a_lo = grp_lo + static_cast<uint32_t>(cb) * (128u * 128u / 16u);
              const uint32_t b_lo = w_lo + (static_cast<uint32_t>(cb) * 3u + slot0) * sub16;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (k < ksteps) {
                  if (cb == 0 && k == 0 && fresh) {
                    if (n > 1) umma_bf16(d0, pack_desc2(hi_a, a_lo), pack_desc2(hi128, b_lo), idn - bn_n, 1u);
                    umma_bf16(d0 + static_cast<uint32_t>((n - 1) * g.bn), pack_desc2(hi_a, a_lo),
                              pack_desc2(hi128, b_lo + static_cast<uint32_t>(n - 1) * sub16), idesc0 + bn_n, 0u);
                  } else {
                    umma_bf16(d0, pack_desc2(hi_a, a_lo + 2 * k), pack_desc2(hi128, b_lo + 2 * k), idn, 1u);
                  }
                }
              }
            }
          }
          umma_commit(&a_empty[st]);
          if (dd == nfr - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++st == g.nst) { st = 0; ph ^= 1; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (g.prof && lane == 0) {
      atomicAdd(&g_halo_prof[0], static_cast<unsigned long long>(w_te));
      atomicAdd(&g_halo_prof[1], static_cast<unsigned long long>(w_a));
      atomicAdd(&g_halo_prof[3], static_cast<unsigned long long>(clock64() - t_start));
    }
  } else {
    // ===================== epilogue: G M tiles = G output frames of the 128-position patch =====================
    // two sets of four warps (one warp per TMEM lane quarter each): set 0 drains the even output frames, set 1 the odd
    const int eset = warp >= 7 ? 1 : 0;
    const int ew = eset ? warp - 3 : warp - 2;   // staging slot 0..7
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < g.m_tiles; tile += gridDim.x) {
      const T3Tile tc = decode_t3(g, tile);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int p = tc.p0 + row;
      const bool valid = p < g.HW;
      for (int j = eset; j < kT3G; j += 2) {
        const int t = tc.t0 + j;
        if (t >= g.T) break;   // warp-uniform
        const long long pos = valid ? (static_cast<long long>(tc.b) * g.T + t) * g.HW + p : 0;
        h16* out_row = e.out + pos * e.out_cs + e.out_coff;
        const h16* mask_row = e.mask ? e.mask + pos * e.mask_cs + e.mask_coff : nullptr;
        const h16* add_row = e.addend ? e.addend + pos * e.add_cs + e.add_coff : nullptr;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc * acc_cols + j * g.bn);
        if (add_row == nullptr)
          epilogue_columns_staged(e, g.bn, 0, taddr, valid, out_row, e.bias, e.cout_store, stage_all + ew * 160, lane,
                                  mask_row);
        else
          epilogue_columns<2>(e, g.bn, 0, taddr, valid, out_row, mask_row, add_row, e.bias);
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

}  // namespace

// row bytes of the last channel block's A tile (= its swizzle mode): 16 channels -> 32 B rows, 32 -> 64 B, else a full box
static int t3_tail_bytes(int cin) {
  const int ctail = cin - 64 * (ceil_div(cin, 64) - 1);
  return ctail == 16 ? 32 : (ctail == 32 ? 64 : 128);
}
// one ring entry: the full blocks at 16 KB each, then the tail block
static int t3_group_bytes(int cin) {
  return (ceil_div(cin, 64) - 1) * 128 * 128 + round_up(128 * t3_tail_bytes(cin), 1024);
}

bool conv_t3_applicable(int cin, int cout_pad, int T) {
  const char* ev = getenv("FAV_T3");   // 0: keep the per-tap kernel (A/B, tests); read at plan time
  const bool on = !(ev && atoi(ev) == 0);
  if (!on || cout_pad > 64 || cout_pad % 16 != 0 || cin % 16 != 0 || T < 2) return false;
  const int cblocks = ceil_div(cin, 64);
  const int w_bytes = round_up(cblocks * 3 * cout_pad * 128, 1024);
  return w_bytes + 3 * t3_group_bytes(cin) + 1024 + 512 + 8 * 32 * 5 * 16 <= 227 * 1024;
}

// x [B,T,HW,x_cs] (channels 0 .. cin), wpk = the per-tap kernel's packed forward weights [cout_pad][3 * cblocks * 64]
int conv_t3_plan(T3Plan* L, int device, const void* x, long long x_cs, int cin, const void* wpk, int cout_pad, int B, int T,
                 int HW, bool f16) {
  FAV_CHECK_ARG(conv_t3_applicable(cin, cout_pad, T), "conv_t3: cin=%d cout=%d T=%d not supported", cin, cout_pad, T);
  memset(L, 0, sizeof(*L));
  T3Geom& g = L->g;
  g.B = B; g.T = T; g.HW = HW; g.cin = cin; g.bn = cout_pad; g.f16 = f16 ? 1 : 0;
  g.cblocks = ceil_div(cin, 64);
  const int ctail = cin - 64 * (g.cblocks - 1);
  g.tail_bytes = t3_tail_bytes(cin);
  g.ktail = ctail == 16 ? 1 : (ctail == 32 ? 2 : ceil_div(ctail, 16));
  g.tp = ceil_div(T, kT3G);
  g.ptiles = ceil_div(HW, 128);
  g.m_tiles = B * g.tp * g.ptiles;
  g.w_bytes = round_up(g.cblocks * 3 * cout_pad * 128, 1024);
  g.grp_bytes = t3_group_bytes(cin);
  g.grp_tx = (g.cblocks - 1) * 128 * 128 + 128 * g.tail_bytes;
  const int fixed = 1024 + g.w_bytes + 512 + 8 * 32 * 5 * 16;
  g.nst = std::min(kT3MaxSt, (227 * 1024 - fixed) / g.grp_bytes);
  FAV_CHECK_ARG(g.nst >= 3, "conv_t3: ring does not fit");
  L->smem_bytes = static_cast<size_t>(fixed) + static_cast<size_t>(g.nst) * g.grp_bytes;
  // activations as [cin][HW][T][B]
  {
    uint64_t dims[4] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(HW), static_cast<uint64_t>(T), static_cast<uint64_t>(B)};
    uint64_t strides[3] = {static_cast<uint64_t>(x_cs) * 2, static_cast<uint64_t>(x_cs) * 2 * HW,
                           static_cast<uint64_t>(x_cs) * 2 * HW * T};
    uint32_t box[4] = {64, 128, 1, 1};
    FAV_TRY(make_tmap_bf16(&L->tmA, x, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
    L->tmAt = L->tmA;
    if (g.tail_bytes != 128) {
      uint32_t boxt[4] = {static_cast<uint32_t>(g.tail_bytes / 2), 128, 1, 1};
      FAV_TRY(make_tmap_bf16(&L->tmAt, x, 4, dims, strides, boxt,
                             g.tail_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B));
    }
  }
  {
    uint64_t bd[2] = {static_cast<uint64_t>(3 * g.cblocks) * 64, static_cast<uint64_t>(cout_pad)};
    uint64_t bs[1] = {bd[0] * 2};
    uint32_t bb[2] = {64, static_cast<uint32_t>(cout_pad)};
    FAV_TRY(make_tmap_bf16(&L->tmB, wpk, 2, bd, bs, bb, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  L->grid = std::max(1, std::min(g.m_tiles, sm_count(device)));
  return FAV_OK;
}

int conv_t3_launch(const T3Plan& L, const ConvEpilogue& e, double flops, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    FAV_CUDA(cudaFuncSetAttribute(conv_t3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  ProfScope ps(PK_CONV_TAP, stream, flops);
  static int prof = -1;
  if (prof < 0) prof = getenv("FAV_TAP_PROF") ? 1 : 0;
  if (prof) {
    T3Geom gp = L.g;
    gp.prof = 1;
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}, r[8];
    cudaMemcpyToSymbol(g_halo_prof, z, sizeof(z));
    conv_t3_kernel<<<L.grid, kT3Threads, L.smem_bytes, stream>>>(L.tmA, L.tmAt, L.tmB, gp, e);
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(r, g_halo_prof, sizeof(r));
    const double n = L.grid;
    fprintf(stderr, "[fav] t3 prof B%d T%d HW%d cin=%d bn=%d ring=%d tiles=%d grid=%d: per-CTA kclk total %.1f, wait tempty %.1f, a_full %.1f\n",
            L.g.B, L.g.T, L.g.HW, L.g.cin, L.g.bn, L.g.nst, L.g.m_tiles, L.grid, r[3] / n / 1e3, r[0] / n / 1e3, r[1] / n / 1e3);
    FAV_COUNT_LAUNCH();
    return FAV_OK;
  }
  FAV_CUDA(launch_pdl(conv_t3_kernel, L.grid, kT3Threads, L.smem_bytes, stream, L.tmA, L.tmAt, L.tmB, L.g, e));
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

}  // namespace fav
