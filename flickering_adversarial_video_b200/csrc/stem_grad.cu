// Stem gradient collapse on the tensor cores: dL/d delta[t,c] = sum_{b,h,w} pass(b,t,h,w,c) * dL/dX(b,t,h,w,c) without
// ever materialising dL/dX and with a cost that does not depend on how many entries the range clip saturated.
//
// Reference: tf.gradients of the loss w.r.t. the perturbation through tf.clip_by_value(x + delta, -1, 1) and the
// Conv3d_1a_7x7 unit (single_video_npy.py:66-84, utils/kinetics_i3d_utils.py:165-171, i3d.py:275-279).
//
//   dL/dX(b,t,h,w,c) = sum over output positions o = (to,ho,wo) and taps k = (kt,kh,kw) with 2*o + k - pad = (t,h,w)
//                      of sum_co g1[b,o,co] * w[k,c,co]
//
// For one output position the 7*7*7*3 = 1029 products P[o,(k,c)] = sum_co g1[o,co] w[k,c,co] are a GEMM row
// (K = 64 output channels, N = 1029): 128 positions x 64 channels of g1 are the A tile (one TMA box of the flat
// [positions, 64] matrix), the weights [N][64] stay resident in shared memory, and the N dimension is walked in 7 chunks
// of 160 columns (one temporal tap kt each: 49 in-plane taps x 3 channels = 147 live columns).  Every column of a chunk
// lands on the same frame t = 2*to + kt - pad_t, so the epilogue reduces a chunk to three numbers per row: it adds
// P[o,col] where the pass bit of the input entry that column touches is set.  The pass bits come from the apply kernel as
// one nibble per pixel (bit c = entry (pixel, c) was not range-clipped) in a zero-padded bitmap, so taps that fall into
// the convolution's zero padding read zeros and need no bounds logic.
//
// Roofline (B=8, T=64, 224x224): 2 * 3.2 M positions * 64 * 1120 = 0.46 TFLOP of bf16 MMA; g1 is read once (411 MB).
#include "conv_umma.cuh"
#include "kernels.cuh"

namespace fav {

namespace {

constexpr int kSgSets = 2;         // epilogue warp sets; set s owns the temporal taps with kt % 2 == s
constexpr int kSgEpiThreads = kSgSets * 128;
constexpr int kSgThreads = 64 + kSgEpiThreads;   // warp 0: TMA, warp 1: MMA issue, then 4 warps (one per TMEM lane quarter) per set
constexpr int kSgChunkN = 160;     // columns per temporal tap: 147 live + 13 zero
constexpr int kSgStages = 2;       // A tiles (and their bitmap windows) in flight
constexpr int kSgAcc = 3;          // TMEM accumulators (3 x 160 columns)
constexpr int kSgABytes = 128 * 128;
constexpr int kSgBBytes = kSgChunkN * 128;

struct StemGradGeom {
  int B, T, To, Ho, Wo;
  int pt, ph, pw;
  int tiles_per_plane;   // ceil(Ho*Wo / 128)
  int m_tiles;           // B * To * tiles_per_plane
  int bits_rows;         // bitmap rows per frame (H + 7: 3 zero rows above, 4 below)
  int bits_pitch;        // 32-bit words per bitmap row (8 zero nibbles left of w = 0; a multiple of 4 words)
  int mrows;             // bitmap rows one tile can touch per frame: 2 * (output rows spanned - 1) + 7
  int mbytes;            // shared-memory bytes of one bitmap stage: KT * mrows * bits_pitch * 4
  int KT, st;            // temporal taps (7: I3D, 3 / 1: torchvision stems) and temporal stride (2 / 1); in-plane 7x7 / 2
  float scale[3];        // per-channel factor of the result (torch stack: 1/std_c, delta enters as delta/std)
  int dbg;               // FAV_SG_DBG (timing experiments only): bit 0 = epilogue skips the accumulator reads, 2 = print MMA-warp wait cycles
};

__device__ unsigned long long g_sg_prof[8];   // FAV_SG_DBG >= 2: MMA-warp cycles [total, wait a_full, wait t_empty]

// contiguous tile range of a CTA: consecutive tiles share (b, to), so the epilogue keeps its sums in registers
__device__ __forceinline__ void sg_tile_range(const StemGradGeom& g, int* first, int* last) {
  const int q = g.m_tiles / gridDim.x, r = g.m_tiles % gridDim.x;
  const int bx = blockIdx.x;
  *first = bx * q + min(bx, r);
  *last = *first + q + (bx < r ? 1 : 0);
}

// one accumulator chunk (one temporal tap of this tile): add the columns whose pass bit is set.  The TMEM loads are
// double-buffered and every channel runs four independent FADD chains: the warp has at most one partner on its scheduler,
// so latencies are only hidden by its own instruction-level parallelism.
__device__ __forceinline__ void sg_chunk(uint32_t taddr, const uint32_t (&m)[7], float (&a)[3]) {
  float ch[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c) ch[c][0] = ch[c][1] = ch[c][2] = ch[c][3] = 0.0f;
  uint32_t v[2][32];
  tmem_ld_32x32(taddr, v[0]);
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    tmem_ld_wait();
    if (j + 1 < 5) tmem_ld_32x32(taddr + (j + 1) * 32, v[(j + 1) & 1]);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int col = j * 32 + i;
      if (col < 147) {
        // columns run (kh, kw, c): consecutive columns test consecutive bits of one mask word (R2P sets 7 predicates at once)
        const int c = col % 3, kw = (col / 3) % 7, kh = col / 21;
        if (m[kh] & (1u << (4 * kw + c))) ch[c][(col / 3) & 3] += __uint_as_float(v[j & 1][i]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) a[c] += (ch[c][0] + ch[c][1]) + (ch[c][2] + ch[c][3]);
}

__global__ void __launch_bounds__(kSgThreads, 1)
stem_grad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const StemGradGeom g,
                 const uint32_t* __restrict__ bits, float* __restrict__ grad) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sB = smem;                                  // KT chunks x [160][64] bf16, SW128
  uint8_t* sA = smem + g.KT * kSgBBytes;               // kSgStages x [128][64] bf16, SW128
  uint8_t* sM = sA + kSgStages * kSgABytes;            // kSgStages x [7][mrows][pitch] bitmap windows
  uint64_t* bars = reinterpret_cast<uint64_t*>(sM + kSgStages * g.mbytes);
  uint64_t* a_full = bars;                 // [3]
  uint64_t* a_empty = bars + 3;            // [3]
  uint64_t* t_full = bars + 6;             // [3]
  uint64_t* t_empty = bars + 9;            // [3]
  uint64_t* m_full = bars + 12;            // [3]
  uint64_t* m_empty = bars + 15;           // [3]
  uint64_t* b_full = bars + 18;            // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);
  float* sacc = reinterpret_cast<float*>(bars + 20);   // [T*3] sums of this CTA
  float* sthr = sacc + ((g.T * 3 + 31) & ~31);         // [8 epilogue warps][4 owned taps][3][32 lanes]: per-thread sums of the current plane

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int tile_first, tile_last;
  sg_tile_range(g, &tile_first, &tile_last);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kSgStages; ++s) {
      mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1);
      mbar_init(&m_full[s], 1); mbar_init(&m_empty[s], kSgSets * 4);
    }
    for (int s = 0; s < kSgAcc; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 4); }
    mbar_init(b_full, 1);
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < g.T * 3; i += kSgThreads) sacc[i] = 0.0f;
  for (int i = threadIdx.x; i < kSgSets * 4 * 4 * 3 * 32; i += kSgThreads) sthr[i] = 0.0f;
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();   // prologue done; global memory only after the previous kernels of the stream have completed

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(b_full, g.KT * kSgBBytes);
      for (int kt = 0; kt < g.KT; ++kt) tma_load_2d(sB + kt * kSgBBytes, &tmB, b_full, 0, kt * kSgChunkN);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_first; tile < tile_last; ++tile) {
        const int plane = tile / g.tiles_per_plane;
        const int pos0 = (tile - plane * g.tiles_per_plane) * 128;
        const int b = plane / g.To, to = plane - b * g.To;
        const int kt_lo = max(0, g.pt - g.st * to), kt_hi = min(g.KT - 1, g.T - 1 + g.pt - g.st * to);
        const int ho0 = pos0 / g.Wo, ho1 = min(g.Ho - 1, (pos0 + 127) / g.Wo);
        const uint32_t wbytes = static_cast<uint32_t>((2 * (ho1 - ho0) + 7) * g.bits_pitch * 4);
        mbar_wait(&a_empty[stage], phase ^ 1);
        mbar_expect_tx(&a_full[stage], kSgABytes);
        tma_load_2d(sA + stage * kSgABytes, &tmA, &a_full[stage], 0, plane * g.Ho * g.Wo + pos0);
        // pass-bit windows of the frames this tile's temporal taps touch: rows 2*ho0 - ph .. 2*ho1 + 6 - ph
        mbar_wait(&m_empty[stage], phase ^ 1);
        mbar_expect_tx(&m_full[stage], wbytes * static_cast<uint32_t>(kt_hi - kt_lo + 1));
        for (int kt = kt_lo; kt <= kt_hi; ++kt) {
          const int t = g.st * to + kt - g.pt;
          const uint32_t* src = bits + (static_cast<long long>(b * g.T + t) * g.bits_rows + (2 * ho0 - g.ph + 3)) * g.bits_pitch;
          bulk_load_1d(sM + stage * g.mbytes + kt * g.mrows * g.bits_pitch * 4, src, wbytes, &m_full[stage]);
        }
        if (++stage == kSgStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Accumulators rotate over the chunks (chunk n -> TMEM buffer n % 3): a buffer's round trip (MMA pipeline latency, two
    // mbarrier hand-offs, the epilogue's reads) is several times the 320 cycles of tensor time per chunk, so all three must
    // be in flight (measured with FAV_SG_DBG: one buffer per epilogue set serialises the round trips: 0.8 ms).
    const uint32_t idesc = umma_idesc_bf16(128, kSgChunkN);
    const uint32_t desc_hi = umma_desc_hi(128);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    mbar_wait(b_full, 0);
    const bool prof = g.dbg >= 2;
    long long w_a = 0, w_t = 0, c0 = 0;
    const long long t_start = clock64();
    for (int tile = tile_first; tile < tile_last; ++tile) {
      const int plane = tile / g.tiles_per_plane;
      const int to = plane % g.To;
      const int kt_lo = max(0, g.pt - g.st * to), kt_hi = min(g.KT - 1, g.T - 1 + g.pt - g.st * to);
      if (prof) c0 = clock64();
      mbar_wait(&a_full[stage], phase);
      if (prof) w_a += clock64() - c0;
      tc_fence_after();
      const uint32_t a_lo = umma_desc_lo(smem_u32(sA + stage * kSgABytes));
      for (int kt = kt_lo; kt <= kt_hi; ++kt) {
        if (prof) c0 = clock64();
        mbar_wait(&t_empty[acc], acc_phase ^ 1);
        if (prof) w_t += clock64() - c0;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_lo = umma_desc_lo(smem_u32(sB + kt * kSgBBytes));
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kSgChunkN);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(d_tmem, make_desc(desc_hi, a_lo + 2 * k), make_desc(desc_hi, b_lo + 2 * k), idesc, k > 0 ? 1u : 0u);
          umma_commit(&t_full[acc]);
          if (kt == kt_hi) umma_commit(&a_empty[stage]);
        }
        __syncwarp();
        if (++acc == kSgAcc) { acc = 0; acc_phase ^= 1; }
      }
      if (++stage == kSgStages) { stage = 0; phase ^= 1; }
    }
    if (prof && lane == 0) {
      atomicAdd(&g_sg_prof[0], static_cast<unsigned long long>(clock64() - t_start));
      atomicAdd(&g_sg_prof[1], static_cast<unsigned long long>(w_a));
      atomicAdd(&g_sg_prof[2], static_cast<unsigned long long>(w_t));
    }
  } else {
    // ===================== epilogue: masked row sums =====================
    const int set = (warp - 2) >> 2;                    // owns the temporal taps with kt % kSgSets == set
    const int quarter = warp & 3;                       // TMEM lanes 32*quarter .. +31 belong to this warp
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    const int HW = g.Ho * g.Wo;
    int nchunk = 0;                                     // chunks issued before this tile (chunk n lives in TMEM buffer n % 3)
    int stage = 0;
    uint32_t phase = 0;
    float* my = sthr + (warp - 2) * (4 * 3 * 32) + lane;   // this thread's slots: my[(j * 3 + c) * 32], j = kt / 2
    int cur_plane = -1;
    auto flush = [&](int plane) {
      const int to = plane % g.To;
      for (int j = 0; j < 4; ++j) {
        const int kt = 2 * j + set;
        const int t = g.st * to + kt - g.pt;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float sum = my[(j * 3 + c) * 32];
          my[(j * 3 + c) * 32] = 0.0f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          if (lane == 0 && kt < g.KT && t >= 0 && t < g.T) atomicAdd(&sacc[t * 3 + c], sum);
        }
      }
    };
    for (int tile = tile_first; tile < tile_last; ++tile) {
      const int plane = tile / g.tiles_per_plane;
      if (plane != cur_plane) {
        if (cur_plane >= 0) flush(cur_plane);
        cur_plane = plane;
      }
      const int pos0 = (tile - plane * g.tiles_per_plane) * 128;
      const int pos = pos0 + r;
      const int to = plane % g.To;
      const bool valid = pos < HW;
      const int ho0 = pos0 / g.Wo;
      const int ho = valid ? pos / g.Wo : ho0;
      const int wo = valid ? pos - ho * g.Wo : 0;
      const int bit0 = 8 * wo + 32 - 4 * g.pw;          // nibble 2*wo - pw + 8 of the bitmap row
      const int shift = bit0 & 31;
      const int kt_lo = max(0, g.pt - g.st * to), kt_hi = min(g.KT - 1, g.T - 1 + g.pt - g.st * to);
      // this thread's 7 x 7 window inside the staged bitmap rows (row 0 of the stage = input row 2*ho0 - ph)
      const uint32_t mrow0 = smem_u32(sM + stage * g.mbytes) + static_cast<uint32_t>((2 * (ho - ho0) * g.bits_pitch + (bit0 >> 5)) * 4);
      mbar_wait(&m_full[stage], phase);
      // Consecutive chunks of one set are at most 3 apart in the running chunk order (kt alternates between the two sets),
      // so a set never waits on an accumulator phase it could confuse with the one two uses earlier.
      for (int kt = set; kt <= kt_hi; kt += kSgSets) {   // rolled: one copy of the chunk code
        if (kt < kt_lo) continue;
        const int n = nchunk + kt - kt_lo;
        const int acc = n % kSgAcc;
        uint32_t m[7];
        const uint32_t mr = mrow0 + static_cast<uint32_t>(kt * g.mrows * g.bits_pitch * 4);
#pragma unroll
        for (int kh = 0; kh < 7; ++kh) {
          uint32_t lo, hi;
          asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%2 + 4];" : "=r"(lo), "=r"(hi) : "r"(mr + kh * g.bits_pitch * 4));
          m[kh] = valid ? (__funnelshift_r(lo, hi, shift) & 0x0fffffffu) : 0u;
        }
        mbar_wait(&t_full[acc], static_cast<uint32_t>(n / kSgAcc) & 1u);
        tc_fence_after();
        float a3[3] = {0.0f, 0.0f, 0.0f};
        if (!(g.dbg & 1)) sg_chunk(tmem_base + lane_addr + static_cast<uint32_t>(acc * kSgChunkN), m, a3);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[acc]);
#pragma unroll
        for (int c = 0; c < 3; ++c) my[((kt >> 1) * 3 + c) * 32] += a3[c];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&m_empty[stage]);
      nchunk += kt_hi - kt_lo + 1;
      if (++stage == kSgStages) { stage = 0; phase ^= 1; }
    }
    if (cur_plane >= 0) flush(cur_plane);
    // the epilogue warps publish the CTA's partial sums
    asm volatile("bar.sync 1, %0;" ::"n"(kSgEpiThreads) : "memory");
    for (int i = threadIdx.x - 64; i < g.T * 3; i += kSgEpiThreads)
      if (sacc[i] != 0.0f) atomicAdd(&grad[i], sacc[i] * g.scale[i % 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// wq: folded fp32 stem weights [KT*7*7][3][C]; dst: [KT][160][64] bf16 rows (kt, (kh*7 + kw)*3 + c), K = co (zero beyond C)
void stem_grad_pack_weights(uint16_t* dst, const float* wq, int KT, int C) {
  memset(dst, 0, static_cast<size_t>(KT) * kSgChunkN * 64 * sizeof(uint16_t));
  for (int kt = 0; kt < KT; ++kt)
    for (int c = 0; c < 3; ++c)
      for (int kh = 0; kh < 7; ++kh)
        for (int kw = 0; kw < 7; ++kw)
          for (int co = 0; co < C; ++co)
            dst[(static_cast<size_t>(kt) * kSgChunkN + (kh * 7 + kw) * 3 + c) * 64 + co] =
                f32_to_bf16_bits(wq[(static_cast<size_t>((kt * 7 + kh) * 7 + kw) * 3 + c) * C + co]);
}

int stem_grad_plan(StemGradLaunch* L, int device, const void* g1, int g1_cs, const void* wpk, const uint32_t* bits, int B,
                   int T, int H, int W, int To, int Ho, int Wo, int KT, int st, int pt, int ph, int pw, const float* scale3,
                   int c_real) {
  memset(L, 0, sizeof(*L));
  FAV_CHECK_ARG(KT >= 1 && KT <= 7 && (st == 1 || st == 2), "stem grad: unsupported temporal kernel %d / stride %d", KT, st);
  FAV_CHECK_ARG(g1_cs % 8 == 0 && g1_cs <= 64, "stem grad: at most 64 stem channels (row stride %d)", g1_cs);
  L->KT = KT; L->st = st;
  for (int c = 0; c < 3; ++c) L->scale[c] = scale3 ? scale3[c] : 1.0f;
  FAV_CHECK_ARG(W % 8 == 0, "stem grad: W=%d must be a multiple of 8", W);
  FAV_CHECK_ARG(ph >= 0 && ph <= 3 && pw >= 0 && pw <= 8, "stem grad: unsupported padding");
  FAV_CHECK_ARG(2 * (Ho - 1) + 6 - ph + 3 < H + 7 && ((8 * (Wo - 1) + 32 - 4 * pw) >> 5) + 2 <= round_up((W + 16) / 8, 4),
                "stem grad: bitmap too small");
  FAV_CHECK_ARG(T * 3 * 4 <= 8192, "stem grad: T=%d too large", T);
  L->B = B; L->T = T; L->To = To; L->Ho = Ho; L->Wo = Wo; L->pt = pt; L->ph = ph; L->pw = pw;
  L->tiles_per_plane = ceil_div(Ho * Wo, 128);
  L->m_tiles = B * To * L->tiles_per_plane;
  L->bits_rows = H + 7;
  L->bits_pitch = round_up((W + 16) / 8, 4);   // rows start 16-byte aligned (bulk copies)
  L->mrows = 2 * ((127 + Wo - 1) / Wo) + 7;
  L->mbytes = round_up(KT * L->mrows * L->bits_pitch * 4, 128);
  L->bits = bits;
  // rows narrower than 64 channels (r2plus1d_18: 48) are zero-filled to the 128-byte swizzle span by TMA
  uint64_t dims[2] = {static_cast<uint64_t>(g1_cs), static_cast<uint64_t>(B) * To * Ho * Wo};
  uint64_t strides[1] = {static_cast<uint64_t>(g1_cs) * 2};
  uint32_t box[2] = {64, 128};
  FAV_TRY(make_tmap_bf16(&L->tmA, g1, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B));
  uint64_t bd[2] = {64, static_cast<uint64_t>(KT) * kSgChunkN};
  uint64_t bstr[1] = {128};
  uint32_t bb[2] = {64, kSgChunkN};
  FAV_TRY(make_tmap_bf16(&L->tmB, wpk, 2, bd, bstr, bb, CU_TENSOR_MAP_SWIZZLE_128B));
  L->smem_bytes = KT * kSgBBytes + kSgStages * (kSgABytes + L->mbytes) + 20 * 8 + static_cast<size_t>(T) * 3 * 4 + 128 +
                  kSgSets * 4 * 4 * 3 * 32 * 4 + 1024 + 64;
  FAV_CHECK_ARG(L->smem_bytes <= 227 * 1024, "stem grad: Wo=%d needs %zu bytes of shared memory", Wo, L->smem_bytes);
  L->grid = std::max(1, std::min(L->m_tiles, sm_count(device)));   // contiguous tile ranges: every CTA gets >= 1 tile
  // algorithmic: the KT*49*3 live columns and the real stem channels (the MMAs run 160 columns x 64 channels)
  L->flops = 2.0 * static_cast<double>(B) * To * Ho * Wo * static_cast<double>(c_real > 0 ? c_real : g1_cs) * KT * 147.0;
  L->bytes = static_cast<double>(B) * To * Ho * Wo * g1_cs * 2.0;
  L->ready = 1;
  return FAV_OK;
}

size_t stem_grad_bitmap_words(int B, int T, int H, int W) {
  return static_cast<size_t>(B) * T * (H + 7) * round_up((W + 16) / 8, 4);
}

int stem_grad_launch(const StemGradLaunch& L, float* grad, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    FAV_CUDA(cudaFuncSetAttribute(stem_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  ProfScope ps(PK_STEM_BWD, stream, L.flops, L.bytes);
  FAV_CUDA(cudaMemsetAsync(grad, 0, static_cast<size_t>(L.T) * 3 * sizeof(float), stream));
  StemGradGeom g;
  g.B = L.B; g.T = L.T; g.To = L.To; g.Ho = L.Ho; g.Wo = L.Wo; g.pt = L.pt; g.ph = L.ph; g.pw = L.pw;
  g.tiles_per_plane = L.tiles_per_plane; g.m_tiles = L.m_tiles; g.bits_rows = L.bits_rows; g.bits_pitch = L.bits_pitch;
  g.mrows = L.mrows; g.mbytes = L.mbytes; g.KT = L.KT; g.st = L.st;
  for (int c = 0; c < 3; ++c) g.scale[c] = L.scale[c];
  static int dbg = -1;
  if (dbg < 0) dbg = getenv("FAV_SG_DBG") ? atoi(getenv("FAV_SG_DBG")) : 0;
  g.dbg = dbg;
  if (dbg >= 2) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0}, r[8];
    cudaMemcpyToSymbol(g_sg_prof, z, sizeof(z));
    stem_grad_kernel<<<L.grid, kSgThreads, L.smem_bytes, stream>>>(L.tmA, L.tmB, g, L.bits, grad);
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(r, g_sg_prof, sizeof(r));
    fprintf(stderr, "[fav] stem grad prof: per-CTA kclk total %.0f, wait a_full %.0f, wait t_empty %.0f (tiles/CTA %.1f)\n",
            r[0] / 1e3 / L.grid, r[1] / 1e3 / L.grid, r[2] / 1e3 / L.grid, static_cast<double>(L.m_tiles) / L.grid);
    FAV_COUNT_LAUNCH();
    return FAV_OK;
  }
  FAV_CUDA(launch_pdl(stem_grad_kernel, L.grid, kSgThreads, L.smem_bytes, stream, L.tmA, L.tmB, g, L.bits, grad));
  FAV_COUNT_LAUNCH();
  FAV_CUDA(cudaGetLastError());
  return FAV_OK;
}

}  // namespace fav
