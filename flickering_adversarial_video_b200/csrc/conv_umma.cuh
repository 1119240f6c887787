// conv_umma.cuh — launch description of the tcgen05 implicit-GEMM Conv3d kernel.
//
// GEMM view (NDHWC 16-bit activations — fp16 forward, bf16 gradients — fp32 accumulation in TMEM):
//   M = output positions (tiles of 128 = bw*bh*bt boxes inside one clip),
//   N = output channels (tile bn <= 256),
//   K = taps x input channels, walked in k-blocks of one tap x 64 channels (generic path) or one
//       (kt,kh) tap x 8 W-positions x 4 channels (stem path, K=32 per block).
// A tiles come straight from the activation tensor with one TMA box per k-block: the tap shift is a
// coordinate offset and SAME zero padding is TMA out-of-bounds fill.  B tiles (weights, K-major,
// BN-folded) come from a 2-D tensor map.
#pragma once
#include "fav_common.cuh"

namespace fav {

// One 16-bit storage element whose format depends on the role of the tensor: IEEE fp16 for forward activations and
// forward weights, bf16 for gradients and data-gradient weights (fav_common.cuh).  ConvGeom::f16 / ConvEpilogue::out_f16 /
// add_f16 say which one a launch reads and writes.
typedef uint16_t h16;

struct ConvGeom {
  int B, T, H, W;      // output positions (= input positions for the stride-1 path)
  int bw, bh, bt;      // M-tile box, bw*bh*bt == 128
  int tw, th, tt;      // tiles per dim
  int kt, kh, kw;      // taps
  int ot, oh, ow;      // coordinate offset of tap 0 (= -pad_before)
  int st, sh, sw;      // A coordinate stride per position (generic path; TMA element strides)
  int oT, oH, oW;      // output tensor dims; output position = pos * es + eo (generic path)
  int est, esh, esw, eot, eoh, eow;
  int cin;             // K channels per tap as seen by the kernel (multiple of 16)
  int cblocks;         // ceil(cin / 64)
  int nkb;             // k-blocks per tile
  int bn, n_tiles;     // N tiling (bn % 16 == 0)
  int m_tiles;
  // K concatenated from up to 3 A tensors (fused 1x1x1 data gradients of an Inception block); nsrc <= 1: plain
  int nsrc;
  int src_blocks[3];   // 64-channel k-blocks per source
  int src_cin[3];      // channels per source (multiple of 16)
  // halo path (3x3x3, stride 1): one A slab per (channel block, dt) holds nrows+2 zero-padded rows of
  // width Wp = W+2; the 9 (dh,dw) taps are 128-row windows of that slab at row offset dw + Wp*dh.
  int halo;            // 1: halo path
  int Wp, nrows;       // padded row pitch and output rows per tile (nrows*Wp <= 128)
  int slab_bytes;      // bytes reserved per slab (>= (130 + 2*Wp) rows, multiple of 1024)
  int slab_tx;         // bytes one slab TMA delivers = (nrows+2)*Wp*128
  int na, nb;          // A-slab ring depth, B ring depth (in groups of bgroup weight tiles)
  int bgroup;          // weight tiles (taps) per mbarrier handshake: 3 (one dh row) or 1
  int mt;              // M tiles (of nrows rows each) that share every B tile load: mt accumulators of bn columns
  int acc_stages;      // 2: accumulators double-buffered (2*mt*bn <= 512), 1: single-buffered
  int swz_base_offset; // 1: put (start>>7)&7 into the descriptor's base_offset field
  int prof;            // 1: the MMA warp records its barrier wait cycles (debug, FAV_HALO_PROF)
  int pair;            // 1: CTA-pair kernel (tcgen05 cta_group::2, conv_halo2.cu)
  int f16;             // 1: A and B operands are fp16 (forward convs); 0: bf16 (data gradients)
};

struct ConvEpilogue {
  h16* out;                  // [B,T,H,W,out_cs], channels out_coff .. out_coff+cout_store
  int out_f16;               // 1: store fp16 (forward activations), 0: bf16 (gradients)
  long long out_cs;
  int out_coff;
  int cout_store;            // multiple of 8
  const float* bias;         // [bias_rows][bias_ld] or nullptr
  int bias_ld;
  int bias_stem;             // 1: bias row = t*16 + hclass*4 + wclass (delta-dependent stem bias)
  int relu;                  // max(.,0) after bias
  const h16* mask;           // multiply by (mask > 0); a forward activation (fp16), same geometry as out
  long long mask_cs;
  int mask_coff;
  const h16* addend;         // added before the mask (may alias out): a forward residual (fp16) or a gradient (bf16)
  int add_f16;
  long long add_cs;
  int add_coff;
  // GEMM column segments routed to different tensors (fused same-input 1x1x1 convs); nseg <= 1: `out` only
  int nseg;
  int seg_n0[4];             // first GEMM column of segment i; seg_n0[nseg] = total columns
  h16* seg_out[3];
  long long seg_cs[3];
  int seg_coff[3];
};

// (3,1,1) stride-1 convs with cout <= 64: input frames shared over the taps (conv_t3.cu)
struct T3Geom {
  int B, T, HW;           // output (= input) positions: frames x flat plane
  int cin, cblocks;       // input channels, 64-channel blocks
  int tail_bytes;         // row bytes of the last block's A tile: 128 (full 64-channel box), 64 or 32 (SW64 / SW32 box)
  int ktail;              // k-steps (16 channels) of the last block
  int bn;                 // output channels (padded), <= 64
  int tp, ptiles, m_tiles;   // frame groups per clip, 128-position tiles per plane, tiles
  int nst;                // A ring entries (one input frame each)
  int grp_bytes, grp_tx;  // shared-memory bytes / TMA bytes per ring entry
  int w_bytes;            // resident weights
  int f16;                // 1: fp16 operands (forward), 0: bf16
  int prof;
};
struct T3Plan {
  CUtensorMap tmA, tmAt, tmB;
  T3Geom g;
  size_t smem_bytes;
  int grid;
};

struct ConvLaunch {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  ConvGeom g;
  ConvEpilogue e;
  int use_t3;             // 1: conv_launch runs conv_t3_kernel with `t3` (and this launch's epilogue)
  T3Plan t3;
  int stages;             // ring entries (per-tap kernel: groups of kg k-blocks)
  int kg;                 // per-tap kernel: k-blocks per barrier hand-off
  int a_bytes, b_bytes, stage_bytes;
  size_t smem_bytes;
  int grid;
  double flops;   // algorithmic FLOPs of this launch (real channels / positions; set by the plan's owner)
};


#ifdef __CUDACC__
// Epilogue of one accumulator row: TMEM -> registers -> bias / addend / ReLU / ReLU-mask -> fp16 or bf16 -> 16-byte
// stores into the NDHWC channel slice.  Columns are processed 64 at a time with every load of the group
// (4 tcgen05.ld, the mask / addend rows, the bias) issued before the first use: the epilogue is a chain of
// memory latencies per group, so the fewer groups the better.
template <int NQ = 4>   // 16-column chunks per batch (4: 64 columns; 2: 32 columns, for kernels with less register room)
__device__ __forceinline__ void epilogue_columns(const ConvEpilogue& e, int bn, int n0, uint32_t taddr, bool valid,
                                                 h16* out_row, const h16* mask_row,
                                                 const h16* add_row, const float* bias_row,
                                                 int cout_store = -1) {
  if (cout_store < 0) cout_store = e.cout_store;
  const bool of16 = e.out_f16 != 0, af16 = e.add_f16 != 0;
  for (int c0 = 0; c0 < bn; c0 += 16 * NQ) {
    uint32_t r[NQ][16];
    uint4 av[NQ][2], mv[NQ][2];
    bool live[NQ][2];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int c = c0 + q * 16;
      if (c < bn) tmem_ld_32x16(taddr + static_cast<uint32_t>(c), r[q]);   // warp-uniform condition
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int n = n0 + c + half * 8;
        live[q][half] = valid && c < bn && n + 8 <= cout_store;
        av[q][half] = make_uint4(0u, 0u, 0u, 0u);
        mv[q][half] = make_uint4(kF16One2, kF16One2, kF16One2, kF16One2);
        if (live[q][half]) {
          if (add_row) av[q][half] = *reinterpret_cast<const uint4*>(add_row + n);
          if (mask_row) mv[q][half] = __ldg(reinterpret_cast<const uint4*>(mask_row + n));
        }
      }
    }
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int c = c0 + q * 16;
      if (c >= bn) break;
      const int n = n0 + c;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[q][j]);
      if (bias_row && valid && n < cout_store) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(bias_row + n + j));
          v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (!live[q][half]) continue;
        float* vv = v + half * 8;
        const uint4 a = av[q][half];
        vv[0] += lo16(a.x, af16); vv[1] += hi16(a.x, af16); vv[2] += lo16(a.y, af16); vv[3] += hi16(a.y, af16);
        vv[4] += lo16(a.z, af16); vv[5] += hi16(a.z, af16); vv[6] += lo16(a.w, af16); vv[7] += hi16(a.w, af16);
        uint32_t pk[4];   // warp-uniform format / ReLU choice: one conversion instruction per pair
        if (of16) {
          if (e.relu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) pk[j] = pack_f16x2_relu(vv[2 * j], vv[2 * j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) pk[j] = pack_f16x2(vv[2 * j], vv[2 * j + 1]);
          }
        } else {
          if (e.relu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) pk[j] = pack_bf16x2_relu(vv[2 * j], vv[2 * j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) pk[j] = pack_bf16x2(vv[2 * j], vv[2 * j + 1]);
          }
        }
        const uint4 mk = mv[q][half];
        uint4 o;
        o.x = pk[0] & relu_mask2(mk.x);   // the mask only zeroes: exact on the packed values
        o.y = pk[1] & relu_mask2(mk.y);
        o.z = pk[2] & relu_mask2(mk.z);
        o.w = pk[3] & relu_mask2(mk.w);
        *reinterpret_cast<uint4*>(out_row + n + half * 8) = o;
      }
    }
  }
}

// Same epilogue for launches without an addend operand, with warp-coalesced stores: every lane's 16-byte store to
// its own row is a separate L1 wavefront (32 per instruction; measured: the 1x1x1 launches were bound by exactly
// that), so the warp stages 32 rows x 32 columns in shared memory and writes 8 rows x 64 contiguous bytes per
// instruction instead.  A ReLU mask (data gradients) is loaded in the same coalesced pattern, before the TMEM loads are
// waited for, and applied to the packed 16-bit values (exact: the mask only zeroes).  `stage` = this warp's 32 x 5 uint4
// scratch.  bias_row may differ per lane (the stem's border classes): it is applied by the row's owner before staging.
__device__ __forceinline__ void epilogue_columns_staged(const ConvEpilogue& e, int ncols, int n0, uint32_t taddr,
                                                        bool valid, h16* out_row, const float* bias_row,
                                                        int cout_store, uint4* stage, int lane,
                                                        const h16* mask_row = nullptr) {
  const bool of16 = e.out_f16 != 0;
  const unsigned long long row_ptr = reinterpret_cast<unsigned long long>(out_row);
  const unsigned long long mask_ptr = reinterpret_cast<unsigned long long>(mask_row);
  const int k = lane & 3;
  // the four (row, 16-byte column slot) pairs this lane stores per 32-column batch
  unsigned long long p4[4], m4[4];
  int ok4[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int rr = (lane >> 2) + 8 * j;
    p4[j] = __shfl_sync(0xffffffffu, row_ptr, rr);
    m4[j] = __shfl_sync(0xffffffffu, mask_ptr, rr);
    ok4[j] = __shfl_sync(0xffffffffu, valid ? 1 : 0, rr);
  }
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    uint32_t r[2][16];
    tmem_ld_32x16(taddr + static_cast<uint32_t>(c0), r[0]);
    if (c0 + 16 < ncols) tmem_ld_32x16(taddr + static_cast<uint32_t>(c0 + 16), r[1]);   // warp-uniform
    const int cs = c0 + 8 * k;
    const bool st_ok = cs < ncols && n0 + cs + 8 <= cout_store;
    uint4 mk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mk[j] = make_uint4(kF16One2, kF16One2, kF16One2, kF16One2);
      if (mask_row && ok4[j] && st_ok)
        mk[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const h16*>(m4[j]) + n0 + cs));
    }
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = c0 + q * 16;
      if (c >= ncols) break;
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[q][j]);
      if (bias_row && n0 + c < cout_store) {   // bias arrays are padded to 16 channels, not to the N tile
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(bias_row + n0 + c + j));
          v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
        }
      }
      // warp-uniform format / ReLU choice: ONE conversion instruction per pair does round + (ReLU) + saturate
      uint32_t pk[8];
      if (of16) {
        if (e.relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = pack_f16x2_relu(v[2 * j], v[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = pack_f16x2(v[2 * j], v[2 * j + 1]);
        }
      } else {
        if (e.relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2_relu(v[2 * j], v[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        }
      }
      stage[lane * 5 + q * 2] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      stage[lane * 5 + q * 2 + 1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rr = (lane >> 2) + 8 * j;
      uint4 val = stage[rr * 5 + k];
      if (mask_row) {
        val.x &= relu_mask2(mk[j].x);
        val.y &= relu_mask2(mk[j].y);
        val.z &= relu_mask2(mk[j].z);
        val.w &= relu_mask2(mk[j].w);
      }
      if (ok4[j] && st_ok) *reinterpret_cast<uint4*>(reinterpret_cast<h16*>(p4[j]) + n0 + cs) = val;
    }
    __syncwarp();
  }
}

#endif

// Pick the M-tile box for a [T,H,W] volume and a kt x kh x kw kernel.
void choose_box(int T, int H, int W, int kt, int kh, int kw, int* bw, int* bh, int* bt);

// General form of the per-tap path.  GEMM rows are the positions of a [B,T,H,W] grid; tap (dt,dh,dw) of position
// (t,h,w) reads A at (t*st + dt + ot, ...) of the [B,aT,aH,aW,x_cs] tensor (out of range = zero), and the result
// row is stored at (t*est + eot, ...) of the [B,oT,oH,oW] output.  Strided forward convs use st/sh/sw, the parity
// classes of a strided data gradient use est/eot (see fav_api.cu::plan_dgrad_classes).
struct ConvSpec {
  const void* x; long long x_cs; int x_coff; int cin;
  int aT, aH, aW;
  const void* wpk; int cout_pad;
  int B, T, H, W;
  int kt, kh, kw, ot, oh, ow, st, sh, sw;
  int oT, oH, oW, est, esh, esw, eot, eoh, eow;
};
int conv_plan_ex(ConvLaunch* L, int device, const ConvSpec& sp);

// Fill geometry-derived fields + tensor maps for a stride-1 SAME conv whose A operand is
// `x` [B,T,H,W,x_cs] (channels x_coff .. x_coff+cin) and whose packed weights are wpk [n_pad][nkb*64].
int conv_plan_generic(ConvLaunch* L, int device, const void* x, long long x_cs, int x_coff, int cin,
                      const void* wpk, int cout_pad, int B, int T, int H, int W, int kt, int kh,
                      int kw, int flat);

// 3x3x3 stride-1 SAME conv with shared-memory halo reuse (same packed weights as conv_plan_generic).
// kt = 3: 3x3x3; kt = 1: (1,3,3) (Conv3DNoTemporal / the spatial half of Conv2Plus1D)
int conv_plan_halo(ConvLaunch* L, int device, const void* x, long long x_cs, int x_coff, int cin,
                   const void* wpk, int cout_pad, int B, int T, int H, int W, int kt = 3);
// true when the halo path applies and beats the per-tap path for this shape
bool conv_halo_applicable(int T, int H, int W, int kt, int kh, int kw);

int conv_launch(const ConvLaunch& L, cudaStream_t stream);
// conv_t3.cu: x [B,T,HW,x_cs] (channels 0 .. cin), wpk = the per-tap kernel's packed weights [cout_pad][3 * cblocks * 64]
bool conv_t3_applicable(int cin, int cout_pad, int T);
int conv_t3_plan(T3Plan* P, int device, const void* x, long long x_cs, int cin, const void* wpk, int cout_pad, int B, int T,
                 int HW, bool f16);
int conv_t3_launch(const T3Plan& P, const ConvEpilogue& e, double flops, cudaStream_t stream);
int conv_launch_halo_pair(const ConvLaunch& L, cudaStream_t stream);   // conv_halo2.cu

// ---- stem (strided 7x7 spatial, Cin = 3): shared-memory halo reuse over kh, see conv_stem.cu ----
struct StemGeom {
  int B, To, Ho, Wo;      // output positions
  int KT, KH;             // taps in T and H (the 7 W taps x RGBX are the K = 32 of one sub-tile)
  int st;                 // temporal stride (2: I3D, 1: torchvision stems); spatial stride is 2
  int pt, ph;             // pad_before in T and H
  int mt;                 // 128-row M tiles (16 w x 8 h) per CTA tile that share every weight load
  int bn;                 // output channels (multiple of 16)
  int th, tw;             // CTA tiles per plane: th = ceil(Ho / (8*mt)), tw = ceil(Wo / 16)
  int m_tiles;            // B * To * th * tw
  int qmin[2], rows[2];   // per H-parity input slab: first q-row relative to the tile, rows in the slab
  int slab_off[2];        // byte offset of each slab inside a stage
  int a_bytes, b_bytes, stage_bytes, stages;
  int nlo_h, nhi_h, nlo_w, nhi_w;   // border classes of the delta-bias table (rows touching the zero padding)
  // raw-row variant (conv_stem_raw_kernel): the A operand is read straight from raw RGBX input rows through a
  // no-swizzle descriptor whose 16-byte leading-dimension offset makes consecutive M rows overlapping windows
  int raw;                // 1: raw-row kernel
  int nf;                 // output frames per CTA tile (the mt = 2*nf M tiles of 8 columns x 16 rows share every weight load)
  int pitch;              // bytes per raw slab row (40 pixels x 8 B)
  int tp;                 // frame groups: ceil(To / nf)
  // temporal-sharing variant (conv_stem_ts_kernel): one M tile = 8 columns x 16 rows of `tsG` consecutive output frames;
  // every input frame's raw rows are read ONCE per (kh, K half) by an MMA whose N spans all output frames the frame
  // feeds (N = up to tsG*bn <= 256).  Input frames of a tile fall into `st` classes (frame index mod st = kt mod st).
  int ts;                 // 1: temporal-sharing kernel
  int tsG;                // output frames per tile
  int ts_nfr[2];          // input frames of each class per tile
  int ts_nslot[2];        // kt taps of each class
  int ts_ktmax[2];        // largest kt of each class
  int ts_slot_bytes;      // bytes per input-frame slot (both H-parity slabs)
  int ts_set_bytes;       // bytes per A set (the frames of one class)
  int ts_khg;             // kh taps per weight block
  int ts_wblk_bytes;      // bytes per weight block = the taps of ts_khg kh's of one class: khg x nslot x bn x 64 B
  int ts_nw;              // weight ring depth
  int prof;               // 1: the MMA / epilogue warps record their barrier wait cycles (debug, FAV_STEM_PROF)
  uint32_t ts_tab[2][8];  // per (class, frame): first / top output frame, accumulator column, weight-row offset (packed)
};
struct StemLaunch {
  CUtensorMap tmA[4];     // [T-parity][H-parity]
  CUtensorMap tmB;        // [32][cout][taps]
  CUtensorMap tmB1;       // same tensor, one (kt,kh) sub-tile per box (temporal-sharing kernel)
  StemGeom g;
  ConvEpilogue e;
  size_t smem_bytes;
  int grid;
  double flops;
};
// x is the padded RGBX buffer [B,T,H,Wp,4] fp16; wpk [KT*KH][bn][32] fp16
int stem_plan(StemLaunch* L, int device, const void* xpad, int B, int T, int H, int Wp, const void* wpk, int bn,
              int To, int Ho, int Wo, int KT, int KH, int st, int pt, int ph);
int stem_launch(const StemLaunch& L, cudaStream_t stream);
// frame schedule of the temporal-sharing kernel (fills ts_nfr / ts_nslot / ts_ktmax / ts_tab from g->tsG)
int stem_ts_schedule(StemGeom* g, int KT, int st, int bn, int* max_nfr, int* max_slot);


// ---- stem gradient collapse on the tensor cores (stem_grad.cu) ----
struct StemGradLaunch {
  CUtensorMap tmA;        // g1 as a flat [positions][64] matrix
  CUtensorMap tmB;        // weights [KT * 160][64]
  const uint32_t* bits;   // pass nibbles written by the apply kernel
  int B, T, To, Ho, Wo, pt, ph, pw, KT, st;
  float scale[3];
  int tiles_per_plane, m_tiles, bits_rows, bits_pitch, mrows, mbytes;
  size_t smem_bytes;
  int grid;
  int ready;
  double flops, bytes;
};
size_t stem_grad_bitmap_words(int B, int T, int H, int W);
void stem_grad_pack_weights(uint16_t* dst /*[KT][160][64]*/, const float* wq /*[KT*49][3][C]*/, int KT, int C);
int stem_grad_plan(StemGradLaunch* L, int device, const void* g1, int g1_cs, const void* wpk, const uint32_t* bits, int B,
                   int T, int H, int W, int To, int Ho, int Wo, int KT, int st, int pt, int ph, int pw, const float* scale3,
                   int c_real = 0);
int stem_grad_launch(const StemGradLaunch& L, float* grad, cudaStream_t stream);

// Host-side weight packing (16-bit patterns in uint16_t): FORWARD operands are fp16, DATA-GRADIENT operands bf16.
// fwd: w [taps][cin_real][cout_real] (TF layout flattened), scale[cout] (BN fold) or nullptr.
void pack_weights_fwd(uint16_t* dst, const float* w, const float* scale, int taps, int cin_real,
                      int cin_k, int cout_real, int n_pad);
// dgrad: roles swapped and taps flipped; K runs over cout_k (>= cout_real), N over cin_pad.
void pack_weights_dgrad(uint16_t* dst, const float* w, const float* scale, int taps, int cin_real,
                        int cout_real, int cout_k, int n_pad);

// general form: packed tap j takes source tap src[j] of w [taps][cin_real][cout_real]; forward: K = cin, N = cout;
// dgrad: K = cout, N = cin (no flip here: the tap list encodes it).  kch = K channels padded to a multiple of 16.
void pack_weights_taps(uint16_t* dst, const float* w, const float* scale, const int* src, int ntaps, int cin_real,
                       int cout_real, int kch, int n_pad, bool dgrad);

uint16_t f32_to_bf16_bits(float f);
float bf16_bits_to_f32(uint16_t b);
uint16_t f32_to_f16_bits(float f);   // round to nearest even, subnormals kept, saturating at +-65504

}  // namespace fav
