// conv_umma.cuh — launch description of the tcgen05 implicit-GEMM Conv3d kernel.
//
// GEMM view (NDHWC bf16 activations, fp32 accumulation in TMEM):
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

struct ConvGeom {
  int B, T, H, W;      // output positions (= input positions for the stride-1 path)
  int bw, bh, bt;      // M-tile box, bw*bh*bt == 128
  int tw, th, tt;      // tiles per dim
  int kt, kh, kw;      // taps
  int ot, oh, ow;      // coordinate offset of tap 0 (= -pad_before)
  int cin;             // K channels per tap as seen by the kernel (multiple of 16)
  int cblocks;         // ceil(cin / 64)
  int nkb;             // k-blocks per tile
  int bn, n_tiles;     // N tiling (bn % 16 == 0)
  int m_tiles;
  int stem;            // 1: stem path (SW64 rows, parity maps)
  int stem_pt, stem_ph;  // pad_before in T and H of the stride-2 stem
  // halo path (3x3x3, stride 1): one A slab per (channel block, dt) holds nrows+2 zero-padded rows of
  // width Wp = W+2; the 9 (dh,dw) taps are 128-row windows of that slab at row offset dw + Wp*dh.
  int halo;            // 1: halo path
  int Wp, nrows;       // padded row pitch and output rows per tile (nrows*Wp <= 128)
  int slab_bytes;      // bytes reserved per slab (>= (130 + 2*Wp) rows, multiple of 1024)
  int slab_tx;         // bytes one slab TMA delivers = (nrows+2)*Wp*128
  int na, nb;          // A-slab ring depth, B-tile ring depth
  int mt;              // M tiles (of nrows rows each) that share every B tile load: mt accumulators of bn columns
  int acc_stages;      // 2: accumulators double-buffered (2*mt*bn <= 512), 1: single-buffered
  int swz_base_offset; // 1: put (start>>7)&7 into the descriptor's base_offset field
};

struct ConvEpilogue {
  __nv_bfloat16* out;        // [B,T,H,W,out_cs], channels out_coff .. out_coff+cout_store
  long long out_cs;
  int out_coff;
  int cout_store;            // multiple of 8
  const float* bias;         // [bias_rows][bias_ld] or nullptr
  int bias_ld;
  int bias_stem;             // 1: bias row = t*16 + hclass*4 + wclass (delta-dependent stem bias)
  int relu;                  // max(.,0) after bias
  const __nv_bfloat16* mask; // multiply by (mask > 0); same geometry as out
  long long mask_cs;
  int mask_coff;
  const __nv_bfloat16* addend;  // added before the mask (may alias out)
  long long add_cs;
  int add_coff;
};

struct ConvLaunch {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  ConvGeom g;
  ConvEpilogue e;
  int stages;
  int a_bytes, b_bytes, stage_bytes;
  size_t smem_bytes;
  int grid;
};

// Pick the M-tile box for a [T,H,W] volume and a kt x kh x kw kernel.
void choose_box(int T, int H, int W, int kt, int kh, int kw, int* bw, int* bh, int* bt);

// Fill geometry-derived fields + tensor maps for a stride-1 SAME conv whose A operand is
// `x` [B,T,H,W,x_cs] (channels x_coff .. x_coff+cin) and whose packed weights are wpk [n_pad][nkb*64].
int conv_plan_generic(ConvLaunch* L, int device, const void* x, long long x_cs, int x_coff, int cin,
                      const void* wpk, int cout_pad, int B, int T, int H, int W, int kt, int kh,
                      int kw, int flat);

// 3x3x3 stride-1 SAME conv with shared-memory halo reuse (same packed weights as conv_plan_generic).
int conv_plan_halo(ConvLaunch* L, int device, const void* x, long long x_cs, int x_coff, int cin,
                   const void* wpk, int cout_pad, int B, int T, int H, int W);
// true when the halo path applies and beats the per-tap path for this shape
bool conv_halo_applicable(int T, int H, int W, int kt, int kh, int kw);

// Stem: x is the padded RGBX buffer [B,T,H,Wp,4]; output positions [B,To,Ho,Wo]; wpk [64][49*32].
int conv_plan_stem(ConvLaunch* L, int device, const void* xpad, int B, int T, int H, int W, int Wp,
                   const void* wpk, int To, int Ho, int Wo, int pt, int ph);

int conv_launch(const ConvLaunch& L, cudaStream_t stream);

// Host-side weight packing (bf16 bits in uint16_t).
// fwd: w [taps][cin_real][cout_real] (TF layout flattened), scale[cout] (BN fold) or nullptr.
void pack_weights_fwd(uint16_t* dst, const float* w, const float* scale, int taps, int cin_real,
                      int cin_k, int cout_real, int n_pad);
// dgrad: roles swapped and taps flipped; K runs over cout_k (>= cout_real), N over cin_pad.
void pack_weights_dgrad(uint16_t* dst, const float* w, const float* scale, int taps, int cin_real,
                        int cout_real, int cout_k, int n_pad);

uint16_t f32_to_bf16_bits(float f);
float bf16_bits_to_f32(uint16_t b);

}  // namespace fav
