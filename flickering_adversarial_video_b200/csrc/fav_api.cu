// fav_api.cu — the C-ABI (include/fav.h) and the static I3D execution plan behind it.
//
// The network is a fixed list of fused-op descriptors built once in fav_create (no graph
// executor): stem (tcgen05, delta enters as an fp32 per-frame bias) -> pools -> 9 Inception blocks
// whose branch epilogues write channel slices of the block output (no concat kernels) -> linear
// head.  The backward plan walks the same descriptors in reverse with the data-gradient GEMMs
// only (weights are frozen: there is no weight-gradient pass anywhere).
#include "conv_umma.cuh"
#include "kernels.cuh"

#include <map>
#include <string>
#include <vector>
#include <memory>
#include <cmath>
#include <cstdlib>
#include <algorithm>

using namespace fav;

#include "resnet_plan.cuh"

namespace {

struct Buf {
  std::string name;
  int T = 0, H = 0, W = 0, C = 0, cs = 0;   // C real channels, cs stored stride
  __half* p = nullptr;                       // activation (fp16)
  __nv_bfloat16* g = nullptr;                // gradient w.r.t. the producer's pre-activation (bf16)
  uint8_t* idx = nullptr;                    // arg-max taps when produced by a max-pool
  long long npos(int B) const { return static_cast<long long>(B) * T * H * W; }
};

struct ConvOp {
  std::string name;       // scope path below RGB/inception_i3d/
  int kt, kh, kw;
  int in, in_coff, cin_real, cin_k;
  int out, out_coff, cout_real, cout_pad;
  bool relu = true, bn = true;
  uint16_t* w_fwd = nullptr;   // device, packed fp16 [cout_pad][nkb*64]
  uint16_t* w_dg = nullptr;    // device, packed bf16 [cin_k][nkb'*64]
  float* bias = nullptr;       // device [cout_pad]
  size_t w_fwd_elems = 0, w_dg_elems = 0;
  ConvLaunch fwd, dg;
};

struct PoolOp {
  int in, out;
  PoolGeom g;
};

struct Block {
  int in, out, pool_buf;
  int b0, b1a, b1b, b2a, b2b, b3b;  // conv ids
  int pool;                          // pool id
  // the three 1x1x1 units that read the block input (Branch_0, Branch_1/0a, Branch_2/0a; i3d.py:196-208) run as ONE
  // GEMM with concatenated output channels, their data gradients as ONE GEMM with concatenated K
  bool fused = false;
  int n_tot = 0;                     // concatenated (16-padded) output channels
  uint16_t* wf = nullptr;            // [n_tot][cblocks(cin)*64]
  uint16_t* wd = nullptr;            // [cin_k][(sum of source k-blocks)*64]
  float* bias = nullptr;             // [n_tot]
  size_t wf_elems = 0, wd_elems = 0;
  ConvLaunch fwd, dg;
};

}  // namespace

struct fav_handle {
  int device = 0;
  fav_net_desc d{};
  int B = 0, T = 0, H = 0, W = 0, K = 0;
  // stem geometry
  int To = 0, Ho = 0, Wo = 0, Wp = 0, pt = 0, ph = 0, pw = 0;
  std::vector<Buf> bufs;
  std::vector<ConvOp> convs;
  std::vector<PoolOp> pools;
  std::vector<Block> blocks;
  std::vector<void*> allocs;
  int64_t bytes = 0;
  bool weights_loaded = false;

  // stem
  __half* xpad = nullptr;         // stem operand x' [B,T,H,Wp,4] fp16 (RGBX)
  uint16_t* stem_w = nullptr;     // packed fp16 [64][49*32]
  float* stem_wc = nullptr;       // class-summed folded weights [7][16][3][64]
  float* stem_bnbias = nullptr;   // [64]
  float* stem_bias_tab = nullptr; // [To][16][64]
  StemLaunch stem_fwd;
  int y1 = -1;                    // buffer id of the stem output
  float last_adv_flag = 1.0f;
  float last_delta_clip = 0.4f;
  const float* last_delta = nullptr;

  // stage ops
  int pool2a = -1, conv2b = -1, conv2c = -1, pool3a = -1, pool4a = -1, pool5a = -1;
  int final_buf = -1;

  // head
  float* head_w = nullptr;   // [1024][K]
  float* head_b = nullptr;   // [K]
  float* feat = nullptr;     // [B][1024]
  float* dfeat = nullptr;    // [B][1024]
  float* logits = nullptr;   // [B][K] (internal copy)
  float* dlogits = nullptr;  // [B][K]

  // torch stack (video ResNets)
  ResNet rn;                       // rn.dx / rn.stem_dg also serve the I3D per-pixel attack
  fav_norm_params nrm{};
  const uint8_t* last_clip_u8 = nullptr;

  // sparse per-pixel attack
  bool pixels_enabled = false;
  std::vector<float> stem_w_host;  // I3D: folded fp32 stem weights [343][3][64] (dense stem data gradient)
  const float* last_delta_px = nullptr;
  float* zero_delta = nullptr;     // [T,3] zeros: the stem bias table without a per-frame delta
  float* pix_partial = nullptr;
  // side streams for the independent branches of an Inception block (FAV_BRANCH_STREAMS=0 disables): the persistent conv
  // kernels end with a partial last wave, and a sibling branch's kernel fills the SMs that fall idle
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
  bool branch_streams = false;
  // programmatic dependent launch (fav_common.cuh): measured -1.5 % on r3d_18 (chains of persistent convs), -2.6 % on I3D
  // with one clip, +1 % on I3D with 8 clips (pools and side streams in between); FAV_PDL=0/1 overrides
  bool pdl = false;
  uint32_t* pass_bits = nullptr;  // pass nibbles of the range clip (stem_grad.cu)
  uint16_t* stem_gw = nullptr;    // stem weights as the [KT*160][64] B operand of the gradient collapse
  StemGradLaunch stem_gd;         // tensor-core gradient collapse through the stem
  bool stem_grad_dense = false;   // FAV_STEM_GRAD_DENSE=1 (tests): dense stem data gradient + masked reduce instead

  // forward-only evaluation handle (fav_create_eval): the plan runs 2 * Bu clips — rows [0, Bu) the clean clips, rows
  // [Bu, 2 Bu) the perturbed ones — through one set of packed forward weights; no gradient buffers, pool codes,
  // data-gradient weights or stem-gradient plans exist.  The two halves see different stem bias tables (delta = 0 /
  // delta), so the stem runs as two half-batch launches.
  bool eval = false;
  int Bu = 0;                       // clips per validation batch (h->B == 2 * Bu)
  StemLaunch stem_fwd2;             // stem of the perturbed half
  float* stem_bias_tab2 = nullptr;  // its delta bias table
};

// an evaluation handle (fav_create_eval) owns no gradient state: the training entry points refuse it
#define FAV_NOT_EVAL(h, what)                                                                        \
  do {                                                                                               \
    if ((h)->eval) {                                                                                 \
      set_error("%s: this is a forward-only evaluation handle (fav_create_eval); use fav_eval_batch", what); \
      return FAV_ERR_STATE;                                                                          \
    }                                                                                                \
  } while (0)

namespace {

template <typename Tp>
int dev_alloc(fav_handle* h, Tp** out, size_t count, bool zero = true) {
  void* p = nullptr;
  size_t bytes = count * sizeof(Tp);
  if (bytes == 0) bytes = 16;
  FAV_CUDA(cudaMalloc(&p, bytes));
  if (zero) FAV_CUDA(cudaMemset(p, 0, bytes));
  h->allocs.push_back(p);
  h->bytes += static_cast<int64_t>(bytes);
  *out = static_cast<Tp*>(p);
  return FAV_OK;
}

int add_buf(fav_handle* h, const std::string& name, int T, int H, int W, int C, bool with_idx) {
  Buf b;
  b.name = name; b.T = T; b.H = H; b.W = W; b.C = C;
  b.cs = round_up(C, 16);
  const size_t n = static_cast<size_t>(b.npos(h->B)) * b.cs;
  if (dev_alloc(h, &b.p, n) != FAV_OK) return -1;
  if (!h->eval) {   // a forward-only plan keeps neither gradients nor pool codes
    if (dev_alloc(h, &b.g, n) != FAV_OK) return -1;
    if (with_idx && dev_alloc(h, &b.idx, n) != FAV_OK) return -1;
  }
  h->bufs.push_back(b);
  return static_cast<int>(h->bufs.size()) - 1;
}

int add_conv(fav_handle* h, const std::string& name, int k, int in, int in_coff, int cin_real, int out,
             int out_coff, int cout_real) {
  ConvOp c;
  c.name = name;
  c.kt = c.kh = c.kw = k;
  c.in = in; c.in_coff = in_coff; c.cin_real = cin_real; c.cin_k = round_up(cin_real, 16);
  c.out = out; c.out_coff = out_coff; c.cout_real = cout_real; c.cout_pad = round_up(cout_real, 16);
  h->convs.push_back(c);
  return static_cast<int>(h->convs.size()) - 1;
}

int add_pool(fav_handle* h, int in, int out, int kt, int kh, int kw, int st, int sh, int sw) {
  PoolOp p;
  p.in = in; p.out = out;
  const Buf& bi = h->bufs[in];
  p.g = make_pool_geom(h->B, bi.T, bi.H, bi.W, bi.cs, kt, kh, kw, st, sh, sw);
  h->pools.push_back(p);
  return static_cast<int>(h->pools.size()) - 1;
}

// i3d.py:194-457 — one Inception block; `in` is the block input buffer
int add_block(fav_handle* h, const std::string& name, int in, int c0, int c1a, int c1b, int c2a, int c2b,
              int c3b, const char* b2b_name = "Conv3d_0b_3x3") {
  const Buf bi = h->bufs[in];
  const int cin = bi.C;
  Block b;
  b.in = in;
  b.out = add_buf(h, name, bi.T, bi.H, bi.W, c0 + c1b + c2b + c3b, false);
  const int t1 = add_buf(h, name + "/b1a", bi.T, bi.H, bi.W, c1a, false);
  const int t2 = add_buf(h, name + "/b2a", bi.T, bi.H, bi.W, c2a, false);
  b.pool_buf = add_buf(h, name + "/pool", bi.T, bi.H, bi.W, cin, true);
  if (b.out < 0 || t1 < 0 || t2 < 0 || b.pool_buf < 0) return -1;
  b.b0 = add_conv(h, name + "/Branch_0/Conv3d_0a_1x1", 1, in, 0, cin, b.out, 0, c0);
  b.b1a = add_conv(h, name + "/Branch_1/Conv3d_0a_1x1", 1, in, 0, cin, t1, 0, c1a);
  b.b1b = add_conv(h, name + "/Branch_1/Conv3d_0b_3x3", 3, t1, 0, c1a, b.out, c0, c1b);
  b.b2a = add_conv(h, name + "/Branch_2/Conv3d_0a_1x1", 1, in, 0, cin, t2, 0, c2a);
  b.b2b = add_conv(h, name + "/Branch_2/" + b2b_name, 3, t2, 0, c2a, b.out, c0 + c1b, c2b);
  b.pool = add_pool(h, in, b.pool_buf, 3, 3, 3, 1, 1, 1);
  b.b3b = add_conv(h, name + "/Branch_3/Conv3d_0b_1x1", 1, b.pool_buf, 0, cin, b.out, c0 + c1b + c2b, c3b);
  h->blocks.push_back(b);
  return b.out;
}

int same_out(int in, int s) { return ceil_div(in, s); }
int same_pad_before(int in, int k, int s) {
  int total = (same_out(in, s) - 1) * s + k - in;
  if (total < 0) total = 0;
  return total / 2;
}

h16* as16(__half* p) { return reinterpret_cast<h16*>(p); }
h16* as16(__nv_bfloat16* p) { return reinterpret_cast<h16*>(p); }
const h16* as16(const __half* p) { return reinterpret_cast<const h16*>(p); }
const h16* as16(const __nv_bfloat16* p) { return reinterpret_cast<const h16*>(p); }

bool use_halo(int T, int H, int W, int kt, int kh, int kw) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* ev = getenv("FAV_DISABLE_HALO");
    disabled = (ev && atoi(ev)) ? 1 : 0;
  }
  return !disabled && conv_halo_applicable(T, H, W, kt, kh, kw);
}

int plan_conv(fav_handle* h, ConvOp& c) {
  const Buf& bi = h->bufs[c.in];
  const Buf& bo = h->bufs[c.out];
  const int taps = c.kt * c.kh * c.kw;
  const int flat = taps == 1;
  // ---- forward ----
  {
    const int cblocks = ceil_div(c.cin_k, 64);
    c.w_fwd_elems = static_cast<size_t>(c.cout_pad) * taps * cblocks * 64;
    FAV_TRY(dev_alloc(h, &c.w_fwd, c.w_fwd_elems));
    FAV_TRY(dev_alloc(h, &c.bias, static_cast<size_t>(c.cout_pad)));
    if (use_halo(bi.T, bi.H, bi.W, c.kt, c.kh, c.kw))
      FAV_TRY(conv_plan_halo(&c.fwd, h->device, bi.p, bi.cs, c.in_coff, c.cin_k, c.w_fwd, c.cout_pad, h->B, bi.T,
                             bi.H, bi.W));
    else
      FAV_TRY(conv_plan_generic(&c.fwd, h->device, bi.p, bi.cs, c.in_coff, c.cin_k, c.w_fwd, c.cout_pad,
                                h->B, bi.T, bi.H, bi.W, c.kt, c.kh, c.kw, flat));
    c.fwd.g.f16 = 1;
    ConvEpilogue& e = c.fwd.e;
    e.out = as16(bo.p); e.out_f16 = 1; e.out_cs = bo.cs; e.out_coff = c.out_coff; e.cout_store = c.cout_pad;
    e.bias = c.bias; e.bias_ld = c.cout_pad; e.bias_stem = 0; e.relu = c.relu ? 1 : 0;
    e.mask = nullptr; e.addend = nullptr;
    c.fwd.flops = 2.0 * static_cast<double>(h->B) * bi.T * bi.H * bi.W * taps * c.cin_real * c.cout_real;
  }
  // ---- backward data: A = grad of out slice, N = cin_k ----
  if (!h->eval) {
    const int cblocks = ceil_div(c.cout_pad, 64);
    c.w_dg_elems = static_cast<size_t>(c.cin_k) * taps * cblocks * 64;
    FAV_TRY(dev_alloc(h, &c.w_dg, c.w_dg_elems));
    if (use_halo(bi.T, bi.H, bi.W, c.kt, c.kh, c.kw))
      FAV_TRY(conv_plan_halo(&c.dg, h->device, bo.g, bo.cs, c.out_coff, c.cout_pad, c.w_dg, c.cin_k, h->B, bi.T,
                             bi.H, bi.W));
    else
      FAV_TRY(conv_plan_generic(&c.dg, h->device, bo.g, bo.cs, c.out_coff, c.cout_pad, c.w_dg, c.cin_k,
                                h->B, bi.T, bi.H, bi.W, c.kt, c.kh, c.kw, flat));
    ConvEpilogue& e = c.dg.e;
    e.out = as16(bi.g); e.out_f16 = 0; e.out_cs = bi.cs; e.out_coff = c.in_coff; e.cout_store = c.cin_k;
    e.bias = nullptr; e.bias_ld = 0; e.bias_stem = 0; e.relu = 0;
    e.mask = nullptr; e.addend = nullptr;
    c.dg.flops = 2.0 * static_cast<double>(h->B) * bi.T * bi.H * bi.W * taps * c.cin_real * c.cout_real;
  }
  return FAV_OK;
}

int make_flat_tmap(CUtensorMap* tm, const void* base, long long cs, int coff, int cin, long long M) {
  uint64_t dims[5] = {static_cast<uint64_t>(cin), static_cast<uint64_t>(M), 1, 1, 1};
  uint64_t strides[4] = {static_cast<uint64_t>(cs) * 2, static_cast<uint64_t>(cs) * 2 * M, static_cast<uint64_t>(cs) * 2 * M,
                         static_cast<uint64_t>(cs) * 2 * M};
  uint32_t box[5] = {64, 128, 1, 1, 1};
  return make_tmap_bf16(tm, reinterpret_cast<const char*>(base) + static_cast<long long>(coff) * 2, 5, dims, strides, box,
                        CU_TENSOR_MAP_SWIZZLE_128B);
}

int plan_block_fused(fav_handle* h, Block& b) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* ev = getenv("FAV_NO_FUSE");
    disabled = (ev && atoi(ev)) ? 1 : 0;
  }
  if (disabled) return FAV_OK;
  const ConvOp& c0 = h->convs[b.b0];
  const ConvOp& c1 = h->convs[b.b1a];
  const ConvOp& c2 = h->convs[b.b2a];
  const Buf& bi = h->bufs[b.in];
  const Buf& bo = h->bufs[b.out];
  const Buf& t1 = h->bufs[c1.out];
  const Buf& t2 = h->bufs[c2.out];
  const long long M = bi.npos(h->B);
  const int cin_k = c0.cin_k;
  b.n_tot = c0.cout_pad + c1.cout_pad + c2.cout_pad;
  // ---- forward: N = [c0 | c1a | c2a] ----
  b.wf_elems = static_cast<size_t>(b.n_tot) * ceil_div(cin_k, 64) * 64;
  FAV_TRY(dev_alloc(h, &b.wf, b.wf_elems));
  FAV_TRY(dev_alloc(h, &b.bias, static_cast<size_t>(b.n_tot)));
  FAV_TRY(conv_plan_generic(&b.fwd, h->device, bi.p, bi.cs, 0, cin_k, b.wf, b.n_tot, h->B, bi.T, bi.H, bi.W, 1, 1, 1, 1));
  b.fwd.g.f16 = 1;
  {
    ConvEpilogue& e = b.fwd.e;
    e.out = as16(bo.p); e.out_f16 = 1; e.out_cs = bo.cs; e.out_coff = 0; e.cout_store = c0.cout_pad;
    e.bias = b.bias; e.bias_ld = b.n_tot; e.bias_stem = 0; e.relu = 1; e.mask = nullptr; e.addend = nullptr;
    e.nseg = 3;
    e.seg_n0[0] = 0; e.seg_n0[1] = c0.cout_pad; e.seg_n0[2] = c0.cout_pad + c1.cout_pad; e.seg_n0[3] = b.n_tot;
    e.seg_out[0] = as16(bo.p); e.seg_cs[0] = bo.cs; e.seg_coff[0] = c0.out_coff;
    e.seg_out[1] = as16(t1.p); e.seg_cs[1] = t1.cs; e.seg_coff[1] = 0;
    e.seg_out[2] = as16(t2.p); e.seg_cs[2] = t2.cs; e.seg_coff[2] = 0;
    b.fwd.flops = c0.fwd.flops + c1.fwd.flops + c2.fwd.flops;
  }
  b.fused = true;
  if (h->eval) return FAV_OK;
  // ---- backward data: K = [g(out)[:, :c0] | g(t1) | g(t2)], N = cin ----
  const int nb0 = ceil_div(c0.cout_pad, 64), nb1 = ceil_div(c1.cout_pad, 64), nb2 = ceil_div(c2.cout_pad, 64);
  b.wd_elems = static_cast<size_t>(cin_k) * (nb0 + nb1 + nb2) * 64;
  FAV_TRY(dev_alloc(h, &b.wd, b.wd_elems));
  FAV_TRY(conv_plan_generic(&b.dg, h->device, bo.g, bo.cs, c0.out_coff, c0.cout_pad, b.wd, cin_k, h->B, bi.T, bi.H, bi.W, 1,
                            1, 1, 1));
  {
    ConvGeom& g = b.dg.g;
    g.nsrc = 3;
    g.src_blocks[0] = nb0; g.src_blocks[1] = nb1; g.src_blocks[2] = nb2;
    g.src_cin[0] = c0.cout_pad; g.src_cin[1] = c1.cout_pad; g.src_cin[2] = c2.cout_pad;
    g.nkb = nb0 + nb1 + nb2;
    g.cblocks = g.nkb;
    FAV_TRY(make_flat_tmap(&b.dg.tmA[1], t1.g, t1.cs, 0, c1.cout_pad, M));
    FAV_TRY(make_flat_tmap(&b.dg.tmA[2], t2.g, t2.cs, 0, c2.cout_pad, M));
    // the weight map must span the concatenated K
    uint64_t bd[2] = {static_cast<uint64_t>(g.nkb) * 64, static_cast<uint64_t>(cin_k)};
    uint64_t bs[1] = {bd[0] * 2};
    uint32_t bb[2] = {64, static_cast<uint32_t>(g.bn)};
    FAV_TRY(make_tmap_bf16(&b.dg.tmB, b.wd, 2, bd, bs, bb, CU_TENSOR_MAP_SWIZZLE_128B));
    ConvEpilogue& e = b.dg.e;
    e.out = as16(bi.g); e.out_f16 = 0; e.out_cs = bi.cs; e.out_coff = 0; e.cout_store = cin_k;
    e.bias = nullptr; e.bias_ld = 0; e.bias_stem = 0; e.relu = 0; e.mask = nullptr; e.addend = nullptr; e.nseg = 0;
    b.dg.flops = c0.dg.flops + c1.dg.flops + c2.dg.flops;
  }
  b.fused = true;
  return FAV_OK;
}

// dgrad launch with epilogue options chosen by the caller
int run_dgrad(fav_handle* h, int conv_id, bool mask_with_input, bool accumulate, cudaStream_t s) {
  ConvOp& c = h->convs[conv_id];
  const Buf& bi = h->bufs[c.in];
  ConvLaunch L = c.dg;
  if (mask_with_input) {
    L.e.mask = as16(bi.p); L.e.mask_cs = bi.cs; L.e.mask_coff = c.in_coff;
  }
  if (accumulate) {
    L.e.addend = as16(bi.g); L.e.add_f16 = 0; L.e.add_cs = bi.cs; L.e.add_coff = c.in_coff;
  }
  return conv_launch(L, s);
}

int i3d_plan_dense_stem(fav_handle* h);   // below (needs plan_dgrad_classes)

// Stem launch plan(s) over the padded RGBX operand.  A training handle runs the whole batch in one launch; an evaluation
// handle runs the clean half and the perturbed half as two launches with their own delta bias tables.
int plan_stem(fav_handle* h, int out_buf, int C1, int KT, int st, int pt, int ph, double flops_per_pos) {
  const Buf& bo = h->bufs[out_buf];
  const int nb = h->eval ? h->Bu : h->B;
  const size_t tab = static_cast<size_t>(h->To) * 16 * C1;
  if (h->eval) FAV_TRY(dev_alloc(h, &h->stem_bias_tab2, tab));
  for (int half = 0; half < (h->eval ? 2 : 1); ++half) {
    StemLaunch& L = half ? h->stem_fwd2 : h->stem_fwd;
    const __half* xin = h->xpad + static_cast<size_t>(half) * nb * h->T * h->H * h->Wp * 4;
    FAV_TRY(stem_plan(&L, h->device, xin, nb, h->T, h->H, h->Wp, h->stem_w, C1, h->To, h->Ho, h->Wo, KT, 7, st, pt, ph));
    ConvEpilogue& e = L.e;
    e.out = as16(bo.p + static_cast<size_t>(half) * bo.npos(nb) * bo.cs); e.out_f16 = 1; e.out_cs = bo.cs; e.out_coff = 0;
    e.cout_store = C1;
    e.bias = half ? h->stem_bias_tab2 : h->stem_bias_tab; e.bias_ld = C1; e.bias_stem = 1; e.relu = 1;
    e.mask = nullptr; e.addend = nullptr;
    L.flops = 2.0 * static_cast<double>(nb) * h->To * h->Ho * h->Wo * flops_per_pos;
  }
  return FAV_OK;
}

int build_i3d(fav_handle* h) {
  const int B = h->B, T = h->T, H = h->H, W = h->W;
  FAV_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && W % 16 == 0, "I3D engine needs even H and W %% 16 == 0 (got %dx%d)", H, W);
  // ---- stem: Conv3d_1a_7x7, 7^3 stride 2 SAME (i3d.py:168-171) ----
  h->To = same_out(T, 2); h->Ho = same_out(H, 2); h->Wo = same_out(W, 2);
  h->pt = same_pad_before(T, 7, 2); h->ph = same_pad_before(H, 7, 2); h->pw = same_pad_before(W, 7, 2);
  FAV_CHECK_ARG(h->Ho >= 4 && h->Wo >= 4, "input too small");
  h->Wp = 2 * (h->Wo - 1) + 8;
  if (h->Wp < W + h->pw) h->Wp = W + h->pw;
  h->Wp = round_up(h->Wp, 2);
  FAV_TRY(dev_alloc(h, &h->xpad, static_cast<size_t>(B) * T * H * h->Wp * 4));
  FAV_TRY(dev_alloc(h, &h->stem_w, static_cast<size_t>(64) * 49 * 32));
  FAV_TRY(dev_alloc(h, &h->stem_wc, static_cast<size_t>(7) * 16 * 3 * 64));
  FAV_TRY(dev_alloc(h, &h->stem_bnbias, 64));
  FAV_TRY(dev_alloc(h, &h->stem_bias_tab, static_cast<size_t>(h->To) * 16 * 64));

  h->y1 = add_buf(h, "Conv3d_1a_7x7", h->To, h->Ho, h->Wo, 64, false);
  if (h->y1 < 0) return FAV_ERR_CUDA;
  FAV_TRY(plan_stem(h, h->y1, 64, 7, 2, h->pt, h->ph, 343.0 * 3.0 * 64.0));
  // ---- MaxPool3d_2a_3x3 [1,3,3]/[1,2,2] (i3d.py:173-175) ----
  const Buf y1 = h->bufs[h->y1];
  int p2a = add_buf(h, "MaxPool3d_2a_3x3", y1.T, same_out(y1.H, 2), same_out(y1.W, 2), 64, true);
  if (p2a < 0) return FAV_ERR_CUDA;
  h->pool2a = add_pool(h, h->y1, p2a, 1, 3, 3, 1, 2, 2);
  // ---- Conv3d_2b_1x1, Conv3d_2c_3x3 (i3d.py:178-185) ----
  const Buf bp2a = h->bufs[p2a];
  int y2b = add_buf(h, "Conv3d_2b_1x1", bp2a.T, bp2a.H, bp2a.W, 64, false);
  int y2c = add_buf(h, "Conv3d_2c_3x3", bp2a.T, bp2a.H, bp2a.W, 192, false);
  if (y2b < 0 || y2c < 0) return FAV_ERR_CUDA;
  h->conv2b = add_conv(h, "Conv3d_2b_1x1", 1, p2a, 0, 64, y2b, 0, 64);
  h->conv2c = add_conv(h, "Conv3d_2c_3x3", 3, y2b, 0, 64, y2c, 0, 192);
  // ---- MaxPool3d_3a_3x3 (i3d.py:188-190) ----
  int p3a = add_buf(h, "MaxPool3d_3a_3x3", bp2a.T, same_out(bp2a.H, 2), same_out(bp2a.W, 2), 192, true);
  if (p3a < 0) return FAV_ERR_CUDA;
  h->pool3a = add_pool(h, y2c, p3a, 1, 3, 3, 1, 2, 2);
  // ---- Mixed_3b, 3c (i3d.py:194-249) ----
  int x = add_block(h, "Mixed_3b", p3a, 64, 96, 128, 16, 32, 32);
  x = add_block(h, "Mixed_3c", x, 128, 128, 192, 32, 96, 64);
  if (x < 0) return FAV_ERR_CUDA;
  // ---- MaxPool3d_4a_3x3 3^3/2^3 (i3d.py:251-253) ----
  {
    const Buf bx = h->bufs[x];
    int p4a = add_buf(h, "MaxPool3d_4a_3x3", same_out(bx.T, 2), same_out(bx.H, 2), same_out(bx.W, 2), bx.C, true);
    if (p4a < 0) return FAV_ERR_CUDA;
    h->pool4a = add_pool(h, x, p4a, 3, 3, 3, 2, 2, 2);
    x = p4a;
  }
  x = add_block(h, "Mixed_4b", x, 192, 96, 208, 16, 48, 64);
  x = add_block(h, "Mixed_4c", x, 160, 112, 224, 24, 64, 64);
  x = add_block(h, "Mixed_4d", x, 128, 128, 256, 24, 64, 64);
  x = add_block(h, "Mixed_4e", x, 112, 144, 288, 32, 64, 64);
  x = add_block(h, "Mixed_4f", x, 256, 160, 320, 32, 128, 128);
  if (x < 0) return FAV_ERR_CUDA;
  // ---- MaxPool3d_5a_2x2 2^3/2^3 (i3d.py:397-399) ----
  {
    const Buf bx = h->bufs[x];
    int p5a = add_buf(h, "MaxPool3d_5a_2x2", same_out(bx.T, 2), same_out(bx.H, 2), same_out(bx.W, 2), bx.C, true);
    if (p5a < 0) return FAV_ERR_CUDA;
    h->pool5a = add_pool(h, x, p5a, 2, 2, 2, 2, 2, 2);
    x = p5a;
  }
  // Mixed_5b's Branch_2 3x3 unit is named Conv3d_0a_3x3 in the reference (i3d.py:418)
  x = add_block(h, "Mixed_5b", x, 256, 160, 320, 32, 128, 128, "Conv3d_0a_3x3");
  x = add_block(h, "Mixed_5c", x, 384, 192, 384, 48, 128, 128);
  if (x < 0) return FAV_ERR_CUDA;
  h->final_buf = x;
  {
    const Buf bx = h->bufs[x];
    FAV_CHECK_ARG(bx.H == 7 && bx.W == 7 && bx.T >= 2,
                  "I3D head (avg_pool [2,7,7] VALID, i3d.py:461) needs a 7x7 final map and >=2 frames; got %dx%dx%d",
                  bx.T, bx.H, bx.W);
  }
  for (auto& c : h->convs) FAV_TRY(plan_conv(h, c));
  for (auto& b : h->blocks) FAV_TRY(plan_block_fused(h, b));
  FAV_TRY(i3d_plan_dense_stem(h));
  // ---- head (i3d.py:459-472) ----
  const int C5 = h->bufs[h->final_buf].C;
  FAV_TRY(dev_alloc(h, &h->head_w, static_cast<size_t>(C5) * h->K));
  FAV_TRY(dev_alloc(h, &h->head_b, static_cast<size_t>(h->K)));
  FAV_TRY(dev_alloc(h, &h->feat, static_cast<size_t>(B) * C5));
  FAV_TRY(dev_alloc(h, &h->dfeat, static_cast<size_t>(B) * C5));
  FAV_TRY(dev_alloc(h, &h->logits, static_cast<size_t>(B) * h->K));
  FAV_TRY(dev_alloc(h, &h->dlogits, static_cast<size_t>(B) * h->K));
  return FAV_OK;
}

struct NamedTensors {
  std::map<std::string, const fav_tensor*> m;
  const fav_tensor* find(const std::string& n) const {
    auto it = m.find(n);
    return it == m.end() ? nullptr : it->second;
  }
};

int64_t numel(const fav_tensor* t) {
  int64_t n = 1;
  for (int i = 0; i < t->ndim; ++i) n *= t->dims[i];
  return n;
}

// BN inference fold (snt.BatchNorm defaults: no scale, eps=1e-3; i3d.py:66-68)
int bn_fold(const NamedTensors& nt, const std::string& unit, int cout, std::vector<float>* scale,
            std::vector<float>* bias) {
  scale->assign(cout, 1.0f);
  bias->assign(cout, 0.0f);
  const fav_tensor* beta = nt.find(unit + "/batch_norm/beta");
  const fav_tensor* mean = nt.find(unit + "/batch_norm/moving_mean");
  const fav_tensor* var = nt.find(unit + "/batch_norm/moving_variance");
  const fav_tensor* gamma = nt.find(unit + "/batch_norm/gamma");
  if (!beta || !mean || !var) {
    set_error("missing batch_norm tensors for %s", unit.c_str());
    return FAV_ERR_MISSING;
  }
  if (numel(beta) != cout || numel(mean) != cout || numel(var) != cout) {
    set_error("batch_norm tensor size mismatch for %s (want %d)", unit.c_str(), cout);
    return FAV_ERR_ARG;
  }
  for (int c = 0; c < cout; ++c) {
    const float s = (gamma ? gamma->data[c] : 1.0f) / std::sqrt(var->data[c] + 1e-3f);
    (*scale)[c] = s;
    (*bias)[c] = beta->data[c] - mean->data[c] * s;
  }
  return FAV_OK;
}

}  // namespace

// fork: the side stream continues from this point of the main stream; join: the main stream waits for it.  Both are plain
// event edges, so a CUDA-graph capture of the step turns the branches into parallel graph nodes.
static int branch_fork(fav_handle* h, cudaStream_t s, int i) {
  FAV_CUDA(cudaEventRecord(h->ev_fork[i], s));
  FAV_CUDA(cudaStreamWaitEvent(h->side[i], h->ev_fork[i], 0));
  return FAV_OK;
}
static int branch_join(fav_handle* h, cudaStream_t s, int i) {
  FAV_CUDA(cudaEventRecord(h->ev_join[i], h->side[i]));
  FAV_CUDA(cudaStreamWaitEvent(s, h->ev_join[i], 0));
  return FAV_OK;
}

#include "resnet_impl.cuh"

namespace {
// Dense data gradient of Conv3d_1a_7x7 (7^3, stride 2, TF SAME pad_before 2) as 8 parity-class GEMMs into
// rn.dx [B,T,H,W,16]: the per-pixel attack needs dL/dX itself.  The flickering attack never materialises it
// (stem_grad.cu); FAV_STEM_GRAD_DENSE=1 routes it through this path + the masked reduce as an independent check.
int i3d_plan_dense_stem(fav_handle* h) {
  ResNet& rn = h->rn;
  h->nrm.lo = -1.0f; h->nrm.hi = 1.0f;
  for (int c = 0; c < 3; ++c) { h->nrm.mean[c] = 0.0f; h->nrm.std[c] = 1.0f; }
  if (h->eval) return FAV_OK;
  FAV_TRY(dev_alloc(h, &rn.dx, static_cast<size_t>(h->B) * h->T * h->H * h->W * 16));
  const Buf& y1 = h->bufs[h->y1];
  FAV_TRY(plan_dgrad_classes(h, &rn.stem_dg, y1.g, y1.cs, 64, y1.T, y1.H, y1.W, rn.dx, 16, 16, h->T, h->H, h->W, 7, 7, 7, 2,
                             2, 2, h->pt, h->ph, h->pw, 3.0, 64.0));
  FAV_TRY(dev_alloc(h, &rn.partial, static_cast<size_t>(h->B) * h->T * stem_dx_reduce_chunks(h->H) * 3));
  FAV_TRY(dev_alloc(h, &h->pass_bits, stem_grad_bitmap_words(h->B, h->T, h->H, h->W)));
  FAV_TRY(dev_alloc(h, &h->stem_gw, static_cast<size_t>(7) * 160 * 64));
  FAV_TRY(stem_grad_plan(&h->stem_gd, h->device, y1.g, y1.cs, h->stem_gw, h->pass_bits, h->B, h->T, h->H, h->W, h->To, h->Ho,
                         h->Wo, 7, 2, h->pt, h->ph, h->pw, nullptr, 64));
  h->stem_grad_dense = getenv("FAV_STEM_GRAD_DENSE") != nullptr;
  return FAV_OK;
}
}  // namespace

// =============================================================================================
// C-ABI
// =============================================================================================
static int create_impl(fav_handle** out, int device, const fav_net_desc* desc, bool eval) {
  if (!out || !desc) {
    set_error("fav_create: null argument");
    return FAV_ERR_ARG;
  }
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("fav_create: no CUDA device (this library has no CPU fallback)");
    return FAV_ERR_NOGPU;
  }
  FAV_CHECK_ARG(device >= 0 && device < ndev, "fav_create: device %d out of range", device);
  cudaDeviceProp prop;
  FAV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("fav_create: device %d is sm_%d%d; libfav is built for sm_100a only", device, prop.major, prop.minor);
    return FAV_ERR_NOGPU;
  }
  FAV_CHECK_ARG(desc->arch >= FAV_NET_I3D && desc->arch <= FAV_NET_R2PLUS1D_18, "fav_create: unsupported arch %d", desc->arch);
  FAV_CHECK_ARG(desc->batch >= 1 && desc->frames >= (desc->arch == FAV_NET_I3D ? 9 : 1) && desc->num_classes >= 1,
                "fav_create: bad shape");
  FAV_CUDA(cudaSetDevice(device));
  std::unique_ptr<fav_handle> h(new fav_handle());
  h->device = device;
  h->d = *desc;
  h->B = desc->batch; h->T = desc->frames; h->H = desc->height; h->W = desc->width; h->K = desc->num_classes;
  if (eval) {   // clean | perturbed halves of every validation batch in one pass
    h->eval = true;
    h->Bu = desc->batch;
    h->B = 2 * desc->batch;
  }
  int st = desc->arch == FAV_NET_I3D ? build_i3d(h.get()) : build_resnet(h.get());
  if (st != FAV_OK) {
    for (void* p : h->allocs) cudaFree(p);
    return st;
  }
  {
    const char* pe = getenv("FAV_PDL");
    // I3D: -2.6 % for one 90-frame clip (small launches, prologues matter), +1 % for 8 x 64 frames
    h->pdl = pe ? atoi(pe) != 0 : (desc->arch != FAV_NET_I3D || h->B * desc->frames <= 192);
  }
  {
    const char* ev = getenv("FAV_BRANCH_STREAMS");
    h->branch_streams = !(ev && atoi(ev) == 0);
    if (h->branch_streams)
      for (int i = 0; i < 2; ++i) {
        FAV_CUDA(cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking));
        FAV_CUDA(cudaEventCreateWithFlags(&h->ev_fork[i], cudaEventDisableTiming));
        FAV_CUDA(cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming));
      }
  }
  *out = h.release();
  return FAV_OK;
}

extern "C" int fav_create(fav_handle** out, int device, const fav_net_desc* desc) {
  return create_impl(out, device, desc, false);
}
extern "C" int fav_create_eval(fav_handle** out, int device, const fav_net_desc* desc) {
  return create_impl(out, device, desc, true);
}

extern "C" int fav_destroy(fav_handle* h) {
  if (!h) return FAV_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (void* p : h->allocs) cudaFree(p);
  for (int i = 0; i < 2; ++i) {
    if (h->side[i]) cudaStreamDestroy(h->side[i]);
    if (h->ev_fork[i]) cudaEventDestroy(h->ev_fork[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
  }
  delete h;
  return FAV_OK;
}

extern "C" int64_t fav_device_bytes(const fav_handle* h) { return h ? h->bytes : 0; }

extern "C" int fav_load_weights(fav_handle* h, const fav_tensor* tensors, int n) {
  FAV_CHECK_ARG(h && tensors && n > 0, "fav_load_weights: null argument");
  FAV_CUDA(cudaSetDevice(h->device));
  NamedTensors nt;
  for (int i = 0; i < n; ++i) nt.m[tensors[i].name] = &tensors[i];
  if (h->d.arch != FAV_NET_I3D) {
    FAV_TRY(load_weights_resnet(h, nt));
    h->weights_loaded = true;
    return FAV_OK;
  }
  const std::string root = "RGB/inception_i3d/";
  std::vector<float> scale, bias;
  std::vector<uint16_t> pk;
  // ---- generic convs ----
  for (auto& c : h->convs) {
    const std::string unit = root + c.name;
    const fav_tensor* w = nt.find(unit + "/conv_3d/w");
    if (!w) {
      set_error("missing tensor %s/conv_3d/w", unit.c_str());
      return FAV_ERR_MISSING;
    }
    const int taps = c.kt * c.kh * c.kw;
    FAV_CHECK_ARG(numel(w) == static_cast<int64_t>(taps) * c.cin_real * c.cout_real,
                  "weight %s has %lld elements, expected %d*%d*%d", unit.c_str(), (long long)numel(w), taps,
                  c.cin_real, c.cout_real);
    FAV_TRY(bn_fold(nt, unit, c.cout_real, &scale, &bias));
    pk.resize(c.w_fwd_elems);
    pack_weights_fwd(pk.data(), w->data, scale.data(), taps, c.cin_real, c.cin_k, c.cout_real, c.cout_pad);
    FAV_CUDA(cudaMemcpy(c.w_fwd, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    if (c.w_dg) {
      pk.resize(c.w_dg_elems);
      pack_weights_dgrad(pk.data(), w->data, scale.data(), taps, c.cin_real, c.cout_real, c.cout_pad, c.cin_k);
      FAV_CUDA(cudaMemcpy(c.w_dg, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    }
    std::vector<float> bpad(c.cout_pad, 0.0f);
    for (int i = 0; i < c.cout_real; ++i) bpad[i] = bias[i];
    FAV_CUDA(cudaMemcpy(c.bias, bpad.data(), bpad.size() * 4, cudaMemcpyHostToDevice));
  }
  // ---- fused 1x1x1 heads of the Inception blocks ----
  for (auto& b : h->blocks) {
    if (!b.fused) continue;
    const int ids[3] = {b.b0, b.b1a, b.b2a};
    const int cin_k = h->convs[b.b0].cin_k, cin_real = h->convs[b.b0].cin_real;
    const int cbl = ceil_div(cin_k, 64);
    const size_t Kf = static_cast<size_t>(cbl) * 64;
    std::vector<uint16_t> wf(b.wf_elems, 0), wd(b.wd_elems, 0);
    std::vector<float> bf(b.n_tot, 0.0f);
    const size_t Kd = b.wd ? b.wd_elems / cin_k : 0;
    int n0 = 0, kb0 = 0;
    for (int i = 0; i < 3; ++i) {
      const ConvOp& c = h->convs[ids[i]];
      const fav_tensor* w = nt.find(root + c.name + "/conv_3d/w");
      FAV_TRY(bn_fold(nt, root + c.name, c.cout_real, &scale, &bias));
      for (int ci = 0; ci < cin_real; ++ci)
        for (int co = 0; co < c.cout_real; ++co) {
          const float v = w->data[static_cast<size_t>(ci) * c.cout_real + co] * scale[co];
          wf[static_cast<size_t>(n0 + co) * Kf + (ci / 64) * 64 + ci % 64] = f32_to_f16_bits(v);
          if (b.wd) wd[static_cast<size_t>(ci) * Kd + (static_cast<size_t>(kb0) + co / 64) * 64 + co % 64] = f32_to_bf16_bits(v);
        }
      for (int co = 0; co < c.cout_real; ++co) bf[n0 + co] = bias[co];
      n0 += c.cout_pad;
      kb0 += ceil_div(c.cout_pad, 64);
    }
    FAV_CUDA(cudaMemcpy(b.wf, wf.data(), wf.size() * 2, cudaMemcpyHostToDevice));
    if (b.wd) FAV_CUDA(cudaMemcpy(b.wd, wd.data(), wd.size() * 2, cudaMemcpyHostToDevice));
    FAV_CUDA(cudaMemcpy(b.bias, bf.data(), bf.size() * 4, cudaMemcpyHostToDevice));
  }
  // ---- stem ----
  {
    const std::string unit = root + "Conv3d_1a_7x7";
    const fav_tensor* w = nt.find(unit + "/conv_3d/w");
    if (!w) {
      set_error("missing tensor %s/conv_3d/w", unit.c_str());
      return FAV_ERR_MISSING;
    }
    FAV_CHECK_ARG(numel(w) == 343 * 3 * 64, "stem weight must be [7,7,7,3,64]");
    FAV_TRY(bn_fold(nt, unit, 64, &scale, &bias));
    std::vector<float> wf(343 * 3 * 64);
    for (int tap = 0; tap < 343; ++tap)
      for (int c = 0; c < 3; ++c)
        for (int co = 0; co < 64; ++co)
          wf[(tap * 3 + c) * 64 + co] = w->data[(tap * 3 + c) * 64 + co] * scale[co];
    // packed fp16 B operand: [tap=(kt,kh)][co][j=(kw8,c4)], kw==7 and c==3 are zero columns
    std::vector<uint16_t> sp(static_cast<size_t>(64) * 49 * 32, 0);
    for (int kt = 0; kt < 7; ++kt)
      for (int kh = 0; kh < 7; ++kh)
        for (int kw = 0; kw < 7; ++kw)
          for (int c = 0; c < 3; ++c)
            for (int co = 0; co < 64; ++co)
              sp[(static_cast<size_t>(kt * 7 + kh) * 64 + co) * 32 + kw * 4 + c] =
                  f32_to_f16_bits(wf[(((kt * 7 + kh) * 7 + kw) * 3 + c) * 64 + co]);
    FAV_CUDA(cudaMemcpy(h->stem_w, sp.data(), sp.size() * 2, cudaMemcpyHostToDevice));
    // The delta path (bias table, gradient collapse, saturation corrections) keeps the folded
    // weights in fp32: delta never passes through a 16-bit rounding.
    const std::vector<float>& wq = wf;
    h->stem_w_host = wf;
    {
      std::vector<uint16_t> dpk;
      for (auto& d : h->rn.stem_dg) {
        dpk.resize(d.elems);
        pack_weights_taps(dpk.data(), wf.data(), nullptr, d.src.data(), static_cast<int>(d.src.size()), 3, 64, 64, 16, true);
        FAV_CUDA(cudaMemcpy(d.w, dpk.data(), dpk.size() * 2, cudaMemcpyHostToDevice));
      }
    }
    if (h->stem_gw) {
      std::vector<uint16_t> gw(static_cast<size_t>(7) * 160 * 64);
      stem_grad_pack_weights(gw.data(), wq.data(), 7, 64);
      FAV_CUDA(cudaMemcpy(h->stem_gw, gw.data(), gw.size() * 2, cudaMemcpyHostToDevice));
    }
    // class-summed weights: Wc[kt][hc][wc][c][co] = sum over kh valid for hc, kw valid for wc
    std::vector<float> wcs(static_cast<size_t>(7) * 16 * 3 * 64, 0.0f);
    auto valid = [](int cls, int n_out, int k, int pad, int n_in) {
      const int o = cls == 0 ? 0 : (cls == 1 ? 1 : (cls == 2 ? n_out - 2 : n_out - 1));
      const int i = 2 * o + k - pad;
      return i >= 0 && i < n_in;
    };
    for (int kt = 0; kt < 7; ++kt)
      for (int hc = 0; hc < 4; ++hc)
        for (int wc = 0; wc < 4; ++wc)
          for (int kh = 0; kh < 7; ++kh) {
            if (!valid(hc, h->Ho, kh, h->ph, h->H)) continue;
            for (int kw = 0; kw < 7; ++kw) {
              if (!valid(wc, h->Wo, kw, h->pw, h->W)) continue;
              for (int c = 0; c < 3; ++c)
                for (int co = 0; co < 64; ++co)
                  wcs[((kt * 16 + hc * 4 + wc) * 3 + c) * 64 + co] += wq[(((kt * 7 + kh) * 7 + kw) * 3 + c) * 64 + co];
            }
          }
    FAV_CUDA(cudaMemcpy(h->stem_wc, wcs.data(), wcs.size() * 4, cudaMemcpyHostToDevice));
    FAV_CUDA(cudaMemcpy(h->stem_bnbias, bias.data(), 64 * 4, cudaMemcpyHostToDevice));
  }
  // ---- head ----
  {
    const std::string unit = root + "Logits/Conv3d_0c_1x1/conv_3d";
    const fav_tensor* w = nt.find(unit + "/w");
    const fav_tensor* b = nt.find(unit + "/b");
    if (!w || !b) {
      set_error("missing tensor %s/{w,b}", unit.c_str());
      return FAV_ERR_MISSING;
    }
    const int C5 = h->bufs[h->final_buf].C;
    FAV_CHECK_ARG(numel(w) == static_cast<int64_t>(C5) * h->K && numel(b) == h->K, "logits weight shape mismatch");
    FAV_CUDA(cudaMemcpy(h->head_w, w->data, static_cast<size_t>(C5) * h->K * 4, cudaMemcpyHostToDevice));
    FAV_CUDA(cudaMemcpy(h->head_b, b->data, static_cast<size_t>(h->K) * 4, cudaMemcpyHostToDevice));
  }
  h->weights_loaded = true;
  return FAV_OK;
}

extern "C" int fav_apply_flicker(fav_handle* h, const void* clip, int in_dtype, const float* delta,
                                 float adv_flag, float delta_clip, uint8_t* adv_u8, float* adv_f32,
                                 void* stream) {
  FAV_CHECK_ARG(h && clip && delta, "fav_apply_flicker: null argument");
  FAV_NOT_EVAL(h, "fav_apply_flicker");
  FAV_CUDA(cudaSetDevice(h->device));   // a handle is bound to its device, whatever the caller's current device is
  g_pdl_on = h->pdl;
  FAV_CHECK_ARG(in_dtype == FAV_U8 || in_dtype == FAV_F32, "fav_apply_flicker: bad dtype");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  h->last_adv_flag = adv_flag;
  h->last_delta_clip = delta_clip;
  h->last_delta = delta;
  h->last_clip_u8 = in_dtype == FAV_U8 ? static_cast<const uint8_t*>(clip) : nullptr;
  if (h->d.arch != FAV_NET_I3D) {
    // torch stack: Perturbation.forward (model.py:80-96); adv_f32 is NCTHW like the reference's tensors
    FAV_CHECK_ARG(in_dtype == FAV_U8 && adv_u8 == nullptr, "torch-stack apply takes a uint8 clip and has no uint8 output");
    // Perturbation.forward returns x untouched when `adversarial` is False (model.py:82-83): the clean forward has no
    // range clamp (dark / bright pixels lie outside the scalar bounds [-1.735, 2.49])
    fav_norm_params nrm = h->nrm;
    if (adv_flag == 0.0f) { nrm.lo = -INFINITY; nrm.hi = INFINITY; }
    FAV_TRY(launch_apply_torch(static_cast<const uint8_t*>(clip), delta, adv_flag, delta_clip, nrm, h->xpad, h->Wp,
                               h->pw, adv_f32, h->pass_bits, h->B, h->T, h->H, h->W, s));
    const int C1 = round_up(h->rn.stem_C, 16);
    float cst[3], ds[3];
    for (int c = 0; c < 3; ++c) {
      cst[c] = (128.0f / 255.0f - h->nrm.mean[c]) / h->nrm.std[c];   // the stem input is u - 128
      ds[c] = 1.0f / h->nrm.std[c];
    }
    FAV_TRY(launch_stem_bias_ex(delta, adv_flag, delta_clip, h->stem_wc, h->stem_bnbias, h->stem_bias_tab, h->T, h->To,
                                h->pt, h->rn.stem_KT, 1, C1, cst, ds, s));
    return FAV_OK;
  }
  FAV_TRY(launch_apply(clip, in_dtype, delta, adv_flag, delta_clip, h->xpad, h->Wp, h->pw, adv_u8, adv_f32, h->pass_bits,
                       h->B, h->T, h->H, h->W, s));
  FAV_TRY(launch_stem_bias(delta, adv_flag, delta_clip, h->stem_wc, h->stem_bnbias, h->stem_bias_tab, h->T,
                           h->To, h->pt, s));
  return FAV_OK;
}

static int run_block_fwd(fav_handle* h, const Block& b, cudaStream_t s) {
  const bool par = h->branch_streams && !g_prof_on;   // per-family timing brackets every launch on one stream
  cudaStream_t s_pool = par ? h->side[0] : s, s_b2 = par ? h->side[1] : s;
  const PoolOp& p = h->pools[b.pool];
  if (par) FAV_TRY(branch_fork(h, s, 0));
  // Branch_3: pool -> 1x1x1 (reads only the block input)
  FAV_TRY(launch_maxpool_fwd(h->bufs[p.in].p, h->bufs[p.out].p, h->bufs[p.out].idx, p.g, s_pool));
  FAV_TRY(conv_launch(h->convs[b.b3b].fwd, s_pool));
  if (b.fused) {
    FAV_TRY(conv_launch(b.fwd, s));
  } else {
    FAV_TRY(conv_launch(h->convs[b.b0].fwd, s));
    FAV_TRY(conv_launch(h->convs[b.b1a].fwd, s));
    FAV_TRY(conv_launch(h->convs[b.b2a].fwd, s));
  }
  if (par) FAV_TRY(branch_fork(h, s, 1));
  FAV_TRY(conv_launch(h->convs[b.b2b].fwd, s_b2));
  FAV_TRY(conv_launch(h->convs[b.b1b].fwd, s));
  if (par) {
    FAV_TRY(branch_join(h, s, 0));
    FAV_TRY(branch_join(h, s, 1));
  }
  return FAV_OK;
}

static int run_pool_fwd(fav_handle* h, int pid, cudaStream_t s) {
  const PoolOp& p = h->pools[pid];
  return launch_maxpool_fwd(h->bufs[p.in].p, h->bufs[p.out].p, h->bufs[p.out].idx, p.g, s);
}

extern "C" int fav_forward(fav_handle* h, float* logits, void* stream) {
  FAV_CHECK_ARG(h, "fav_forward: null handle");
  FAV_CUDA(cudaSetDevice(h->device));
  g_pdl_on = h->pdl;
  if (!h->weights_loaded) {
    set_error("fav_forward: weights not loaded");
    return FAV_ERR_STATE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (h->d.arch != FAV_NET_I3D) {
    FAV_TRY(resnet_forward(h, s));
    if (logits)
      FAV_CUDA(cudaMemcpyAsync(logits, h->logits, static_cast<size_t>(h->B) * h->K * 4, cudaMemcpyDeviceToDevice, s));
    return FAV_OK;
  }
  FAV_TRY(stem_launch(h->stem_fwd, s));
  if (h->eval) FAV_TRY(stem_launch(h->stem_fwd2, s));
  FAV_TRY(run_pool_fwd(h, h->pool2a, s));
  FAV_TRY(conv_launch(h->convs[h->conv2b].fwd, s));
  FAV_TRY(conv_launch(h->convs[h->conv2c].fwd, s));
  FAV_TRY(run_pool_fwd(h, h->pool3a, s));
  FAV_TRY(run_block_fwd(h, h->blocks[0], s));
  FAV_TRY(run_block_fwd(h, h->blocks[1], s));
  FAV_TRY(run_pool_fwd(h, h->pool4a, s));
  for (int i = 2; i < 7; ++i) FAV_TRY(run_block_fwd(h, h->blocks[i], s));
  FAV_TRY(run_pool_fwd(h, h->pool5a, s));
  FAV_TRY(run_block_fwd(h, h->blocks[7], s));
  FAV_TRY(run_block_fwd(h, h->blocks[8], s));
  const Buf& fb = h->bufs[h->final_buf];
  FAV_TRY(launch_head_fwd(fb.p, h->B, fb.T, fb.H * fb.W, fb.C, h->feat, h->head_w, h->head_b, h->K,
                          h->logits, s));
  if (logits)
    FAV_CUDA(cudaMemcpyAsync(logits, h->logits, static_cast<size_t>(h->B) * h->K * 4, cudaMemcpyDeviceToDevice, s));
  return FAV_OK;
}

extern "C" int fav_loss(fav_handle* h, const int64_t* labels, const fav_loss_params* p, float* probs,
                        float* scalars, void* stream) {
  FAV_CHECK_ARG(h && labels && p && scalars, "fav_loss: null argument");
  FAV_NOT_EVAL(h, "fav_loss");
  FAV_CUDA(cudaSetDevice(h->device));
  return launch_loss(h->logits, labels, *p, h->B, h->K, probs, h->dlogits, scalars,
                     static_cast<cudaStream_t>(stream));
}

// ---- fused evaluation pass -------------------------------------------------------------
// kinetics_i3d.evaluate (utils/kinetics_i3d_utils.py:217-250) and the torch stack's validation phase (model.py:697-713)
// run every validation batch through the network twice (adv_flag 0 and 1); here both versions of the batch go through
// one forward-only pass and the miss / valid counters are accumulated on the device.
extern "C" int fav_eval_batch(fav_handle* h, const void* clip_clean, const void* clip_adv, int in_dtype, const float* delta,
                              float delta_clip, const int64_t* labels, int n_clips, int targeted, int64_t target_class,
                              int exclude_misclassify, const fav_loss_params* loss, int64_t* counts, float* probs,
                              float* scalars, void* stream) {
  FAV_CHECK_ARG(h && clip_clean && delta && labels && counts, "fav_eval_batch: null argument");
  if (!h->eval) {
    set_error("fav_eval_batch: needs a handle from fav_create_eval");
    return FAV_ERR_STATE;
  }
  if (!h->weights_loaded) {
    set_error("fav_eval_batch: weights not loaded");
    return FAV_ERR_STATE;
  }
  FAV_CHECK_ARG(in_dtype == FAV_U8 || in_dtype == FAV_F32, "fav_eval_batch: bad dtype");
  FAV_CHECK_ARG(n_clips >= 0 && n_clips <= h->Bu, "fav_eval_batch: n_clips %d outside [0, %d]", n_clips, h->Bu);
  FAV_CHECK_ARG((loss == nullptr) == (scalars == nullptr), "fav_eval_batch: loss and scalars go together");
  FAV_CUDA(cudaSetDevice(h->device));
  g_pdl_on = h->pdl;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!clip_adv) clip_adv = clip_clean;
  const int Bu = h->Bu;
  __half* x_adv = h->xpad + static_cast<size_t>(Bu) * h->T * h->H * h->Wp * 4;
  if (h->d.arch != FAV_NET_I3D) {
    FAV_CHECK_ARG(in_dtype == FAV_U8, "torch-stack evaluation takes uint8 clips");
    fav_norm_params clean = h->nrm;   // Perturbation.forward returns x untouched when adversarial is False (model.py:82-83)
    clean.lo = -INFINITY; clean.hi = INFINITY;
    FAV_TRY(launch_apply_torch(static_cast<const uint8_t*>(clip_clean), delta, 0.0f, delta_clip, clean, h->xpad, h->Wp, h->pw,
                               nullptr, nullptr, Bu, h->T, h->H, h->W, s));
    FAV_TRY(launch_apply_torch(static_cast<const uint8_t*>(clip_adv), delta, 1.0f, delta_clip, h->nrm, x_adv, h->Wp, h->pw,
                               nullptr, nullptr, Bu, h->T, h->H, h->W, s));
    const int C1 = round_up(h->rn.stem_C, 16);
    float cst[3], ds[3];
    for (int c = 0; c < 3; ++c) {
      cst[c] = (128.0f / 255.0f - h->nrm.mean[c]) / h->nrm.std[c];
      ds[c] = 1.0f / h->nrm.std[c];
    }
    FAV_TRY(launch_stem_bias_ex(delta, 0.0f, delta_clip, h->stem_wc, h->stem_bnbias, h->stem_bias_tab, h->T, h->To, h->pt,
                                h->rn.stem_KT, 1, C1, cst, ds, s));
    FAV_TRY(launch_stem_bias_ex(delta, 1.0f, delta_clip, h->stem_wc, h->stem_bnbias, h->stem_bias_tab2, h->T, h->To, h->pt,
                                h->rn.stem_KT, 1, C1, cst, ds, s));
  } else {
    FAV_TRY(launch_apply(clip_clean, in_dtype, delta, 0.0f, delta_clip, h->xpad, h->Wp, h->pw, nullptr, nullptr, nullptr, Bu,
                         h->T, h->H, h->W, s));
    FAV_TRY(launch_apply(clip_adv, in_dtype, delta, 1.0f, delta_clip, x_adv, h->Wp, h->pw, nullptr, nullptr, nullptr, Bu,
                         h->T, h->H, h->W, s));
    FAV_TRY(launch_stem_bias(delta, 0.0f, delta_clip, h->stem_wc, h->stem_bnbias, h->stem_bias_tab, h->T, h->To, h->pt, s));
    FAV_TRY(launch_stem_bias(delta, 1.0f, delta_clip, h->stem_wc, h->stem_bnbias, h->stem_bias_tab2, h->T, h->To, h->pt, s));
  }
  FAV_TRY(fav_forward(h, nullptr, stream));
  FAV_TRY(launch_eval_counts(h->logits, labels, Bu, n_clips, h->K, targeted, target_class, exclude_misclassify, counts, probs,
                             s));
  // the validation phase also logs the adversarial loss of the perturbed rows (model.py:706)
  if (loss)
    FAV_TRY(launch_loss(h->logits + static_cast<size_t>(Bu) * h->K, labels, *loss, Bu, h->K, nullptr, h->dlogits, scalars, s));
  return FAV_OK;
}

static int run_block_bwd(fav_handle* h, const Block& b, cudaStream_t s) {
  // gradient of the block output (already masked by out > 0) lives in bufs[b.out].g
  const bool par = h->branch_streams && !g_prof_on;
  cudaStream_t s_b2 = par ? h->side[0] : s, s_b3 = par ? h->side[1] : s;
  if (par) {
    FAV_TRY(branch_fork(h, s, 0));
    FAV_TRY(branch_fork(h, s, 1));
  }
  FAV_TRY(run_dgrad(h, b.b2b, true, false, s_b2));                 // -> g(b2a), masked by b2a > 0
  FAV_TRY(run_dgrad(h, b.b3b, false, false, s_b3));                // -> g(pool)
  FAV_TRY(run_dgrad(h, b.b1b, /*mask*/ true, /*acc*/ false, s));   // -> g(b1a), masked by b1a > 0
  if (par) {
    FAV_TRY(branch_join(h, s, 0));
    FAV_TRY(branch_join(h, s, 1));
  }
  if (b.fused) {
    FAV_TRY(conv_launch(b.dg, s));                                 // -> g(in): the three 1x1x1 data gradients at once
  } else {
    FAV_TRY(run_dgrad(h, b.b0, false, false, s));                  // -> g(in)  (first writer)
    FAV_TRY(run_dgrad(h, b.b1a, false, true, s));                  // +=
    FAV_TRY(run_dgrad(h, b.b2a, false, true, s));                  // +=
  }
  const PoolOp& p = h->pools[b.pool];
  const Buf& bi = h->bufs[b.in];
  // g(in) = (g(in) + maxpool^T(g(pool))) * (in > 0)
  FAV_TRY(launch_maxpool_bwd(h->bufs[p.out].g, h->bufs[p.out].idx, bi.g, bi.p, bi.g, p.g, s));
  return FAV_OK;
}

static int run_pool_bwd(fav_handle* h, int pid, cudaStream_t s) {
  const PoolOp& p = h->pools[pid];
  const Buf& bi = h->bufs[p.in];
  return launch_maxpool_bwd(h->bufs[p.out].g, h->bufs[p.out].idx, nullptr, bi.p, bi.g, p.g, s, h->bufs[p.out].p);
}

static int i3d_backward_to_stem(fav_handle* h, cudaStream_t s);

extern "C" int fav_backward_delta(fav_handle* h, float* grad, void* stream) {
  FAV_CHECK_ARG(h && grad, "fav_backward_delta: null argument");
  FAV_NOT_EVAL(h, "fav_backward_delta");
  FAV_CUDA(cudaSetDevice(h->device));
  g_pdl_on = h->pdl;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (h->d.arch != FAV_NET_I3D) return resnet_backward(h, grad, s);
  FAV_TRY(i3d_backward_to_stem(h, s));
  // stem: collapse over B,H,W without materialising dL/dx (stem_grad.cu)
  if (!h->stem_grad_dense) return stem_grad_launch(h->stem_gd, grad, s);
  // test switch: dense stem data gradient + masked reduce (an independent formulation of the same sum)
  FAV_CHECK_ARG(h->last_clip_u8 != nullptr, "FAV_STEM_GRAD_DENSE needs a uint8 clip");
  for (const DgradClass& d : h->rn.stem_dg) FAV_TRY(conv_launch(d.L, s));
  return launch_stem_dx_reduce(h->rn.dx, h->last_clip_u8, h->last_delta, h->last_adv_flag, h->last_delta_clip, h->nrm, 0,
                               h->rn.partial, grad, h->B, h->T, h->H, h->W, s);
}

static int i3d_backward_to_stem(fav_handle* h, cudaStream_t s) {
  const Buf& fb = h->bufs[h->final_buf];
  FAV_TRY(launch_head_bwd(h->dlogits, h->head_w, h->K, fb.p, fb.g, h->dfeat, h->B, fb.T, fb.H * fb.W, fb.C, s));
  FAV_TRY(run_block_bwd(h, h->blocks[8], s));
  FAV_TRY(run_block_bwd(h, h->blocks[7], s));
  FAV_TRY(run_pool_bwd(h, h->pool5a, s));
  for (int i = 6; i >= 2; --i) FAV_TRY(run_block_bwd(h, h->blocks[i], s));
  FAV_TRY(run_pool_bwd(h, h->pool4a, s));
  FAV_TRY(run_block_bwd(h, h->blocks[1], s));
  FAV_TRY(run_block_bwd(h, h->blocks[0], s));
  FAV_TRY(run_pool_bwd(h, h->pool3a, s));
  FAV_TRY(run_dgrad(h, h->conv2c, true, false, s));
  FAV_TRY(run_dgrad(h, h->conv2b, false, false, s));
  FAV_TRY(run_pool_bwd(h, h->pool2a, s));
  return FAV_OK;
}

// ---- sparse per-pixel attack (kinetics_i3d_L12, utils/kinetics_i3d_utils.py:308-521; torch attack_type "L12") ----
extern "C" int fav_pixels_enable(fav_handle* h) {
  FAV_CHECK_ARG(h, "fav_pixels_enable: null handle");
  FAV_NOT_EVAL(h, "fav_pixels_enable");
  if (h->pixels_enabled) return FAV_OK;
  if (!h->weights_loaded) {
    set_error("fav_pixels_enable: weights not loaded");
    return FAV_ERR_STATE;
  }
  FAV_CUDA(cudaSetDevice(h->device));
  FAV_TRY(dev_alloc(h, &h->zero_delta, static_cast<size_t>(h->T) * 3));
  FAV_TRY(dev_alloc(h, &h->pix_partial, static_cast<size_t>(pixels_partial_floats(h->T, h->H, h->W))));
  h->pixels_enabled = true;
  return FAV_OK;
}

extern "C" int fav_apply_pixels(fav_handle* h, const void* clip_u8, const float* delta_px, float adv_flag,
                                float delta_clip, float* adv_f32, void* stream) {
  FAV_CHECK_ARG(h && clip_u8 && delta_px, "fav_apply_pixels: null argument");
  FAV_CUDA(cudaSetDevice(h->device));
  g_pdl_on = h->pdl;
  if (!h->pixels_enabled) {
    set_error("fav_apply_pixels: call fav_pixels_enable first");
    return FAV_ERR_STATE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int torch_mode = h->d.arch != FAV_NET_I3D;
  h->last_adv_flag = adv_flag;
  h->last_delta_clip = delta_clip;
  h->last_delta_px = delta_px;
  h->last_delta = nullptr;
  h->last_clip_u8 = static_cast<const uint8_t*>(clip_u8);
  fav_norm_params nrm = h->nrm;
  if (torch_mode && adv_flag == 0.0f) { nrm.lo = -INFINITY; nrm.hi = INFINITY; }   // clean forward: no clamp (model.py:82-83)
  FAV_TRY(launch_apply_pixels(h->last_clip_u8, delta_px, adv_flag, delta_clip, nrm, torch_mode, h->xpad, h->Wp, h->pw,
                              adv_f32, h->B, h->T, h->H, h->W, s));
  if (torch_mode) {
    const int C1 = round_up(h->rn.stem_C, 16);
    float cst[3], ds[3];
    for (int c = 0; c < 3; ++c) {
      cst[c] = (128.0f / 255.0f - h->nrm.mean[c]) / h->nrm.std[c];
      ds[c] = 1.0f / h->nrm.std[c];
    }
    FAV_TRY(launch_stem_bias_ex(h->zero_delta, 0.0f, 1.0f, h->stem_wc, h->stem_bnbias, h->stem_bias_tab, h->T, h->To, h->pt,
                                h->rn.stem_KT, 1, C1, cst, ds, s));
  } else {
    FAV_TRY(launch_stem_bias(h->zero_delta, 0.0f, 1.0f, h->stem_wc, h->stem_bnbias, h->stem_bias_tab, h->T, h->To, h->pt, s));
  }
  return FAV_OK;
}

extern "C" int fav_backward_pixels(fav_handle* h, float* grad_px, void* stream) {
  FAV_CHECK_ARG(h && grad_px, "fav_backward_pixels: null argument");
  FAV_CUDA(cudaSetDevice(h->device));
  g_pdl_on = h->pdl;
  if (!h->pixels_enabled || !h->last_delta_px || !h->last_clip_u8) {
    set_error("fav_backward_pixels: no per-pixel apply preceded this call");
    return FAV_ERR_STATE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int torch_mode = h->d.arch != FAV_NET_I3D;
  if (torch_mode) {
    FAV_TRY(resnet_backward_to_dx(h, s));
  } else {
    FAV_TRY(i3d_backward_to_stem(h, s));
    FAV_TRY(run_dgrad_classes(h->rn.stem_dg, nullptr, nullptr, s, h));
  }
  return launch_stem_dx_pixels(h->rn.dx, h->last_clip_u8, h->last_delta_px, h->last_adv_flag, h->last_delta_clip, h->nrm,
                               torch_mode, grad_px, h->B, h->T, h->H, h->W, s);
}

extern "C" int fav_pixels_update(fav_handle* h, float* delta_px, const float* grad_px, float* m, float* v, int64_t* step,
                                 float reg_weight, float delta_clip, const fav_adam_params* adam, float* scalars,
                                 void* stream) {
  FAV_CHECK_ARG(h && delta_px && grad_px && m && v && step && adam && scalars, "fav_pixels_update: null argument");
  FAV_CUDA(cudaSetDevice(h->device));
  if (!h->pixels_enabled) {
    set_error("fav_pixels_update: call fav_pixels_enable first");
    return FAV_ERR_STATE;
  }
  return launch_pixels_update(delta_px, grad_px, m, v, step, h->pix_partial, reg_weight, delta_clip, *adam, scalars, h->T,
                              h->H, h->W, static_cast<cudaStream_t>(stream));
}

extern "C" int fav_delta_update(fav_handle* h, float* delta, const float* grad, float* m, float* v,
                                int64_t* step, const fav_reg_params* reg, const fav_adam_params* adam,
                                float* scalars, void* stream) {
  FAV_CHECK_ARG(h && delta && grad && m && v && step && reg && adam && scalars, "fav_delta_update: null argument");
  FAV_CUDA(cudaSetDevice(h->device));
  return launch_delta_update(delta, grad, m, v, step, *reg, *adam, h->last_adv_flag, scalars, h->T,
                             static_cast<cudaStream_t>(stream));
}

extern "C" uint16_t fav_debug_f32_to_f16(float f) { return f32_to_f16_bits(f); }

extern "C" int64_t fav_debug_read(fav_handle* h, const char* name, float* out, int64_t capacity, void* stream) {
  if (!h || !name || !out) {
    set_error("fav_debug_read: null argument");
    return FAV_ERR_ARG;
  }
  std::string n(name);
  bool grad = false;
  if (n.rfind("grad:", 0) == 0) {
    grad = true;
    n = n.substr(5);
  }
  for (const Buf& b : h->bufs) {
    if (b.name != n) continue;
    if (grad && !b.g) {
      set_error("fav_debug_read: '%s' has no gradient buffer (evaluation handle)", name);
      return FAV_ERR_STATE;
    }
    const long long npos = b.npos(h->B);
    const int64_t count = npos * b.C;
    if (count > capacity) {
      set_error("fav_debug_read: capacity %lld < %lld", (long long)capacity, (long long)count);
      return FAV_ERR_ARG;
    }
    int st = grad ? launch_bf16_to_f32(b.g, b.cs, 0, b.C, npos, out, static_cast<cudaStream_t>(stream))
                  : launch_f16_to_f32(b.p, b.cs, 0, b.C, npos, out, static_cast<cudaStream_t>(stream));
    return st == FAV_OK ? count : st;
  }
  set_error("fav_debug_read: unknown buffer '%s'", name);
  return FAV_ERR_ARG;
}

// =============================================================================================
// op-level entry points (parity tests)
// =============================================================================================
extern "C" int fav_op_conv3d(int device, const void* x, int64_t x_cs, int64_t x_coff, const float* w,
                             const float* bias, int kt, int kh, int kw, int cin, int cout, void* y,
                             int64_t y_cs, int64_t y_coff, int B, int T, int H, int W, int relu, int dgrad,
                             const void* relu_src, int64_t relu_cs, int64_t relu_coff, void* stream) {
  FAV_CHECK_ARG(x && w && y, "fav_op_conv3d: null argument");
  FAV_CUDA(cudaSetDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int taps = kt * kh * kw;
  // GEMM roles: forward K = cin, N = cout; dgrad K = cout, N = cin
  const int kc_real = dgrad ? cout : cin;
  const int n_real = dgrad ? cin : cout;
  const int kc = round_up(kc_real, 16);
  const int n_pad = round_up(n_real, 16);
  FAV_CHECK_ARG(kc == kc_real, "fav_op_conv3d: GEMM-K channel count %d must be a multiple of 16", kc_real);
  FAV_CHECK_ARG(n_real % 8 == 0, "fav_op_conv3d: GEMM-N channel count %d must be a multiple of 8", n_real);
  const size_t welems = static_cast<size_t>(n_pad) * taps * ceil_div(kc, 64) * 64;
  std::vector<uint16_t> pk(welems);
  if (dgrad) pack_weights_dgrad(pk.data(), w, nullptr, taps, cin, cout, kc, n_pad);
  else pack_weights_fwd(pk.data(), w, nullptr, taps, cin, kc, cout, n_pad);
  uint16_t* dw = nullptr;
  float* db = nullptr;
  FAV_CUDA(cudaMalloc(&dw, welems * 2));
  FAV_CUDA(cudaMemcpyAsync(dw, pk.data(), welems * 2, cudaMemcpyHostToDevice, s));
  if (bias && !dgrad) {
    std::vector<float> bp(n_pad, 0.0f);
    for (int i = 0; i < n_real; ++i) bp[i] = bias[i];
    FAV_CUDA(cudaMalloc(&db, n_pad * 4));
    FAV_CUDA(cudaMemcpyAsync(db, bp.data(), n_pad * 4, cudaMemcpyHostToDevice, s));
    FAV_CUDA(cudaStreamSynchronize(s));
  }
  ConvLaunch L;
  int st;
  if (use_halo(T, H, W, kt, kh, kw))
    st = conv_plan_halo(&L, device, x, x_cs, static_cast<int>(x_coff), kc, dw, n_pad, B, T, H, W, kt);
  else
    st = conv_plan_generic(&L, device, x, x_cs, static_cast<int>(x_coff), kc, dw, n_pad, B, T, H, W, kt, kh, kw,
                           taps == 1);
  if (st == FAV_OK) {
    // forward: fp16 in / fp16 out; data gradient: bf16 in / bf16 out, ReLU mask from an fp16 activation
    L.g.f16 = dgrad ? 0 : 1;
    L.e.out = static_cast<h16*>(y); L.e.out_f16 = dgrad ? 0 : 1; L.e.out_cs = y_cs; L.e.out_coff = static_cast<int>(y_coff);
    L.e.cout_store = n_real;
    L.e.bias = db; L.e.bias_ld = n_pad; L.e.bias_stem = 0; L.e.relu = relu;
    L.e.mask = static_cast<const h16*>(relu_src); L.e.mask_cs = relu_cs; L.e.mask_coff = static_cast<int>(relu_coff);
    L.e.addend = nullptr;
    st = conv_launch(L, s);
  }
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(dw);
  if (db) cudaFree(db);
  if (st != FAV_OK) return st;
  if (e != cudaSuccess) {
    set_error("fav_op_conv3d: kernel failed: %s", cudaGetErrorString(e));
    return FAV_ERR_CUDA;
  }
  return FAV_OK;
}

extern "C" int fav_op_maxpool3d(int device, const void* x, void* y, uint8_t* idx, int B, int T, int H, int W,
                                int C, int kt, int kh, int kw, int st, int sh, int sw, void* stream) {
  FAV_CHECK_ARG(x && y, "fav_op_maxpool3d: null argument");
  FAV_CUDA(cudaSetDevice(device));
  PoolGeom g = make_pool_geom(B, T, H, W, C, kt, kh, kw, st, sh, sw);
  return launch_maxpool_fwd(static_cast<const __half*>(x), static_cast<__half*>(y), idx, g,
                            static_cast<cudaStream_t>(stream));
}

extern "C" int fav_op_maxpool3d_bwd(int device, const void* dy, const uint8_t* idx, const void* add,
                                    const void* relu_src, void* dx, int B, int T, int H, int W, int C, int kt,
                                    int kh, int kw, int st, int sh, int sw, void* stream) {
  FAV_CHECK_ARG(dy && idx && dx, "fav_op_maxpool3d_bwd: null argument");
  FAV_CUDA(cudaSetDevice(device));
  PoolGeom g = make_pool_geom(B, T, H, W, C, kt, kh, kw, st, sh, sw);
  return launch_maxpool_bwd(static_cast<const __nv_bfloat16*>(dy), idx, static_cast<const __nv_bfloat16*>(add),
                            static_cast<const __half*>(relu_src), static_cast<__nv_bfloat16*>(dx), g,
                            static_cast<cudaStream_t>(stream));
}

extern "C" int fav_op_loss(int device, const float* logits, const int64_t* labels, const fav_loss_params* p, int B,
                           int K, float* probs, float* dlogits, float* scalars, void* stream) {
  FAV_CHECK_ARG(logits && labels && p && dlogits && scalars, "fav_op_loss: null argument");
  FAV_CUDA(cudaSetDevice(device));
  return launch_loss(logits, labels, *p, B, K, probs, dlogits, scalars, static_cast<cudaStream_t>(stream));
}

extern "C" int fav_op_delta_update(int device, float* delta, const float* grad, float* m, float* v, int64_t* step,
                                   const fav_reg_params* reg, const fav_adam_params* adam, float adv_flag,
                                   float* scalars, int T, void* stream) {
  FAV_CHECK_ARG(delta && grad && m && v && step && reg && adam && scalars, "fav_op_delta_update: null argument");
  FAV_CUDA(cudaSetDevice(device));
  return launch_delta_update(delta, grad, m, v, step, *reg, *adam, adv_flag, scalars, T,
                             static_cast<cudaStream_t>(stream));
}
