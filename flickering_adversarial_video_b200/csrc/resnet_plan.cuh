// resnet_plan.cuh — static execution plan of the torchvision video ResNets the reference's torch stack
// attacks (r3d_18 / mc3_18 / r2plus1d_18; utils_cv/action_recognition/model.py:403-441 builds them with
// torchvision.models.video).  Included by fav_api.cu only (shares its handle and helpers).
//
// Every Conv3d + BatchNorm3d (+ReLU) is one tcgen05 implicit-GEMM launch (BN folded at load, eps 1e-5, with
// gamma); the BasicBlock tail adds the residual in the epilogue before the ReLU.  The backward pass is the
// data-gradient chain only: stride-1 convs run the tap-flipped conv, strided convs run one stride-1 GEMM per
// input-parity class whose epilogue scatters into the class's positions (no zero-stuffed transposed conv).
#pragma once

namespace {

struct DimClass {
  int par = 0;        // input index parity (i % s)
  int nk = 0;         // taps of this class
  int off0 = 0;       // A coordinate of tap 0 for class position 0
  int Q = 0;          // class positions
  int ksrc[8] = {0};  // source kernel index of packed tap d
};

// input i = s*q + par receives from taps k == (par + pad) mod s; output index o = q + (par + pad - k)/s
static std::vector<DimClass> dim_classes(int K, int s, int pad, int I) {
  std::vector<DimClass> v;
  for (int par = 0; par < s; ++par) {
    DimClass d;
    d.par = par;
    const int c = (par + pad) % s;
    d.nk = c < K ? (K - c + s - 1) / s : 0;
    const int e = (par + pad - c) / s;
    d.off0 = e - (d.nk - 1);
    d.Q = par < I ? (I - par + s - 1) / s : 0;
    for (int k = 0; k < d.nk && k < 8; ++k) d.ksrc[k] = c + s * (d.nk - 1 - k);
    v.push_back(d);
  }
  return v;
}

struct DgradClass {
  ConvLaunch L;
  uint16_t* w = nullptr;
  size_t elems = 0;
  std::vector<int> src;   // source tap (kt,kh,kw flattened) of each packed tap
  bool class0 = false;    // parity (0,0,0): the only class a strided 1x1x1 shortcut reaches
};

struct RConv {
  std::string wname, bnname;   // state_dict prefixes
  int kt = 1, kh = 1, kw = 1, st = 1, sh = 1, sw = 1, pt = 0, ph = 0, pw = 0;
  int in = -1, out = -1;
  int cin_real = 0, cin_k = 0, cout_real = 0, cout_pad = 0;
  bool relu = true;
  int residual = -1;           // buffer added before the ReLU (BasicBlock tail)
  int grad_src = -1;           // buffer whose gradient feeds this conv's dgrad (shortcut convs: the block output)
  uint16_t* w_fwd = nullptr;
  size_t w_fwd_elems = 0;
  float* bias = nullptr;
  ConvLaunch fwd;
  bool halo_dg = false;
  std::vector<DgradClass> dg;
};

struct RBlock {
  std::vector<int> chain;      // conv ids; ReLU after each but the last, which adds the residual first
  int ds = -1;                 // downsample conv id (1x1x1 strided + BN) or -1
  int in = -1, out = -1;
};

struct ResNet {
  int arch = 0;
  int stem_KT = 3, stem_C = 64, stem_pt = 1;
  int stem_out = -1;           // buffer of the strided RGB stem conv
  std::string stem_w, stem_bn;
  std::vector<RConv> convs;
  std::vector<int> pre;        // convs between the stem and layer1 (r2plus1d: the (3,1,1) 45->64 conv)
  std::vector<RBlock> blocks;
  int final_buf = -1;
  // dense stem data gradient
  __nv_bfloat16* dx = nullptr; // [B,T,H,W,16]
  std::vector<DgradClass> stem_dg;
  float* partial = nullptr;
  std::vector<float> stem_w_tf; // host copy [KT*7*7][3][C] folded (x-space) for the class packers
};

}  // namespace
