// resnet_impl.cuh — builder, weight ingest and forward / backward runners of the torchvision video ResNets
// (included by fav_api.cu after the handle definition; see resnet_plan.cuh).
#pragma once

namespace {

int conv_out(int in, int k, int s, int p) { return (in + 2 * p - k) / s + 1; }

// ---- data-gradient planning: one stride-1 GEMM per input-parity class ------------------------------------
int plan_dgrad_classes(fav_handle* h, std::vector<DgradClass>* out, const __nv_bfloat16* gout, long long gout_cs,
                       int kch, int To, int Ho, int Wo, __nv_bfloat16* gin, long long gin_cs, int n_pad, int T, int H,
                       int W, int kt, int kh, int kw, int st, int sh, int sw, int pt, int ph, int pw,
                       double flops_cin = 0.0, double flops_cout = 0.0) {
  const std::vector<DimClass> ct = dim_classes(kt, st, pt, T), chh = dim_classes(kh, sh, ph, H),
                              cw = dim_classes(kw, sw, pw, W);
  for (const DimClass& a : ct)
    for (const DimClass& b : chh)
      for (const DimClass& c : cw) {
        if (a.nk == 0 || b.nk == 0 || c.nk == 0 || a.Q == 0 || b.Q == 0 || c.Q == 0) continue;
        DgradClass d;
        d.class0 = a.par == 0 && b.par == 0 && c.par == 0;
        for (int i = 0; i < a.nk; ++i)
          for (int j = 0; j < b.nk; ++j)
            for (int k = 0; k < c.nk; ++k) d.src.push_back((a.ksrc[i] * kh + b.ksrc[j]) * kw + c.ksrc[k]);
        const int ntaps = static_cast<int>(d.src.size());
        d.elems = static_cast<size_t>(n_pad) * ntaps * ceil_div(kch, 64) * 64;
        FAV_TRY(dev_alloc(h, &d.w, d.elems));
        ConvSpec sp{};
        sp.x = gout; sp.x_cs = gout_cs; sp.x_coff = 0; sp.cin = kch;
        sp.aT = To; sp.aH = Ho; sp.aW = Wo;
        sp.wpk = d.w; sp.cout_pad = n_pad;
        sp.B = h->B; sp.T = a.Q; sp.H = b.Q; sp.W = c.Q;
        sp.kt = a.nk; sp.kh = b.nk; sp.kw = c.nk;
        sp.ot = a.off0; sp.oh = b.off0; sp.ow = c.off0;
        sp.st = sp.sh = sp.sw = 1;
        sp.oT = T; sp.oH = H; sp.oW = W;
        sp.est = st; sp.esh = sh; sp.esw = sw;
        sp.eot = a.par; sp.eoh = b.par; sp.eow = c.par;
        FAV_TRY(conv_plan_ex(&d.L, h->device, sp));
        ConvEpilogue& e = d.L.e;
        e.out = as16(gin); e.out_f16 = 0; e.out_cs = gin_cs; e.out_coff = 0; e.cout_store = n_pad;
        e.bias = nullptr; e.bias_ld = 0; e.bias_stem = 0; e.relu = 0; e.mask = nullptr; e.addend = nullptr;
        d.L.flops = 2.0 * static_cast<double>(h->B) * a.Q * b.Q * c.Q * ntaps * flops_cin * flops_cout;
        out->push_back(std::move(d));
      }
  return FAV_OK;
}

// The parity classes of a strided data gradient write disjoint output positions: with a handle they are spread over the
// main and the two side streams (parallel nodes of a captured graph) — each class is a small launch that rarely fills
// the GPU on its own.
int run_dgrad_classes(const std::vector<DgradClass>& cls, const Buf* mask, const Buf* addend, cudaStream_t s,
                      fav_handle* h = nullptr) {
  const bool par = h && h->branch_streams && !g_prof_on && cls.size() >= 2;
  if (par) {
    FAV_TRY(branch_fork(h, s, 0));
    if (cls.size() >= 3) FAV_TRY(branch_fork(h, s, 1));
  }
  const int nstreams = par ? (cls.size() >= 3 ? 3 : 2) : 1;
  int i = 0;
  for (const DgradClass& d : cls) {
    ConvLaunch L = d.L;
    if (mask) { L.e.mask = as16(mask->p); L.e.mask_cs = mask->cs; L.e.mask_coff = 0; }
    if (addend) { L.e.addend = as16(addend->g); L.e.add_f16 = 0; L.e.add_cs = addend->cs; L.e.add_coff = 0; }
    const int k = i++ % nstreams;
    FAV_TRY(conv_launch(L, k == 0 ? s : h->side[k - 1]));
  }
  if (par) {
    FAV_TRY(branch_join(h, s, 0));
    if (cls.size() >= 3) FAV_TRY(branch_join(h, s, 1));
  }
  return FAV_OK;
}

// ---- plan ------------------------------------------------------------------------------------------------
int rn_add_conv(fav_handle* h, const std::string& wname, const std::string& bnname, int kt, int kh, int kw, int st,
                int sh, int sw, int pt, int ph, int pw, int in, int cout, bool relu, const std::string& bufname) {
  const Buf bi = h->bufs[in];
  RConv c;
  c.wname = wname; c.bnname = bnname;
  c.kt = kt; c.kh = kh; c.kw = kw; c.st = st; c.sh = sh; c.sw = sw; c.pt = pt; c.ph = ph; c.pw = pw;
  c.in = in;
  c.cin_real = bi.C; c.cin_k = round_up(bi.C, 16);
  c.cout_real = cout; c.cout_pad = round_up(cout, 16);
  c.relu = relu;
  c.out = add_buf(h, bufname, conv_out(bi.T, kt, st, pt), conv_out(bi.H, kh, sh, ph), conv_out(bi.W, kw, sw, pw), cout,
                  false);
  if (c.out < 0) return -1;
  h->rn.convs.push_back(c);
  return static_cast<int>(h->rn.convs.size()) - 1;
}

int rn_plan_conv(fav_handle* h, RConv& c) {
  const Buf& bi = h->bufs[c.in];
  const Buf& bo = h->bufs[c.out];
  const Buf& bg = h->bufs[c.grad_src >= 0 ? c.grad_src : c.out];   // where the incoming gradient lives
  const int taps = c.kt * c.kh * c.kw;
  const bool s1 = c.st == 1 && c.sh == 1 && c.sw == 1;
  const bool k333 = c.kt == 3 && c.kh == 3 && c.kw == 3 && c.pt == 1 && c.ph == 1 && c.pw == 1;
  const bool k133 = c.kt == 1 && c.kh == 3 && c.kw == 3 && c.pt == 0 && c.ph == 1 && c.pw == 1;
  const bool halo = s1 && (k333 || k133) && use_halo(bi.T, bi.H, bi.W, c.kt, 3, 3);
  // ---- forward ----
  c.w_fwd_elems = static_cast<size_t>(c.cout_pad) * taps * ceil_div(c.cin_k, 64) * 64;
  FAV_TRY(dev_alloc(h, &c.w_fwd, c.w_fwd_elems));
  FAV_TRY(dev_alloc(h, &c.bias, static_cast<size_t>(c.cout_pad)));
  if (halo) {
    FAV_TRY(conv_plan_halo(&c.fwd, h->device, bi.p, bi.cs, 0, c.cin_k, c.w_fwd, c.cout_pad, h->B, bi.T, bi.H, bi.W, c.kt));
  } else {
    ConvSpec sp{};
    sp.x = bi.p; sp.x_cs = bi.cs; sp.x_coff = 0; sp.cin = c.cin_k;
    sp.aT = bi.T; sp.aH = bi.H; sp.aW = bi.W;
    sp.wpk = c.w_fwd; sp.cout_pad = c.cout_pad;
    sp.B = h->B; sp.T = bo.T; sp.H = bo.H; sp.W = bo.W;
    sp.kt = c.kt; sp.kh = c.kh; sp.kw = c.kw;
    sp.ot = -c.pt; sp.oh = -c.ph; sp.ow = -c.pw;
    sp.st = c.st; sp.sh = c.sh; sp.sw = c.sw;
    sp.oT = bo.T; sp.oH = bo.H; sp.oW = bo.W;
    sp.est = sp.esh = sp.esw = 1;
    FAV_TRY(conv_plan_ex(&c.fwd, h->device, sp));
    // Conv2Plus1D's temporal half with few output channels (the stem's 45 -> 64, layer1's 144 -> 64): every input frame
    // is loaded once for its three taps (conv_t3.cu) instead of three times through L2
    if (c.kt == 3 && c.kh == 1 && c.kw == 1 && s1 && c.pt == 1 && conv_t3_applicable(c.cin_k, c.cout_pad, bo.T)) {
      FAV_TRY(conv_t3_plan(&c.fwd.t3, h->device, bi.p, bi.cs, c.cin_k, c.w_fwd, c.cout_pad, h->B, bo.T, bo.H * bo.W, true));
      c.fwd.use_t3 = 1;
    }
  }
  {
    c.fwd.g.f16 = 1;
    ConvEpilogue& e = c.fwd.e;
    e.out = as16(bo.p); e.out_f16 = 1; e.out_cs = bo.cs; e.out_coff = 0; e.cout_store = c.cout_pad;
    e.bias = c.bias; e.bias_ld = c.cout_pad; e.bias_stem = 0;
    e.relu = (c.relu || c.residual >= 0) ? 1 : 0;
    e.mask = nullptr;
    e.addend = nullptr;
    if (c.residual >= 0) {
      const Buf& br = h->bufs[c.residual];
      FAV_CHECK_ARG(br.T == bo.T && br.H == bo.H && br.W == bo.W && br.cs == bo.cs, "residual shape mismatch at %s",
                    c.wname.c_str());
      e.addend = as16(br.p); e.add_f16 = 1; e.add_cs = br.cs; e.add_coff = 0;
    }
    c.fwd.flops = 2.0 * static_cast<double>(h->B) * bo.T * bo.H * bo.W * taps * c.cin_real * c.cout_real;
  }
  // ---- backward data ----
  c.halo_dg = halo;
  if (h->eval) return FAV_OK;
  if (halo) {
    DgradClass d;
    d.class0 = true;
    for (int j = 0; j < taps; ++j) d.src.push_back(taps - 1 - j);
    d.elems = static_cast<size_t>(c.cin_k) * taps * ceil_div(c.cout_pad, 64) * 64;
    FAV_TRY(dev_alloc(h, &d.w, d.elems));
    FAV_TRY(conv_plan_halo(&d.L, h->device, bg.g, bg.cs, 0, c.cout_pad, d.w, c.cin_k, h->B, bi.T, bi.H, bi.W, c.kt));
    ConvEpilogue& e = d.L.e;
    e.out = as16(bi.g); e.out_f16 = 0; e.out_cs = bi.cs; e.out_coff = 0; e.cout_store = c.cin_k;
    e.bias = nullptr; e.bias_ld = 0; e.bias_stem = 0; e.relu = 0; e.mask = nullptr; e.addend = nullptr;
    d.L.flops = 2.0 * static_cast<double>(h->B) * bo.T * bo.H * bo.W * taps * c.cin_real * c.cout_real;
    c.dg.push_back(std::move(d));
  } else {
    FAV_TRY(plan_dgrad_classes(h, &c.dg, bg.g, bg.cs, c.cout_pad, bo.T, bo.H, bo.W, bi.g, bi.cs, c.cin_k, bi.T, bi.H,
                               bi.W, c.kt, c.kh, c.kw, c.st, c.sh, c.sw, c.pt, c.ph, c.pw, c.cin_real, c.cout_real));
  }
  return FAV_OK;
}

int build_resnet(fav_handle* h) {
  const int B = h->B, T = h->T, H = h->H, W = h->W;
  ResNet& rn = h->rn;
  rn.arch = h->d.arch;
  FAV_CHECK_ARG(H % 2 == 0 && W % 16 == 0, "video ResNet engine needs even H and W %% 16 == 0 (got %dx%d)", H, W);
  const bool r21 = rn.arch == FAV_NET_R2PLUS1D_18;
  rn.stem_KT = r21 ? 1 : 3;
  rn.stem_pt = r21 ? 0 : 1;
  rn.stem_C = r21 ? 45 : 64;
  rn.stem_w = "stem.0"; rn.stem_bn = "stem.1";
  const int C1 = round_up(rn.stem_C, 16);
  // ---- stem: Conv3d(3, C, (KT,7,7), stride (1,2,2), padding (pt,3,3)) + BN + ReLU ----
  h->To = conv_out(T, rn.stem_KT, 1, rn.stem_pt);
  h->Ho = conv_out(H, 7, 2, 3);
  h->Wo = conv_out(W, 7, 2, 3);
  h->pt = rn.stem_pt; h->ph = 3; h->pw = 3;
  h->Wp = round_up(std::max(2 * (h->Wo - 1) + 8, W + h->pw), 2);
  FAV_TRY(dev_alloc(h, &h->xpad, static_cast<size_t>(B) * T * H * h->Wp * 4));
  FAV_TRY(dev_alloc(h, &h->stem_w, static_cast<size_t>(C1) * rn.stem_KT * 7 * 32));
  FAV_TRY(dev_alloc(h, &h->stem_wc, static_cast<size_t>(rn.stem_KT) * 16 * 3 * C1));
  FAV_TRY(dev_alloc(h, &h->stem_bnbias, static_cast<size_t>(C1)));
  FAV_TRY(dev_alloc(h, &h->stem_bias_tab, static_cast<size_t>(h->To) * 16 * C1));
  rn.stem_out = add_buf(h, "stem.conv", h->To, h->Ho, h->Wo, rn.stem_C, false);
  if (rn.stem_out < 0) return FAV_ERR_CUDA;
  FAV_TRY(plan_stem(h, rn.stem_out, C1, rn.stem_KT, 1, rn.stem_pt, 3, rn.stem_KT * 49.0 * 3.0 * rn.stem_C));
  int x = rn.stem_out;
  if (r21) {   // R2Plus1dStem: Conv3d(45, 64, (3,1,1), padding (1,0,0)) + BN + ReLU
    const int id = rn_add_conv(h, "stem.3", "stem.4", 3, 1, 1, 1, 1, 1, 1, 0, 0, x, 64, true, "stem");
    if (id < 0) return FAV_ERR_CUDA;
    rn.pre.push_back(id);
    x = rn.convs[id].out;
  }
  // ---- layers ----
  const int planes[4] = {64, 128, 256, 512};
  int inplanes = 64;
  for (int l = 0; l < 4; ++l) {
    for (int i = 0; i < 2; ++i) {
      const int s = (l > 0 && i == 0) ? 2 : 1;
      const std::string pre = "layer" + std::to_string(l + 1) + "." + std::to_string(i);
      RBlock blk;
      blk.in = x;
      int y = x;
      const bool no_temporal = rn.arch == FAV_NET_MC3_18 && l > 0;
      const bool need_ds = s != 1 || inplanes != planes[l];
      for (int cv = 1; cv <= 2; ++cv) {
        const int cs_ = cv == 1 ? s : 1;
        const std::string cp = pre + ".conv" + std::to_string(cv);
        const bool last = cv == 2;
        if (r21) {
          const int mid = (inplanes * planes[l] * 27) / (inplanes * 9 + 3 * planes[l]);
          int id = rn_add_conv(h, cp + ".0.0", cp + ".0.1", 1, 3, 3, 1, cs_, cs_, 0, 1, 1, y, mid, true, cp + ".0.0");
          if (id < 0) return FAV_ERR_CUDA;
          blk.chain.push_back(id);
          y = rn.convs[id].out;
          id = rn_add_conv(h, cp + ".0.3", cp + ".1", 3, 1, 1, cs_, 1, 1, 1, 0, 0, y, planes[l], !last,
                           last ? pre : cp);
          if (id < 0) return FAV_ERR_CUDA;
          blk.chain.push_back(id);
          y = rn.convs[id].out;
        } else {
          int id;
          if (no_temporal)
            id = rn_add_conv(h, cp + ".0", cp + ".1", 1, 3, 3, 1, cs_, cs_, 0, 1, 1, y, planes[l], !last, last ? pre : cp);
          else
            id = rn_add_conv(h, cp + ".0", cp + ".1", 3, 3, 3, cs_, cs_, cs_, 1, 1, 1, y, planes[l], !last,
                             last ? pre : cp);
          if (id < 0) return FAV_ERR_CUDA;
          blk.chain.push_back(id);
          y = rn.convs[id].out;
        }
      }
      int res = x;
      if (need_ds) {
        const int dt = no_temporal ? 1 : s;
        const int id = rn_add_conv(h, pre + ".downsample.0", pre + ".downsample.1", 1, 1, 1, dt, s, s, 0, 0, 0, x,
                                   planes[l], false, pre + ".downsample");
        if (id < 0) return FAV_ERR_CUDA;
        blk.ds = id;
        res = rn.convs[id].out;
      }
      rn.convs[blk.chain.back()].residual = res;
      if (blk.ds >= 0) rn.convs[blk.ds].grad_src = y;   // the shortcut's output gradient is the block output's
      blk.out = y;
      rn.blocks.push_back(blk);
      x = y;
      inplanes = planes[l];
    }
  }
  rn.final_buf = x;
  h->final_buf = x;
  for (auto& c : rn.convs) FAV_TRY(rn_plan_conv(h, c));
  // ---- dense stem data gradient: dX [B,T,H,W,16] ----
  if (!h->eval) {
    FAV_TRY(dev_alloc(h, &rn.dx, static_cast<size_t>(B) * T * H * W * 16));
    const Buf& bs = h->bufs[rn.stem_out];
    FAV_TRY(plan_dgrad_classes(h, &rn.stem_dg, bs.g, bs.cs, C1, bs.T, bs.H, bs.W, rn.dx, 16, 16, T, H, W, rn.stem_KT, 7, 7,
                               1, 2, 2, rn.stem_pt, 3, 3, 3.0, rn.stem_C));
    FAV_TRY(dev_alloc(h, &rn.partial, static_cast<size_t>(B) * T * stem_dx_reduce_chunks(H) * 3));
  }
  // torch-stack defaults (dataset.py:28-29; Perturbation scalar bounds model.py:72-75)
  const float mean[3] = {0.43216f, 0.394666f, 0.37645f}, sd[3] = {0.22803f, 0.22145f, 0.216989f};
  float lo = -1e30f, hi = 1e30f;
  for (int c = 0; c < 3; ++c) {
    h->nrm.mean[c] = mean[c]; h->nrm.std[c] = sd[c];
    lo = std::max(lo, (0.0f - mean[c]) / sd[c]);
    hi = std::min(hi, (1.0f - mean[c]) / sd[c]);
  }
  h->nrm.lo = lo; h->nrm.hi = hi;
  // ---- flicker gradient through the stem without dX (stem_grad.cu); delta enters the network as delta / std_c ----
  if (!h->eval) {
    const Buf& bs = h->bufs[rn.stem_out];
    const float sc[3] = {1.0f / h->nrm.std[0], 1.0f / h->nrm.std[1], 1.0f / h->nrm.std[2]};
    FAV_TRY(dev_alloc(h, &h->pass_bits, stem_grad_bitmap_words(B, T, H, W)));
    FAV_TRY(dev_alloc(h, &h->stem_gw, static_cast<size_t>(rn.stem_KT) * 160 * 64));
    FAV_TRY(stem_grad_plan(&h->stem_gd, h->device, bs.g, bs.cs, h->stem_gw, h->pass_bits, B, T, H, W, h->To, h->Ho, h->Wo,
                           rn.stem_KT, 1, rn.stem_pt, 3, 3, sc, rn.stem_C));
    h->stem_grad_dense = getenv("FAV_STEM_GRAD_DENSE") != nullptr;
  }
  // ---- head: AdaptiveAvgPool3d(1) + Linear(512, K) ----
  const int C5 = h->bufs[rn.final_buf].C;
  FAV_TRY(dev_alloc(h, &h->head_w, static_cast<size_t>(C5) * h->K));
  FAV_TRY(dev_alloc(h, &h->head_b, static_cast<size_t>(h->K)));
  FAV_TRY(dev_alloc(h, &h->feat, static_cast<size_t>(B) * C5));
  FAV_TRY(dev_alloc(h, &h->dfeat, static_cast<size_t>(B) * C5));
  FAV_TRY(dev_alloc(h, &h->logits, static_cast<size_t>(B) * h->K));
  FAV_TRY(dev_alloc(h, &h->dlogits, static_cast<size_t>(B) * h->K));
  return FAV_OK;
}

// ---- weights ---------------------------------------------------------------------------------------------
// torchvision BatchNorm3d inference fold: y = gamma * (x - mean) / sqrt(var + 1e-5) + beta
int bn_fold_torch(const NamedTensors& nt, const std::string& bn, int cout, std::vector<float>* scale,
                  std::vector<float>* bias) {
  const fav_tensor* g = nt.find(bn + ".weight");
  const fav_tensor* b = nt.find(bn + ".bias");
  const fav_tensor* m = nt.find(bn + ".running_mean");
  const fav_tensor* v = nt.find(bn + ".running_var");
  if (!g || !b || !m || !v) {
    set_error("missing BatchNorm3d tensors for %s", bn.c_str());
    return FAV_ERR_MISSING;
  }
  if (numel(g) != cout || numel(b) != cout || numel(m) != cout || numel(v) != cout) {
    set_error("BatchNorm3d size mismatch for %s (want %d)", bn.c_str(), cout);
    return FAV_ERR_ARG;
  }
  scale->assign(cout, 1.0f);
  bias->assign(cout, 0.0f);
  for (int c = 0; c < cout; ++c) {
    const float s = g->data[c] / std::sqrt(v->data[c] + 1e-5f);
    (*scale)[c] = s;
    (*bias)[c] = b->data[c] - m->data[c] * s;
  }
  return FAV_OK;
}

// torch Conv3d weight [cout][cin][kt][kh][kw] -> [tap][cin][cout]
int torch_weight_to_taps(const NamedTensors& nt, const std::string& name, int taps, int cin, int cout,
                         std::vector<float>* out) {
  const fav_tensor* w = nt.find(name + ".weight");
  if (!w) {
    set_error("missing tensor %s.weight", name.c_str());
    return FAV_ERR_MISSING;
  }
  if (numel(w) != static_cast<int64_t>(taps) * cin * cout) {
    set_error("weight %s.weight has %lld elements, expected %d*%d*%d", name.c_str(), (long long)numel(w), cout, cin, taps);
    return FAV_ERR_ARG;
  }
  out->resize(static_cast<size_t>(taps) * cin * cout);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int tp = 0; tp < taps; ++tp)
        (*out)[(static_cast<size_t>(tp) * cin + ci) * cout + co] = w->data[(static_cast<size_t>(co) * cin + ci) * taps + tp];
  return FAV_OK;
}

int load_weights_resnet(fav_handle* h, const NamedTensors& nt) {
  ResNet& rn = h->rn;
  std::vector<float> scale, bias, wt;
  std::vector<uint16_t> pk;
  for (auto& c : rn.convs) {
    const int taps = c.kt * c.kh * c.kw;
    FAV_TRY(torch_weight_to_taps(nt, c.wname, taps, c.cin_real, c.cout_real, &wt));
    FAV_TRY(bn_fold_torch(nt, c.bnname, c.cout_real, &scale, &bias));
    std::vector<int> ident(taps);
    for (int j = 0; j < taps; ++j) ident[j] = j;
    pk.resize(c.w_fwd_elems);
    pack_weights_taps(pk.data(), wt.data(), scale.data(), ident.data(), taps, c.cin_real, c.cout_real, c.cin_k, c.cout_pad,
                      false);
    FAV_CUDA(cudaMemcpy(c.w_fwd, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    for (auto& d : c.dg) {
      pk.resize(d.elems);
      pack_weights_taps(pk.data(), wt.data(), scale.data(), d.src.data(), static_cast<int>(d.src.size()), c.cin_real,
                        c.cout_real, c.cout_pad, c.cin_k, true);
      FAV_CUDA(cudaMemcpy(d.w, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    }
    std::vector<float> bpad(c.cout_pad, 0.0f);
    for (int i = 0; i < c.cout_real; ++i) bpad[i] = bias[i];
    FAV_CUDA(cudaMemcpy(c.bias, bpad.data(), bpad.size() * 4, cudaMemcpyHostToDevice));
  }
  // ---- stem ----
  {
    const int KT = rn.stem_KT, C = rn.stem_C, C1 = round_up(C, 16), taps = KT * 49;
    FAV_TRY(torch_weight_to_taps(nt, rn.stem_w, taps, 3, C, &wt));
    FAV_TRY(bn_fold_torch(nt, rn.stem_bn, C, &scale, &bias));
    for (int tp = 0; tp < taps; ++tp)
      for (int c = 0; c < 3; ++c)
        for (int co = 0; co < C; ++co) wt[(static_cast<size_t>(tp) * 3 + c) * C + co] *= scale[co];
    // forward operand in uint8 units: [tap=(kt,kh)][co][kw*4+c], scaled by 1/(255*std_c)
    std::vector<uint16_t> sp(static_cast<size_t>(KT) * 7 * C1 * 32, 0);
    for (int kt = 0; kt < KT; ++kt)
      for (int kh = 0; kh < 7; ++kh)
        for (int kw = 0; kw < 7; ++kw)
          for (int c = 0; c < 3; ++c)
            for (int co = 0; co < C; ++co)
              sp[(static_cast<size_t>(kt * 7 + kh) * C1 + co) * 32 + kw * 4 + c] = f32_to_f16_bits(
                  wt[(static_cast<size_t>((kt * 7 + kh) * 7 + kw) * 3 + c) * C + co] / (255.0f * h->nrm.std[c]));
    FAV_CUDA(cudaMemcpy(h->stem_w, sp.data(), sp.size() * 2, cudaMemcpyHostToDevice));
    // class-summed weights for the (delta - mean)/std bias table (fp32)
    const StemGeom& sg = h->stem_fwd.g;
    auto first_of = [](int cls, int n, int nlo, int nhi) { return cls < nlo ? cls : (cls == nlo ? nlo : n - nhi + (cls - nlo - 1)); };
    std::vector<float> wcs(static_cast<size_t>(KT) * 16 * 3 * C1, 0.0f);
    for (int kt = 0; kt < KT; ++kt)
      for (int hc = 0; hc < 4; ++hc)
        for (int wc = 0; wc < 4; ++wc) {
          if (hc > sg.nlo_h + sg.nhi_h || wc > sg.nlo_w + sg.nhi_w) continue;
          const int ho = first_of(hc, h->Ho, sg.nlo_h, sg.nhi_h), wo = first_of(wc, h->Wo, sg.nlo_w, sg.nhi_w);
          for (int kh = 0; kh < 7; ++kh) {
            const int ih = 2 * ho + kh - h->ph;
            if (ih < 0 || ih >= h->H) continue;
            for (int kw = 0; kw < 7; ++kw) {
              const int iw = 2 * wo + kw - h->pw;
              if (iw < 0 || iw >= h->W) continue;
              for (int c = 0; c < 3; ++c)
                for (int co = 0; co < C; ++co)
                  wcs[((static_cast<size_t>(kt) * 16 + hc * 4 + wc) * 3 + c) * C1 + co] +=
                      wt[(static_cast<size_t>((kt * 7 + kh) * 7 + kw) * 3 + c) * C + co];
            }
          }
        }
    FAV_CUDA(cudaMemcpy(h->stem_wc, wcs.data(), wcs.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> bpad(C1, 0.0f);
    for (int i = 0; i < C; ++i) bpad[i] = bias[i];
    FAV_CUDA(cudaMemcpy(h->stem_bnbias, bpad.data(), bpad.size() * 4, cudaMemcpyHostToDevice));
    if (h->stem_gw) {
      std::vector<uint16_t> gw(static_cast<size_t>(KT) * 160 * 64);
      stem_grad_pack_weights(gw.data(), wt.data(), KT, C);
      FAV_CUDA(cudaMemcpy(h->stem_gw, gw.data(), gw.size() * 2, cudaMemcpyHostToDevice));
    }
    // dense data gradient (x-space weights)
    for (auto& d : rn.stem_dg) {
      pk.resize(d.elems);
      pack_weights_taps(pk.data(), wt.data(), nullptr, d.src.data(), static_cast<int>(d.src.size()), 3, C, C1, 16, true);
      FAV_CUDA(cudaMemcpy(d.w, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    }
  }
  // ---- head ----
  {
    const fav_tensor* w = nt.find("fc.weight");
    const fav_tensor* b = nt.find("fc.bias");
    if (!w || !b) {
      set_error("missing tensor fc.{weight,bias}");
      return FAV_ERR_MISSING;
    }
    const int C5 = h->bufs[rn.final_buf].C;
    FAV_CHECK_ARG(numel(w) == static_cast<int64_t>(C5) * h->K && numel(b) == h->K, "fc shape mismatch");
    std::vector<float> wtr(static_cast<size_t>(C5) * h->K);
    for (int k = 0; k < h->K; ++k)
      for (int c = 0; c < C5; ++c) wtr[static_cast<size_t>(c) * h->K + k] = w->data[static_cast<size_t>(k) * C5 + c];
    FAV_CUDA(cudaMemcpy(h->head_w, wtr.data(), wtr.size() * 4, cudaMemcpyHostToDevice));
    FAV_CUDA(cudaMemcpy(h->head_b, b->data, static_cast<size_t>(h->K) * 4, cudaMemcpyHostToDevice));
  }
  return FAV_OK;
}

// ---- runners ---------------------------------------------------------------------------------------------
int resnet_forward(fav_handle* h, cudaStream_t s) {
  ResNet& rn = h->rn;
  FAV_TRY(stem_launch(h->stem_fwd, s));
  if (h->eval) FAV_TRY(stem_launch(h->stem_fwd2, s));
  for (int id : rn.pre) FAV_TRY(conv_launch(rn.convs[id].fwd, s));
  const bool par = h->branch_streams && !g_prof_on;
  for (const RBlock& b : rn.blocks) {
    // the 1x1x1 downsample shortcut only meets the main branch in the epilogue of the block's last conv
    const bool side = par && b.ds >= 0 && b.chain.size() >= 2;
    if (side) FAV_TRY(branch_fork(h, s, 0));
    if (b.ds >= 0) FAV_TRY(conv_launch(rn.convs[b.ds].fwd, side ? h->side[0] : s));
    for (size_t i = 0; i < b.chain.size(); ++i) {
      if (side && i + 1 == b.chain.size()) FAV_TRY(branch_join(h, s, 0));
      FAV_TRY(conv_launch(rn.convs[b.chain[i]].fwd, s));
    }
  }
  const Buf& fb = h->bufs[rn.final_buf];
  FAV_TRY(launch_head_fwd(fb.p, h->B, -fb.T, fb.H * fb.W, fb.C, h->feat, h->head_w, h->head_b, h->K, h->logits, s));
  return FAV_OK;
}

int resnet_backward_to_dx(fav_handle* h, cudaStream_t s, bool with_dx = true) {
  ResNet& rn = h->rn;
  const Buf& fb = h->bufs[rn.final_buf];
  FAV_TRY(launch_head_bwd(h->dlogits, h->head_w, h->K, fb.p, fb.g, h->dfeat, h->B, -fb.T, fb.H * fb.W, fb.C, s));
  for (int bi = static_cast<int>(rn.blocks.size()) - 1; bi >= 0; --bi) {
    const RBlock& b = rn.blocks[bi];
    const Buf& bin = h->bufs[b.in];
    const Buf& bout = h->bufs[b.out];
    for (int i = static_cast<int>(b.chain.size()) - 1; i >= 1; --i) {
      const RConv& c = rn.convs[b.chain[i]];
      FAV_TRY(run_dgrad_classes(c.dg, &h->bufs[c.in], nullptr, s, h));   // masked by the producer's ReLU
    }
    const RConv& c0 = rn.convs[b.chain[0]];
    if (b.ds >= 0) {
      // shortcut gradient first (reaches the class-(0,0,0) positions only), everything else accumulates on it
      FAV_CUDA(cudaMemsetAsync(bin.g, 0, static_cast<size_t>(bin.npos(h->B)) * bin.cs * 2, s));
      FAV_TRY(run_dgrad_classes(rn.convs[b.ds].dg, nullptr, nullptr, s, h));
      FAV_TRY(run_dgrad_classes(c0.dg, &bin, &bin, s, h));
    } else {
      FAV_TRY(run_dgrad_classes(c0.dg, &bin, &bout, s, h));   // identity shortcut: + g(out)
    }
  }
  for (int i = static_cast<int>(rn.pre.size()) - 1; i >= 0; --i) {
    const RConv& c = rn.convs[rn.pre[i]];
    FAV_TRY(run_dgrad_classes(c.dg, &h->bufs[c.in], nullptr, s, h));
  }
  if (with_dx) FAV_TRY(run_dgrad_classes(rn.stem_dg, nullptr, nullptr, s, h));   // dense dL/d(adv) -> rn.dx
  return FAV_OK;
}

int resnet_backward(fav_handle* h, float* grad, cudaStream_t s) {
  if (!h->stem_grad_dense) {
    FAV_TRY(resnet_backward_to_dx(h, s, /*with_dx=*/false));
    return stem_grad_launch(h->stem_gd, grad, s);
  }
  // test switch: dense stem data gradient + masked reduce (an independent formulation of the same sum)
  FAV_TRY(resnet_backward_to_dx(h, s, true));
  FAV_CHECK_ARG(h->last_clip_u8 != nullptr && h->last_delta != nullptr, "fav_backward_delta: apply a uint8 clip first");
  FAV_TRY(launch_stem_dx_reduce(h->rn.dx, h->last_clip_u8, h->last_delta, h->last_adv_flag, h->last_delta_clip, h->nrm, 1,
                                h->rn.partial, grad, h->B, h->T, h->H, h->W, s));
  return FAV_OK;
}

}  // namespace
