"""ORACLE (test infrastructure, never the product path) — CPU restatement of the reference's
I3D flickering-attack step in plain PyTorch fp32/fp64.

PARITY: PARTLY PINNED.  The reference ships no tests, golden vectors or checkpoints for this path (SURVEY.md §4,
§8c) and TensorFlow 1.15 / dm-sonnet 1.23 cannot be installed here, so this file restates the published semantics
of those dependencies and follows the reference call sites line by line.  The FORWARD semantics (TF `SAME` padding
of strided Conv3D and max_pool3d, snt.BatchNorm inference, the VALID average pool and the mean over T' of the head,
the InceptionI3d topology) are pinned against an independent implementation: OpenCV's DNN module running the same
network as an ONNX graph agrees to 1e-4 relative on the logits of a 17-frame clip, and per op exactly / to 1e-5
(tests/test_cpu_oracle_opencv_pin.py).  The backward pass is torch autograd of that forward, checked
against central finite differences of the loss in float64 (tests/test_cpu_oracle.py).  The loss assembly is
pinned where the reference's two stacks coincide: probability-mode margin loss, untargeted CE loss, the three
regulariser terms and the Adam trajectory (up to the epsilon placement) against vectors produced by the reference's
own torch classes (tests/test_cpu_golden.py).  Still unpinned (restated from the TF documentation only):
tf.clip_by_value's gradient at the bounds, tf.train.AdamOptimizer's epsilon placement, the logits-mode margin and the
`logits - one_hot` selection quirk (utils/kinetics_i3d_utils.py:169, :271), which exist only in the TF stack.

  i3d.py:32-71        Unit3D  = Conv3D(SAME, no bias) -> BatchNorm(inference, eps 1e-3, no gamma) -> ReLU
  i3d.py:144-479      InceptionI3d topology, TF SAME max-pools, VALID avg-pool head, mean over T'
  utils/kinetics_i3d_utils.py:76-307   attack graph: delta clip, apply, softmax, selections,
                                       regularisers, metrics, improve/ce adversarial losses
  utils/pre_process_rgb_flow.py:234    x = uint8/128 - 1
  utils/stats_and_plot/stats_plots.py:57   uint8 view ((adv+1.0)*127.5).astype(uint8)
  i3d_adversarial_main_single_video_npy.py:56-59,79-84   loss assembly and Adam on delta

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import it.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# (name, c0, c1a, c1b, c2a, c2b, c3b) — i3d.py:194-457
BLOCKS = [
    ("Mixed_3b", 64, 96, 128, 16, 32, 32),
    ("Mixed_3c", 128, 128, 192, 32, 96, 64),
    ("Mixed_4b", 192, 96, 208, 16, 48, 64),
    ("Mixed_4c", 160, 112, 224, 24, 64, 64),
    ("Mixed_4d", 128, 128, 256, 24, 64, 64),
    ("Mixed_4e", 112, 144, 288, 32, 64, 64),
    ("Mixed_4f", 256, 160, 320, 32, 128, 128),
    ("Mixed_5b", 256, 160, 320, 32, 128, 128),
    ("Mixed_5c", 384, 192, 384, 48, 128, 128),
]
ROOT = "RGB/inception_i3d/"


def unit_list():
    """All 57 BN units as (scope, kernel, cin, cout) in execution order (i3d.py:168-457)."""
    units = [("Conv3d_1a_7x7", 7, 3, 64), ("Conv3d_2b_1x1", 1, 64, 64), ("Conv3d_2c_3x3", 3, 64, 192)]
    cin = 192
    for name, c0, c1a, c1b, c2a, c2b, c3b in BLOCKS:
        b2b = "Conv3d_0a_3x3" if name == "Mixed_5b" else "Conv3d_0b_3x3"   # i3d.py:418 naming quirk
        units += [
            (f"{name}/Branch_0/Conv3d_0a_1x1", 1, cin, c0),
            (f"{name}/Branch_1/Conv3d_0a_1x1", 1, cin, c1a),
            (f"{name}/Branch_1/Conv3d_0b_3x3", 3, c1a, c1b),
            (f"{name}/Branch_2/Conv3d_0a_1x1", 1, cin, c2a),
            (f"{name}/Branch_2/{b2b}", 3, c2a, c2b),
            (f"{name}/Branch_3/Conv3d_0b_1x1", 1, cin, c3b),
        ]
        cin = c0 + c1b + c2b + c3b
    return units


def same_pads(n, k, s):
    """TF SAME: out = ceil(n/s); pad_total = max((out-1)*s + k - n, 0); before = total//2."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _pad_ndhwc_as_ncdhw(x, k, s, value=0.0):
    T, H, W = x.shape[2:]
    pt, ph, pw = same_pads(T, k[0], s[0]), same_pads(H, k[1], s[1]), same_pads(W, k[2], s[2])
    return F.pad(x, (pw[0], pw[1], ph[0], ph[1], pt[0], pt[1]), value=value)


def conv3d_same(x, w_tf, stride=(1, 1, 1)):
    """x NCDHW; w_tf [kt,kh,kw,Cin,Cout] (snt.Conv3D layout, i3d.py:61-65)."""
    w = w_tf.permute(4, 3, 0, 1, 2).contiguous()
    k = tuple(w.shape[2:])
    return F.conv3d(_pad_ndhwc_as_ncdhw(x, k, stride), w, stride=stride)


def maxpool3d_same(x, k, s):
    """tf.nn.max_pool3d SAME: padding never wins (-inf)."""
    return F.max_pool3d(_pad_ndhwc_as_ncdhw(x, k, s, value=float("-inf")), kernel_size=k, stride=s)


STORAGE = {"fp16": torch.float16, "bf16": torch.bfloat16}


def _round_st(t, fmt=torch.bfloat16):
    """value-rounding to a 16-bit storage format with a straight-through gradient (the engine rounds
    activations between layers; its backward treats the rounding as identity)."""
    return t + (t.to(fmt).to(t.dtype) - t).detach()


def _bf16_round(t):
    return _round_st(t, torch.bfloat16)


class OracleI3D:
    """InceptionI3d(final_endpoint='Logits') forward, differentiable w.r.t. its input.

    emulate="fp16" (the engine's forward format since round 2) or "bf16" (round 1; `emulate_bf16=True` is
    kept as an alias) restates the ENGINE's arithmetic instead of the reference's fp32: BN folded into the
    weights, folded weights and every layer output rounded to the 16-bit storage format, fp32 accumulation,
    delta entering the stem in fp32.  It exists to separate kernel correctness from the precision effect of
    16-bit storage (ReLU masks of a random-init network flip under the rounding noise); see DESIGN.md §4."""

    def __init__(self, weights, dtype=torch.float32, emulate_bf16=False, emulate=None):
        self.dtype = dtype
        if emulate is None and emulate_bf16:
            emulate = "bf16"
        self.emulate = emulate or False
        self.fmt = STORAGE[emulate] if emulate else None
        self.w = {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in weights.items()}

    def folded(self, scope):
        w = self.w[ROOT + scope + "/conv_3d/w"]
        beta = self.w[ROOT + scope + "/batch_norm/beta"].reshape(-1)
        mean = self.w[ROOT + scope + "/batch_norm/moving_mean"].reshape(-1)
        var = self.w[ROOT + scope + "/batch_norm/moving_variance"].reshape(-1)
        scale = 1.0 / torch.sqrt(var + 1e-3)
        return w * scale, beta - mean * scale

    def unit(self, x, scope, stride=(1, 1, 1), delta_img=None):
        if self.emulate:
            wf, bias = self.folded(scope)
            wq = wf.to(self.fmt).to(self.dtype)
            y = conv3d_same(x, wq, stride) + bias.reshape(1, -1, 1, 1, 1)
            if delta_img is not None:   # stem: delta contributes through the unrounded folded weights
                y = y + conv3d_same(delta_img, wf, stride)
            return _round_st(F.relu(y), self.fmt)
        w = self.w[ROOT + scope + "/conv_3d/w"]
        y = conv3d_same(x, w, stride)
        beta = self.w[ROOT + scope + "/batch_norm/beta"].reshape(1, -1, 1, 1, 1)
        mean = self.w[ROOT + scope + "/batch_norm/moving_mean"].reshape(1, -1, 1, 1, 1)
        var = self.w[ROOT + scope + "/batch_norm/moving_variance"].reshape(1, -1, 1, 1, 1)
        y = (y - mean) / torch.sqrt(var + 1e-3) + beta     # snt.BatchNorm: eps=1e-3, no scale
        return F.relu(y)

    def forward_split(self, x_clean, delta, adv_flag=1.0, delta_clip=0.4, endpoints=None, raw_endpoints=None):
        """Engine-style evaluation (emulate modes only): the clean clip goes through the 16-bit stem
        operand, delta through the fp32 side path; saturated pixels carry clip(x+d)-d."""
        assert self.emulate
        d = adv_flag * torch.clamp(delta.reshape(-1, 1, 1, 3).to(self.dtype), -delta_clip, delta_clip)
        s = x_clean.to(self.dtype) + d
        adv = torch.clamp(s, -1.0, 1.0)
        sat = (s < -1.0) | (s > 1.0)
        # x' = x where the clip did not fire, clip(x+d) - d elsewhere; d/d(delta) of (x' + d) is the clip mask
        xprime = torch.where(sat, (adv - d).detach(), x_clean.to(self.dtype))
        xprime = xprime.to(self.fmt).to(self.dtype)
        dimg = torch.where(sat, d.detach().expand_as(s), d.expand_as(s))
        return self.forward(xprime, endpoints=endpoints, delta_img=dimg, raw_endpoints=raw_endpoints)

    def forward(self, x_ndhwc, endpoints=None, delta_img=None, raw_endpoints=None):
        """x [B,T,H,W,3] -> logits [B,400]; optionally records end points (NDHWC views in
        `endpoints`; the graph tensors themselves, NCDHW, in `raw_endpoints` for retain_grad)."""
        x = x_ndhwc.to(self.dtype).permute(0, 4, 1, 2, 3)
        if delta_img is not None:
            delta_img = delta_img.to(self.dtype).permute(0, 4, 1, 2, 3)

        def rec(name, t):
            if endpoints is not None:
                endpoints[name] = t.permute(0, 2, 3, 4, 1)
            if raw_endpoints is not None:
                raw_endpoints[name] = t
            return t

        net = rec("Conv3d_1a_7x7", self.unit(x, "Conv3d_1a_7x7", (2, 2, 2), delta_img=delta_img))
        net = rec("MaxPool3d_2a_3x3", maxpool3d_same(net, (1, 3, 3), (1, 2, 2)))
        net = rec("Conv3d_2b_1x1", self.unit(net, "Conv3d_2b_1x1"))
        net = rec("Conv3d_2c_3x3", self.unit(net, "Conv3d_2c_3x3"))
        net = rec("MaxPool3d_3a_3x3", maxpool3d_same(net, (1, 3, 3), (1, 2, 2)))
        for name, *_ in BLOCKS:
            if name == "Mixed_4b":
                net = rec("MaxPool3d_4a_3x3", maxpool3d_same(net, (3, 3, 3), (2, 2, 2)))
            if name == "Mixed_5b":
                net = rec("MaxPool3d_5a_2x2", maxpool3d_same(net, (2, 2, 2), (2, 2, 2)))
            b2b = "Conv3d_0a_3x3" if name == "Mixed_5b" else "Conv3d_0b_3x3"
            b0 = self.unit(net, f"{name}/Branch_0/Conv3d_0a_1x1")
            b1 = self.unit(self.unit(net, f"{name}/Branch_1/Conv3d_0a_1x1"), f"{name}/Branch_1/Conv3d_0b_3x3")
            b2 = self.unit(self.unit(net, f"{name}/Branch_2/Conv3d_0a_1x1"), f"{name}/Branch_2/{b2b}")
            b3 = self.unit(maxpool3d_same(net, (3, 3, 3), (1, 1, 1)), f"{name}/Branch_3/Conv3d_0b_1x1")
            net = rec(name, torch.cat([b0, b1, b2, b3], 1))
        # Logits (i3d.py:459-472): avg_pool [2,7,7] VALID s1, 1x1x1 conv + bias, squeeze, mean over T'
        net = F.avg_pool3d(net, kernel_size=(2, 7, 7), stride=1)
        w = self.w[ROOT + "Logits/Conv3d_0c_1x1/conv_3d/w"]
        b = self.w[ROOT + "Logits/Conv3d_0c_1x1/conv_3d/b"].reshape(-1)
        logits = conv3d_same(net, w) + b.reshape(1, -1, 1, 1, 1)
        logits = logits.squeeze(4).squeeze(3)          # [B,K,T']
        return logits.mean(dim=2)


# ------------------------------------------------------------------------------------------------
# attack graph (utils/kinetics_i3d_utils.py:76-307)
# ------------------------------------------------------------------------------------------------
def normalize_u8(clip_u8):
    """utils/pre_process_rgb_flow.py:234 — float32(uint8)/128 - 1."""
    return clip_u8.to(torch.float32) / 128.0 - 1.0


def apply_flicker(x, delta, adv_flag=1.0, delta_clip=0.4):
    """kinetics_i3d_utils.py:104-105,122,139-142 (mask == 1 at the default _IND_START/_IND_END).
    x [B,T,H,W,3], delta [T,1,1,3] (or [T,3])."""
    d = torch.clamp(delta.reshape(-1, 1, 1, 3), -delta_clip, delta_clip)
    return torch.clamp(x + adv_flag * d, -1.0, 1.0)


def quantize_u8(adv):
    """utils/stats_and_plot/stats_plots.py:57 — ((adv+1.0)*127.5).astype(np.uint8) in fp32."""
    a = adv.detach().to(torch.float32).numpy()
    return ((a + np.float32(1.0)) * np.float32(127.5)).astype(np.uint8)


def regularizers(delta):
    """kinetics_i3d_utils.py:177-200 on perturbation = eps_rgb [T,1,1,3] (raw, unclipped)."""
    p = delta.reshape(-1, 3)
    right = torch.roll(p, 1, 0)
    left = torch.roll(p, -1, 0)
    norm_reg = (p ** 2).mean() + 1e-12
    diff_norm_reg = ((p - right) ** 2).mean() + 1e-12
    laplacian_norm_reg = ((-2 * p + right + left) ** 2).mean() + 1e-12
    roughness = (p - right).abs().mean()
    thickness = p.abs().mean()
    return norm_reg, diff_norm_reg, laplacian_norm_reg, thickness, roughness


def selections(logits, labels):
    """kinetics_i3d_utils.py:152-169 (incl. the `logits - one_hot` quirk at :169)."""
    softmax = F.softmax(logits, dim=-1)
    one_hot = F.one_hot(labels, logits.shape[-1]).to(logits.dtype)
    label_prob = (softmax * one_hot).sum(-1)
    label_logits = (logits * one_hot).sum(-1)
    max_non_label_prob = (softmax - one_hot).max(-1).values
    max_non_label_logits = (logits - one_hot).max(-1).values
    return softmax, label_prob, label_logits, max_non_label_prob, max_non_label_logits


def improve_adversarial_loss(logits, labels, margin=0.05, targeted=False, use_logits=False):
    """kinetics_i3d_utils.py:253-288.  Returns (loss_total, to_min_prob, to_max_prob)."""
    softmax, label_prob, label_logits, mnl_prob, mnl_logits = selections(logits, labels)
    if targeted:
        if use_logits:
            to_min, to_max = mnl_logits, label_logits
            m = torch.log(1.0 + margin * (1.0 / label_prob))
        else:
            to_min, to_max, m = mnl_prob, label_prob, margin
        to_min_prob, to_max_prob = mnl_prob, label_prob
    else:
        if use_logits:
            to_min, to_max = label_logits, mnl_logits
            m = torch.log(1.0 + margin * (1.0 / (0.00001 + mnl_prob)))
        else:
            to_min, to_max, m = label_prob, mnl_prob, margin
        to_min_prob, to_max_prob = label_prob, mnl_prob
    l_2 = ((to_min - (to_max - m)) ** 2) / m
    l_3 = to_min - (to_max - m)
    adv = torch.maximum(torch.zeros_like(l_3), torch.minimum(l_2, l_3))
    return adv.sum(), to_min_prob, to_max_prob


def ce_adversarial_loss(logits, labels, targeted=False):
    """kinetics_i3d_utils.py:290-307."""
    softmax, label_prob, _, mnl_prob, _ = selections(logits, labels)
    if targeted:
        ce = F.cross_entropy(logits, labels, reduction="none")
        return ce.mean(), mnl_prob, label_prob
    ce = -torch.log(1 - label_prob + 1e-6)
    return ce.mean(), label_prob, mnl_prob


class TFAdam:
    """tf.train.AdamOptimizer(lr, 0.9, 0.999, 1e-8): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    var -= lr_t * m / (sqrt(v) + eps)   (SURVEY App. B.6)."""

    def __init__(self, shape, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, dtype=torch.float32):
        self.m = torch.zeros(shape, dtype=dtype)
        self.v = torch.zeros(shape, dtype=dtype)
        self.t = 0
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps

    def step(self, var, grad, lr=None):
        lr = self.lr if lr is None else lr
        self.t += 1
        self.m = self.b1 * self.m + (1 - self.b1) * grad
        self.v = self.b2 * self.v + (1 - self.b2) * grad * grad
        lr_t = lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        return var - lr_t * self.m / (torch.sqrt(self.v) + self.eps)


def attack_step(model, x, labels, delta, cfg, opt=None, data_grad_only=False):
    """One `sess.run([train_op, loss, ...])` of i3d_adversarial_main_single_video_npy.py:213-215.

    x [B,T,H,W,3] float (normalised), delta [T,3]; cfg keys: improve_loss, targeted, use_logits,
    margin, beta0..beta3, lr, delta_clip.  Returns a dict with the fetched scalars, the data
    gradient [T,3], the total gradient and (if opt) the updated delta."""
    dtype = model.dtype
    d = delta.detach().clone().to(dtype).requires_grad_(True)
    if getattr(model, "emulate", False):
        logits = model.forward_split(x, d, cfg.get("adv_flag", 1.0), cfg.get("delta_clip", 0.4))
    else:
        adv = apply_flicker(x.to(dtype), d, cfg.get("adv_flag", 1.0), cfg.get("delta_clip", 0.4))
        logits = model.forward(adv)
    if cfg.get("improve_loss", True):
        adv_loss, p_min, p_max = improve_adversarial_loss(
            logits, labels, cfg.get("margin", 0.05), cfg.get("targeted", False), cfg.get("use_logits", False))
    else:
        adv_loss, p_min, p_max = ce_adversarial_loss(logits, labels, cfg.get("targeted", False))
    (g_data,) = torch.autograd.grad(adv_loss, d, retain_graph=False)
    out = {"logits": logits.detach(), "softmax": F.softmax(logits.detach(), -1), "adv_loss": float(adv_loss.detach()),
           "grad_data": g_data.detach().reshape(-1, 3), "to_min_prob": p_min.detach(), "to_max_prob": p_max.detach()}
    if data_grad_only:
        return out
    d2 = delta.detach().clone().to(dtype).requires_grad_(True)
    norm_reg, diff_reg, lap_reg, thick, rough = regularizers(d2)
    reg = cfg["beta1"] * norm_reg + cfg["beta2"] * diff_reg + cfg["beta3"] * lap_reg
    (g_reg,) = torch.autograd.grad(cfg["beta0"] * reg, d2)
    g = g_data.reshape(-1, 3) + g_reg.reshape(-1, 3)
    out.update(norm_reg=float(norm_reg), diff_norm_reg=float(diff_reg), laplacian_norm_reg=float(lap_reg),
               thickness=float(thick), roughness=float(rough), reg_loss=float(reg),
               total_loss=float(adv_loss) + cfg["beta0"] * float(reg), grad=g.detach())
    pred = logits.detach().argmax(-1)
    out["is_adversarial"] = bool((pred == labels).all()) if cfg.get("targeted", False) else bool((pred != labels).all())
    if opt is not None:
        out["delta_new"] = opt.step(delta.detach().to(dtype).reshape(-1, 3), g.detach(), cfg.get("lr", 1e-3))
    return out


def sparse_attack_step(model, x, labels, delta_thwc, beta1=0.5, margin=0.05, lr=1e-3, opt=None):
    """kinetics_i3d_L12 (utils/kinetics_i3d_utils.py:308-521): eps [T,H,W,3] without the +-0.4 clip,
    adv = clip(x + eps, -1, 1), loss = adv_loss + beta_1 * (sum_t sqrt(mean_{h,w,c} eps_t^2) + 1e-12)
    (i3d_adversarial_main_universal.py:133)."""
    d = delta_thwc.detach().clone().requires_grad_(True)
    adv = torch.clamp(x + d.unsqueeze(0), -1.0, 1.0)
    logits = model.forward(adv)
    adv_loss, _, _ = improve_adversarial_loss(logits, labels, margin, False, False)
    l12 = torch.sum(torch.sqrt(torch.mean(d ** 2, dim=[1, 2, 3]))) + 1e-12
    (g_data,) = torch.autograd.grad(adv_loss, d, retain_graph=True)
    (g_tot,) = torch.autograd.grad(adv_loss + beta1 * l12, d)
    opt = opt or TFAdam(tuple(d.shape), lr=lr)
    return dict(logits=logits.detach(), adv_loss=float(adv_loss.detach()), l12=float(l12.detach()), grad_data=g_data.detach(),
                grad_total=g_tot.detach(), delta_new=opt.step(delta_thwc.detach(), g_tot.detach(), lr))
