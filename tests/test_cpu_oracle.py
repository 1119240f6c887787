"""not gpu: pin the oracle's own building blocks (the reference ships no golden vectors: SURVEY §8c)
against independent restatements and closed-form cases."""
import math

import numpy as np
import pytest
import torch

from flickering_adversarial_video_b200 import synthetic
from oracle import oracle_i3d as O


def test_same_pads_match_tf_rule():
    # TF SAME: out = ceil(n/s), extra padding goes at the END
    assert O.same_pads(224, 7, 2) == (2, 3)
    assert O.same_pads(90, 7, 2) == (2, 3)
    assert O.same_pads(112, 3, 2) == (0, 1)
    assert O.same_pads(45, 3, 2) == (1, 1)
    assert O.same_pads(23, 2, 2) == (0, 1)
    assert O.same_pads(14, 2, 2) == (0, 0)
    assert O.same_pads(28, 3, 1) == (1, 1)


def test_unit_list_matches_reference_topology():
    units = O.unit_list()
    assert len(units) == 57                       # 57 BN units + the logits conv = 58 convs (i3d.py)
    assert ("Mixed_5b/Branch_2/Conv3d_0a_3x3", 3, 32, 128) in units   # naming quirk i3d.py:418
    macs = 0
    shapes = {"Conv3d_1a_7x7": (45, 112, 112), "Conv3d_2b_1x1": (45, 56, 56), "Conv3d_2c_3x3": (45, 56, 56)}
    for scope, k, cin, cout in units:
        if scope in shapes:
            t, h, w = shapes[scope]
        elif scope.startswith("Mixed_3"):
            t, h, w = 45, 28, 28
        elif scope.startswith("Mixed_4"):
            t, h, w = 23, 14, 14
        else:
            t, h, w = 12, 7, 7
        macs += t * h * w * k ** 3 * cin * cout
    macs += 11 * 1024 * 400
    assert macs == 157_061_958_656                # BASELINE.md forward MACs per 90-frame clip


def test_uint8_quantisation_known_answers():
    # ((adv+1.0)*127.5).astype(uint8): -1 -> 0, 1 -> 255, truncation toward zero
    adv = torch.tensor([-1.0, 1.0, 0.0, 0.9921875, -0.9921875, 0.5])
    assert O.quantize_u8(adv).tolist() == [0, 255, 127, 254, 0, 191]
    x = O.normalize_u8(torch.tensor([0, 128, 255], dtype=torch.uint8))
    assert x.tolist() == [-1.0, 0.0, 0.9921875]


def test_apply_clip_and_gradient_masks_are_inclusive():
    # tf.clip_by_value passes the gradient at the bounds
    x = torch.tensor([-1.0, 0.0, 0.9921875]).reshape(1, 1, 1, 3, 1).expand(1, 1, 1, 3, 3).clone()
    d = torch.zeros(1, 3, requires_grad=True)
    adv = O.apply_flicker(x, d)
    adv.sum().backward()
    assert d.grad.tolist() == [[3.0, 3.0, 3.0]]   # x == -1 still passes
    d2 = torch.tensor([[0.4, -0.4, 0.5]], requires_grad=True)
    O.apply_flicker(torch.zeros(1, 1, 1, 1, 3), d2).sum().backward()
    assert d2.grad.tolist() == [[1.0, 1.0, 0.0]]  # |delta| == 0.4 passes, 0.5 does not


def test_regularisers_closed_form():
    T = 6
    d = torch.arange(T * 3, dtype=torch.float32).reshape(T, 3) * 0.01
    nr, dr, lr, th, ro = O.regularizers(d)
    p = d.numpy().astype(np.float64)
    right = np.roll(p, 1, 0)
    left = np.roll(p, -1, 0)
    assert math.isclose(float(nr), (p ** 2).mean() + 1e-12, rel_tol=1e-5)
    assert math.isclose(float(dr), ((p - right) ** 2).mean() + 1e-12, rel_tol=1e-5)
    assert math.isclose(float(lr), ((-2 * p + right + left) ** 2).mean() + 1e-12, rel_tol=1e-5)
    assert math.isclose(float(th), np.abs(p).mean(), rel_tol=1e-5)
    assert math.isclose(float(ro), np.abs(p - right).mean(), rel_tol=1e-5)


def test_improve_loss_branches():
    # d = p_y - (p_other - m):  d<0 -> 0 ; 0<d<m -> d^2/m ; d>=m -> d
    def loss_for(py, po, m=0.05):
        rest = (1 - py - po) / 398
        p = torch.full((1, 400), rest)
        p[0, 3] = py
        p[0, 7] = po
        logits = torch.log(p)
        l, pmin, pmax = O.improve_adversarial_loss(logits, torch.tensor([3]), margin=m)
        return float(l), float(pmin), float(pmax)
    l, pmin, pmax = loss_for(0.2, 0.5)
    assert l == 0.0 and math.isclose(pmin, 0.2, rel_tol=1e-5) and math.isclose(pmax, 0.5, rel_tol=1e-5)
    l, _, _ = loss_for(0.30, 0.32)          # d = 0.03 < m
    assert math.isclose(l, 0.03 ** 2 / 0.05, rel_tol=1e-3)
    l, _, _ = loss_for(0.6, 0.1)            # d = 0.55 >= m
    assert math.isclose(l, 0.55, rel_tol=1e-4)


def test_tf_adam_first_step_is_lr_times_sign():
    opt = O.TFAdam((4,), lr=1e-3)
    v = opt.step(torch.zeros(4), torch.tensor([1.0, -2.0, 0.5, -1e-3]))
    assert torch.allclose(v, torch.tensor([-1e-3, 1e-3, -1e-3, 1e-3]), rtol=1e-3)


@pytest.fixture(scope="module")
def small():
    w = synthetic.i3d_weights(0)
    clip = synthetic.clips_u8(1, 16, seed=1001)
    return w, clip


def test_head_is_linear_in_features(small):
    """logits = b + W^T * weighted mean — the identity the engine's head kernel relies on"""
    w, _ = small
    g = torch.Generator().manual_seed(1)
    y = torch.rand((1, 1024, 3, 7, 7), generator=g)
    net = torch.nn.functional.avg_pool3d(y, (2, 7, 7), 1)
    wl = torch.as_tensor(w[O.ROOT + "Logits/Conv3d_0c_1x1/conv_3d/w"]).reshape(1024, 400)
    bl = torch.as_tensor(w[O.ROOT + "Logits/Conv3d_0c_1x1/conv_3d/b"])
    ref = (torch.einsum("bcthw,ck->bkt", net, wl) + bl.reshape(1, -1, 1)).mean(2)
    coef = torch.tensor([1.0, 2.0, 1.0]).reshape(1, 1, 3, 1, 1)
    feat = (y * coef).sum((2, 3, 4)) / (49 * 2 * 2)
    assert torch.allclose(feat @ wl + bl, ref, rtol=1e-4, atol=1e-5)


def test_forward_and_autograd_agree_with_fp64(small):
    w, clip = small
    x = O.normalize_u8(clip)
    delta = synthetic.delta_uniform(16, seed=7, lo=-0.05, hi=0.05)
    m32, m64 = O.OracleI3D(w), O.OracleI3D(w, dtype=torch.float64)
    with torch.no_grad():
        labels = m32.forward(x).argmax(-1)
    cfg = dict(improve_loss=True, margin=0.05, beta0=1.0, beta1=0.5, beta2=0.5, beta3=0.5)
    a = O.attack_step(m32, x, labels, delta, cfg, data_grad_only=True)
    b = O.attack_step(m64, x.double(), labels, delta.double(), cfg, data_grad_only=True)
    assert float((a["logits"] - b["logits"].float()).abs().max()) < 1e-3
    ga, gb = a["grad_data"], b["grad_data"].float()
    assert float((ga * gb).sum() / (ga.norm() * gb.norm())) > 0.9999


def test_bf16_storage_limits_gradient_cosine(small):
    """Documents why the engine's 0.999 gradient gate is taken against the precision-matched
    oracle: the fp32 oracle re-run with bf16-rounded activations/weights (pure CPU, no engine)
    already sits at ~0.98-0.99 cosine against itself on a random-init network."""
    w, clip = small
    x = O.normalize_u8(clip)
    delta = synthetic.delta_uniform(16, seed=7, lo=-0.05, hi=0.05)
    m32, mq = O.OracleI3D(w), O.OracleI3D(w, emulate_bf16=True)
    with torch.no_grad():
        labels = m32.forward(x).argmax(-1)
    cfg = dict(improve_loss=True, margin=0.05, beta0=1.0, beta1=0.5, beta2=0.5, beta3=0.5)
    a = O.attack_step(m32, x, labels, delta, cfg, data_grad_only=True)
    b = O.attack_step(mq, x, labels, delta, cfg, data_grad_only=True)
    rel = float((a["logits"] - b["logits"]).abs().max() / a["logits"].abs().max())
    cos = float((a["grad_data"] * b["grad_data"]).sum() / (a["grad_data"].norm() * b["grad_data"].norm()))
    print(f"fp32 vs bf16-storage oracle: logits rel {rel:.3e}, gradient cosine {cos:.4f}")
    assert rel < 1e-2
    assert 0.95 < cos < 0.9995


def test_fp16_storage_gradient_cosine_bound(small):
    """The engine's round-2 arithmetic restated on the CPU (fp16 forward storage and weights, fp32 accumulation):
    10 mantissa bits put the end-to-end dL/d-delta cosine against the strict-fp32 gradient near 0.996 for one 16-frame
    clip - ten times closer than bf16 storage (0.968 above) and where cuDNN's TF32 convolutions put the reference's own
    network on a B200 (0.9955, profiles/r02_open_precision_1x16.txt) - but still short of 0.999: the direction of the
    gradient is a small residual of a cancelling sum over H x W on this random-init fixture."""
    w, clip = small
    x = O.normalize_u8(clip)
    delta = synthetic.delta_uniform(16, seed=7, lo=-0.05, hi=0.05)
    m32, mq = O.OracleI3D(w), O.OracleI3D(w, emulate="fp16")
    with torch.no_grad():
        labels = m32.forward(x).argmax(-1)
    cfg = dict(improve_loss=True, margin=0.05, beta0=1.0, beta1=0.5, beta2=0.5, beta3=0.5)
    a = O.attack_step(m32, x, labels, delta, cfg, data_grad_only=True)
    b = O.attack_step(mq, x, labels, delta, cfg, data_grad_only=True)
    rel = float((a["logits"] - b["logits"]).abs().max() / a["logits"].abs().max())
    cos = float((a["grad_data"] * b["grad_data"]).sum() / (a["grad_data"].norm() * b["grad_data"].norm()))
    print(f"fp32 vs fp16-storage oracle: logits rel {rel:.3e}, gradient cosine {cos:.4f}")
    assert rel < 2e-3
    assert 0.99 < cos < 0.9995


def test_split_stem_evaluation_equals_plain_evaluation(small):
    """conv(x' ) + conv_fp32(delta) == conv(clip(x+delta)) when nothing is rounded"""
    w, clip = small
    x = O.normalize_u8(synthetic.clips_u8_extreme(1, 16))
    delta = synthetic.delta_uniform(16, seed=3, lo=-0.3, hi=0.3)
    m = O.OracleI3D(w, dtype=torch.float64)
    ep_a, ep_b = {}, {}
    with torch.no_grad():
        m.forward(O.apply_flicker(x.double(), delta.double()), endpoints=ep_a)
        d = torch.clamp(delta.double().reshape(-1, 1, 1, 3), -0.4, 0.4)
        s = x.double() + d
        adv = torch.clamp(s, -1, 1)
        sat = (s < -1) | (s > 1)
        xprime = torch.where(sat, adv - d, x.double())
        y = O.conv3d_same((xprime + d.expand_as(s)).permute(0, 4, 1, 2, 3),
                          m.w[O.ROOT + "Conv3d_1a_7x7/conv_3d/w"], (2, 2, 2))
        y_ref = O.conv3d_same(adv.permute(0, 4, 1, 2, 3), m.w[O.ROOT + "Conv3d_1a_7x7/conv_3d/w"], (2, 2, 2))
    assert float((y - y_ref).abs().max()) < 1e-9


def test_framework_default_init_fixture():
    """the second fixture of SURVEY §8(d): Sonnet / torchvision default initialisers (small-signal regime)"""
    import numpy as np
    from flickering_adversarial_video_b200 import synthetic
    w, he = synthetic.i3d_weights_framework_default(0), synthetic.i3d_weights(0)
    assert set(w) == set(he) and all(w[k].shape == he[k].shape for k in w)
    k = "RGB/inception_i3d/Conv3d_2c_3x3/conv_3d/w"
    sigma = (27 * 64) ** -0.5
    assert np.abs(w[k]).max() <= 2 * sigma + 1e-7 and abs(w[k].std() / sigma - 0.880) < 0.01     # truncated at 2 sigma
    assert (w["RGB/inception_i3d/Conv3d_2c_3x3/batch_norm/moving_variance"] == 1).all()
    m = synthetic.resnet_model_framework_default("r3d_18")
    assert not m.training and float(m.fc.weight.std()) < 0.05


def test_data_gradient_matches_finite_differences(small):
    """the oracle's dL/d-delta (what every engine gradient is compared with) against central finite differences of its own
    loss in float64 along two random directions: pins the backward chain (clip gradients, ReLU / max-pool routing, the
    sum over H x W) independently of autograd"""
    w, clip = small
    x = O.normalize_u8(clip).double()
    delta = synthetic.delta_uniform(16, seed=7, lo=-0.05, hi=0.05).double()
    m64 = O.OracleI3D(w, dtype=torch.float64)
    with torch.no_grad():
        labels = m64.forward(x).argmax(-1)
    cfg = dict(improve_loss=True, margin=0.05, beta0=1.0, beta1=0.5, beta2=0.5, beta3=0.5)
    g = O.attack_step(m64, x, labels, delta, cfg, data_grad_only=True)["grad_data"]

    def loss(d):
        with torch.no_grad():
            logits = m64.forward(O.apply_flicker(x, d))
            return float(O.improve_adversarial_loss(logits, labels, 0.05, False, False)[0])

    gen = torch.Generator().manual_seed(3)
    for _ in range(2):
        u = torch.randn((16, 3), generator=gen, dtype=torch.float64)
        u /= u.norm()
        h = 2e-6                     # few ReLU / pool / clip decisions flip between the two points; float64 keeps the difference exact enough
        fd = (loss(delta + h * u) - loss(delta - h * u)) / (2 * h)
        an = float((g * u).sum())
        print(f"finite difference {fd:.6e} vs analytic {an:.6e}")
        assert abs(fd - an) <= 1e-2 * max(abs(an), float(g.norm()) * 0.05), (fd, an)
