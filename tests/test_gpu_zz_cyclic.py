"""gpu: the cyclic perturbation attack (torch stack `cyclic_pert`, model.py:91-92; TF stack `cyclic_pert_flag`,
utils/kinetics_i3d_utils.py:130-137): FlickerAttack.step_rolled against the oracle's autograd through the rolled
perturbation.  (Sorts last: added after the last GPU session of round 1.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_step_rolled_matches_oracle():
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200.attack import FlickerAttack
    from oracle import oracle_resnet
    B, T, max_norm, shift = 2, 8, 0.1, 3
    model = synthetic.resnet_model("r3d_18", seed=0)
    clip = synthetic.clips_u8(B, T, 112, 112, seed=1003)
    delta = synthetic.delta_uniform(T, seed=9, lo=-0.12, hi=0.12)         # exceeds max_norm: the clamp fires
    with torch.no_grad():
        labels = model(oracle_resnet.normalize_u8(clip)).argmax(-1)
    # d loss(roll(delta, s)) / d delta = roll^-1 of the gradient w.r.t. the rolled perturbation
    ref = oracle_resnet.attack_step(model, clip, labels, torch.roll(delta, shift, 0), max_norm=max_norm)
    want = torch.roll(ref["grad_data"], -shift, 0)

    cfg = {"LAMBDA": 1.0, "BETA_1": 0.5, "IMPROVE_ADV_LOSS": True, "PROB_MARGIN": 0.05}
    atk = FlickerAttack(model.state_dict(), B, T, cfg, arch="r3d_18", delta_clip=max_norm)
    atk.delta.copy_(delta)
    atk.step_rolled(clip.cuda(), labels.cuda(), shift)
    torch.cuda.synchronize()
    logits = atk.eng.logits.cpu()
    rel = float((logits - ref["logits"]).abs().max() / ref["logits"].abs().max())
    g = atk.grad.cpu()
    cos = float((g * want).sum() / (g.norm() * want.norm() + 1e-30))
    # reported only: how different the un-rolled gradient is (a temporally smooth gradient can be close to its roll)
    cos_wrong = float((g * ref["grad_data"]).sum() / (g.norm() * ref["grad_data"].norm() + 1e-30))
    print(f"step_rolled: logits rel err {rel:.3e}, cosine {cos:.6f} (against the un-rolled gradient {cos_wrong:.3f})")
    assert rel <= 1e-2 and cos >= 0.99
    # the updated delta: Adam's first step moves every unclamped entry by lr against the sign of its total gradient
    d_new = atk.delta.cpu()
    assert float((d_new - delta).abs().max()) <= 1e-3 * 1.001

    # shift = 0 is the plain step
    a0 = FlickerAttack(model.state_dict(), B, T, cfg, arch="r3d_18", delta_clip=max_norm)
    a0.delta.copy_(delta)
    a0.step(clip.cuda(), labels.cuda())
    a1 = FlickerAttack(model.state_dict(), B, T, cfg, arch="r3d_18", delta_clip=max_norm)
    a1.delta.copy_(delta)
    a1.step_rolled(clip.cuda(), labels.cuda(), 0)
    torch.cuda.synchronize()
    # (the stem-gradient kernel flushes per-plane sums with atomics: equal up to fp32 summation order)
    assert torch.allclose(a0.grad, a1.grad, rtol=1e-3, atol=1e-4 * float(a0.grad.abs().max()))
    assert float((a0.delta - a1.delta).abs().max()) <= 1e-5
    # predict(shift) evaluates the rolled perturbation
    p = atk.predict(clip.cuda(), adv_flag=1.0, shift=shift).cpu()
    atk2 = FlickerAttack(model.state_dict(), B, T, cfg, arch="r3d_18", delta_clip=max_norm)
    atk2.delta.copy_(torch.roll(atk.delta, shift, 0))
    assert torch.allclose(p, atk2.predict(clip.cuda(), adv_flag=1.0).cpu(), rtol=0, atol=1e-6)


def test_frame_range_mask():
    """`_IND_START` / `_IND_END` (utils/kinetics_i3d_utils.py:14-15,107-113): delta acts on a frame range only"""
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200.attack import FlickerAttack
    T, lo, hi = 16, 4, 9
    weights = synthetic.i3d_weights(seed=0)
    clips = synthetic.clips_u8(1, T, seed=1001).cuda()
    delta = synthetic.delta_uniform(T, seed=7, lo=-0.05, hi=0.05).cuda()
    mask = torch.zeros((T, 1), device="cuda")
    mask[lo:hi + 1] = 1.0
    atk = FlickerAttack(weights, 1, T, {}, frame_range=(lo, hi))
    atk.delta.copy_(delta)
    plain = FlickerAttack(weights, 1, T, {})
    plain.delta.copy_(delta * mask)
    assert torch.equal(atk.adversarial_video(clips, as_uint8=True), plain.adversarial_video(clips, as_uint8=True))
    assert FlickerAttack(weights, 1, T, {}, frame_range=(0, T)).frame_mask is None        # the reference's default
    labels = plain.predict(clips, adv_flag=0.0).argmax(-1)
    atk.step(clips, labels)
    torch.cuda.synchronize()
    g = atk.grad.cpu()
    assert float(g[:lo].abs().max()) == 0.0 and float(g[hi + 1:].abs().max()) == 0.0 and float(g[lo:hi + 1].abs().sum()) > 0.0


def test_on_demand_handles_of_the_tf_mirror():
    """`softmax_clean` and `adversarial_inputs_rgb` of kinetics_i3d (utils/kinetics_i3d_utils.py:139-149): evaluated for
    the input of the last run"""
    import numpy as np
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200.kinetics_i3d import kinetics_i3d
    T = 16
    k = kinetics_i3d(ckpt_path="", batch_size=1, frames=T, weights=synthetic.i3d_weights(0))
    with pytest.raises(AttributeError):
        k.softmax_clean
    clip = synthetic.clips_u8(1, T, seed=1001).numpy()
    clean = k(clip, adv_flag=0)
    k.improve_adversarial_loss(margin=0.05, targeted=False, logits=False)
    k.train_step(clip, [int(clean.argmax())], learning_rate=1e-3, beta_0=1.0, beta_1=0.5, beta_2=0.5, beta_3=0.5)
    assert np.allclose(k.softmax_clean, clean, atol=1e-6)
    adv = k.adversarial_inputs_rgb
    assert adv.shape == (1, T, 224, 224, 3) and adv.dtype == np.float32
    x = clip.astype(np.float32) / 128.0 - 1.0
    d = np.clip(k.eps_rgb.reshape(1, T, 1, 1, 3), -0.4, 0.4)
    assert np.allclose(adv, np.clip(x + d, -1.0, 1.0), atol=1e-6)
    k.close()
