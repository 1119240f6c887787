"""-m gpu: the fused evaluation pass (include/fav.h fav_create_eval / fav_eval_batch) against
(1) the CPU oracle, (2) the training handle's two-forward path it replaces (kinetics_i3d.evaluate's two sess.run per
batch, utils/kinetics_i3d_utils.py:217-250; the torch validation phase, model.py:697-713) and (3) the counting rules of
the reference restated on the host."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _host_counts(clean_logits, adv_logits, labels, n, targeted=False, target=0, exclude=True):
    """utils/kinetics_i3d_utils.py:226-243 / model.py:293-323 on host arrays"""
    pred, pred_clean = adv_logits.argmax(-1), clean_logits.argmax(-1)
    miss_cond = (pred == target) if targeted else (pred != labels)
    valid = (pred_clean == labels) if exclude else np.ones_like(miss_cond)
    return int(np.logical_and(miss_cond, valid)[:n].sum()), int(valid[:n].sum())


def _delta(T, amp, seed):
    g = torch.Generator().manual_seed(seed)
    return ((torch.rand((T, 3), generator=g) - 0.5) * 2 * amp).cuda()


def test_i3d_eval_pass_matches_training_handle_and_oracle():
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200 import _lib as L
    from flickering_adversarial_video_b200.attack import FlickerAttack
    from oracle import oracle_i3d as O
    B, T = 2, 16
    w = synthetic.i3d_weights(0)
    atk = FlickerAttack(w, B, T)
    atk.delta.copy_(_delta(T, 0.3, 5))
    clips = synthetic.clips_u8(B, T, seed=1001).cuda()
    rolled = torch.roll(clips, 5, dims=1).contiguous()
    clean = atk.predict(clips, adv_flag=0.0).clone()
    adv = atk.predict(clips, adv_flag=1.0).clone()
    adv_rolled = atk.predict(rolled, adv_flag=1.0).clone()
    cl, ad, adr = clean.cpu().numpy(), adv.cpu().numpy(), adv_rolled.cpu().numpy()
    lab_np = cl.argmax(-1).copy()
    lab_np[1] = (lab_np[1] + 1) % 400                     # clip 1 is "misclassified when clean": excluded from the ratio
    labels = torch.as_tensor(lab_np).cuda()
    ev = atk.evaluator()
    # forward-only arena: 2B clips of activations, no gradients / pool codes / data-gradient weights
    print(f"device bytes: training handle (B={B}) {atk.eng.device_bytes / 2**20:.0f} MiB, "
          f"evaluation handle (2x{B} clips) {ev.device_bytes / 2**20:.0f} MiB")
    assert ev.device_bytes < 0.95 * atk.eng.device_bytes
    # (2) same probabilities as the two-forward path (same kernels, other batch size): tight tolerance
    atk.eval_batch(clips, labels, want_probs=True)
    p = ev.probs.cpu().numpy()
    assert np.abs(p[:B] - cl).max() < 2e-5 and np.abs(p[B:] - ad).max() < 2e-5
    assert np.array_equal(p[:B].argmax(-1), cl.argmax(-1)) and np.array_equal(p[B:].argmax(-1), ad.argmax(-1))
    assert atk.eval_counts() == _host_counts(cl, ad, lab_np, B)
    # (3) the counting rules: no exclusion, targeted, ragged batch, cyclic (only the perturbed input is rolled)
    atk.eval_batch(clips, labels, exclude_misclassify=False)
    assert atk.eval_counts() == _host_counts(cl, ad, lab_np, B, exclude=False)
    tgt = int(ad[0].argmax())
    atk.eval_batch(clips, labels, targeted=True, target_class=tgt)
    assert atk.eval_counts() == _host_counts(cl, ad, lab_np, B, targeted=True, target=tgt)
    atk.eval_batch(clips, labels, n_clips=1)
    atk.eval_batch(clips, labels, clips_adv=rolled)      # counters accumulate over batches
    a, b = _host_counts(cl, ad, lab_np, 1), _host_counts(cl, adr, lab_np, B)
    assert atk.eval_counts() == (a[0] + b[0], a[1] + b[1])
    assert atk.eval_counts() == (0, 0)
    # (1) the CPU oracle on the same clips and perturbation
    model = O.OracleI3D(w)
    x = O.normalize_u8(clips.cpu())
    with torch.no_grad():
        ref_clean = torch.softmax(model.forward(x), -1).numpy()
        ref_adv = torch.softmax(model.forward(O.apply_flicker(x, atk.delta.cpu())), -1).numpy()
    err = max(np.abs(p[:B] - ref_clean).max(), np.abs(p[B:] - ref_adv).max())
    print(f"fused evaluation pass vs oracle: max |dprob| {err:.3e}")
    assert err < 2e-3 and np.array_equal(p[:B].argmax(-1), ref_clean.argmax(-1))
    assert np.array_equal(p[B:].argmax(-1), ref_adv.argmax(-1))
    assert _host_counts(ref_clean, ref_adv, lab_np, B) == _host_counts(cl, ad, lab_np, B)
    # the training entry points refuse an evaluation handle
    st = ev.lib.fav_apply_flicker(ev.h, L.ptr(clips), L.FAV_U8, L.ptr(atk.delta), 1.0, 0.4, None, None,
                                  L.stream_ptr(None, ev.device))
    assert st == -3 and "evaluation handle" in L.last_error()          # FAV_ERR_STATE
    atk.close()


def test_kinetics_i3d_evaluate_uses_the_fused_pass():
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200.kinetics_i3d import kinetics_i3d
    T = 16
    k = kinetics_i3d(ckpt_path="", batch_size=1, frames=T, weights=synthetic.i3d_weights(0))
    k._atk.delta.copy_(_delta(T, 0.35, 7))
    clips = [synthetic.clips_u8(1, T, seed=1001 + i).numpy() for i in range(3)]
    clean = [k(c, adv_flag=0) for c in clips]
    adv = [k(c, adv_flag=1) for c in clips]
    labels = [int(c.argmax()) for c in clean]
    labels[2] = (labels[2] + 3) % 400
    want_miss = sum(int(a.argmax() != l) for a, l, c in zip(adv, labels, clean) if int(c.argmax()) == l)
    want_total = sum(int(c.argmax()) == l for c, l in zip(clean, labels))
    rate, total = k.evaluate(iter([(c, [l]) for c, l in zip(clips, labels)]))
    assert total == want_total == 2 and abs(rate - want_miss / want_total) < 1e-12
    rate, total = k.evaluate(iter([(c, [l]) for c, l in zip(clips, labels)]), exclude_misclassify=False)
    assert total == 3 and abs(rate - sum(int(a.argmax() != l) for a, l in zip(adv, labels)) / 3) < 1e-12
    assert k._atk._eval is not None
    k.close()


@pytest.mark.parametrize("arch", ["r3d_18", "r2plus1d_18"])
def test_torch_stack_eval_pass(arch):
    """validation phase of the torch stack (model.py:697-713): clean forward without the range clamp, perturbed forward,
    adversarial loss of the perturbed rows"""
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200 import _lib as L
    from flickering_adversarial_video_b200.attack import FlickerAttack
    B, T = 2, 8
    sd = synthetic.resnet_model(arch, seed=0).state_dict()
    atk = FlickerAttack(sd, B, T, {"LAMBDA": 1.0, "BETA_1": 0.5}, arch=arch, delta_clip=0.1)
    atk.delta.copy_(_delta(T, 0.12, 3))                   # partly outside +-0.1: the clamp of the perturbation acts
    clips = synthetic.clips_u8(B, T, 112, 112, seed=1001).cuda()
    clips[0, :, :8] = 0
    clips[1, :, :8] = 255                                 # pixels outside the scalar bounds: the clean forward keeps them
    clean = atk.predict(clips, adv_flag=0.0).cpu().numpy()
    adv = atk.predict(clips, adv_flag=1.0, shift=3).cpu().numpy()
    lab_np = clean.argmax(-1).copy()
    labels = torch.as_tensor(lab_np).cuda()
    sc_ref = atk.eng.loss(labels, improve_loss=True, targeted=False, use_logits=False, margin=0.05,
                          stack=L.FAV_STACK_TORCH).clone()
    atk.eval_batch(clips, labels, shift=3, with_loss=True, want_probs=True)
    ev = atk.evaluator()
    p = ev.probs.cpu().numpy()
    assert np.abs(p[:B] - clean).max() < 2e-5 and np.abs(p[B:] - adv).max() < 2e-5
    assert atk.eval_counts() == _host_counts(clean, adv, lab_np, B)
    assert abs(float(ev.scalars[L.S_ADV_LOSS]) - float(sc_ref[L.S_ADV_LOSS])) < 1e-4 * max(1.0, abs(float(sc_ref[L.S_ADV_LOSS])))
    assert ev.device_bytes < 0.95 * atk.eng.device_bytes
    atk.close()
