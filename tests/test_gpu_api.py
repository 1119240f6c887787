"""-m gpu: the host-side mirror of the reference's operator surface (kinetics_i3d object, drivers,
result layouts) and attack-outcome parity against the oracle."""
import os
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
T = 16


@pytest.fixture(scope="module")
def k_i3d():
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200.kinetics_i3d import kinetics_i3d
    k = kinetics_i3d(ckpt_path="", batch_size=1, frames=T, weights=synthetic.i3d_weights(0))
    yield k
    k.close()


def test_reference_attribute_surface(k_i3d):
    from flickering_adversarial_video_b200 import synthetic
    clip = synthetic.clips_u8(1, T, seed=1001).numpy()
    prob = k_i3d(clip, adv_flag=0)
    assert prob.shape == (1, 400) and abs(prob.sum() - 1) < 1e-4
    k_i3d.improve_adversarial_loss(margin=0.05, targeted=False, logits=False)
    k_i3d.reset()
    out = k_i3d.train_step(clip, [int(prob.argmax())], learning_rate=1e-3, beta_0=1.0, beta_1=0.5, beta_2=0.5, beta_3=0.5)
    # names the reference drivers fetch (i3d_adversarial_main_single_video_npy.py:61-77, universal.py:90-109)
    for name in ("softmax", "model_logits", "norm_reg", "diff_norm_reg", "laplacian_norm_reg", "thickness", "roughness",
                 "thickness_relative", "roughness_relative", "to_min_prob", "to_max_prob"):
        assert getattr(k_i3d, name) is not None and name in out
    assert k_i3d.eps_rgb.shape == (T, 1, 1, 3) and k_i3d.perturbation.shape == (T, 1, 1, 3)
    assert np.abs(k_i3d.eps_rgb).max() <= 1.001e-3          # first Adam step moves every element by <= lr
    assert len(k_i3d.get_kinetics_classes()) == 400
    adv8 = k_i3d.adversarial_video_uint8(clip)
    advf = k_i3d.adversarial_inputs_rgb_of(clip)
    assert adv8.dtype == np.uint8 and np.array_equal(adv8, ((advf + np.float32(1.0)) * np.float32(127.5)).astype(np.uint8))


def test_single_video_driver_and_pkl_layout(k_i3d, tmp_path):
    from flickering_adversarial_video_b200 import config, synthetic
    from flickering_adversarial_video_b200.drivers import single_video_attack
    cfg = config.default_config().SINGLE_VIDEO_ATTACK
    cfg.MAX_NUM_STEP = 3
    clip_u8 = synthetic.clips_u8(1, T, seed=1001)
    rgb_sample = (clip_u8.float() / 128.0 - 1.0).numpy()    # the .npy clips are stored normalised (single_video_npy.py:121)
    label = int(k_i3d(rgb_sample, adv_flag=0).argmax())
    res = single_video_attack(k_i3d, rgb_sample, label, cfg, result_path=str(tmp_path), max_extra_steps=4)
    assert res is not None and os.path.exists(res["pkl_path"])
    with open(res["pkl_path"], "rb") as f:
        d = pickle.load(f)
    # keys written by the reference (single_video_npy.py:177-181,314-334) and read by its viewer
    for key in ("correct_cls_prob", "correct_cls", "correct_cls_id", "softmax_init", "rgb_sample", "total_loss_l",
                "adv_loss_l", "reg_loss_l", "norm_reg_loss_l", "diff_norm_reg_loss_l", "perturbation", "adv_video",
                "softmax", "total_steps", "beta_0", "beta_1", "beta_2", "beta_3", "fatness", "smoothness"):
        assert key in d, key
    assert d["adv_video"].shape == (1, T, 224, 224, 3) and d["perturbation"][-1].shape == (T, 1, 1, 3)
    assert len(d["total_loss_l"]) == d["total_steps"] + 1
    # wrong label -> the reference skips the clip
    assert single_video_attack(k_i3d, rgb_sample, (label + 1) % 400, cfg) is None


def test_attack_outcome_parity(k_i3d):
    """north_star gate: same fooled / not-fooled outcome, thickness and roughness within 5 % of the
    oracle after the same number of Adam steps on the same clip."""
    from flickering_adversarial_video_b200 import synthetic
    from oracle import oracle_i3d as O
    steps = 60
    clip_u8 = synthetic.clips_u8(1, T, seed=1001)
    x = O.normalize_u8(clip_u8)
    model = O.OracleI3D(synthetic.i3d_weights(0))
    with torch.no_grad():
        label = model.forward(x).argmax(-1)
    cfg = dict(improve_loss=True, margin=0.05, beta0=1.0, beta1=0.5, beta2=0.5, beta3=0.5, lr=1e-3)
    opt = O.TFAdam((T, 3))
    delta = torch.zeros((T, 3))
    for _ in range(steps):
        out = O.attack_step(model, x, label, delta, cfg, opt=opt)
        delta = out["delta_new"]
    _, _, _, th_ref, ro_ref = O.regularizers(delta)
    with torch.no_grad():
        fooled_ref = bool(model.forward(O.apply_flicker(x, delta)).argmax(-1) != label)
    k_i3d.improve_adversarial_loss(margin=0.05)
    k_i3d.reset()
    for _ in range(steps):
        k_i3d.train_step(clip_u8.numpy(), label.numpy(), learning_rate=1e-3, beta_0=1.0, beta_1=0.5, beta_2=0.5, beta_3=0.5)
    d = torch.tensor(k_i3d.eps_rgb.reshape(T, 3))
    _, _, _, th, ro = O.regularizers(d)
    fooled = bool(k_i3d(clip_u8.numpy(), adv_flag=1).argmax(-1)[0] != int(label))
    print(f"after {steps} steps: thickness engine {float(th):.5f} oracle {float(th_ref):.5f}; roughness engine "
          f"{float(ro):.5f} oracle {float(ro_ref):.5f}; fooled engine {fooled} oracle {fooled_ref}")
    assert fooled == fooled_ref
    assert abs(float(th) / float(th_ref) - 1) < 0.05
    assert abs(float(ro) / float(ro_ref) - 1) < 0.05


def test_class_gen_driver(k_i3d):
    from flickering_adversarial_video_b200 import config, synthetic
    from flickering_adversarial_video_b200.drivers import class_gen_attack
    cfg = config.default_config().CLASS_GEN_ATTACK
    cfg.MAX_NUM_STEP = 2
    clips = [synthetic.clips_u8(1, T, seed=1001 + i).numpy() for i in range(2)]
    labels = [[int(k_i3d(c, adv_flag=0).argmax())] for c in clips]
    batches = lambda: iter(list(zip(clips, labels)))
    res = class_gen_attack(k_i3d, batches, batches, cfg)
    for key in ("total_loss_l", "adv_loss_l", "reg_loss_l", "norm_reg_loss_l", "diff_norm_reg_loss_l", "perturbation",
                "total_steps", "beta_1", "beta_2", "fatness", "smoothness", "fool_rate"):
        assert key in res, key
    assert res["total_steps"] == 2 and len(res["fool_rate"]) == 2


def test_single_video_90_frames_parity():
    """BASELINE.json configs[0]: one 1x90x224x224x3 clip (odd temporal sizes 45 -> 23 -> 12 through the pools)."""
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200.engine import FlickerEngine
    from oracle import oracle_i3d
    B, T = 1, 90
    weights = synthetic.i3d_weights(seed=0)
    clip = synthetic.clips_u8(B, T, seed=1090)
    delta = synthetic.delta_uniform(T, seed=17, lo=-0.05, hi=0.05)
    model = oracle_i3d.OracleI3D(weights)
    x = oracle_i3d.normalize_u8(clip)
    with torch.no_grad():
        labels = model.forward(x).argmax(-1)
    cfg = dict(improve_loss=True, margin=0.05, beta0=1.0, beta1=0.5, beta2=0.5, beta3=0.5, lr=1e-3)
    ref = oracle_i3d.attack_step(model, x, labels, delta, cfg, data_grad_only=True)
    eng = FlickerEngine(B, T)
    eng.load_weights(weights)
    eng.apply(clip.cuda(), delta.cuda())
    logits = eng.forward().cpu()
    eng.loss(labels.cuda(), improve_loss=True, margin=0.05)
    g = eng.backward().cpu()
    torch.cuda.synchronize()
    rel = float((logits - ref["logits"]).abs().max() / ref["logits"].abs().max())
    gr = ref["grad_data"]
    cos = float((g * gr).sum() / (g.norm() * gr.norm() + 1e-30))
    print(f"T=90 single video: logits rel err {rel:.3e}, top1 {logits.argmax(-1).tolist()} vs {ref['logits'].argmax(-1).tolist()}, "
          f"dL/d-delta cosine {cos:.5f}")
    assert rel <= 1e-2 and logits.argmax(-1).tolist() == ref["logits"].argmax(-1).tolist()
    # one clip, 270 gradient entries of ~1e-3: the reference's own network with TF32 convolutions on this GPU has
    # 0.987 here (profiles/r02_open_precision_1x90.txt; bf16 storage in round 1: 0.881); kernels are gated stage by
    # stage elsewhere
    assert cos >= 0.97
    eng.close()


def test_sparse_universal_driver_with_kinetics_i3d_L12(tmp_path):
    """FLICKERING_ATTACK=False: the reference's kinetics_i3d_L12 object (per-pixel eps_rgb [T,224,224,3], initialised to
    1e-8, loss = adv + beta_0 * beta_1 * L12) behind the universal driver (i3d_adversarial_main_universal.py:126-135)."""
    from flickering_adversarial_video_b200 import config, synthetic
    from flickering_adversarial_video_b200.drivers import universal_attack
    from flickering_adversarial_video_b200.kinetics_i3d import kinetics_i3d_L12
    k = kinetics_i3d_L12(ckpt_path="", batch_size=1, frames=T, weights=synthetic.i3d_weights(0))
    try:
        assert k.flickering is False
        assert k.eps_rgb.shape == (T, 224, 224, 3) and np.allclose(k.eps_rgb, 1e-8)
        cfg = config.default_config().UNIVERSAL_ATTACK
        clips = [synthetic.clips_u8(1, T, seed=2001 + i).numpy() for i in range(2)]
        labels = [[int(k(c, adv_flag=0).argmax())] for c in clips]
        batches = lambda: iter(list(zip(clips, labels)))
        res = universal_attack(k, batches, batches, cfg, max_steps=3, summary_dir=str(tmp_path))
        assert res["total_steps"] == 3 and res["perturbation"].shape == (T, 224, 224, 3)
        from flickering_adversarial_video_b200.records import read_scalars
        ev = [f for f in os.listdir(tmp_path / "train") if f.startswith("events.out.tfevents.")]
        scal = read_scalars(str(tmp_path / "train" / ev[0]))
        assert ("Loss/total" in {t for _, t, _ in scal}) and ("ACC: 1- FOOLING_RATIO" in {t for _, t, _ in scal})
        assert np.abs(res["perturbation"]).max() > 1e-6          # Adam moved the pixels
        step0 = dict((t, v[0][1]) for t, v in res["scalars"].items())
        assert abs(step0["Loss/total"] - (step0["Loss/adversarial_loss"] + step0["Loss/regularizer_loss"])) < 1e-5
        assert 0.0 <= res["fool_rate"][-1][1] <= 1.0
    finally:
        k.close()


def test_staged_graph_replay_matches_eager_steps():
    """End-to-end path (pinned host clips in, host scalars out): replaying the per-slot CUDA graphs of
    `capture_staged()` must give the same perturbation as the eager launches."""
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200.attack import FlickerAttack
    w = synthetic.i3d_weights(0)
    host = [synthetic.clips_u8(1, T, seed=3001 + i).pin_memory() for i in range(2)]
    labels = [torch.tensor([5 + i], dtype=torch.int64).pin_memory() for i in range(2)]
    deltas, losses = [], []
    for graphs in (False, True):
        atk = FlickerAttack(w, 1, T, {})
        for i in range(2):
            atk.step_staged(atk.prefetch(host[i], labels[i]))
        if graphs:
            atk.capture_staged()
        ls = []
        for i in range(4):
            out, ev = atk.step_staged(atk.prefetch(host[i % 2], labels[i % 2]))
            ev.synchronize()
            ls.append(float(out[9]))
        torch.cuda.synchronize()
        deltas.append(atk.perturbation.detach().cpu().clone())
        losses.append(ls)
        atk.close()
    assert torch.allclose(deltas[0], deltas[1], rtol=0, atol=1e-6), float((deltas[0] - deltas[1]).abs().max())
    assert np.allclose(losses[0], losses[1], rtol=1e-4, atol=1e-6)
