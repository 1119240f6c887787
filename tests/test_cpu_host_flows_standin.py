"""not gpu: the host flows above the engine — kinetics_i3d (train_step, cyclic roll, on-demand handles, evaluate), the TF
drivers (single video / class-gen / universal with TensorBoard scalars) and the torch learner's single-video loops — run
end to end on the CPU with the stand-in engine of test_cpu_attack_host_logic.py.  Checks control flow, layouts and
bookkeeping (every statement of these paths executes); numerics of the real engine are the GPU suite's business."""
import os
import pickle

import numpy as np
import pytest
import torch

from flickering_adversarial_video_b200 import _lib as L
from flickering_adversarial_video_b200 import attack
from test_cpu_attack_host_logic import HW, K, StandInEngine, StandInEvalEngine, T


class TFStandIn(StandInEngine):
    """same double; fills the optional outputs of apply the TF-stack helpers ask for"""

    def apply(self, clips, delta, adv_flag=1.0, delta_clip=0.4, adv_u8=None, adv_f32=None, stream=None):
        super().apply(clips, delta, adv_flag, delta_clip)
        x = clips.float() / 128.0 - 1.0 if clips.dtype == torch.uint8 else clips
        adv = (x + adv_flag * delta.clamp(-delta_clip, delta_clip).reshape(1, -1, 1, 1, 3)).clamp(-1, 1)
        if adv_f32 is not None:
            adv_f32.copy_(adv)
        if adv_u8 is not None:
            adv_u8.copy_(((adv + 1.0) * 127.5).to(torch.uint8))


class TFEvalStandIn(StandInEvalEngine, TFStandIn):
    pass


@pytest.fixture()
def standin(monkeypatch):
    from flickering_adversarial_video_b200 import kinetics_i3d as ki
    monkeypatch.setattr(attack, "FlickerEngine", TFStandIn)
    monkeypatch.setattr(attack, "EvalEngine", TFEvalStandIn)
    monkeypatch.setattr(ki, "_IMAGE_SIZE", HW)
    return ki


def _clip(seed):
    return torch.randint(0, 256, (1, T, HW, HW, 3), generator=torch.Generator().manual_seed(seed), dtype=torch.uint8).numpy()


def test_kinetics_i3d_mirror_flow(standin):
    k = standin.kinetics_i3d(ckpt_path="", batch_size=1, frames=T, weights={}, cyclic_pert_flag_default_c=0.0)
    with pytest.raises(AttributeError):
        k.softmax_clean
    clip = _clip(1)
    clean = k(clip, adv_flag=0)
    label = int(clean.argmax())
    k.improve_adversarial_loss(margin=0.05, targeted=False, logits=False)
    out = k.train_step(clip, [label], learning_rate=1e-3, beta_0=1.0, beta_1=0.5, beta_2=0.5, beta_3=0.5)
    for key in ("loss", "adversarial_loss", "regularizer_loss", "norm_reg", "diff_norm_reg", "thickness", "roughness",
                "thickness_relative", "to_min_prob", "to_max_prob", "model_logits", "softmax"):
        assert key in out
    assert k.thickness == out["thickness"] and k.softmax.shape == (1, 400)          # attribute handles of the last run
    assert np.allclose(k.softmax_clean, clean, atol=1e-6)
    adv = k.adversarial_inputs_rgb
    assert adv.shape == (1, T, HW, HW, 3) and np.abs(adv).max() <= 1.0
    # cyclic perturbation: the engine sees the rolled delta
    before = k._atk.delta.clone()
    k._rng = np.random.RandomState(5)
    k.train_step(clip, [label], cyclic_pert_flag=1.0)
    shift = int(np.random.RandomState(5).randint(0, T))
    assert torch.equal(k._atk.eng.applied[-1], torch.roll(before, shift, 0)) or shift == 0
    miss, total = k.evaluate(iter([(clip, [label]), (_clip(2), [label])]), exclude_misclassify=True)
    assert 0.0 <= miss <= 1.0 and isinstance(total, int) and total <= 2
    miss, total = k.evaluate(iter([(clip, [label])]), exclude_misclassify=False)
    assert total == 1
    k.reset()
    assert float(k._atk.delta.abs().max()) == 0.0
    k.close()


def test_tf_drivers_flow(standin, tmp_path):
    from flickering_adversarial_video_b200 import config, drivers
    from flickering_adversarial_video_b200.records import read_scalars
    k = standin.kinetics_i3d(ckpt_path="", batch_size=1, frames=T, weights={})
    clip = _clip(3)
    label = int(k(clip, adv_flag=0).argmax())
    cfg = config.default_config()
    sv = cfg.SINGLE_VIDEO_ATTACK
    sv.MAX_NUM_STEP = 2
    res = drivers.single_video_attack(k, clip, label, sv, result_path=str(tmp_path / "sv"), max_extra_steps=3)
    assert res is not None and res["total_steps"] >= 3 and os.path.exists(res["pkl_path"])
    saved = pickle.load(open(res["pkl_path"], "rb"))
    for key in ("correct_cls_prob", "correct_cls", "correct_cls_id", "softmax_init", "rgb_sample", "total_loss_l", "adv_loss_l",
                "reg_loss_l", "norm_reg_loss_l", "diff_norm_reg_loss_l", "perturbation", "adv_video", "softmax", "total_steps",
                "beta_0", "beta_1", "beta_2", "beta_3", "fatness", "smoothness"):
        assert key in saved, key
    assert saved["perturbation"][-1].shape == (T, 1, 1, 3) and saved["adv_video"].shape == (1, T, HW, HW, 3)
    assert drivers.single_video_attack(k, clip, (label + 1) % 400, sv) is None      # clean-misclassified clips are skipped

    batches = lambda: iter([(clip, [label]), (_clip(4), [label])])
    cg = cfg.CLASS_GEN_ATTACK
    cg.MAX_NUM_STEP = 3
    res = drivers.class_gen_attack(k, batches, batches, cg, result_path=str(tmp_path / "cg"), epochs=5)
    assert res["total_steps"] == 3 and os.path.exists(str(tmp_path / "cg" / "res.pkl")) and len(res["fool_rate"]) == 3

    ua = cfg.UNIVERSAL_ATTACK
    res = drivers.universal_attack(k, batches, batches, ua, max_steps=4, eval_every=2, summary_dir=str(tmp_path / "ua"))
    assert res["total_steps"] == 4 and res["perturbation"].shape == (T, 1, 1, 3) and len(res["fool_rate"]) == 3
    ev = [f for f in os.listdir(str(tmp_path / "ua" / "train")) if "tfevents" in f]
    tags = {t for _, t, _ in read_scalars(os.path.join(str(tmp_path / "ua" / "train"), ev[0]))}
    assert {"Loss/total", "Perturbation/thickness_%", "Probability/prob_to_min", "ACC: 1- FOOLING_RATIO"} <= tags
    k.close()


def test_torch_learner_single_video_flows(monkeypatch, tmp_path):
    from flickering_adversarial_video_b200 import torch_stack as ts
    monkeypatch.setattr(attack, "FlickerEngine", StandInEngine)
    monkeypatch.setattr(attack, "EvalEngine", StandInEvalEngine)

    class Learner(ts.VideoLearnerAdversarial):
        def __init__(self):
            self.results, self.dataset, self.batch_size, self.sample_length = [], None, 1, T
            self.model_name, self.attack_type, self.num_classes = "r3d_18", "flickering", K
            self._weights, self._device, self._atk, self._rng = {}, 0, None, np.random.RandomState(0)
            self.label_id_to_text = {i: f"class {i}" for i in range(K)}
            self.pert_model = ts.Perturbation((3, T, 1, 1), device="cpu", max_norm=0.1)

    lp = {"lambda_": 1.0, "beta_1": 0.5, "targeted_attack": False, "target_class_id": None, "improve_loss": True,
          "use_logits": False}
    lrn = Learner()
    clip = torch.from_numpy(_clip(7))[0]                                       # [T,H,W,3]
    probe = attack.FlickerAttack({}, 1, T, {}, num_classes=K, arch="r3d_18")
    label = int(probe.predict(clip[None], adv_flag=0.0).argmax())
    res = lrn.fit_single_video(1e-2, 3, clip, label, video_name="v", class_name="c", model_dir=str(tmp_path),
                               loss_params_dict=lp, max_restarts=1, restart_after=6)
    assert res is not None and 3 <= len(res["is_adversarial"]) <= 8 and res["perturbation"][-1].shape == (3, T, 1, 1)
    assert res["prob_clean_input"].shape == (1, K) and res["label"].tolist() == [label]
    assert isinstance(float(res["perturbation/inf_norm"]), float) and os.path.exists(str(tmp_path / "v_@c.npy"))
    assert lrn.fit_single_video(1e-2, 3, clip, (label + 1) % K, loss_params_dict=lp) is None
    # engine + Adam state reuse across videos (the reference's single optimizer), and the reset switch
    first = lrn._atk
    steps_before = int(first.step_count)
    lrn.pert_model.perturbation = torch.zeros((3, T, 1, 1))
    lrn.fit_single_video(1e-2, 2, clip, label, loss_params_dict=lp, max_restarts=1, restart_after=4, reuse_attack=first)
    assert lrn._atk is first and int(first.step_count) > steps_before
    lrn.fit_single_video(1e-2, 2, clip, label, loss_params_dict=lp, max_restarts=1, restart_after=4, reuse_attack=first,
                         reset_optimizer=True)
    assert int(first.step_count) <= 6
    # cyclic perturbation inside the single-video loop
    lrn.pert_model.cyclic_pert = True
    lrn.fit_single_video(1e-2, 2, clip, label, loss_params_dict=lp, max_restarts=1, restart_after=4)
    # the whole fit_many_videos protocol with the real per-video attack
    lrn2 = Learner()
    vids = [(clip, label, "root/c/vid_a"), (torch.from_numpy(_clip(8))[0], label, "root/c/vid_b")]
    out = lrn2.fit_many_videos(1e-2, model_dir=str(tmp_path / "many"), save_model=True, loss_params_dict=lp, n_iter=2,
                               videos=vids, max_restarts=1, restart_after=4)
    assert set(out) == {"vid_a", "vid_b"}
    assert sorted(os.listdir(str(tmp_path / "many"))) == [f"vid_a_@class_{label}.npy", f"vid_b_@class_{label}.npy"]


class SparseStandIn(StandInEngine):
    """the per-pixel entry points of FlickerEngine (pixels_enable / apply_pixels / backward_pixels / update_pixels)"""

    def pixels_enable(self):
        self.enabled = True

    def apply_pixels(self, clip_u8, delta_px, adv_flag=1.0, delta_clip=0.0, adv_f32=None, stream=None):
        self._clips, self._px, self._flag, self._clip = clip_u8, delta_px.detach().clone(), adv_flag, delta_clip

    def forward(self, stream=None):
        from test_cpu_attack_host_logic import _normalize
        d = self._px.clone().requires_grad_(True)
        self._d = d
        pc = d.clamp(-self._clip, self._clip) if self._clip > 0 else d
        x = _normalize(self._clips) if self.torch_stack else (self._clips.float() / 128.0 - 1.0).permute(0, 4, 1, 2, 3)
        self._logits_graph = self.net(x + self._flag * pc.permute(3, 0, 1, 2)[None])
        self.logits.copy_(self._logits_graph.detach())
        return self.logits

    def backward_pixels(self, grad_px, stream=None):
        (g,) = torch.autograd.grad(self._loss, self._d)
        grad_px.copy_(g)
        return grad_px

    def update_pixels(self, delta_px, grad_px, m, v, step, reg_weight, delta_clip=0.0, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8,
                      stack=L.FAV_STACK_TF, stream=None):
        step += 1
        t = int(step)
        m.mul_(b1).add_((1 - b1) * grad_px)
        v.mul_(b2).add_((1 - b2) * grad_px * grad_px)
        delta_px.sub_((lr / (1 - b1 ** t)) * m / (v.sqrt() / (1 - b2 ** t) ** 0.5 + eps))
        self.scalars[L.S_NORM_REG] = float(delta_px.pow(2).mean((1, 2, 3)).sqrt().sum())
        self.scalars[L.S_TOTAL_LOSS] = self.scalars[L.S_ADV_LOSS] + reg_weight * self.scalars[L.S_NORM_REG]
        return self.scalars


def test_sparse_flows(monkeypatch, tmp_path):
    """kinetics_i3d_L12 behind the universal driver (FLICKERING_ATTACK = False) and the torch learner with
    attack_type "L12": SparseAttack's host logic (initial values, step, predict, replica check, result layouts)"""
    from flickering_adversarial_video_b200 import config, drivers, kinetics_i3d as ki, torch_stack as ts
    monkeypatch.setattr(attack, "FlickerEngine", SparseStandIn)
    monkeypatch.setattr(ki, "_IMAGE_SIZE", HW)
    k = ki.kinetics_i3d_L12(ckpt_path="", batch_size=1, frames=T, weights={})
    assert k.flickering is False and k.eps_rgb.shape == (T, HW, HW, 3) and np.allclose(k.eps_rgb, 1e-8)
    clip = _clip(11)
    label = int(k(clip, adv_flag=0).argmax())
    batches = lambda: iter([(clip, [label]), (_clip(12), [label])])
    res = drivers.universal_attack(k, batches, batches, config.default_config().UNIVERSAL_ATTACK, max_steps=3)
    assert res["total_steps"] == 3 and res["perturbation"].shape == (T, HW, HW, 3)
    assert k.loss_L12 >= 0 and not np.allclose(k.eps_rgb, 1e-8)
    k.close()

    class Learner(ts.VideoLearnerAdversarial):
        def __init__(self):
            self.results, self.dataset, self.batch_size, self.sample_length = [], None, 1, T
            self.model_name, self.attack_type, self.num_classes = "r3d_18", "L12", K
            self._weights, self._device, self._atk, self._rng = {}, 0, None, np.random.RandomState(0)
            self.pert_model = ts.Perturbation((3, T, HW, HW), device="cpu", max_norm=0.2)

    lp = {"lambda_": 1.0, "beta_1": 0.5, "targeted_attack": False, "target_class_id": None, "improve_loss": True,
          "use_logits": False}
    lrn = Learner()
    vid = torch.from_numpy(clip)[0]
    probe = attack.SparseAttack({}, 1, T, {}, num_classes=K, arch="r3d_18")
    assert probe.world == 1 and probe.delta.shape == (T, HW, HW, 3) and float(probe.delta.abs().max()) <= 1e-6
    probe.check_replicas()
    lab = int(probe.predict(vid[None], adv_flag=0.0).argmax())
    res = lrn.fit_single_video(1e-2, 2, vid, lab, loss_params_dict=lp, max_restarts=1, restart_after=4)
    assert res is not None and res["perturbation"][-1].shape == (3, T, HW, HW)
    lrn.pert_model.cyclic_pert = True
    with pytest.raises(NotImplementedError):
        lrn.fit_single_video(1e-2, 2, vid, lab, loss_params_dict=lp)


def test_torch_fit_with_the_real_attack_factory(monkeypatch, tmp_path):
    """VideoLearnerAdversarial.fit through the real `_attack` (value-bound check, sharded flag, delta hand-over)"""
    from flickering_adversarial_video_b200 import torch_stack as ts
    monkeypatch.setattr(attack, "FlickerEngine", StandInEngine)
    monkeypatch.setattr(attack, "EvalEngine", StandInEvalEngine)

    class Learner(ts.VideoLearnerAdversarial):
        def __init__(self):
            self.results, self.dataset, self.batch_size, self.sample_length = [], None, 2, T
            self.model_name, self.attack_type, self.num_classes = "r3d_18", "flickering", K
            self._weights, self._device, self._atk, self._rng = {}, 0, None, np.random.RandomState(0)
            self.pert_model = ts.Perturbation((3, T, 1, 1), device="cpu", max_norm=0.1)

    lp = {"lambda_": 1.0, "beta_1": 0.5, "targeted_attack": False, "target_class_id": None, "improve_loss": True,
          "use_logits": False}
    g = torch.Generator().manual_seed(1)
    clips = torch.randint(0, 256, (2, T, HW, HW, 3), generator=g, dtype=torch.uint8)
    lrn = Learner()
    probe = attack.FlickerAttack({}, 2, T, {}, num_classes=K, arch="r3d_18")
    labels = probe.predict(clips, adv_flag=0.0).argmax(-1)
    batches = lambda: [(clips, labels)] * 2
    res = lrn.fit(1e-2, 2, str(tmp_path), save_model=True, loss_params_dict=lp, train_batches=batches, valid_batches=batches)
    assert len(res) == 2 and sorted(f for f in os.listdir(str(tmp_path)) if f.endswith(".npy")) == ["r3d_18_001.npy", "r3d_18_002.npy"]
    final = lrn._atk.state_dict()
    # restart after epoch 1 (r2plus1d_main_universal_attack.py:197-216: INIT_PERT_FROM_LAST_CKPT + CONTINUE_TRAIN) from a fresh
    # learner: the epoch numbering continues and, with the state sidecar, so do the Adam moments -> bit-identical end state
    d2 = tmp_path / "resume"
    first = Learner()
    first.fit(1e-2, 1, str(d2), save_model=True, loss_params_dict=lp, train_batches=batches, valid_batches=batches)
    cont = Learner()
    start = cont.resume_from(str(d2))
    assert start == 2
    res2 = cont.fit(1e-2, 2, str(d2), save_model=True, loss_params_dict=lp, train_batches=batches, valid_batches=batches,
                    start_epoch=start, lr_step_size=100)
    assert len(res2) == 1 and os.path.exists(str(d2 / "r3d_18_002.npy"))
    got = cont._atk.state_dict()
    ref_run = Learner()
    ref_run.fit(1e-2, 2, str(tmp_path / "ref"), loss_params_dict=lp, train_batches=batches, valid_batches=batches, lr_step_size=100)
    ref = ref_run._atk.state_dict()
    for key in ("delta", "m", "v"):
        assert torch.equal(got[key], ref[key]), key
    assert got["step"] == ref["step"] == final["step"]
    assert 0.0 <= res[-1]["valid/fooling_ratio"] <= 1.0 and res[-1]["train/pert_thickness"] > 0
    lrn.pert_model.cyclic_pert = True
    lrn.fit(1e-2, 1, str(tmp_path / "cyc"), loss_params_dict=lp, train_batches=batches, valid_batches=batches)
    bad = Learner()
    bad.pert_model.max_value = 1.0
    with pytest.raises(NotImplementedError):
        bad.fit(1e-2, 1, str(tmp_path), loss_params_dict=lp, train_batches=batches, valid_batches=batches)


def test_universal_and_class_gen_resume_bit_identically(standin, tmp_path):
    """Checkpoint / resume of the TF drivers (i3d_adversarial_main_universal.py:310-348: model_dir with
    save_checkpoints_steps / keep_checkpoint_max / latest_checkpoint; single_class_gen.py:192-197,214,373: `model_step_XXXXX`
    files): a run interrupted at a checkpoint and restarted in a NEW process state continues bit-identically."""
    from flickering_adversarial_video_b200 import checkpoint as fckpt, config, drivers
    cfg = config.default_config()
    clips = [_clip(10 + i) for i in range(4)]

    def make():
        k = standin.kinetics_i3d(ckpt_path="", batch_size=1, frames=T, weights={})
        label = 7
        return k, (lambda: iter([(c, [label]) for c in clips]))

    ua = cfg.UNIVERSAL_ATTACK
    k, batches = make()
    full = drivers.universal_attack(k, batches, batches, ua, max_steps=8)
    ref = k._atk.state_dict()
    k.close()
    md = str(tmp_path / "ua_model")
    k, batches = make()
    drivers.universal_attack(k, batches, batches, ua, max_steps=4, model_dir=md, save_checkpoints_steps=2, keep_checkpoint_max=2)
    k.close()
    saved = sorted(os.listdir(md))
    assert saved == ["model_step_00002.npz", "model_step_00004.npz"], saved          # keep_checkpoint_max pruned nothing yet
    k, batches = make()                                                               # fresh object: zero delta, zero Adam
    res = drivers.universal_attack(k, batches, batches, ua, max_steps=8, model_dir=md, save_checkpoints_steps=2,
                                   keep_checkpoint_max=2)
    got = k._atk.state_dict()
    assert res["total_steps"] == 8 and got["step"] == ref["step"] == 8
    for key in ("delta", "m", "v"):
        assert torch.equal(got[key], ref[key]), key
    assert np.array_equal(res["perturbation"], full["perturbation"])
    assert sorted(os.listdir(md)) == ["model_step_00006.npz", "model_step_00008.npz"]   # newest two kept
    k.close()

    # class-generalisation driver: checkpoint at the start and after every pass; restart continues from the file's step
    cg = cfg.CLASS_GEN_ATTACK
    cg.MAX_NUM_STEP = 8
    k, batches = make()
    drivers.class_gen_attack(k, batches, batches, cg, epochs=2)
    ref = k._atk.state_dict()
    k.close()
    prefix = str(tmp_path / "cg") + os.sep
    k, batches = make()
    drivers.class_gen_attack(k, batches, batches, cg, epochs=1, ckpt_prefix=prefix)
    k.close()
    assert fckpt.latest_checkpoint(prefix)[0] == 4
    k, batches = make()
    res = drivers.class_gen_attack(k, batches, batches, cg, epochs=5, ckpt_prefix=prefix)
    got = k._atk.state_dict()
    assert res["total_steps"] == 8
    for key in ("delta", "m", "v"):
        assert torch.equal(got[key], ref[key]), key
    k.close()
