"""-m gpu: the torch-stack operator surface (flickering_adversarial_video_b200/torch_stack.py mirrors
utils_cv/action_recognition/model.py) — result layouts of the universal and single-video loops, and attack
outcome parity (same fooled flag, thickness and roughness within 5 %) against the oracle loop on the CPU."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LP = {"lambda_": 1.0, "beta_1": 0.5, "targeted_attack": False, "target_class_id": None, "target_class_name": None,
      "improve_loss": True, "use_logits": False}


def test_universal_fit_layout(tmp_path):
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200.torch_stack import VideoLearnerAdversarial
    from oracle import oracle_resnet
    B, T = 2, 8
    model = synthetic.resnet_model("mc3_18", seed=0)
    clips = [synthetic.clips_u8(B, T, 112, 112, seed=1100 + i) for i in range(2)]
    with torch.no_grad():
        labels = [model(oracle_resnet.normalize_u8(c)).argmax(-1) for c in clips]
    batches = lambda: [(c.cuda(), l.cuda()) for c, l in zip(clips, labels)]
    learner = VideoLearnerAdversarial(num_classes=400, base_model="mc3_18", sample_length=T, l_inf_pert_norm=0.1,
                                      attack_type="flickering", weights=model.state_dict(), batch_size=B)
    res = learner.fit(lr=1e-3, epochs=2, train_batches=batches, valid_batches=batches, model_dir=str(tmp_path), save_model=True,
                      model_name="mc3_18", loss_params_dict=LP)
    assert len(res) == 2
    for key in ("time", "loss", "fooling_ratio", "pert_thickness", "pert_roughness", "inf_norm", "perturbation"):
        assert f"train/{key}" in res[-1] and f"valid/{key}" in res[-1]
    assert res[-1]["valid/perturbation"].shape == (3, T, 1, 1)
    assert 0.0 <= res[-1]["valid/fooling_ratio"] <= 1.0
    assert res[-1]["valid/inf_norm"] <= 0.1 + 1e-6
    saved = np.load(os.path.join(str(tmp_path), "mc3_18_002.npy"), allow_pickle=True)
    assert saved[-1]["valid/perturbation"].shape == (3, T, 1, 1)      # resume reads [-1]['valid/perturbation']


def test_single_video_outcome_parity():
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200.torch_stack import VideoLearnerAdversarial
    from oracle import oracle_resnet, oracle_torchstack as ots
    T, n_iter, max_norm = 8, 40, 0.2
    model = synthetic.resnet_model("r3d_18", seed=0)
    clip = synthetic.clips_u8(1, T, 112, 112, seed=1200)
    with torch.no_grad():
        label = int(model(oracle_resnet.normalize_u8(clip)).argmax(-1))
    learner = VideoLearnerAdversarial(num_classes=400, base_model="r3d_18", sample_length=T, l_inf_pert_norm=max_norm,
                                      attack_type="flickering", weights=model.state_dict())
    learner.pert_model.perturbation.zero_()
    res = learner.fit_single_video(lr=1e-3, n_iter=n_iter, clip_u8=clip.cuda(), label=label, loss_params_dict=LP,
                                   restart_after=n_iter, max_restarts=1)
    assert res is not None and len(res["perturbation"]) >= n_iter
    # oracle loop (model.py:1046-1116) from the same zero start
    delta = torch.zeros((T, 3))
    opt = ots.TorchAdam((T, 3), lr=1e-3)
    labels = torch.tensor([label])
    fooled = False
    for _ in range(n_iter):
        ref = oracle_resnet.attack_step(model, clip, labels, delta, max_norm=max_norm, opt=opt)
        fooled = int(ref["logits"].argmax(-1)) != label
        delta = ref["delta_new"]
    dc = delta.clamp(-max_norm, max_norm)
    th_ref = float(dc.abs().mean())
    ro_ref = float((torch.roll(dc, 1, 0) - dc).abs().mean())
    th, ro = float(res["perturbation/thickness"][n_iter - 1]), float(res["perturbation/roughness"][n_iter - 1])
    print(f"single-video r3d_18 after {n_iter} steps: fooled engine {res['is_adversarial'][n_iter - 1]} oracle {fooled}; "
          f"thickness {th:.5f} vs {th_ref:.5f}; roughness {ro:.5f} vs {ro_ref:.5f}")
    assert res["is_adversarial"][n_iter - 1] == fooled
    assert abs(th - th_ref) <= 0.05 * th_ref and abs(ro - ro_ref) <= 0.05 * ro_ref
