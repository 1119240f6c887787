"""not gpu: host logic of the torch-stack mirror (flickering_adversarial_video_b200/torch_stack.py) against golden
vectors produced by the REFERENCE's own classes (tests/golden/make_torch_stack_golden.py)."""
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "torch_stack_golden.npz"))


def test_accuracy_for_eval_matches_reference():
    from flickering_adversarial_video_b200.torch_stack import Adversarial_metrics
    met = Adversarial_metrics(targeted=False, target_class=None)
    miss, num = met.accuracy_for_eval(torch.tensor(G["metrics/adv_out"]), torch.tensor(G["metrics/gt"]), topk=(1,),
                                      clean_pred=torch.tensor(G["metrics/clean_out"]))
    assert float(miss) == float(G["metrics/miss"]) and float(num) == float(G["metrics/num"])


def test_regularisers_match_reference():
    from flickering_adversarial_video_b200.torch_stack import Losses
    l12 = Losses(attack_type="L12").L12_regularization_loss(torch.tensor(G["l12/pert"]))
    assert np.isclose(float(l12), G["l12/value"], rtol=1e-6)
    reg = Losses(beta_1=0.3, lambda_=2.0).flickering_regularization_loss(torch.tensor(G["perturbation"]))
    assert np.isclose(float(reg), G["improve_prob/reg"], rtol=1e-6)


def test_perturbation_layouts_and_bounds():
    from flickering_adversarial_video_b200.torch_stack import Perturbation
    p = Perturbation(size=(3, 16, 1, 1), device="cpu", max_norm=0.1)
    assert np.isclose(p.min_value, G["pert/min_value"]) and np.isclose(p.max_value, G["pert/max_value"])
    p.init_perturbation(G["pert/param"])
    clamped, raw = p.get_perturbation()
    assert float(clamped.abs().max()) <= 0.1 + 1e-7 and torch.equal(raw, torch.tensor(G["pert/param"]))
    th, ro = p.metric_calc()
    assert np.isclose(float(th), G["pert/thickness"], rtol=1e-5) and np.isclose(float(ro), G["pert/roughness"], rtol=1e-5)
    e = p.as_engine()                      # [T,3]
    assert tuple(e.shape) == (16, 3) and torch.equal(e[:, 1], raw[1, :, 0, 0])
    p.from_engine(e)
    assert torch.equal(p.perturbation, raw)
    with pytest.raises(Exception):
        p.forward([torch.zeros(1), True])  # no engine bound: the product path never falls back to the CPU


def test_learner_rejects_unknown_model_and_missing_weights():
    from flickering_adversarial_video_b200.torch_stack import VideoLearnerAdversarial
    with pytest.raises(ValueError):
        VideoLearnerAdversarial(base_model="r2plus1d_34", weights={})
    with pytest.raises(ValueError):
        VideoLearnerAdversarial(base_model="r3d_18", weights=None)


def test_fit_many_videos_file_protocol(tmp_path):
    """model.py:925-946, 974-979: result file naming, skip rules and the None placeholder — host logic only (the
    per-video attack itself is stubbed; it is covered on the GPU by test_gpu_torch_api.py)"""
    import numpy as np
    import torch
    from flickering_adversarial_video_b200 import torch_stack as ts

    class Pert:
        size, device, max_norm, dynamic_max_norm = (3, 4, 1, 1), "cpu", 0.2, 0.5
        perturbation = None

    calls = []

    class Learner(ts.VideoLearnerAdversarial):
        def __init__(self):
            self.pert_model, self.dataset, self.label_id_to_text = Pert(), None, {0: "riding a bike", 1: "yoga"}

        def fit_single_video(self, lr, n_iter, clip_u8, label, **kw):
            assert kw["max_restarts"] == 4 and kw["restart_after"] == 3000
            calls.append((label, float(self.pert_model.perturbation.abs().max()), self.pert_model.dynamic_max_norm))
            if label == 1:
                return None                                  # clean clip misclassified
            return {"is_adversarial": [False, True], "label": label}

    vids = [(torch.zeros((4, 8, 8, 3), dtype=torch.uint8), 0, "root/cls/vid_a"),
            (torch.zeros((4, 8, 8, 3), dtype=torch.uint8), 1, "root/cls/vid_b"),
            (torch.zeros((4, 8, 8, 3), dtype=torch.uint8), 0, "root/cls/vid_c")]
    np.save(str(tmp_path / "vid_c_@riding_a_bike.npy"), {"is_adversarial": [True]}, allow_pickle=True)   # already done
    lp = {"lambda_": 1.0, "beta_1": 0.5, "targeted_attack": False, "target_class_id": None, "improve_loss": True,
          "use_logits": False}
    out = Learner().fit_many_videos(1e-3, model_dir=str(tmp_path), save_model=True, loss_params_dict=lp, videos=vids)
    assert [c[0] for c in calls] == [0, 1] and set(out) == {"vid_a", "vid_b"}
    assert all(0 < c[1] <= 0.005 and c[2] == 0.2 for c in calls)          # re-drawn U(-1,1)*0.005, norm reset
    a = np.load(str(tmp_path / "vid_a_@riding_a_bike.npy"), allow_pickle=True).tolist()
    assert a["label"] == 0 and a["is_adversarial"] == [False, True]
    assert np.load(str(tmp_path / "vid_b_@yoga.npy"), allow_pickle=True).tolist() is None     # placeholder stays
    # second run: everything is skipped (success / placeholder / success)
    calls.clear()
    assert Learner().fit_many_videos(1e-3, model_dir=str(tmp_path), save_model=True, loss_params_dict=lp, videos=vids) == {}
    assert calls == []
