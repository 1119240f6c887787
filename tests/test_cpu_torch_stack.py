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
