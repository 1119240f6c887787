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
            self._atk = None

        def fit_single_video(self, lr, n_iter, clip_u8, label, **kw):
            assert kw["max_restarts"] == 4 and kw["restart_after"] == 3000
            # one engine / one Adam state for all videos (model.py:868): the attack of the previous video is handed back
            assert kw["reuse_attack"] is (None if not calls else self._atk) and kw["reset_optimizer"] is False
            self._atk = getattr(self, "_atk", None) or object()
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


def test_fit_loop_epochs_schedule_and_files(tmp_path):
    """model.py:587-623 on a stub engine: epochs start_epoch..epochs inclusive, StepLR learning rates handed to the
    step, `{model_name}_{e:03d}.npy` per epoch, result keys — host logic only (the engine is stubbed)."""
    import numpy as np
    import torch
    from flickering_adversarial_video_b200 import _lib as L
    from flickering_adversarial_video_b200 import torch_stack as ts
    T, B, K = 4, 2, 10
    lrs, evals = [], []

    class Eng:
        logits = torch.zeros((B, K))
        scalars = torch.zeros(L.S_COUNT)

        def loss(self, lab, **kw):
            sc = torch.zeros(L.S_COUNT)
            sc[L.S_ADV_LOSS] = 0.25
            return sc

    class Atk:
        device, pg, world, delta_clip = "cpu", None, 1, 0.1

        def __init__(self):
            self.delta, self.eng = torch.zeros((T, 3)), Eng()

        def predict(self, clips, adv_flag=1.0):
            z = torch.zeros((B, K))
            z[:, 3] = 5.0 if adv_flag == 0.0 else -5.0          # clean: class 3; perturbed: anything else
            self.eng.logits = z
            return torch.softmax(z, -1)

        def check_replicas(self):
            pass

        # fused validation pass (row f3): counters on the "device", read once per phase
        def eval_batch(self, clips, labels, shift=0, with_loss=False, **kw):
            self._counts = getattr(self, "_counts", [0, 0])
            self._counts[0] += B
            self._counts[1] += B
            evals.append(shift)

        def eval_counts(self, reset=True):
            c = tuple(getattr(self, "_counts", [0, 0]))
            if reset:
                self._counts = [0, 0]
            return c

        def evaluator(self):
            return self.eng

        def state_dict(self):
            z = torch.zeros((T, 3))
            return {"delta": self.delta.clone(), "m": z, "v": z, "step": len(lrs)}

        def step(self, clips, lab, lr=None):
            lrs.append(lr)
            self.delta += 0.01
            self.predict(clips, 1.0)
            sc = torch.zeros(L.S_COUNT)
            sc[L.S_TOTAL_LOSS], sc[L.S_ADV_LOSS] = 1.5, 1.0
            return sc

    class Learner(ts.VideoLearnerAdversarial):
        def __init__(self):
            self.results, self.dataset, self.batch_size, self.sample_length = [], None, B, T
            self.model_name, self.attack_type = "r3d_18", "flickering"
            self.pert_model = ts.Perturbation((3, T, 1, 1), device="cpu", max_norm=0.1)

        def _attack(self, lr, lp, batch, sharded=True):
            return Atk()

    batches = lambda: [(torch.zeros((B, T, 8, 8, 3), dtype=torch.uint8), torch.tensor([3, 3]))] * 3
    lp = {"lambda_": 1.0, "beta_1": 0.5, "targeted_attack": False, "target_class_id": None, "improve_loss": True,
          "use_logits": False}
    lrn = Learner()
    res = lrn.fit(1e-2, 6, str(tmp_path), None, save_model=True, loss_params_dict=lp, start_epoch=2,
                  train_batches=batches, valid_batches=batches)
    assert len(res) == 5                                                      # epochs 2..6 inclusive
    # per epoch: the reference's result file and the sidecar for an exact restart (delta, Adam moments, Adam step)
    assert sorted(os.listdir(str(tmp_path))) == sorted([f"r3d_18_{e:03d}.npy" for e in range(2, 7)] +
                                                       [f"r3d_18_{e:03d}.state.npz" for e in range(2, 7)])
    assert int(np.load(str(tmp_path / "r3d_18_006.state.npz"))["adam_step"]) == 15
    # StepLR(step_size=ceil(2/3*6)=4, gamma=0.1), restarted at lr on this call: 4 epochs at 1e-2, then 1e-3
    assert np.allclose(lrs, [1e-2] * 12 + [1e-3] * 3)
    last = np.load(str(tmp_path / "r3d_18_006.npy"), allow_pickle=True)[-1]
    assert last["valid/perturbation"].shape == (3, T, 1, 1) and last["train/fooling_ratio"] == 1.0
    assert abs(last["train/loss"] - 1.5) < 1e-6 and abs(last["valid/pert_thickness"] - 0.1) < 1e-6     # clamped at 0.1
    assert last["valid/fooling_ratio"] == 1.0 and len(evals) == 15            # every validation batch: one fused pass
    # cyclic_pert (model.py:91-92): every adversarial forward sees the perturbation rolled by a fresh random shift
    rolled = []
    Atk.step_rolled = lambda self, clips, lab, shift, lr=None: (rolled.append(("train", shift)), Atk.step(self, clips, lab, lr))[1]
    plain_predict = Atk.predict
    Atk.predict = lambda self, clips, adv_flag=1.0, shift=0: (rolled.append(("valid", shift)) if shift else None,
                                                                plain_predict(self, clips, adv_flag))[1]
    plain_eval = Atk.eval_batch
    Atk.eval_batch = lambda self, clips, labels, shift=0, **kw: (rolled.append(("valid", shift)) if shift else None,
                                                                  plain_eval(self, clips, labels, shift=shift, **kw))[1]
    cyc = Learner()
    cyc.pert_model.cyclic_pert, cyc._rng = True, np.random.RandomState(0)
    lrs.clear()
    cyc.fit(1e-2, 2, str(tmp_path / "cyc"), loss_params_dict=lp, train_batches=batches, valid_batches=batches)
    want = np.random.RandomState(0).randint(0, T, size=1000)                  # one draw per batch, train and valid alike
    drawn = [int(w) for w in want[:12]]
    assert [s for _, s in rolled] == [d for d in drawn if d] and len(lrs) == 6
    assert {p for p, _ in rolled} == {"train", "valid"}
    with pytest.raises(NotImplementedError):
        lrn.fit(1e-2, 1, use_one_cycle_policy=True, loss_params_dict=lp, train_batches=batches, valid_batches=batches)
    with pytest.raises(ValueError):
        lrn.fit(1e-2, 1, loss_params_dict=lp)                                 # neither batch sources nor a dataset
