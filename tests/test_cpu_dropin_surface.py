"""not gpu: the host mirror keeps the reference's operator surface — every class / method the attack drivers call exists
with the same positional parameter names in the same order (extra parameters only after them or keyword-only).  The
reference's lists are committed in tests/golden/reference_signatures.json (generator: make_reference_signatures.py —
introspection of utils_cv/action_recognition/{model,dataset}.py, syntax tree of utils/kinetics_i3d_utils.py)."""
import inspect
import json
import os

import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_signatures.json")))


def _resolve(key):
    from flickering_adversarial_video_b200 import kinetics_i3d, torch_stack, video_dataset
    mod, *path = key.split(".")
    obj = {"model": torch_stack, "dataset": video_dataset, "kinetics_i3d_utils": kinetics_i3d}[mod]
    for p in path:
        obj = getattr(obj, p)
    return obj


@pytest.mark.parametrize("key", sorted(GOLD))
def test_signature_is_a_superset_of_the_reference(key):
    ref = [name for name, _ in GOLD[key]]
    sig = inspect.signature(_resolve(key))
    if any(p.kind == p.VAR_POSITIONAL for p in sig.parameters.values()):       # Perturbation.forward(self, *input)
        assert [n for n in ref if n != "self"] == [n for n in sig.parameters if n != "self"]
        return
    mine = [k for k, v in sig.parameters.items() if v.kind != v.KEYWORD_ONLY]
    assert mine[:len(ref)] == ref, (key, ref, mine)


def test_defaults_the_drivers_rely_on():
    """defaults that change behaviour when a reference script omits the argument"""
    want = {("model.Losses.__init__", "margin"): "0.05", ("model.Perturbation.__init__", "max_norm"): "1.0",
            ("model.VideoLearnerAdversarial.fit", "start_epoch"): "1", ("model.VideoLearnerAdversarial.fit", "lr_gamma"): "0.1",
            ("model.VideoLearnerAdversarial.fit", "save_model"): "False",
            ("kinetics_i3d_utils.kinetics_i3d.__init__", "batch_size"): None,
            ("kinetics_i3d_utils.kinetics_i3d.evaluate", "exclude_misclassify"): "True",
            ("dataset.VideoDataset.__init__", "sample_length"): "8", ("dataset.VideoDataset.__init__", "batch_size"): "8"}
    for (key, name), expect in want.items():
        ref = dict((n, d) for n, d in GOLD[key])[name]
        if expect is not None:
            assert ref == expect, (key, name, ref)
        p = inspect.signature(_resolve(key)).parameters[name]
        if ref is None:
            assert p.default is inspect._empty, (key, name)
        else:
            assert repr(p.default) == ref, (key, name, p.default, ref)
