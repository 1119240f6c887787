"""Record the positional parameter lists of the reference's torch-stack operator surface
(utils_cv/action_recognition/model.py, dataset.py) so that the drop-in check runs without /root/reference.

    python tests/golden/make_reference_signatures.py   ->  tests/golden/reference_signatures.json
"""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_torch_stack_golden import import_reference   # noqa: E402


def params(fn):
    return [[k, None if v.default is inspect._empty else repr(v.default)] for k, v in inspect.signature(fn).parameters.items()]


def main():
    ref = import_reference()
    import utils_cv.action_recognition.dataset as ds
    out = {}
    for cls, methods in (("Perturbation", ["__init__", "forward", "clamp_perturbation", "apply_perturbation", "metric_calc",
                                           "init_perturbation", "get_perturbation"]),
                         ("Losses", ["__init__", "__call__", "flickering_regularization_loss", "L12_regularization_loss"]),
                         ("Adversarial_metrics", ["__init__", "accuracy_for_eval"]),
                         ("VideoLearnerAdversarial", ["__init__", "fit", "fit_many_videos"])):
        for m in methods:
            out[f"model.{cls}.{m}"] = params(getattr(getattr(ref, cls), m))
    for m in ("__init__", "split_by_folder", "split_with_file", "__len__", "__getitem__"):
        out[f"dataset.VideoDataset.{m}"] = params(getattr(ds.VideoDataset, m))
    out["dataset.VideoRecord.__init__"] = params(ds.VideoRecord.__init__)
    out["dataset.get_transforms"] = params(ds.get_transforms)
    # the TF stack cannot be imported (tensorflow 1.15 / sonnet are absent): its signatures are read from the syntax tree
    import ast
    tree = ast.parse(open("/root/reference/utils/kinetics_i3d_utils.py").read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in ("kinetics_i3d", "kinetics_i3d_L12"):
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name in ("__init__", "__call__", "get_kinetics_classes", "evaluate",
                                                                   "improve_adversarial_loss", "ce_adversarial_loss"):
                    a = fn.args
                    defaults = [None] * (len(a.args) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
                    out[f"kinetics_i3d_utils.{node.name}.{fn.name}"] = [[x.arg, d] for x, d in zip(a.args, defaults)]
    path = os.path.join(HERE, "reference_signatures.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
