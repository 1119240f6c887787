"""Generate golden vectors for the torch-stack loss / regulariser / Perturbation semantics by running
the REFERENCE's own classes (utils_cv/action_recognition/model.py:58-250) on the CPU of this container.

    python tests/golden/make_torch_stack_golden.py     ->  tests/golden/torch_stack_golden.npz

The reference cannot travel to the GPU box, so the vectors are committed.  Shims: stub
`decord`, `IPython`, `matplotlib` (imported at module top), and route `.to('cuda')` to the CPU because
`Losses.__init__` hard-codes it (model.py:146).
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "torch_stack_golden.npz")


class _Stub(types.ModuleType):
    def __getattr__(self, attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return type(attr, (), {})


def _stub(name):
    m = _Stub(name)
    m.__file__ = "<stub>"
    m.__path__ = []
    sys.modules[name] = m
    return m


def import_reference():
    for n in ("decord", "IPython", "IPython.display", "matplotlib", "matplotlib.pyplot"):
        _stub(n)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["IPython"].display = sys.modules["IPython.display"]
    _orig_to = torch.Tensor.to

    def _to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        if "device" in k and isinstance(k["device"], str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return _orig_to(self, *a, **k)
    torch.Tensor.to = _to
    sys.path.insert(0, REF)
    import utils_cv.action_recognition.model as ref_model   # noqa: E402
    return ref_model


def main():
    ref = import_reference()
    g = torch.Generator().manual_seed(20260)
    out = {}
    B, K, T = 6, 400, 16
    logits = torch.randn((B, K), generator=g) * 3.0
    labels = torch.randint(0, K, (B,), generator=g)
    # make clip 0 confidently correct, clip 1 barely correct, clip 2 already fooled
    logits[0, labels[0]] = 12.0
    logits[1, labels[1]] = logits[1].max() + 0.02
    logits[2, labels[2]] = -5.0
    out["logits"], out["labels"] = logits.numpy(), labels.numpy()
    pert = (torch.rand((3, T, 1, 1), generator=g) - 0.5) * 0.3
    out["perturbation"] = pert.numpy()
    for name, kw in [("improve_prob", dict(improve_loss=True, logits=False)),
                     ("improve_logits", dict(improve_loss=True, logits=True)),
                     ("ce", dict(improve_loss=False, logits=False))]:
        lg = logits.clone().requires_grad_(True)
        p = pert.clone().requires_grad_(True)
        prob = torch.softmax(lg, dim=1)
        losses = ref.Losses(beta_1=0.3, lambda_=2.0, targeted=False, target_class=None, margin=0.05,
                            attack_type="flickering", **kw)
        loss, adv, reg = losses(labels, lg, prob, p)
        loss.backward()
        out[f"{name}/loss"] = np.float32(loss.item())
        out[f"{name}/adv"] = np.float32(adv.item())
        out[f"{name}/reg"] = np.float32(reg.item())
        out[f"{name}/dlogits"] = lg.grad.numpy()
        out[f"{name}/dpert"] = p.grad.numpy()
    # Perturbation.forward: clamp -> /std -> add -> clamp (model.py:80-96)
    pm = ref.Perturbation(size=(3, T, 1, 1), device="cpu", max_norm=0.1)
    with torch.no_grad():
        pm.perturbation.copy_((torch.rand((3, T, 1, 1), generator=g) - 0.5) * 0.4)   # beyond +-0.1
    x = torch.randn((2, 3, T, 8, 8), generator=g) * 1.5
    y = pm([x, True])
    y.sum().backward()
    out["pert/param"] = pm.perturbation.detach().numpy()
    out["pert/x"] = x.numpy()
    out["pert/y"] = y.detach().numpy()
    out["pert/grad"] = pm.perturbation.grad.numpy()
    out["pert/min_value"], out["pert/max_value"] = np.float64(pm.min_value), np.float64(pm.max_value)
    th, ro = pm.metric_calc()
    out["pert/thickness"], out["pert/roughness"] = np.float32(th.item()), np.float32(ro.item())
    # torch.optim.Adam trajectory on delta with the reference's regulariser (model.py:542, 198-209)
    d = torch.nn.Parameter(((torch.rand((3, T, 1, 1), generator=g) - 0.5) * 0.25))
    opt = torch.optim.Adam([d], lr=1e-3)
    losses = ref.Losses(beta_1=0.3, lambda_=2.0, improve_loss=True, logits=False, attack_type="flickering")
    gd = torch.randn((5, 3, T, 1, 1), generator=g) * 0.01
    traj = [d.detach().clone().numpy()]
    max_norm = 0.1
    for i in range(5):
        opt.zero_grad()
        dc = d.clamp(-max_norm, max_norm)
        reg = losses.flickering_regularization_loss(dc)
        (2.0 * reg + (dc * gd[i]).sum()).backward()     # data term: gradient gd[i] w.r.t. the clamped delta
        opt.step()
        traj.append(d.detach().clone().numpy())
    # Adversarial_metrics.accuracy_for_eval (model.py:293-323) and Losses.L12_regularization_loss (:211-214)
    met = ref.Adversarial_metrics(targeted=False, target_class=None)
    adv_out = torch.randn((B, K), generator=g)
    clean_out = torch.randn((B, K), generator=g)
    gt = torch.randint(0, K, (B,), generator=g)
    clean_out[torch.arange(4), gt[:4]] = 20.0          # four clips clean-correct
    adv_out[torch.arange(2), gt[:2]] = 20.0            # two of them survive the attack
    miss, num = met.accuracy_for_eval(adv_out, gt, topk=(1,), clean_pred=clean_out)
    out["metrics/adv_out"], out["metrics/clean_out"], out["metrics/gt"] = adv_out.numpy(), clean_out.numpy(), gt.numpy()
    out["metrics/miss"], out["metrics/num"] = np.float32(miss.item()), np.float32(num.item())
    psp = (torch.rand((3, 5, 6, 7), generator=g) - 0.5) * 0.3
    l12 = ref.Losses(attack_type="L12").L12_regularization_loss(psp)
    out["l12/pert"], out["l12/value"] = psp.numpy(), np.float32(l12.item())
    out["adam/data_grads"] = gd.numpy()
    out["adam/traj"] = np.stack(traj)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
