"""not gpu: the torch-stack oracle against golden vectors produced by the REFERENCE's own classes
(tests/golden/torch_stack_golden.npz, generator tests/golden/make_torch_stack_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_torchstack as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "torch_stack_golden.npz"))


@pytest.mark.parametrize("name,kw", [("improve_prob", dict(improve_loss=True, use_logits=False)),
                                     ("improve_logits", dict(improve_loss=True, use_logits=True)),
                                     ("ce", dict(improve_loss=False))])
def test_losses_match_reference(name, kw):
    logits = torch.tensor(G["logits"], requires_grad=True)
    labels = torch.tensor(G["labels"])
    pert = torch.tensor(G["perturbation"], requires_grad=True)
    loss, adv, reg = O.losses(labels, logits, torch.softmax(logits, 1), pert, beta_1=0.3, lambda_=2.0, margin=0.05, **kw)
    loss.backward()
    assert np.isclose(float(loss), G[f"{name}/loss"], rtol=1e-6)
    assert np.isclose(float(adv), G[f"{name}/adv"], rtol=1e-6)
    assert np.isclose(float(reg), G[f"{name}/reg"], rtol=1e-6)
    assert np.allclose(logits.grad.numpy(), G[f"{name}/dlogits"], rtol=1e-5, atol=1e-8)
    assert np.allclose(pert.grad.numpy(), G[f"{name}/dpert"], rtol=1e-5, atol=1e-9)


def test_perturbation_forward_matches_reference():
    lo, hi = O.value_bounds()
    assert np.isclose(lo, G["pert/min_value"]) and np.isclose(hi, G["pert/max_value"])
    p = torch.tensor(G["pert/param"], requires_grad=True)
    y = O.perturbation_forward(torch.tensor(G["pert/x"]), p, max_norm=0.1)
    y.sum().backward()
    assert np.allclose(y.detach().numpy(), G["pert/y"], rtol=1e-6, atol=1e-7)
    assert np.allclose(p.grad.numpy(), G["pert/grad"], rtol=1e-5, atol=1e-6)
    assert np.isclose(float(p.detach().abs().mean() * 100), G["pert/thickness"], rtol=1e-5)
    assert np.isclose(float((torch.roll(p.detach(), 1, 1) - p.detach()).abs().mean() * 100), G["pert/roughness"], rtol=1e-5)


def test_adam_trajectory_matches_reference():
    traj, gd = G["adam/traj"], torch.tensor(G["adam/data_grads"])
    d = torch.tensor(traj[0])
    opt = O.TorchAdam(d.shape)
    for i in range(5):
        dd = d.clone().requires_grad_(True)
        dc = dd.clamp(-0.1, 0.1)
        (2.0 * O.flickering_regularization_loss(dc, 0.3) + (dc * gd[i]).sum()).backward()
        d = opt.step(d, dd.grad)
        assert np.allclose(d.numpy(), traj[i + 1], rtol=1e-5, atol=1e-8), f"step {i}"
