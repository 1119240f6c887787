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


# ---- the TF-stack oracle (oracle/oracle_i3d.py) where its formulas COINCIDE with the reference's torch classes -----------
# The two stacks of the reference implement the same paper: in probability mode the margin loss, the untargeted CE loss
# and the three regulariser terms are the same expressions (they differ only in logits mode and in how
# the terms are weighted).  The TF stack itself cannot run here, so this pins oracle_i3d's loss assembly against vectors
# the reference's OWN torch code produced.
def test_tf_oracle_losses_match_reference_where_the_stacks_coincide():
    from oracle import oracle_i3d as TF
    labels = torch.tensor(G["labels"])
    for name, fn in (("improve_prob", lambda z: TF.improve_adversarial_loss(z, labels, 0.05, False, False)[0]),
                     ("ce", lambda z: TF.ce_adversarial_loss(z, labels, False)[0])):
        z = torch.tensor(G["logits"], requires_grad=True)
        adv = fn(z)
        adv.backward()
        assert np.isclose(float(adv), G[f"{name}/adv"], rtol=1e-6), name
        assert np.allclose(z.grad.numpy(), G[f"{name}/dlogits"], rtol=1e-5, atol=1e-8), name
    # regularisers: the reference's torch loss is beta_1*thick + (1-beta_1)*(diff + lap) of the same three terms
    # (model.py:198-209, beta_1 = 0.3 in the generator) as kinetics_i3d_utils.py:177-190
    pert = torch.tensor(G["perturbation"])                                   # [3,T,1,1]
    delta = pert.reshape(3, -1).t().contiguous()                             # the TF stack's [T,3]
    thick, diff, lap, _, _ = TF.regularizers(delta)
    assert np.isclose(float(0.3 * thick + 0.7 * (diff + lap)), G["improve_prob/reg"], rtol=1e-6)


def test_tf_adam_tracks_the_reference_adam():
    """tf.train.AdamOptimizer differs from torch.optim.Adam only in where epsilon enters: on the
    reference's own trajectory the TF restatement must stay within the O(eps / sqrt(v)) distance that implies — a wrong
    bias correction or moment update would be off by orders of magnitude more."""
    from oracle import oracle_i3d as TF
    traj, gd = G["adam/traj"], torch.tensor(G["adam/data_grads"])
    d = torch.tensor(traj[0])
    opt = TF.TFAdam(tuple(d.shape), lr=1e-3)
    for i in range(5):
        dd = d.clone().requires_grad_(True)
        dc = dd.clamp(-0.1, 0.1)
        (2.0 * O.flickering_regularization_loss(dc, 0.3) + (dc * gd[i]).sum()).backward()
        d = opt.step(d, dd.grad)
        step = np.abs(traj[i + 1] - traj[i]).max()
        assert np.abs(d.numpy() - traj[i + 1]).max() < 5e-4 * step, i       # measured 1.0-1.4e-4 (eps / |g| for the smallest |g|)
