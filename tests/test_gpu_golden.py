"""-m gpu: the loss kernel and the delta-update kernel (kernel c) in torch-stack mode against golden
vectors produced by the reference's own classes (tests/golden/torch_stack_golden.npz), and in TF
mode against the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "torch_stack_golden.npz"))


@pytest.mark.parametrize("name,kw", [("improve_prob", dict(improve_loss=True, use_logits=False)),
                                     ("improve_logits", dict(improve_loss=True, use_logits=True)),
                                     ("ce", dict(improve_loss=False))])
def test_loss_kernel_torch_stack_golden(name, kw):
    from flickering_adversarial_video_b200 import _lib as L
    from flickering_adversarial_video_b200.engine import op_loss
    logits = torch.tensor(G["logits"]).cuda()
    labels = torch.tensor(G["labels"]).cuda()
    probs, dlogits, sc = op_loss(logits, labels, margin=0.05, stack=L.FAV_STACK_TORCH, **kw)
    assert np.isclose(float(sc[L.S_ADV_LOSS]), G[f"{name}/adv"], rtol=2e-5)
    assert np.allclose(dlogits.cpu().numpy(), G[f"{name}/dlogits"], rtol=2e-4, atol=1e-7)
    assert np.allclose(probs.cpu().numpy(), torch.softmax(torch.tensor(G["logits"]), 1).numpy(), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("kw", [dict(improve_loss=True, use_logits=False), dict(improve_loss=True, use_logits=True),
                                dict(improve_loss=False), dict(improve_loss=True, targeted=True),
                                dict(improve_loss=False, targeted=True)])
def test_loss_kernel_tf_stack_oracle(kw):
    from flickering_adversarial_video_b200 import _lib as L
    from flickering_adversarial_video_b200.engine import op_loss
    from oracle import oracle_i3d as O
    logits = torch.tensor(G["logits"], requires_grad=True)
    labels = torch.tensor(G["labels"])
    if kw.get("improve_loss", True):
        loss, pmin, pmax = O.improve_adversarial_loss(logits, labels, 0.05, kw.get("targeted", False), kw.get("use_logits", False))
    else:
        loss, pmin, pmax = O.ce_adversarial_loss(logits, labels, kw.get("targeted", False))
    loss.backward()
    _, dlogits, sc = op_loss(logits.detach().cuda(), labels.cuda(), margin=0.05, stack=L.FAV_STACK_TF, **kw)
    assert np.isclose(float(sc[L.S_ADV_LOSS]), float(loss), rtol=2e-5, atol=1e-7)
    assert np.allclose(dlogits.cpu().numpy(), logits.grad.numpy(), rtol=3e-4, atol=1e-7)
    assert np.isclose(float(sc[L.S_SUM_P_MIN]), float(pmin.sum()), rtol=1e-5)
    assert np.isclose(float(sc[L.S_SUM_P_MAX]), float(pmax.sum()), rtol=1e-5)
    pred = logits.detach().argmax(-1)
    fooled = (pred == labels).sum() if kw.get("targeted", False) else (pred != labels).sum()
    assert int(sc[L.S_FOOLED]) == int(fooled)


def test_delta_update_torch_stack_golden():
    """clamped-delta regulariser (beta1, 1-beta1 weights) + clamp mask + torch.optim.Adam"""
    from flickering_adversarial_video_b200 import _lib as L
    from flickering_adversarial_video_b200.engine import op_delta_update
    traj, gd = G["adam/traj"], G["adam/data_grads"]          # [6,3,T,1,1], [5,3,T,1,1]
    T = traj.shape[2]
    to_ours = lambda a: torch.tensor(a.reshape(3, T).T.copy()).cuda()   # [3,T,1,1] -> [T,3]
    d = to_ours(traj[0])
    m, v = torch.zeros_like(d), torch.zeros_like(d)
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    for i in range(5):
        op_delta_update(d, to_ours(gd[i]), m, v, step, beta0=2.0, beta1=0.3, beta2=0.7, beta3=0.7, lr=1e-3,
                        delta_clip=0.1, stack=L.FAV_STACK_TORCH)
        assert np.allclose(d.cpu().numpy(), traj[i + 1].reshape(3, T).T, rtol=1e-5, atol=1e-8), f"step {i}"
