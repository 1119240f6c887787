"""ORACLE (test infrastructure) — CPU restatement of the reference's PyTorch attack stack
(utils_cv/action_recognition/model.py:58-250) in plain torch.  Pinned against outputs of the
reference's own classes run in the build container: tests/golden/torch_stack_golden.npz
(generator: tests/golden/make_torch_stack_golden.py)."""
import numpy as np
import torch

DEFAULT_MEAN = (0.43216, 0.394666, 0.37645)     # utils_cv/action_recognition/dataset.py:28
DEFAULT_STD = (0.22803, 0.22145, 0.216989)      # :29


def value_bounds():
    """Perturbation.__init__ (model.py:72-75): scalar clamp bounds of the normalised input."""
    mean, std = np.array(DEFAULT_MEAN), np.array(DEFAULT_STD)
    return float(np.max((0.0 - mean) / std)), float(np.min((1.0 - mean) / std))


def perturbation_forward(x, pert, max_norm, adversarial=True):
    """Perturbation.forward (model.py:80-96): clamp(+-max_norm) -> /std -> add -> clamp(min,max).
    x [N,3,T,H,W], pert [3,T,1,1]."""
    if not adversarial:
        return x
    lo, hi = value_bounds()
    pc = pert.clamp(-max_norm, max_norm)
    std = torch.tensor(DEFAULT_STD, dtype=x.dtype).reshape(3, 1, 1, 1)
    return (x + pc / std).clamp(lo, hi)


def flickering_regularization_loss(pert, beta_1):
    """Losses.flickering_regularization_loss (model.py:198-209); rolls along dim 1 (time)."""
    norm_reg = (pert ** 2).mean() + 1e-12
    right, left = torch.roll(pert, 1, 1), torch.roll(pert, -1, 1)
    diff = ((pert - right) ** 2).mean() + 1e-12
    lap = ((-2 * pert + right + left) ** 2).mean() + 1e-12
    return beta_1 * norm_reg + (1 - beta_1) * (diff + lap)


def improve_adversarial_loss(labels, logits, prob, margin=0.05, use_logits=False):
    """Losses.improve_adversarial_loss, untargeted (model.py:216-250): true exclusion of the label for
    the max over other classes; the logits margin uses label_prob."""
    B, K = prob.shape
    label_prob = prob.gather(1, labels.view(-1, 1))
    non_label = torch.ones_like(prob, dtype=torch.bool)
    non_label[torch.arange(B), labels] = False
    max_non_label_prob = prob[non_label].reshape(B, -1).max(1)[0].reshape(B, 1)
    if use_logits:
        to_min = logits.gather(1, labels.view(-1, 1))
        to_max = logits[non_label].reshape(B, -1).max(1)[0].reshape(B, 1)
        m = torch.log(1.0 + margin * (1.0 / (0.00001 + label_prob)))
    else:
        to_min, to_max, m = label_prob, max_non_label_prob, margin
    l_2 = ((to_min - (to_max - m)) ** 2) / m
    l_3 = to_min - (to_max - m)
    return torch.max(torch.zeros_like(l_3), torch.min(l_2, l_3)).sum()


def ce_adversarial_loss(labels, prob):
    """Losses.ce_adversarial_loss, untargeted (model.py:177-196)."""
    label_prob = prob.gather(1, labels.view(-1, 1))
    return (-torch.log(1 - label_prob + 1e-6)).mean()


def losses(labels, logits, prob, pert, beta_1=0.5, lambda_=1.0, margin=0.05, improve_loss=True, use_logits=False):
    """Losses.__call__ (model.py:169-175) -> [loss, adv_loss, reg_loss]."""
    reg = flickering_regularization_loss(pert, beta_1)
    adv = improve_adversarial_loss(labels, logits, prob, margin, use_logits) if improve_loss else ce_adversarial_loss(labels, prob)
    return adv + lambda_ * reg, adv, reg


class TorchAdam:
    """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8): step = lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)."""

    def __init__(self, shape, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
        self.m, self.v, self.t = torch.zeros(shape), torch.zeros(shape), 0
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps

    def step(self, var, grad):
        self.t += 1
        self.m = self.b1 * self.m + (1 - self.b1) * grad
        self.v = self.b2 * self.v + (1 - self.b2) * grad * grad
        denom = self.v.sqrt() / (1 - self.b2 ** self.t) ** 0.5 + self.eps
        return var - (self.lr / (1 - self.b1 ** self.t)) * self.m / denom
