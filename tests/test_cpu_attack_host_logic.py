"""not gpu: the host logic of attack.FlickerAttack (step, step_rolled, predict(shift), frame-range mask, sharded=False) run
on the CPU against a STAND-IN engine — a test double with FlickerEngine's interface built from the pinned torch-stack
restatement (oracle/oracle_torchstack.py) over a tiny 3D CNN.  The double exists only here; the product has no CPU path.
Each result is compared with an independent autograd computation of the same quantity."""
import pytest
import torch

from oracle import oracle_torchstack as ots
from flickering_adversarial_video_b200 import _lib as L
from flickering_adversarial_video_b200 import attack

T, B, HW, K = 6, 2, 12, 7


def _net(k=K):
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv3d(3, 5, (3, 3, 3), padding=1), torch.nn.ReLU(),
                              torch.nn.AdaptiveAvgPool3d((T, 1, 1)), torch.nn.Flatten(), torch.nn.Linear(5 * T, k))
    for p in net.parameters():
        p.requires_grad_(False)
    return net.eval()


def _normalize(clips_u8):
    x = clips_u8.permute(0, 4, 1, 2, 3).float() / 255.0
    mean = torch.tensor(ots.DEFAULT_MEAN).reshape(1, 3, 1, 1, 1)
    std = torch.tensor(ots.DEFAULT_STD).reshape(1, 3, 1, 1, 1)
    return (x - mean) / std


def _adv_loss(net, clips_u8, delta_t3, labels, delta_clip):
    """logits and margin loss for the perturbation AS THE NETWORK SEES IT (delta_t3 [T,3])"""
    lo, hi = ots.value_bounds()
    std = torch.tensor(ots.DEFAULT_STD).reshape(1, 3, 1, 1, 1)
    pc = delta_t3.clamp(-delta_clip, delta_clip).t().reshape(1, 3, T, 1, 1)
    logits = net((_normalize(clips_u8) + pc / std).clamp(lo, hi))
    return logits, ots.improve_adversarial_loss(labels, logits, torch.softmax(logits, 1), 0.05, False)


class StandInEngine:
    """FlickerEngine's interface (engine.py) for the torch stack, eager torch on the CPU"""

    def __init__(self, batch, frames, height=None, width=None, num_classes=K, device=0, arch="r3d_18"):
        self.B, self.T, self.H, self.W, self.K = batch, frames, HW, HW, num_classes
        self.device, self.torch_stack, self.arch = torch.device("cpu"), arch != "i3d", arch
        self.net = _net(num_classes)
        self.logits, self.probs = torch.zeros((batch, num_classes)), torch.zeros((batch, num_classes))
        self.scalars, self.grad = torch.zeros(L.S_COUNT), torch.zeros((frames, 3))
        self.applied = []

    def load_weights(self, weights):
        pass

    def close(self):
        pass

    def apply(self, clips, delta, adv_flag=1.0, delta_clip=0.4, adv_u8=None, adv_f32=None, stream=None):
        self._clips, self._delta, self._flag, self._clip = clips, delta.detach().clone(), adv_flag, delta_clip
        self.applied.append(self._delta.clone())

    def forward(self, stream=None):
        d = (self._flag * self._delta).requires_grad_(True)
        self._d = d
        self._logits_graph, _ = _adv_loss(self.net, self._clips, d, torch.zeros(self.B, dtype=torch.int64), self._clip)
        self.logits.copy_(self._logits_graph.detach())
        return self.logits

    def loss(self, labels, improve_loss=True, targeted=False, use_logits=False, margin=0.05, grad_scale=1.0,
             global_batch=0, stack=L.FAV_STACK_TORCH, stream=None):
        prob = torch.softmax(self._logits_graph, 1)
        self._loss = ots.improve_adversarial_loss(labels, self._logits_graph, prob, margin, use_logits)
        self.probs.copy_(prob.detach())
        self.scalars[L.S_ADV_LOSS] = float(self._loss.detach())
        return self.scalars

    def backward(self, stream=None):
        (g,) = torch.autograd.grad(self._loss, self._d)
        # like fav_backward_delta: the data term w.r.t. the CLAMPED perturbation, before the |delta| <= clip mask
        inside = (self._d.detach().abs() <= self._clip).float()
        self.grad.copy_(torch.where(inside > 0, g, torch.zeros_like(g)) + 0.0)
        return self.grad

    def update(self, delta, grad, m, v, step, beta0, beta1, beta2, beta3, lr=1e-3, delta_clip=0.4, b1=0.9, b2=0.999,
               eps=1e-8, stack=L.FAV_STACK_TORCH, stream=None):
        d = delta.detach().clone().requires_grad_(True)
        pc = d.clamp(-delta_clip, delta_clip)
        reg = beta0 * ots.flickering_regularization_loss(pc.t().reshape(3, T, 1, 1), beta1)
        (g_reg,) = torch.autograd.grad(reg, d)
        g = grad * (delta.abs() <= delta_clip).float() + g_reg
        step += 1
        t = int(step)
        m.mul_(b1).add_((1 - b1) * g)
        v.mul_(b2).add_((1 - b2) * g * g)
        delta.sub_((lr / (1 - b1 ** t)) * m / (v.sqrt() / (1 - b2 ** t) ** 0.5 + eps))
        return self.scalars


class StandInEvalEngine(StandInEngine):
    """EvalEngine's interface (engine.py; fav_create_eval / fav_eval_batch): clean + perturbed rows of one validation
    batch, counters accumulated in `counts` = [miss, valid]"""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.counts = torch.zeros(2, dtype=torch.int64)
        self.probs2 = torch.zeros((2 * self.B, self.K))
        self.calls = []

    def reset_counts(self):
        self.counts.zero_()

    def eval_batch(self, clips, delta, labels, clips_adv=None, n_clips=None, delta_clip=0.4, targeted=False,
                   target_class=0, exclude_misclassify=True, loss=None, want_probs=False, stream=None):
        self.calls.append(dict(delta=delta.clone(), clips_adv=clips_adv, n_clips=n_clips, loss=loss))
        n = self.B if n_clips is None else n_clips
        self.apply(clips, delta, adv_flag=0.0, delta_clip=delta_clip)
        clean = self.forward().clone()
        self.apply(clips if clips_adv is None else clips_adv, delta, adv_flag=1.0, delta_clip=delta_clip)
        adv = self.forward().clone()
        if loss is not None:
            self.loss(labels, improve_loss=loss["improve_loss"], targeted=loss["targeted"], use_logits=loss["use_logits"],
                      margin=loss["margin"], stack=loss["stack"])
        valid = (clean.argmax(-1) == labels) if exclude_misclassify else torch.ones_like(labels, dtype=torch.bool)
        miss = (adv.argmax(-1) == int(target_class or 0)) if targeted else (adv.argmax(-1) != labels)
        self.counts[0] += int((valid & miss)[:n].sum())
        self.counts[1] += int(valid[:n].sum())
        if want_probs:
            self.probs2.copy_(torch.softmax(torch.cat([clean, adv]), -1))
        return self.counts


@pytest.fixture()
def make_attack(monkeypatch):
    monkeypatch.setattr(attack, "FlickerEngine", StandInEngine)
    monkeypatch.setattr(attack, "EvalEngine", StandInEvalEngine)

    def make(**kw):
        return attack.FlickerAttack({}, B, T, {"LAMBDA": 1.0, "BETA_1": 0.5}, num_classes=K, arch="r3d_18", delta_clip=0.1, **kw)
    return make


def _data():
    g = torch.Generator().manual_seed(3)
    clips = torch.randint(0, 256, (B, T, HW, HW, 3), generator=g, dtype=torch.uint8)
    delta = (torch.rand((T, 3), generator=g) - 0.5) * 0.16                  # mostly inside +-0.1, so gradients flow
    with torch.no_grad():                       # the clean prediction: the margin loss is active for the predicted class
        labels = _net()(_normalize(clips)).argmax(-1)
    return clips, delta, labels


def test_step_rolled_gradient_and_shift_zero(make_attack):
    clips, delta, labels = _data()
    for shift in (1, 4):
        atk = make_attack()
        atk.delta.copy_(delta)
        atk.step_rolled(clips, labels, shift)
        assert torch.equal(atk.eng.applied[-1], torch.roll(delta, shift, 0))          # the network saw the rolled delta
        d = delta.clone().requires_grad_(True)                                        # independent: autograd through the roll
        _, loss = _adv_loss(_net(), clips, torch.roll(d, shift, 0), labels, 0.1)
        (want,) = torch.autograd.grad(loss, d)
        assert want.abs().max() > 0 and torch.allclose(atk.grad, want, rtol=1e-5, atol=1e-9)
    a0, a1 = make_attack(), make_attack()
    a0.delta.copy_(delta)
    a1.delta.copy_(delta)
    a0.step(clips, labels)
    a1.step_rolled(clips, labels, 0)
    assert torch.equal(a0.delta, a1.delta) and torch.equal(a0.grad, a1.grad) and int(a0.step_count) == 1
    assert not torch.equal(a0.delta, delta)                                           # Adam moved it
    # predict(shift) evaluates the rolled perturbation
    p = a0.predict(clips, adv_flag=1.0, shift=2)
    ref, _ = _adv_loss(_net(), clips, torch.roll(a0.delta, 2, 0), labels, 0.1)
    assert torch.allclose(p, torch.softmax(ref, 1), atol=1e-7)


def test_frame_range_mask_on_the_stand_in(make_attack):
    clips, delta, labels = _data()
    atk = make_attack(frame_range=(2, 4))
    atk.delta.copy_(delta)
    atk.step(clips, labels)
    seen = atk.eng.applied[-1]
    assert torch.equal(seen[2:5], delta[2:5]) and float(seen[:2].abs().max()) == 0 and float(seen[5:].abs().max()) == 0
    assert float(atk.grad[:2].abs().max()) == 0 and float(atk.grad[5:].abs().max()) == 0 and float(atk.grad[2:5].abs().max()) > 0
    d = delta.clone().requires_grad_(True)                                            # independent: d/d delta of loss(mask * delta)
    mask = torch.zeros((T, 1))
    mask[2:5] = 1
    _, loss = _adv_loss(_net(), clips, d * mask, labels, 0.1)
    (want,) = torch.autograd.grad(loss, d)
    assert torch.allclose(atk.grad, want, rtol=1e-5, atol=1e-9)
    assert make_attack(frame_range=(0, T)).frame_mask is None and make_attack(frame_range=(0, T - 1)).frame_mask is None
    # the frame mask and the cyclic roll compose: roll(mask * delta)
    atk2 = make_attack(frame_range=(2, 4))
    atk2.delta.copy_(delta)
    atk2.step_rolled(clips, labels, 3)
    assert torch.equal(atk2.eng.applied[-1], torch.roll(delta * mask, 3, 0))


def test_replica_attacks_never_join_a_collective(make_attack):
    atk = make_attack(sharded=False)
    assert atk.world == 1 and atk.global_batch == B
    atk.check_replicas()                                                              # no-op without ranks


def test_eval_batch_host_logic(make_attack):
    """FlickerAttack.eval_batch / eval_counts (row f3): the evaluation handle is created lazily with the attack's shape,
    sees the masked / rolled perturbation, and its device counters are read and reset by eval_counts."""
    clips, delta, labels = _data()
    atk = make_attack(frame_range=(1, 4))
    atk.delta.copy_(delta)
    assert atk._eval is None
    atk.eval_batch(clips, labels, shift=2, with_loss=True)
    ev = atk.evaluator()
    assert (ev.B, ev.T, ev.K) == (B, T, K) and ev is atk.evaluator()
    mask = torch.zeros((T, 1))
    mask[1:5] = 1
    assert torch.equal(ev.calls[-1]["delta"], torch.roll(delta * mask, 2, 0))
    assert ev.calls[-1]["loss"]["stack"] == L.FAV_STACK_TORCH and ev.calls[-1]["loss"]["margin"] == atk.margin
    # independent count for the same perturbation
    with torch.no_grad():
        adv_logits, _ = _adv_loss(_net(), clips, torch.roll(delta * mask, 2, 0), labels, 0.1)
        clean_pred = _net()(_normalize(clips)).argmax(-1)
    valid = clean_pred == labels
    want = (int((valid & (adv_logits.argmax(-1) != labels)).sum()), int(valid.sum()))
    atk.eval_batch(clips, labels, shift=2, n_clips=1)                    # ragged batch: only the first clip counts
    want2 = (want[0] + int((valid & (adv_logits.argmax(-1) != labels))[:1].sum()), want[1] + int(valid[:1].sum()))
    assert atk.eval_counts(reset=True) == want2
    assert atk.eval_counts() == (0, 0)
    atk.close()
    assert atk._eval is None
