"""ORACLE (test infrastructure) — the torch stack's attack step on the CPU in fp32: the torchvision video
ResNet itself (the reference's model, utils_cv/action_recognition/model.py:421) wrapped by the restatement
of Perturbation / Losses / Adam in oracle_torchstack.py (pinned against the reference's own classes by
tests/golden/torch_stack_golden.npz).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this."""
import torch

from . import oracle_torchstack as ots


def normalize_u8(clips_u8):
    """[B,T,H,W,3] uint8 -> [B,3,T,H,W] float32, (u/255 - mean)/std
    (references/functional_video.py:65-97 to_tensor + normalize; dataset.py:28-29)."""
    x = clips_u8.permute(0, 4, 1, 2, 3).to(torch.float32) / 255.0
    mean = torch.tensor(ots.DEFAULT_MEAN, dtype=torch.float32, device=x.device).reshape(1, 3, 1, 1, 1)
    std = torch.tensor(ots.DEFAULT_STD, dtype=torch.float32, device=x.device).reshape(1, 3, 1, 1, 1)
    return (x - mean) / std


def attack_step(model, clips_u8, labels, delta_t3, max_norm=0.1, beta_1=0.5, lambda_=1.0, margin=0.05,
                improve_loss=True, use_logits=False, lr=1e-3, opt=None, endpoints=None, data_grad_only=False):
    """One iteration of the reference's loop (model.py:697-735): adversarial forward, Losses, backward
    to the perturbation, Adam.  delta_t3 is [T,3] (the engine's layout of the reference's [3,T,1,1]).  Runs on the
    device of its inputs (the BASELINE-shape GPU tests run this same code in strict fp32 on cuda); data_grad_only skips
    the regulariser backward and the Adam step."""
    T = delta_t3.shape[0]
    x = normalize_u8(clips_u8)
    pert = delta_t3.t().reshape(3, T, 1, 1).clone().requires_grad_(True)
    pc = pert.clamp(-max_norm, max_norm)
    pc.retain_grad()
    lo, hi = ots.value_bounds()
    std = torch.tensor(ots.DEFAULT_STD, dtype=torch.float32, device=x.device).reshape(3, 1, 1, 1)
    adv = (x + pc / std).clamp(lo, hi)
    hooks = []
    if endpoints is not None:
        def keep(name):
            def fn(_m, _i, out):
                endpoints[name] = out.detach().permute(0, 2, 3, 4, 1).contiguous()
            return fn
        for name, mod in model.named_modules():
            if name in ("stem", "layer1.0", "layer1.1", "layer2.0", "layer2.1", "layer3.0", "layer3.1", "layer4.0",
                        "layer4.1"):
                hooks.append(mod.register_forward_hook(keep(name)))
    logits = model(adv)
    for hk in hooks:
        hk.remove()
    prob = torch.softmax(logits, dim=1)
    loss, adv_loss, reg_loss = ots.losses(labels, logits, prob, pc, beta_1, lambda_, margin, improve_loss, use_logits)
    adv_loss.backward(retain_graph=True)
    grad_data = pc.grad.detach().reshape(3, T).t().clone()          # d adv_loss / d clamped delta, [T,3]
    if data_grad_only:
        return dict(adv=adv.detach(), logits=logits.detach(), prob=prob.detach(), adv_loss=float(adv_loss),
                    grad_data=grad_data)
    pert.grad = None
    pc.grad = None
    loss.backward()
    grad_total = pert.grad.detach().reshape(3, T).t().clone()
    opt = opt or ots.TorchAdam((T, 3), lr=lr)
    delta_new = opt.step(delta_t3, grad_total)
    return dict(adv=adv.detach(), logits=logits.detach(), prob=prob.detach(), loss=float(loss), adv_loss=float(adv_loss),
                reg_loss=float(reg_loss), grad_data=grad_data, grad_total=grad_total, delta_new=delta_new)


def sparse_attack_step(model, clips_u8, labels, delta_thwc, max_norm=0.2, lambda_=1.0, margin=0.05, lr=1e-3, opt=None):
    """attack_type "L12" (model.py:383-384, 211-214): per-pixel perturbation [3,T,H,W] (here [T,H,W,3]),
    loss = adv + lambda_ * sum_t sqrt(mean_{c,h,w} clamp(delta)^2)."""
    x = normalize_u8(clips_u8)
    pert = delta_thwc.permute(3, 0, 1, 2).clone().requires_grad_(True)      # [3,T,H,W]
    pc = pert.clamp(-max_norm, max_norm)
    pc.retain_grad()
    lo, hi = ots.value_bounds()
    std = torch.tensor(ots.DEFAULT_STD, dtype=torch.float32).reshape(3, 1, 1, 1)
    adv = (x + pc / std).clamp(lo, hi)
    logits = model(adv)
    prob = torch.softmax(logits, dim=1)
    adv_loss = ots.improve_adversarial_loss(labels, logits, prob, margin, False)
    reg = torch.sum(torch.sqrt(torch.mean(pc ** 2, [0, 2, 3]))) + 1e-12
    loss = adv_loss + lambda_ * reg
    adv_loss.backward(retain_graph=True)
    grad_data = pc.grad.detach().permute(1, 2, 3, 0).clone()
    pert.grad = None
    pc.grad = None
    loss.backward()
    grad_total = pert.grad.detach().permute(1, 2, 3, 0).clone()
    opt = opt or ots.TorchAdam(tuple(delta_thwc.shape), lr=lr)
    delta_new = opt.step(delta_thwc, grad_total)
    return dict(logits=logits.detach(), adv_loss=float(adv_loss.detach()), reg_loss=float(reg.detach()),
                grad_data=grad_data, grad_total=grad_total, delta_new=delta_new)
