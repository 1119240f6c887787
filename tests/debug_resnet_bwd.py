"""Layer-wise backward comparison of the ResNet engine against torch autograd (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flickering_adversarial_video_b200 import synthetic, _lib as L
from flickering_adversarial_video_b200.engine import FlickerEngine
from oracle import oracle_resnet, oracle_torchstack as ots

arch = sys.argv[1] if len(sys.argv) > 1 else "r3d_18"
B, T, max_norm = 2, 8, 0.1
model = synthetic.resnet_model(arch, seed=0)
clip = synthetic.clips_u8(B, T, 112, 112, seed=1003)
delta = synthetic.delta_uniform(T, seed=9, lo=-0.12, hi=0.12)
x = oracle_resnet.normalize_u8(clip)
with torch.no_grad():
    labels = model(x).argmax(-1)
pert = delta.t().reshape(3, T, 1, 1).clone().requires_grad_(True)
pc = pert.clamp(-max_norm, max_norm)
lo, hi = ots.value_bounds()
std = torch.tensor(ots.DEFAULT_STD).reshape(3, 1, 1, 1)
adv = (x + pc / std).clamp(lo, hi)
adv.retain_grad()
acts = {}
def keep(name):
    def fn(_m, _i, out):
        out.retain_grad()
        acts[name] = out
    return fn
names = ["stem"] + [f"layer{l}.{i}" for l in range(1, 5) for i in range(2)]
extra = [n for n, _ in model.named_modules() if n.endswith(("conv1", "conv2", "conv1.0.0", "conv2.0.0", "conv1.0.2", "conv2.0.2", "stem.2"))]
for name, mod in model.named_modules():
    if name in names or name in extra:
        mod.register_forward_hook(keep(name))
logits = model(adv)
prob = torch.softmax(logits, 1)
adv_loss = ots.improve_adversarial_loss(labels, logits, prob, 0.05, False)
adv_loss.backward()

eng = FlickerEngine(B, T, arch=arch)
eng.load_weights(model.state_dict())
d = delta.cuda()
eng.apply(clip.cuda(), d, delta_clip=max_norm)
eng.forward()
eng.loss(labels.cuda(), improve_loss=True, margin=0.05, stack=L.FAV_STACK_TORCH)
g = eng.backward().cpu()
torch.cuda.synchronize()
def cmp(engname, ref, what):
    try:
        got = eng.read(engname, tuple(ref.shape)).cpu()
    except Exception as ex:
        print(f"{what:28s} -> {engname}: {ex}")
        return
    rel = float((got - ref).norm() / (ref.norm() + 1e-30))
    cos = float((got * ref).sum() / (got.norm() * ref.norm() + 1e-30))
    print(f"{what:28s} rel_l2={rel:.4e} cos={cos:.6f} |ref|={float(ref.norm()):.3e} |got|={float(got.norm()):.3e}")
for name in reversed(names):
    a = acts[name]
    gref = (a.grad * (a > 0)).permute(0, 2, 3, 4, 1).contiguous()
    en = name if not (name == "stem" and arch != "r2plus1d_18") else "stem.conv"
    cmp("grad:" + en, gref, "grad " + name)
for name in extra:
    a = acts[name]
    if a.grad is None:
        continue
    gref = (a.grad * (a > 0)).permute(0, 2, 3, 4, 1).contiguous()
    cmp("grad:" + name, gref, "grad " + name)
gd = adv.grad  # [B,3,T,H,W]
print("dX ref norm", float(gd.norm()))
gr = pc.grad if pc.grad is not None else None
print("engine g", g[:3].tolist())
