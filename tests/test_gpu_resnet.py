"""-m gpu: torch-stack parity — the engine (C-ABI) on the torchvision video ResNets against the reference's own
model class (torchvision) + the pinned restatement of Perturbation / Losses / Adam, same seeded inputs.
Gates (north_star): logits within 1e-2 relative with identical top-1; dL/d-delta cosine (logged; bf16 storage
limits it the same way as on I3D, see test_gpu_i3d.py); delta after one Adam step."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

T_CLIP = 8


def _report(line):
    print(line)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "resnet_parity.log"), "a") as f:
            f.write(line + "\n")


@pytest.mark.parametrize("arch", ["r3d_18", "mc3_18", "r2plus1d_18"])
def test_resnet_step_parity(arch):
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200 import _lib as L
    from flickering_adversarial_video_b200.engine import FlickerEngine
    from oracle import oracle_resnet
    B, T, max_norm = 2, T_CLIP, 0.1
    model = synthetic.resnet_model(arch, seed=0)
    clip = synthetic.clips_u8(B, T, 112, 112, seed=1003)
    delta = synthetic.delta_uniform(T, seed=9, lo=-0.12, hi=0.12)     # exceeds max_norm: the delta clamp fires
    with torch.no_grad():
        labels = model(oracle_resnet.normalize_u8(clip)).argmax(-1)
    eps = {}
    ref = oracle_resnet.attack_step(model, clip, labels, delta, max_norm=max_norm, endpoints=eps)

    eng = FlickerEngine(B, T, arch=arch)
    eng.load_weights(model.state_dict())
    d = delta.cuda()
    adv = torch.zeros((B, 3, T, 112, 112), dtype=torch.float32, device="cuda")
    eng.apply(clip.cuda(), d, delta_clip=max_norm, adv_f32=adv)
    logits = eng.forward().cpu()
    sc = eng.loss(labels.cuda(), improve_loss=True, margin=0.05, stack=L.FAV_STACK_TORCH)
    g = eng.backward().cpu()
    torch.cuda.synchronize()
    # adversarial input: same fp32 op sequence as torch
    err_adv = float((adv.cpu() - ref["adv"]).abs().max())
    _report(f"[{arch}] adversarial input max abs err {err_adv:.3e}")
    assert err_adv <= 1e-5
    names = {"stem": "stem" if arch == "r2plus1d_18" else "stem.conv"}
    for name, r in eps.items():
        got = eng.read(names.get(name, name), tuple(r.shape)).cpu()
        rel = float((got - r).norm() / (r.norm() + 1e-12))
        _report(f"[{arch}] layer {name:10s} rel_l2={rel:.4e} ref_rms={float(r.pow(2).mean().sqrt()):.4g}")
        assert rel < 2e-2, f"{arch} {name}: relative L2 error {rel}"
    rel = float((logits - ref["logits"]).abs().max() / ref["logits"].abs().max())
    _report(f"[{arch}] logits max rel err {rel:.4e}; top1 engine {logits.argmax(-1).tolist()} oracle "
            f"{ref['logits'].argmax(-1).tolist()}; logits std {float(ref['logits'].std()):.3g}")
    assert rel <= 1e-2
    assert logits.argmax(-1).tolist() == ref["logits"].argmax(-1).tolist()
    gr = ref["grad_data"]
    cos = float((g * gr).sum() / (g.norm() * gr.norm() + 1e-30))
    _report(f"[{arch}] adv_loss engine {float(sc.cpu()[0]):.6f} oracle {ref['adv_loss']:.6f}; |g| engine {float(g.norm()):.4e} "
            f"oracle {float(gr.norm()):.4e}; dL/d-delta cosine {cos:.6f}")
    # the margin loss is quadratic in a difference of probabilities: a 3e-3 logit error moves it by a few percent
    assert abs(float(sc.cpu()[0]) - ref["adv_loss"]) <= max(2e-3, 5e-2 * abs(ref["adv_loss"]))
    assert cos >= 0.995
    # Adam step with the torch rules (regulariser on the clamped delta, clamp mask, eps inside the bias correction)
    m = torch.zeros_like(d)
    v = torch.zeros_like(d)
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    eng.update(d, eng.grad, m, v, step, 1.0, 0.5, 0.5, 0.5, lr=1e-3, delta_clip=max_norm, stack=L.FAV_STACK_TORCH)
    torch.cuda.synchronize()
    dd = float((d.cpu() - ref["delta_new"]).abs().max())
    _report(f"[{arch}] max |delta - oracle| after one Adam step {dd:.3e}")
    assert dd <= 2.5e-3
    eng.close()


@pytest.mark.parametrize("arch", ["r3d_18", "r2plus1d_18"])
def test_stem_gradient_collapse_matches_dense_data_gradient(arch, monkeypatch):
    """Torch stack: dL/d-delta through the stem from the tensor-core collapse (stem_grad.cu, pass bitmap written by the
    apply kernel) against the dense stem data gradient + masked reduce (FAV_STEM_GRAD_DENSE=1), on a clip whose dark and
    bright pixels hit the scalar clamp bounds of Perturbation.forward (model.py:72-75)."""
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200 import _lib as L
    from flickering_adversarial_video_b200.engine import FlickerEngine
    B, T = 2, T_CLIP
    model = synthetic.resnet_model(arch, seed=0)
    clip = synthetic.clips_u8(B, T, 112, 112, seed=77)
    clip[:, :, :40] = (clip[:, :, :40] // 16)            # dark band: clamps at the lower bound
    clip[:, :, 80:] = 255 - (clip[:, :, 80:] // 16)      # bright band: clamps at the upper bound
    delta = synthetic.delta_uniform(T, seed=5, lo=-0.1, hi=0.1)
    labels = None
    grads = []
    for dense in (False, True):
        if dense:
            monkeypatch.setenv("FAV_STEM_GRAD_DENSE", "1")
        else:
            monkeypatch.delenv("FAV_STEM_GRAD_DENSE", raising=False)
        eng = FlickerEngine(B, T, arch=arch)
        eng.load_weights(model.state_dict())
        eng.apply(clip.cuda(), delta.cuda(), delta_clip=0.1)
        logits = eng.forward()
        if labels is None:
            labels = logits.argmax(-1).clone()
        eng.loss(labels, improve_loss=True, margin=0.05, stack=L.FAV_STACK_TORCH)
        grads.append(eng.backward().clone().cpu())
        torch.cuda.synchronize()
        eng.close()
    a, b = grads
    cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
    rel = float((a - b).norm() / (b.norm() + 1e-30))
    _report(f"[{arch}] stem gradient collapse vs dense data gradient + masked reduce: cosine {cos:.6f}, rel L2 {rel:.3e}, "
            f"|g| {float(a.norm()):.4e}")
    assert float(b.norm()) > 0
    assert cos >= 0.9999 and rel <= 1e-2


def test_clean_forward_is_not_range_clamped():
    """Perturbation.forward returns x untouched when `adversarial` is False (model.py:82-83): the clean prediction
    (adv_flag = 0) of a clip with black and white pixels — whose normalised values lie outside the scalar clamp bounds
    [-1.735, 2.49] of the adversarial path — must match torchvision on the unclamped input, and differ from the
    clamped one."""
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200.engine import FlickerEngine
    from oracle import oracle_resnet, oracle_torchstack as ots
    B, T = 1, T_CLIP
    model = synthetic.resnet_model("r3d_18", seed=0)
    clip = synthetic.clips_u8(B, T, 112, 112, seed=1011)
    clip[:, :, :30] = 0
    clip[:, :, 80:] = 255
    x = oracle_resnet.normalize_u8(clip)
    lo, hi = ots.value_bounds()
    with torch.no_grad():
        ref = model(x)
        ref_clamped = model(x.clamp(lo, hi))
    eng = FlickerEngine(B, T, arch="r3d_18")
    eng.load_weights(model.state_dict())
    adv = torch.zeros((B, 3, T, 112, 112), dtype=torch.float32, device="cuda")
    eng.apply(clip.cuda(), torch.zeros((T, 3), device="cuda"), adv_flag=0.0, delta_clip=0.1, adv_f32=adv)
    logits = eng.forward().cpu()
    torch.cuda.synchronize()
    assert float((adv.cpu() - x).abs().max()) <= 1e-5, "the clean input must be x itself"
    rel = float((logits - ref).abs().max() / ref.abs().max())
    rel_clamped = float((logits - ref_clamped).abs().max() / ref_clamped.abs().max())
    _report(f"[clean forward] logits rel err vs unclamped torchvision {rel:.3e}, vs clamped {rel_clamped:.3e}")
    assert rel <= 1e-2 and logits.argmax(-1).tolist() == ref.argmax(-1).tolist()
    assert rel < rel_clamped
    eng.close()
