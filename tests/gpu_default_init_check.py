"""Manual GPU check (not collected by pytest): engine vs oracle on the SECOND fixture of SURVEY §8(d) — the frameworks'
default initialisers ("random-init" taken literally, synthetic.*_framework_default).  Activations shrink by ~2^-11 through
I3D and the softmax is near-uniform, so this reports the small-signal behaviour (relative logit error, top-1 agreement,
dL/d-delta cosine) instead of asserting margins.      python tests/gpu_default_init_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flickering_adversarial_video_b200 import synthetic, _lib as L   # noqa: E402
from flickering_adversarial_video_b200.engine import FlickerEngine    # noqa: E402
from oracle import oracle_i3d, oracle_resnet                          # noqa: E402


def report(tag, logits, ref_logits, g, gr):
    rel = float((logits - ref_logits).abs().max() / ref_logits.abs().max())
    cos = float((g * gr).sum() / (g.norm() * gr.norm() + 1e-30))
    print(f"{tag}: logits std {float(ref_logits.std()):.3e}, max rel err {rel:.3e}, top-1 engine "
          f"{logits.argmax(-1).tolist()} oracle {ref_logits.argmax(-1).tolist()}, |g| engine {float(g.norm()):.3e} oracle "
          f"{float(gr.norm()):.3e}, dL/d-delta cosine {cos:.5f}")


def main():
    B, T = 1, 16
    w = synthetic.i3d_weights_framework_default(0)
    clip = synthetic.clips_u8(B, T, seed=1001)
    delta = synthetic.delta_uniform(T, seed=7, lo=-0.05, hi=0.05)
    model = oracle_i3d.OracleI3D(w)
    x = oracle_i3d.normalize_u8(clip)
    with torch.no_grad():
        labels = model.forward(x).argmax(-1)
    cfg = dict(improve_loss=True, margin=0.05, beta0=1.0, beta1=0.5, beta2=0.5, beta3=0.5, lr=1e-3)
    ref = oracle_i3d.attack_step(model, x, labels, delta, cfg, data_grad_only=True)
    eng = FlickerEngine(B, T)
    eng.load_weights(w)
    eng.apply(clip.cuda(), delta.cuda())
    logits = eng.forward().cpu()
    eng.loss(labels.cuda(), improve_loss=True, margin=0.05)
    g = eng.backward().cpu()
    report("I3D, Sonnet default init", logits, ref["logits"], g, ref["grad_data"])
    eng.close()

    for arch in ("r3d_18", "mc3_18", "r2plus1d_18"):
        T = 8
        m = synthetic.resnet_model_framework_default(arch, seed=0)
        clip = synthetic.clips_u8(2, T, 112, 112, seed=1003)
        delta = synthetic.delta_uniform(T, seed=9, lo=-0.08, hi=0.08)
        with torch.no_grad():
            labels = m(oracle_resnet.normalize_u8(clip)).argmax(-1)
        ref = oracle_resnet.attack_step(m, clip, labels, delta, max_norm=0.1)
        eng = FlickerEngine(2, T, arch=arch)
        eng.load_weights(m.state_dict())
        eng.apply(clip.cuda(), delta.cuda(), delta_clip=0.1)
        logits = eng.forward().cpu()
        eng.loss(labels.cuda(), improve_loss=True, margin=0.05, stack=L.FAV_STACK_TORCH)
        g = eng.backward().cpu()
        report(f"{arch}, torchvision default init", logits, ref["logits"], g, ref["grad_data"])
        eng.close()


if __name__ == "__main__":
    main()
