import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flickering_adversarial_video_b200 import synthetic
from flickering_adversarial_video_b200.engine import FlickerEngine
from oracle import oracle_i3d
weights = synthetic.i3d_weights(seed=0)
for B, T in [(1, 32), (1, 64), (1, 90), (1, 89)]:
    clip = synthetic.clips_u8(B, T, seed=1090)
    delta = synthetic.delta_uniform(T, seed=17, lo=-0.05, hi=0.05)
    model = oracle_i3d.OracleI3D(weights)
    modelq = oracle_i3d.OracleI3D(weights, emulate_bf16=True)
    x = oracle_i3d.normalize_u8(clip)
    with torch.no_grad():
        labels = model.forward(x).argmax(-1)
    cfg = dict(improve_loss=True, margin=0.05, beta0=1.0, beta1=0.5, beta2=0.5, beta3=0.5, lr=1e-3)
    ref = oracle_i3d.attack_step(model, x, labels, delta, cfg, data_grad_only=True)
    refq = oracle_i3d.attack_step(modelq, x, labels, delta, cfg, data_grad_only=True)
    eng = FlickerEngine(B, T)
    eng.load_weights(weights)
    eng.apply(clip.cuda(), delta.cuda())
    logits = eng.forward().cpu()
    eng.loss(labels.cuda(), improve_loss=True, margin=0.05)
    g = eng.backward().cpu()
    torch.cuda.synchronize()
    cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
    gr, gq = ref["grad_data"], refq["grad_data"]
    # per-frame cosine to localise
    pf = [(t, round(cos(g[t], gr[t]), 3)) for t in range(T)]
    worst = sorted(pf, key=lambda z: z[1])[:6]
    print(f"B={B} T={T}: cos engine/fp32 {cos(g, gr):.4f}  engine/bf16-oracle {cos(g, gq):.4f}  bf16-oracle/fp32 {cos(gq, gr):.4f}  |g| {float(g.norm()):.3e}/{float(gr.norm()):.3e}; worst frames {worst}")
    eng.close()
