"""Measured precision table for DESIGN.md §4 (manual GPU script, not collected by pytest).

For one BASELINE.json shape it prints, against the fp32 CPU oracle (oracle/oracle_i3d.py, the restatement of the
reference's fp32 arithmetic): the engine's logits error / top-1 / dL/d-delta cosine, and the same three numbers for the
oracle network run by plain PyTorch on the same GPU in strict fp32 and with TF32 convolutions allowed (torch's default
for cuDNN, and what TF 1.15 NGC builds do on Ampere and later) — i.e. what the reference itself would compute on this
GPU.  A 10-bit-mantissa tensor-core format (TF32, fp16) is the ceiling any single-pass tensor-core implementation has.

    python tests/gpu_precision_report.py 8 64        # BASELINE configs[1]
    python tests/gpu_precision_report.py 1 90        # configs[0]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from flickering_adversarial_video_b200 import synthetic  # noqa: E402
from flickering_adversarial_video_b200.engine import FlickerEngine  # noqa: E402
from oracle import oracle_i3d as O  # noqa: E402

CFG = dict(improve_loss=True, margin=0.05, beta0=1.0, beta1=0.5, beta2=0.5, beta3=0.5, lr=1e-3)


def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-300))


def oracle_per_clip(model, x, labels, delta, dev):
    """Margin loss = sum over clips (kinetics_i3d_utils.py:285) -> gradient = sum of per-clip gradients."""
    logits, g = [], torch.zeros((x.shape[1], 3), dtype=torch.float64)
    for b in range(x.shape[0]):
        out = O.attack_step(model, x[b:b + 1].to(dev), labels[b:b + 1].to(dev), delta.to(dev), CFG, data_grad_only=True)
        logits.append(out["logits"].cpu())
        g += out["grad_data"].double().cpu()
    return torch.cat(logits), g


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1001
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    weights = synthetic.i3d_weights(seed=0)
    clip = synthetic.clips_u8(B, T, seed=seed)
    delta = synthetic.delta_uniform(T, seed=7, lo=-0.05, hi=0.05)
    x = O.normalize_u8(clip)
    m_cpu = O.OracleI3D(weights)
    with torch.no_grad():
        labels = torch.cat([m_cpu.forward(x[b:b + 1]).argmax(-1) for b in range(B)])
    ref_logits, ref_g = oracle_per_clip(m_cpu, x, labels, delta, torch.device("cpu"))
    rows = {}

    def row(name, logits, g):
        rel = float((logits.cpu() - ref_logits).abs().max() / ref_logits.abs().max())
        rows[name] = {"logits_rel": rel, "top1_equal": bool((logits.cpu().argmax(-1) == ref_logits.argmax(-1)).all()),
                      "cosine": cos(g, ref_g)}
        print(f"{name:34s} logits rel {rel:.3e}  top-1 equal {rows[name]['top1_equal']}  dL/d-delta cosine {rows[name]['cosine']:.6f}",
              flush=True)

    # ---- engine ----
    eng = FlickerEngine(B, T)
    eng.load_weights(weights)
    eng.apply(clip.cuda(), delta.cuda())
    logits = eng.forward().clone()
    eng.loss(labels.cuda(), improve_loss=True, margin=0.05)
    g = eng.backward().clone()
    torch.cuda.synchronize()
    row("engine (libfav)", logits, g)
    eng.close()
    # ---- the same network in plain PyTorch on this GPU ----
    dev = torch.device("cuda", 0)
    m_gpu = O.OracleI3D(weights)
    m_gpu.w = {k: v.to(dev) for k, v in m_gpu.w.items()}
    for name, tf32 in (("torch cuda fp32 (TF32 off)", False), ("torch cuda TF32 convolutions", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        lg, gg = oracle_per_clip(m_gpu, x, labels, delta, dev)
        row(name, lg, gg)
    print(json.dumps({"tool": "gpu_precision_report", "batch": B, "frames": T, "seed": seed, "rows": rows}))


if __name__ == "__main__":
    main()
