"""'Reference as a user would run it on the same GPU' (SURVEY.md §8(d)(iii)): the attack step in plain PyTorch on cuda —
cuDNN fp32 convolutions (TF32 off, as on the reference's Titan-X), autograd to the perturbation, torch Adam — for the
bench workloads.  A reported comparison, never a product path and never a bench value; the network is the oracle's
restatement (I3D) or torchvision (video ResNets), because /root/reference cannot travel to the GPU box.

    python tests/gpu_torch_reference_timing.py i3d 8 64        # BASELINE.json configs[1]
    python tests/gpu_torch_reference_timing.py r3d_18 16 16    # the per-GPU shard of configs[3]
Prints one JSON line per run: ms/step, clip-frames/s, peak memory.  `--tf32` allows TF32 tensor cores (what torch does
by default for cuDNN convolutions), `--amp` runs the forward under bf16 autocast."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from flickering_adversarial_video_b200 import synthetic  # noqa: E402
from oracle import oracle_i3d, oracle_resnet, oracle_torchstack as ots  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    arch = args[0] if args else "i3d"
    B = int(args[1]) if len(args) > 1 else (8 if arch == "i3d" else 16)
    T = int(args[2]) if len(args) > 2 else (64 if arch == "i3d" else 16)
    tf32, amp = "--tf32" in sys.argv, "--amp" in sys.argv
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    cpu = "--cpu" in sys.argv                  # dry run of this script without a GPU (wall-clock timing)
    dev = torch.device("cpu") if cpu else torch.device("cuda", 0)
    if arch == "i3d":
        model = oracle_i3d.OracleI3D(synthetic.i3d_weights(seed=0))
        model.w = {k: v.to(dev) for k, v in model.w.items()}
        clips = synthetic.clips_u8(B, T, seed=1).to(dev)
        delta = torch.zeros((T, 3), device=dev, requires_grad=True)

        def loss_of(labels):
            x = oracle_i3d.normalize_u8(clips)
            logits = model.forward(oracle_i3d.apply_flicker(x, delta))
            adv, _, _ = oracle_i3d.improve_adversarial_loss(logits.float(), labels)
            thick, diff, lap, _, _ = oracle_i3d.regularizers(delta)
            return adv + 1.0 * (0.5 * thick + 0.5 * diff + 0.5 * lap), logits
    else:
        net = synthetic.resnet_model(arch, seed=0).to(dev).eval()
        for p in net.parameters():
            p.requires_grad_(False)
        clips = synthetic.clips_u8(B, T, 112, 112, seed=1).to(dev)
        delta = ((torch.rand((3, T, 1, 1), device=dev) * 2 - 1) * 1e-6).requires_grad_(True)
        std = torch.tensor(ots.DEFAULT_STD, device=dev).reshape(3, 1, 1, 1)
        lo, hi = ots.value_bounds()

        def loss_of(labels):
            x = oracle_resnet.normalize_u8(clips)
            pc = delta.clamp(-0.1, 0.1)
            logits = net((x + pc / std).clamp(lo, hi)).float()
            adv = ots.improve_adversarial_loss(labels, logits, torch.softmax(logits, 1), 0.05, False)
            return adv + 1.0 * ots.flickering_regularization_loss(pc, 0.5), logits
    opt = torch.optim.Adam([delta], lr=1e-3)
    with torch.no_grad(), torch.autocast(dev.type, dtype=torch.bfloat16, enabled=amp):
        labels = loss_of(torch.zeros(B, dtype=torch.int64, device=dev))[1].argmax(-1)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=amp):
            loss, _ = loss_of(labels)
        loss.backward()
        opt.step()

    import time
    sync = (lambda: None) if cpu else torch.cuda.synchronize
    for _ in range(3):
        step()
    sync()
    if not cpu:
        torch.cuda.reset_peak_memory_stats()
    n = 2 if cpu else 10
    if cpu:
        t0 = time.perf_counter()
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    for _ in range(n):
        step()
    if cpu:
        ms = (time.perf_counter() - t0) * 1e3 / n
    else:
        e1.record()
        sync()
        ms = e0.elapsed_time(e1) / n
    print(json.dumps({"tool": "gpu_torch_reference_timing", "arch": arch, "batch": B, "frames": T, "tf32": tf32, "bf16_autocast": amp,
                      "ms_per_step": ms, "clip_frames_per_sec": B * T / ms * 1e3,
                      "peak_mem_gib": None if cpu else torch.cuda.max_memory_allocated() / 2 ** 30, "device": dev.type,
                      "note": "plain PyTorch (cuDNN) attack step on the same GPU; a comparison, not a bench value"}))


if __name__ == "__main__":
    main()
