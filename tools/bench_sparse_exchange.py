"""Per-pixel universal attack on several GPUs: step time and the cost of its gradient exchange (one packed sum-all-reduce of
[T,H,W,3] fp32 + scalars: 38.5 MB for I3D at T = 64, 2.4 MB for the 112x112 nets).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_sparse_exchange.py [arch] [batch_per_gpu] [frames]
Prints one JSON line on rank 0 (CUDA-event times, max over ranks)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from flickering_adversarial_video_b200 import synthetic
from flickering_adversarial_video_b200.attack import SparseAttack


def main():
    arch = sys.argv[1] if len(sys.argv) > 1 else "i3d"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    T = int(sys.argv[3]) if len(sys.argv) > 3 else (64 if arch == "i3d" else 16)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    side = 224 if arch == "i3d" else 112
    weights = synthetic.i3d_weights(0) if arch == "i3d" else synthetic.resnet_model(arch, 0).state_dict()
    atk = SparseAttack(weights, B, T, {"LAMBDA": 1.0, "BETA_1": 0.5}, device=local, arch=arch)
    clips = synthetic.clips_u8(B, T, side, side, seed=500 + rank, device=f"cuda:{local}")
    labels = atk.predict(clips, adv_flag=0.0).argmax(-1)

    def timed(fn, n):
        for _ in range(3):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{local}")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    step_ms = timed(lambda: atk.step(clips, labels), 10)
    ex_ms = timed(lambda: dist.all_reduce(atk.comm), 20) if world > 1 else 0.0
    atk.check_replicas()
    if rank == 0:
        nbytes = atk.comm.numel() * 4
        print(json.dumps({"tool": "bench_sparse_exchange", "arch": arch, "n_gpus": world, "batch_per_gpu": B, "frames": T,
                          "ms_per_step": step_ms, "exchange_ms": ex_ms, "exchange_bytes": nbytes,
                          "exchange_algbw_gbs": (nbytes / ex_ms / 1e6) if ex_ms else None,
                          "clip_frames_per_sec": world * B * T / step_ms * 1e3}))
    atk.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
