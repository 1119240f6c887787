"""Time the stand-alone pool kernels at the I3D shapes of the bench workload (B=8, T=64):
python tools/bench_ops.py   (CUDA events, 5 reps after 2 warm-ups, inputs > L2 in aggregate)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from flickering_adversarial_video_b200.engine import op_maxpool3d, op_maxpool3d_bwd

CASES = [
    ("pool2a", (8, 32, 112, 112, 64), (1, 3, 3), (1, 2, 2)),
    ("pool3a", (8, 32, 56, 56, 192), (1, 3, 3), (1, 2, 2)),
    ("3b", (8, 32, 28, 28, 192), (3, 3, 3), (1, 1, 1)),
    ("3c", (8, 32, 28, 28, 256), (3, 3, 3), (1, 1, 1)),
    ("pool4a", (8, 32, 28, 28, 480), (3, 3, 3), (2, 2, 2)),
    ("4b", (8, 16, 14, 14, 480), (3, 3, 3), (1, 1, 1)),
    ("4f", (8, 16, 14, 14, 528), (3, 3, 3), (1, 1, 1)),
    ("pool5a", (8, 16, 14, 14, 832), (2, 2, 2), (2, 2, 2)),
    ("5b", (8, 8, 7, 7, 832), (3, 3, 3), (1, 1, 1)),
]


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


tot_f = tot_b = 0.0
for name, shape, k, s in CASES:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(shape, generator=g, device="cuda").clamp_min(0).to(torch.float16)
    y, idx = op_maxpool3d(x, k, s)
    dy = torch.randn(y.shape, generator=g, device="cuda").to(torch.bfloat16)
    add = torch.randn(shape, generator=g, device="cuda").to(torch.bfloat16)
    tf = timeit(lambda: op_maxpool3d(x, k, s))
    tb = timeit(lambda: op_maxpool3d_bwd(dy, idx, shape, k, s, add=add, relu_src=x))
    fb = (x.numel() * 2 + y.numel() * 3) / 1e6
    bb = (y.numel() * 3 + x.numel() * 6) / 1e6
    print(f"{name:7s} fwd {tf:8.1f} us ({fb / tf * 1e3:6.0f} GB/s)   bwd {tb:8.1f} us ({bb / tb * 1e3:6.0f} GB/s)")
