"""FAV_HALO_PROF=1 [FAV_HALO_MT=m FAV_HALO_NT=n] python tools/halo_sweep.py : per-launch cycle counts of the halo
conv at the Inception 3x3x3 shapes of the bench workload (B=8, T=64), forward and data gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flickering_adversarial_video_b200.engine import op_conv3d

CASES = [  # name, (T,H,W), cin, cout
    ("2c", (32, 56, 56), 64, 192),
    ("3b.b1b", (32, 28, 28), 96, 128),
    ("3b.b2b", (32, 28, 28), 16, 32),
    ("3c.b1b", (32, 28, 28), 128, 192),
    ("3c.b2b", (32, 28, 28), 32, 96),
    ("4b.b1b", (16, 14, 14), 96, 208),
    ("4c.b1b", (16, 14, 14), 112, 224),
    ("4e.b1b", (16, 14, 14), 144, 288),
    ("4f.b1b", (16, 14, 14), 160, 320),
    ("4f.b2b", (16, 14, 14), 32, 128),
    ("x.c32", (32, 28, 28), 32, 96),
    ("x.c64", (32, 28, 28), 64, 96),
    ("x.c128", (32, 28, 28), 128, 96),
    ("x.c16", (32, 28, 28), 16, 96),
    ("y.n32", (32, 28, 28), 32, 32),
    ("y.n64", (32, 28, 28), 32, 64),
    ("y.n128", (32, 28, 28), 32, 128),
    ("y.n192", (32, 28, 28), 32, 192),
    ("y.k64n32", (32, 28, 28), 64, 32),
    ("y.k64n192", (32, 28, 28), 64, 192),
]
only = sys.argv[1:] or None
g = torch.Generator(device="cuda").manual_seed(0)
B = 8
for name, (T, H, W), cin, cout in CASES:
    if only and not any(name.startswith(o) for o in only):
        continue
    for dgrad in (False, True):
        kc = cout if dgrad else cin
        x = torch.randn((B, T, H, W, kc), generator=g, device="cuda").to(torch.bfloat16 if dgrad else torch.float16)
        w = torch.randn((3, 3, 3, cin, cout), generator=g, device="cuda") * 0.05
        sys.stderr.write(f"== {name} {'dgrad' if dgrad else 'fwd'} cin={cin} cout={cout}\n")
        sys.stderr.flush()
        for _ in range(2):
            op_conv3d(x, w, relu=not dgrad, dgrad=dgrad)
        torch.cuda.synchronize()
