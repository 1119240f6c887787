"""FAV_HALO_PROF=1 python tools/halo_prof.py : wait-cycle breakdown of the halo conv on the conv2c shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flickering_adversarial_video_b200.engine import op_conv3d
g = torch.Generator(device="cuda").manual_seed(0)
B, T, H, W = 8, 32, 56, 56
for cin, cout, dgrad in [(64, 192, False), (64, 192, True), (128, 192, False)]:
    if cin == 128:
        H = W = 28
    kc = cout if dgrad else cin
    x = torch.randn((B, T, H, W, kc), generator=g, device="cuda").to(torch.bfloat16 if dgrad else torch.float16)
    w = torch.randn((3, 3, 3, cin, cout), generator=g, device="cuda") * 0.05
    for _ in range(2):
        op_conv3d(x, w, relu=not dgrad, dgrad=dgrad)
    torch.cuda.synchronize()
