import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flickering_adversarial_video_b200 import synthetic
from flickering_adversarial_video_b200.attack import FlickerAttack
arch = sys.argv[1] if len(sys.argv) > 1 else "r2plus1d_18"
B, T = 16, 16
model = synthetic.resnet_model(arch, seed=0)
atk = FlickerAttack(model.state_dict(), B, T, {"LAMBDA": 1.0, "BETA_1": 0.5}, arch=arch)
clips = synthetic.clips_u8(B, T, 112, 112, seed=1, device="cuda")
labels = atk.predict(clips, adv_flag=0.0).argmax(-1)
atk.step(clips, labels)
torch.cuda.synchronize()
print("=== profiled step", flush=True)
atk.step(clips, labels)
torch.cuda.synchronize()
