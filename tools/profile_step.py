"""One attack step of the bench workload between cudaProfilerStart/Stop (for ncu
--profile-from-start off).  Usage: python tools/profile_step.py [batch] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from flickering_adversarial_video_b200 import synthetic
from flickering_adversarial_video_b200.attack import FlickerAttack

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = {"IMPROVE_ADV_LOSS": True, "PROB_MARGIN": 0.05, "LAMBDA": 10.0, "BETA_1": 0.5, "BETA_2": 0.5}
atk = FlickerAttack(synthetic.i3d_weights(seed=0), B, T, cfg)
clips = synthetic.clips_u8(B, T, seed=1000, device="cuda")
labels = atk.predict(clips, adv_flag=0.0).argmax(-1)
for _ in range(2):
    atk.step(clips, labels)
torch.cuda.synchronize()
torch.cuda.profiler.start()
atk.step(clips, labels)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step; scalars", atk.scalars[:10].tolist())
