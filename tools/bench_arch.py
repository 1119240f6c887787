"""Step time of the torch-stack archs (BASELINE.json configs[3]/[4] shapes): python tools/bench_arch.py [arch] [B] [T]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flickering_adversarial_video_b200 import synthetic, _lib
from flickering_adversarial_video_b200.attack import FlickerAttack

archs = [sys.argv[1]] if len(sys.argv) > 1 else ["r2plus1d_18", "r3d_18", "mc3_18"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T = int(sys.argv[3]) if len(sys.argv) > 3 else 16
FLOP = {"r3d_18": 162.79e9, "mc3_18": 173.37e9, "r2plus1d_18": 162.08e9}   # per clip-iteration at 16x112x112
lib = _lib.load()
for arch in archs:
    model = synthetic.resnet_model(arch, seed=0)
    atk = FlickerAttack(model.state_dict(), B, T, {"LAMBDA": 1.0, "BETA_1": 0.5}, arch=arch)
    clips = synthetic.clips_u8(B, T, 112, 112, seed=1, device="cuda")
    labels = atk.predict(clips, adv_flag=0.0).argmax(-1)
    g = atk.capture(clips, labels)
    for _ in range(3):
        atk.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 20
    for _ in range(n):
        atk.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    lib.fav_profile_begin()
    for _ in range(3):
        atk.step(clips, labels)
    buf = (ctypes.c_double * (4 * len(_lib.PROF_KINDS)))()
    lib.fav_profile_end(buf, len(buf))
    parts = {k: round(buf[4 * i] / 3, 3) for i, k in enumerate(_lib.PROF_KINDS) if buf[4 * i + 1] > 0}
    tf = B * FLOP[arch] * (T / 16) / (ms * 1e-3) / 1e12
    print(f"{arch}: B={B} T={T}: {ms:.3f} ms/step, {B * T / ms * 1e3:.0f} clip-frames/s, {tf:.0f} TFLOP/s algorithmic "
          f"({100 * tf / 1391.3:.1f} % of sustained bf16 peak); ms per family {parts}")
    atk.close()
