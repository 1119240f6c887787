"""Where the end-to-end dL/d-delta cosine of a 16-bit forward goes (manual CPU script, not collected by pytest; it imports
the oracle, so it lives under tests/).  The fp32 oracle is re-run with the engine's fp16 arithmetic restated on the CPU
(OracleI3D(emulate="fp16")) while the rounding of the ACTIVATIONS is switched on for one layer group at a time; the weights
are always fp16-rounded.  Three 16-frame clips, random-init fixture.  Result: fp16 WEIGHTS alone cost
1.3e-3 of cosine (0.9987 < 0.999) and every layer group's activation rounding another 0.2-0.7e-3; all together 4.5e-3
(0.9955).  Deciding the max-pool arg-max on unrounded activations recovers about half of the activation part (0.9977).

    python tests/cpu_precision_budget.py
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flickering_adversarial_video_b200 import synthetic
from oracle import oracle_i3d as O
torch.set_num_threads(len(os.sched_getaffinity(0)))

class Var(O.OracleI3D):
    def __init__(self, w, keep_fp32=(), only=None, pool_raw=False):
        super().__init__(w, emulate="fp16")
        self.keep, self.only, self.pool_raw = keep_fp32, only, pool_raw
    def unit(self, x, scope, stride=(1, 1, 1), delta_img=None):
        wf, bias = self.folded(scope)
        wq = wf.to(self.fmt).to(self.dtype)
        y = O.conv3d_same(x, wq, stride) + bias.reshape(1, -1, 1, 1, 1)
        if delta_img is not None:
            y = y + O.conv3d_same(delta_img, wf, stride)
        y = F.relu(y)
        grp = scope.split("/")[0]
        rnd = (grp not in self.keep) if self.only is None else (grp in self.only)
        out = O._round_st(y, self.fmt) if rnd else y
        out._raw = y
        return out

def run(model, x, labels, delta):
    cfg = dict(improve_loss=True, margin=0.05, beta0=1.0, beta1=0.5, beta2=0.5, beta3=0.5)
    return O.attack_step(model, x, labels, delta, cfg, data_grad_only=True)
def cos(a, b): return float((a.double() * b.double()).sum() / (a.double().norm() * b.double().norm()))

w = synthetic.i3d_weights(0)
groups = ["Conv3d_1a_7x7", "Conv3d_2b_1x1", "Conv3d_2c_3x3", "Mixed_3b", "Mixed_3c", "Mixed_4b", "Mixed_4c", "Mixed_4d", "Mixed_4e", "Mixed_4f", "Mixed_5b", "Mixed_5c"]
res = {}
for seed in (1001, 1002, 1003):
    clip = synthetic.clips_u8(1, 16, seed=seed)
    x = O.normalize_u8(clip)
    delta = synthetic.delta_uniform(16, seed=7, lo=-0.05, hi=0.05)
    m32 = O.OracleI3D(w)
    with torch.no_grad():
        labels = m32.forward(x).argmax(-1)
    ref = run(m32, x, labels, delta)["grad_data"]
    res.setdefault("all rounded", []).append(cos(run(Var(w), x, labels, delta)["grad_data"], ref))
    for g in groups:
        res.setdefault("only " + g, []).append(cos(run(Var(w, only=(g,)), x, labels, delta)["grad_data"], ref))
    res.setdefault("none rounded (weights only)", []).append(cos(run(Var(w, only=()), x, labels, delta)["grad_data"], ref))
for k, v in res.items():
    print(f"{k:32s} " + " ".join(f"{c:.5f}" for c in v) + f"   mean 1-cos {sum(1-c for c in v)/len(v):.2e}")
