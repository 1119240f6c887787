"""Seeded synthetic inputs shared by the engine, the oracle, the tests and bench.py.

There is no network for checkpoints or datasets, so weights are random-init in the reference's
native layout and naming (`RGB/inception_i3d/<unit>/conv_3d/w` [kt,kh,kw,Cin,Cout],
`…/batch_norm/{beta,moving_mean,moving_variance}`, `Logits/Conv3d_0c_1x1/conv_3d/{w,b}`;
cf. utils/kinetics_i3d_utils.py:41-62 and i3d.py:168-469) and clips are seeded uint8 noise with
spatial/temporal correlation so that a realistic ~1 % of pixels sit at the range limits.
"""
import numpy as np
import torch

ROOT = "RGB/inception_i3d/"

# (name, c0, c1a, c1b, c2a, c2b, c3b) — i3d.py:194-457
I3D_BLOCKS = [
    ("Mixed_3b", 64, 96, 128, 16, 32, 32),
    ("Mixed_3c", 128, 128, 192, 32, 96, 64),
    ("Mixed_4b", 192, 96, 208, 16, 48, 64),
    ("Mixed_4c", 160, 112, 224, 24, 64, 64),
    ("Mixed_4d", 128, 128, 256, 24, 64, 64),
    ("Mixed_4e", 112, 144, 288, 32, 64, 64),
    ("Mixed_4f", 256, 160, 320, 32, 128, 128),
    ("Mixed_5b", 256, 160, 320, 32, 128, 128),
    ("Mixed_5c", 384, 192, 384, 48, 128, 128),
]


def i3d_units():
    """All BN units as (scope, kernel, cin, cout) in execution order."""
    units = [("Conv3d_1a_7x7", 7, 3, 64), ("Conv3d_2b_1x1", 1, 64, 64), ("Conv3d_2c_3x3", 3, 64, 192)]
    cin = 192
    for name, c0, c1a, c1b, c2a, c2b, c3b in I3D_BLOCKS:
        b2b = "Conv3d_0a_3x3" if name == "Mixed_5b" else "Conv3d_0b_3x3"   # i3d.py:418
        units += [
            (f"{name}/Branch_0/Conv3d_0a_1x1", 1, cin, c0),
            (f"{name}/Branch_1/Conv3d_0a_1x1", 1, cin, c1a),
            (f"{name}/Branch_1/Conv3d_0b_3x3", 3, c1a, c1b),
            (f"{name}/Branch_2/Conv3d_0a_1x1", 1, cin, c2a),
            (f"{name}/Branch_2/{b2b}", 3, c2a, c2b),
            (f"{name}/Branch_3/Conv3d_0b_1x1", 1, cin, c3b),
        ]
        cin = c0 + c1b + c2b + c3b
    return units


def i3d_weights(seed=0, num_classes=400, head_std=0.02):
    """He-init conv weights, non-trivial BN statistics (so the fold is exercised), small logits
    bias.  Returns {tf_variable_name: np.float32 array}."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for scope, k, cin, cout in i3d_units():
        fan_in = k * k * k * cin
        w = torch.randn((k, k, k, cin, cout), generator=g) * (2.0 / fan_in) ** 0.5
        out[ROOT + scope + "/conv_3d/w"] = w.numpy()
        out[ROOT + scope + "/batch_norm/beta"] = (torch.randn((1, 1, 1, 1, cout), generator=g) * 0.1).numpy()
        out[ROOT + scope + "/batch_norm/moving_mean"] = (torch.randn((1, 1, 1, 1, cout), generator=g) * 0.1).numpy()
        out[ROOT + scope + "/batch_norm/moving_variance"] = (
            torch.rand((1, 1, 1, 1, cout), generator=g) + 0.5).numpy()
    out[ROOT + "Logits/Conv3d_0c_1x1/conv_3d/w"] = (
        torch.randn((1, 1, 1, 1024, num_classes), generator=g) * head_std).numpy()
    out[ROOT + "Logits/Conv3d_0c_1x1/conv_3d/b"] = (torch.randn((num_classes,), generator=g) * 0.01).numpy()
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


def _lowpass(x):
    """separable [1,2,1]/4 over T,H,W with circular borders (x: [B,T,H,W,3] float)."""
    for dim in (1, 2, 3):
        x = 0.25 * torch.roll(x, 1, dim) + 0.5 * x + 0.25 * torch.roll(x, -1, dim)
    return x


def clips_u8(batch, frames, height=224, width=224, seed=1000, device="cpu", std=52.0):
    """uint8 clips [B,T,H,W,3]: clamp(round(127.5 + std*unit-variance low-passed noise))."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((batch, frames, height, width, 3), dtype=torch.uint8, device=device)
    for b in range(batch):
        z = torch.randn((1, frames, height, width, 3), generator=g, device=device)
        z = _lowpass(z) / (0.375 ** 1.5)   # sum of squared taps per axis = 0.375
        out[b] = torch.clamp(torch.round(127.5 + std * z), 0, 255).to(torch.uint8)[0]
    return out


def clips_u8_extreme(batch, frames, height=224, width=224, seed=2000, device="cpu"):
    """iid uniform uint8 with ~6 % of pixels forced to 0 or 255: exercises the range-clip mask."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randint(0, 256, (batch, frames, height, width, 3), generator=g, device=device, dtype=torch.int32)
    r = torch.rand((batch, frames, height, width, 3), generator=g, device=device)
    x = torch.where(r < 0.03, torch.zeros_like(x), x)
    x = torch.where(r > 0.97, torch.full_like(x, 255), x)
    return x.to(torch.uint8)


def delta_uniform(frames, seed=7, lo=-0.5, hi=0.5):
    """delta ~ U(lo,hi) [T,3]; the default range exceeds +-0.4 so the inner clip fires."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand((frames, 3), generator=g) * (hi - lo) + lo


# ---- torch stack: torchvision video ResNets (utils_cv/action_recognition/model.py:403-441) -------------
def i3d_weights_framework_default(seed=0, num_classes=400):
    """The second fixture of SURVEY §8(d): the frameworks' DEFAULT initialisers, "random-init" taken literally —
    Sonnet's Conv3D default is a truncated normal (±2σ) with σ = 1/√fan_in and zero bias, snt.BatchNorm starts with
    β = 0, moving mean 0, moving variance 1 [dep: dm-sonnet 1.23].  Through ~22 ReLU layers the activations shrink by
    ≈ 2⁻¹¹ and the softmax is near-uniform, so this fixture probes the small-signal regime rather than parity margins."""
    g = torch.Generator().manual_seed(seed)

    def trunc(shape, std):
        w = torch.empty(shape)
        torch.nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=g)
        return w

    out = {}
    for scope, k, cin, cout in i3d_units():
        out[ROOT + scope + "/conv_3d/w"] = trunc((k, k, k, cin, cout), (k * k * k * cin) ** -0.5).numpy()
        out[ROOT + scope + "/batch_norm/beta"] = np.zeros((1, 1, 1, 1, cout), np.float32)
        out[ROOT + scope + "/batch_norm/moving_mean"] = np.zeros((1, 1, 1, 1, cout), np.float32)
        out[ROOT + scope + "/batch_norm/moving_variance"] = np.ones((1, 1, 1, 1, cout), np.float32)
    out[ROOT + "Logits/Conv3d_0c_1x1/conv_3d/w"] = trunc((1, 1, 1, 1024, num_classes), 1024 ** -0.5).numpy()
    out[ROOT + "Logits/Conv3d_0c_1x1/conv_3d/b"] = np.zeros((num_classes,), np.float32)
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


def resnet_model_framework_default(arch, seed=0, num_classes=400):
    """torchvision.models.video.<arch>(weights=None) exactly as torchvision initialises it (kaiming fan_out convs,
    BatchNorm weight 1 / bias 0 / running stats 0 and 1, Linear N(0, 0.01)), eval mode — the torch half of the second
    fixture of SURVEY §8(d)."""
    import torchvision
    with torch.random.fork_rng():
        torch.manual_seed(seed)
        model = getattr(torchvision.models.video, arch)(weights=None, num_classes=num_classes)
    return model.eval()


def resnet_model(arch, seed=0, num_classes=400, head_gain=10.0):
    """Random-init torchvision.models.video.<arch> in eval mode: torchvision's own initialisers under a
    fixed seed, plus non-trivial BatchNorm statistics (so the fold is exercised) and a head scaled so the
    softmax is not degenerate.  Its `state_dict()` is what `FlickerEngine.load_weights` ingests."""
    import torchvision
    with torch.random.fork_rng():
        torch.manual_seed(seed)
        model = getattr(torchvision.models.video, arch)(weights=None, num_classes=num_classes)
    g = torch.Generator().manual_seed(seed + 1)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm3d):
            n = mod.num_features
            mod.weight.data = 0.8 + 0.4 * torch.rand(n, generator=g)
            mod.bias.data = 0.1 * torch.randn(n, generator=g)
            mod.running_mean = 0.1 * torch.randn(n, generator=g)
            mod.running_var = 0.5 + torch.rand(n, generator=g)
    model.fc.weight.data *= head_gain
    return model.eval()
