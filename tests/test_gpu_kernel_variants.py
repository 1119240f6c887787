"""-m gpu: the kernels added in the second half of round 2 against the kernels they replace, through the C-ABI.

Each variant is selected at plan time by an environment switch, so the same engine call runs both:
  * FAV_STEM_TS=0     raw-row stem kernel instead of the temporal-sharing one (conv_stem.cu)
  * FAV_T3=0          per-tap kernel instead of the shared-frame (3,1,1) kernel (conv_t3.cu)
  * FAV_TAP_KG=1      one k-block per barrier hand-off in the per-tap kernel (conv_umma.cu)
Both sides multiply the same fp16 operands into fp32 accumulators; only the order of the accumulation differs, so the
stem output must agree to fp16 rounding and everything downstream to the usual mask-flip noise.  The clip lengths are
chosen so that the last group of output frames is ragged (fewer than 4 frames) and so that both temporal parities of the
I3D stem see out-of-range input frames."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(arch, B, T, H, W, env, monkeypatch, read=None, seed=0, labels=None):
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200 import _lib as L
    from flickering_adversarial_video_b200.engine import FlickerEngine
    for k in ("FAV_STEM_TS", "FAV_T3", "FAV_TAP_KG"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    if arch == "i3d":
        eng = FlickerEngine(B, T, H, W)
        eng.load_weights(synthetic.i3d_weights(seed=seed))
        stack = L.FAV_STACK_TF
    else:
        eng = FlickerEngine(B, T, H, W, arch=arch)
        eng.load_weights(synthetic.resnet_model(arch, seed=seed).state_dict())
        stack = L.FAV_STACK_TORCH
    clip = synthetic.clips_u8(B, T, H, W, seed=41)
    clip[:, :, : H // 4] = clip[:, :, : H // 4] // 16                 # dark band: the range clip fires
    delta = synthetic.delta_uniform(T, seed=3, lo=-0.08, hi=0.08)
    eng.apply(clip.cuda(), delta.cuda(), delta_clip=0.1)
    logits = eng.forward().clone()
    if labels is None:
        labels = logits.argmax(-1)                                      # the predicted class: the margin loss is active
    eng.loss(labels.cuda(), improve_loss=True, margin=0.05, stack=stack)
    grad = eng.backward().clone().cpu()
    out = {"logits": logits.cpu(), "grad": grad, "labels": labels.cpu()}
    if read is not None:
        name, shape = read
        out["layer"] = eng.read(name, shape).cpu()
    torch.cuda.synchronize()
    eng.close()
    return out


def _close(a, b, what, rel_tol):
    rel = float((a - b).norm() / (b.norm() + 1e-30))
    print(f"{what}: rel L2 {rel:.3e}")
    assert rel <= rel_tol, f"{what}: relative L2 {rel}"
    return rel


@pytest.mark.parametrize("T", [16, 10, 9])
def test_i3d_temporal_sharing_stem_matches_raw_rows(T, monkeypatch):
    """Conv3d_1a_7x7 (i3d.py:168-171): To = ceil(T/2) = 8 / 5 / 5 output frames, i.e. full and ragged groups of 4, even
    and odd clip lengths (TF SAME padding differs between them)."""
    B, H, W = 1, 224, 224
    To = (T + 1) // 2
    shape = (B, To, H // 2, W // 2, 64)
    a = _run("i3d", B, T, H, W, {}, monkeypatch, read=("Conv3d_1a_7x7", shape))
    b = _run("i3d", B, T, H, W, {"FAV_STEM_TS": "0"}, monkeypatch, read=("Conv3d_1a_7x7", shape))
    assert float(b["layer"].abs().max()) > 0
    _close(a["layer"], b["layer"], f"I3D T={T} stem output, temporal sharing vs raw rows", 2e-3)
    _close(a["logits"], b["logits"], f"I3D T={T} logits", 2e-2)


@pytest.mark.parametrize("arch,T", [("r3d_18", 6), ("mc3_18", 5), ("r3d_18", 8)])
def test_torchvision_stem_temporal_sharing_matches_raw_rows(arch, T, monkeypatch):
    """(3,7,7) stride (1,2,2) stems (torchvision BasicStem): stride 1 in T, one frame class, ragged last group."""
    B, H, W = 2, 112, 112
    shape = (B, T, 56, 56, 64)
    a = _run(arch, B, T, H, W, {}, monkeypatch, read=("stem.conv", shape))
    b = _run(arch, B, T, H, W, {"FAV_STEM_TS": "0"}, monkeypatch, read=("stem.conv", shape))
    assert float(b["layer"].abs().max()) > 0
    _close(a["layer"], b["layer"], f"{arch} T={T} stem output, temporal sharing vs raw rows", 2e-3)
    _close(a["logits"], b["logits"], f"{arch} T={T} logits", 2e-2)


@pytest.mark.parametrize("T", [8, 6, 3])
def test_r2plus1d_shared_frame_temporal_conv_matches_per_tap(T, monkeypatch):
    """Conv2Plus1D's (3,1,1) halves with 64 output channels (stem 45 -> 64, layer1 144 -> 64, with and without the
    residual addend): conv_t3_kernel vs the per-tap kernel, T a multiple of 4, ragged, and shorter than one group."""
    B, H, W = 2, 112, 112
    a = _run("r2plus1d_18", B, T, H, W, {}, monkeypatch, read=("layer1.1", (B, T, 56, 56, 64)))
    b = _run("r2plus1d_18", B, T, H, W, {"FAV_T3": "0"}, monkeypatch, read=("layer1.1", (B, T, 56, 56, 64)),
             labels=a["labels"])
    assert float(b["layer"].abs().max()) > 0
    _close(a["layer"], b["layer"], f"r2plus1d_18 T={T} layer1 output, shared frames vs per tap", 5e-3)
    _close(a["logits"], b["logits"], f"r2plus1d_18 T={T} logits", 2e-2)
    cos = float((a["grad"] * b["grad"]).sum() / (a["grad"].norm() * b["grad"].norm() + 1e-30))
    print(f"r2plus1d_18 T={T} dL/d-delta cosine between the two forward kernels {cos:.6f}")
    assert float(b["grad"].norm()) > 0 and cos >= 0.995


@pytest.mark.parametrize("arch", ["r2plus1d_18", "mc3_18"])
def test_grouped_k_blocks_match_single_hand_offs(arch, monkeypatch):
    """Per-tap kernel: k-blocks handed over in groups (planner's choice) vs one per barrier (FAV_TAP_KG=1); the MMA order is
    the same, so the network's results are bit-identical (the [T,3] gradient is summed with atomics by the stem gradient
    kernel: equal up to the order of those fp32 additions)."""
    B, T, H, W = 2, 8, 112, 112
    a = _run(arch, B, T, H, W, {"FAV_T3": "0"}, monkeypatch)
    b = _run(arch, B, T, H, W, {"FAV_T3": "0", "FAV_TAP_KG": "1"}, monkeypatch, labels=a["labels"])
    assert float(a["grad"].norm()) > 0
    assert torch.equal(a["logits"], b["logits"])
    assert torch.allclose(a["grad"], b["grad"], rtol=1e-4, atol=1e-6 * float(a["grad"].abs().max()))
