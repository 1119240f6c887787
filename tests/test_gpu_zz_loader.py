"""gpu: the loader's resize + centre-crop kernel (fav_op_resize_crop, csrc/loader.cu) against the numpy oracle and the
golden vectors of the reference's own transform chain; then a decoded-video batch through VideoDataset into the
attack engine.  (File name sorts last on purpose: the kernel was added after the last GPU session of round 1.)"""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_loader as ol
from flickering_adversarial_video_b200 import video_dataset as vd

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loader_golden.npz"))


def _check(clip, im_scale, input_size, exact_f32_tol=0.0):
    T = clip.shape[0]
    x = torch.from_numpy(clip).cuda()
    got_u8 = vd.transform(x, im_scale, input_size).cpu().numpy()
    got_f32 = vd.transform(x, im_scale, input_size, frames_per_clip=T, normalized=True).cpu().numpy()
    v = ol.resize_crop(clip, im_scale, input_size)
    ref_u8, ref_f32 = ol.quantize(v), ol.normalize_ncthw(v)
    assert got_u8.shape == ref_u8.shape and got_f32.shape == (1,) + ref_f32.shape
    # same float32 operations in the same order, no FMA contraction: bit-exact against the restatement.  The gate
    # leaves room for one float32 ulp (a tie flipping one uint8 level) so that a different rounding of 1/255 on some
    # toolchain cannot turn the suite red; the exact counts are printed.
    d8 = np.abs(got_u8.astype(np.int32) - ref_u8.astype(np.int32))
    df = np.abs(got_f32[0] - ref_f32)
    print(f"resize_crop {clip.shape} -> {got_u8.shape}: uint8 mismatches {int((d8 > 0).sum())} / {d8.size}, "
          f"max fp32 error {float(df.max()):.3e}")
    assert int(d8.max()) <= 1 and float((d8 > 0).mean()) <= 1e-3
    assert float(df.max()) <= max(exact_f32_tol, 1e-6)
    return got_u8, got_f32[0]


@pytest.mark.parametrize("k", range(len(GOLD["cases"])))
def test_resize_crop_matches_oracle_and_reference(k):
    H, W, im_scale, input_size, T = (int(v) for v in GOLD["cases"][k])
    _, got_f32 = _check(GOLD[f"clip{k}"], im_scale, input_size)
    # and against the reference's own output (not bit-reproducible itself, see test_cpu_loader.py)
    assert float(np.abs(got_f32 - GOLD[f"norm{k}"]).max()) <= 2.5e-6


def test_resize_crop_full_size_and_batched():
    g = torch.Generator().manual_seed(77)
    clips = (torch.rand((2, 16, 256, 340, 3), generator=g) * 255).round().to(torch.uint8).numpy()
    a, _ = _check(clips[0], 128, 112)
    # two clips in one launch: the normalised output is laid out per clip [B,3,T,h,w]
    x = torch.from_numpy(clips).cuda().reshape(32, 256, 340, 3)
    both = vd.transform(x, 128, 112, frames_per_clip=16, normalized=True).cpu().numpy()
    assert both.shape == (2, 3, 16, 112, 112)
    for b in range(2):
        ref = ol.normalize_ncthw(ol.resize_crop(clips[b], 128, 112))
        assert float(np.abs(both[b] - ref).max()) <= 1e-6
    u8 = vd.transform(x, 128, 112).cpu().numpy()
    assert np.array_equal(u8[:16], a)


def test_resize_crop_argument_errors():
    from flickering_adversarial_video_b200 import _lib as L
    x = torch.zeros((4, 100, 100, 3), dtype=torch.uint8, device="cuda")
    with pytest.raises(AssertionError):                 # crop larger than the resized frame (functional_video.py:56-58)
        vd.transform(x, 64, 112)
    with pytest.raises(ValueError):
        vd.transform(x, 128, 112, frames_per_clip=3, normalized=True)
    out = torch.empty((4, 112, 112, 3), dtype=torch.uint8, device="cuda")
    with pytest.raises(L.FavError, match="multiple of frames_per_clip"):       # the C-ABI's own argument check
        vd.launch_resize_crop(L.ptr(x), 4, 100, 100, L.ptr(out), None, frames_per_clip=3)
    with pytest.raises(ValueError):
        vd.transform(x.float(), 128, 112)


def test_video_files_into_the_attack(tmp_path):
    """mp4 files -> VideoDataset batches (decode, GPU transform) -> one r3d_18 attack step"""
    cv2 = pytest.importorskip("cv2")
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200.attack import FlickerAttack
    root = tmp_path / "videos"
    rng = np.random.RandomState(5)
    for cls in ("a", "b"):
        (root / cls).mkdir(parents=True)
        for i in range(2):
            wr = cv2.VideoWriter(str(root / cls / f"{cls}{i}.mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 25.0, (172, 130))
            if not wr.isOpened():
                pytest.skip("OpenCV cannot encode mp4v here")
            base = rng.randint(0, 255, (130, 172, 3)).astype(np.uint8)
            for t in range(12):
                wr.write(np.roll(base, 3 * t, axis=1))
            wr.release()
    ds = vd.VideoDataset(str(root), seed=1, train_pct=1.0, sample_length=8, sample_step=1, batch_size=2)
    batches = list(ds.train_batches())
    assert len(batches) == 2
    clips, labels = batches[0]
    assert clips.shape == (2, 8, 112, 112, 3) and clips.dtype == torch.uint8 and clips.is_cuda
    assert labels.dtype == torch.int64 and labels.shape == (2,)
    # the batch equals the oracle's transform of the same decoded frames
    idx = ds.train_range[0]
    frames, lab, _ = ds.load_frames(idx)
    ref = ol.quantize(ol.resize_crop(frames[0], 128, 112))
    d = np.abs(clips[0].cpu().numpy().astype(np.int32) - ref.astype(np.int32))
    assert int(d.max()) <= 1 and float((d > 0).mean()) <= 1e-3 and int(labels[0]) == lab
    cfg = {"LAMBDA": 1.0, "BETA_1": 0.5, "TARGETED_ATTACK": False, "IMPROVE_ADV_LOSS": True, "USE_LOGITS": False,
           "PROB_MARGIN": 0.05}
    model = synthetic.resnet_model("r3d_18", seed=0)
    atk = FlickerAttack(model.state_dict(), 2, 8, cfg, arch="r3d_18", delta_clip=0.1)
    pred = atk.predict(clips, adv_flag=0.0).argmax(-1)
    sc = atk.step(clips, pred)
    assert torch.isfinite(sc).all()


def test_fit_many_videos_over_a_video_folder(tmp_path):
    """VideoDataset -> VideoLearnerAdversarial.fit_many_videos: one bounded single-video attack per file, results in
    the reference's `{video}_@{class}.npy` layout (model.py:925-979)"""
    cv2 = pytest.importorskip("cv2")
    from flickering_adversarial_video_b200 import synthetic
    from flickering_adversarial_video_b200.torch_stack import VideoLearnerAdversarial
    root = tmp_path / "videos"
    rng = np.random.RandomState(6)
    for cls in ("a", "b"):
        (root / cls).mkdir(parents=True)
        wr = cv2.VideoWriter(str(root / cls / f"{cls}0.mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 25.0, (160, 120))
        if not wr.isOpened():
            pytest.skip("OpenCV cannot encode mp4v here")
        base = rng.randint(0, 255, (120, 160, 3)).astype(np.uint8)
        for t in range(10):
            wr.write(np.roll(base, 2 * t, axis=0))
        wr.release()
    ds = vd.VideoDataset(str(root), seed=1, train_pct=1.0, sample_length=8, batch_size=1)
    model = synthetic.resnet_model("r3d_18", seed=0)
    # label the videos with the model's own clean prediction so that the attacks start (clean-misclassified ones are skipped)
    from oracle import oracle_resnet
    names = {}
    for i in ds.train_range:
        clip, _, path = ds[i]
        with torch.no_grad():
            pred = int(model(oracle_resnet.normalize_u8(clip.cpu()[None])).argmax())
        ds.video_records[i]._data[1] = pred
        names[pred] = f"class {pred}"
    learner = VideoLearnerAdversarial(dataset=ds, num_classes=400, base_model="r3d_18", sample_length=8,
                                      l_inf_pert_norm=0.2, attack_type="flickering", labaels_id_to_text=names,
                                      weights=model.state_dict(), batch_size=1)
    lp = {"lambda_": 1.0, "beta_1": 0.5, "targeted_attack": False, "target_class_id": None, "target_class_name": None,
          "improve_loss": True, "use_logits": False}
    out = learner.fit_many_videos(1e-2, model_dir=str(tmp_path / "res"), save_model=True, loss_params_dict=lp, n_iter=2,
                                  max_restarts=1, restart_after=40)
    assert set(out) == {"a0", "b0"}
    for name, res in out.items():
        files = [f for f in os.listdir(str(tmp_path / "res")) if f.startswith(name + "_@class_")]
        assert len(files) == 1
        if res is not None:          # bf16 engine and fp32 torchvision can disagree on a near-tie clean prediction
            assert len(res["is_adversarial"]) >= 2 and res["perturbation"][-1].shape == (3, 8, 1, 1)
            saved = np.load(os.path.join(str(tmp_path / "res"), files[0]), allow_pickle=True).tolist()
            assert saved["label"] == res["label"]
