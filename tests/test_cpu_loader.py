"""not gpu: the clip loader's oracle (oracle/oracle_loader.py) against golden vectors produced by the reference's own
transform classes and VideoDataset sampling (tests/golden/make_loader_golden.py), and the host logic of
video_dataset.py (geometry, sampling, mp4 decode through OpenCV, split files)."""
import os

import numpy as np
import pytest

from oracle import oracle_loader as ol
from flickering_adversarial_video_b200 import video_dataset as vd

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loader_golden.npz"))


@pytest.mark.parametrize("k", range(len(GOLD["cases"])))
def test_oracle_transform_matches_reference(k):
    H, W, im_scale, input_size, T = (int(v) for v in GOLD["cases"][k])
    clip, ref = GOLD[f"clip{k}"], GOLD[f"norm{k}"]
    rh, rw, _, _ = ol.resize_geometry(H, W, im_scale)
    assert (rh, rw) == tuple(GOLD[f"resized_shape{k}"])
    z = ol.normalize_ncthw(ol.resize_crop(clip, im_scale, input_size))
    assert z.shape == ref.shape == (3, T, input_size, input_size)
    # torch's CPU bilinear kernel is not bit-reproducible across thread counts / memory formats (it picks between two
    # summation orders); the restatement fixes one order, so the pin is "within 2 float32 ulps of a [0,1] value
    # through the 1/std ~ 4.6 scale": 4 * 1.2e-7 * 4.6
    assert float(np.abs(z - ref).max()) <= 2.5e-6
    if min(H, W) == im_scale:          # nothing is interpolated: exact copy + crop + normalise
        assert np.array_equal(z, ref)


def test_oracle_sampling_matches_reference():
    for i, (n, length, step, samples) in enumerate(GOLD["sampling_cases"]):
        offs = ol.sample_offsets(int(n), int(length), int(step), int(samples))
        assert np.array_equal(offs, GOLD[f"offsets{i}"])
        idx = np.array([ol.frame_indices(int(n), int(o), int(length), int(step)) for o in offs])
        assert np.array_equal(idx, GOLD[f"indices{i}"])
        assert np.array_equal(vd.sample_offsets(int(n), int(length), int(step), int(samples)), offs)


def test_host_geometry_matches_oracle():
    for H, W in [(240, 320), (256, 340), (360, 480), (128, 171), (480, 270), (112, 112), (113, 300), (720, 1280)]:
        for size in (128, 32, 112):
            a, b = vd.resize_geometry(H, W, size), ol.resize_geometry(H, W, size)
            assert a[:2] == b[:2] and a[2] == b[2] and a[3] == b[3]
            crop = min(a[0], a[1], 112)
            assert vd.center_crop_origin(a[0], a[1], crop, crop) == ol.center_crop_origin(a[0], a[1], crop, crop)
    # the reference's torchvision-style tuple / keep_ratio=False variants (transforms_video.py:31-45)
    assert vd.resize_geometry(100, 200, (50, 50), True)[:2] == (25, 50)
    assert vd.resize_geometry(100, 200, (50, 60), False)[:2] == (50, 60)
    assert vd.resize_geometry(100, 200, 64, False)[:2] == (64, 64)
    with pytest.raises(AssertionError):
        vd.center_crop_origin(100, 128, 112, 112)


def test_quantisation_is_nearest_uint8():
    v = np.array([0.0, 1.0, 0.5, 127.5 / 255, 128.4999 / 255, 254.51 / 255], np.float32)
    assert ol.quantize(v).tolist() == [0, 255, 128, 128, 128, 255]


def _write_video(path, n, h=48, w=64):
    cv2 = pytest.importorskip("cv2")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 25.0, (w, h))
    if not wr.isOpened():
        pytest.skip("OpenCV cannot encode mp4v here")
    for i in range(n):
        wr.write(np.full((h, w, 3), 8 * i + 4, np.uint8))          # frame id in the grey level
    wr.release()


def _ids(clip):
    return [int(round((float(f.mean()) - 4.0) / 8.0)) for f in clip]


def test_read_clip_frame_selection(tmp_path):
    path = str(tmp_path / "v.mp4")
    _write_video(path, 30)
    assert vd.video_num_frames(path) == 30
    for offset, length, step in [(0, 8, 1), (5, 8, 2), (3, 6, 3), (20, 16, 1), (22, 8, 2), (29, 4, 1)]:
        clip = vd.read_clip(path, offset, length, step)
        assert clip.shape == (length, 48, 64, 3) and clip.dtype == np.uint8
        assert _ids(clip) == ol.frame_indices(30, offset, length, step), (offset, length, step)
    with pytest.raises(IOError):
        vd.read_clip(str(tmp_path / "missing.mp4"), 0, 4)


def test_dataset_splits_and_decode(tmp_path):
    root = tmp_path / "videos"
    for cls, n in (("dancing", 3), ("cooking", 2)):
        (root / cls).mkdir(parents=True)
        for i in range(n):
            _write_video(str(root / cls / f"{cls}_{i}.mp4"), 24)
    ds = vd.VideoDataset(str(root), seed=3, train_pct=0.6, sample_length=8, sample_step=2, batch_size=2)
    assert len(ds) == 5 and sorted(ds.classes) == ["cooking", "dancing"]
    assert len(ds.test_range) == 2 and len(ds.train_range) == 3
    assert sorted(ds.train_range + ds.test_range) == list(range(5))
    clips, label, path = ds.load_frames(ds.train_range[0])
    assert clips.shape == (1, 8, 48, 64, 3) and os.path.basename(path).startswith(ds.classes[label])
    # 24 frames, presample 16 -> uniform offset int(9 / 2) = 4 (dataset.py:522-531)
    assert _ids(clips[0]) == list(range(4, 20, 2))

    train, test = tmp_path / "train.txt", tmp_path / "test.txt"
    train.write_text("dancing/dancing_0,0,dancing\ndancing/dancing_1,0,dancing\ncooking/cooking_0,1,cooking\n")
    test.write_text("cooking/cooking_1,1\n")
    ds = vd.VideoDataset(str(root), sample_length=4, train_split_file=str(train), test_split_file=str(test))
    assert ds.train_range == [0, 1, 2] and ds.test_range == [3]
    assert ds.video_records[2].label == 1 and ds.video_records[2].label_name == "cooking"
    assert ds.video_records[3].label_name is None
    assert ds.video_path(ds.video_records[3]).endswith("cooking/cooking_1.mp4")
    with pytest.raises(NotImplementedError):
        vd.VideoDataset(str(root), temporal_jitter=True)


def test_transform_without_gpu_is_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ValueError):
        vd.transform(torch.zeros((2, 8, 8, 3), dtype=torch.uint8))


def test_host_batches_prefetch_thread(tmp_path):
    """full batches in order, early exit stops the decode thread, decode errors reach the consumer"""
    import threading
    root = tmp_path / "videos" / "c"
    root.mkdir(parents=True)
    for i in range(5):
        _write_video(str(root / f"v{i}.mp4"), 10 + i)
    split = tmp_path / "s.txt"
    split.write_text("".join(f"c/v{i},{i}\n" for i in range(5)))
    ds = vd.VideoDataset(str(tmp_path / "videos"), sample_length=4, batch_size=2, prefetch=1,
                         train_split_file=str(split), test_split_file=str(split))
    got = list(ds._host_batches(ds.train_range))
    assert [[lab for _, lab, _ in b] for b in got] == [[0, 1], [2, 3]]              # the odd fifth video is dropped
    assert all(c.shape == (1, 4, 48, 64, 3) for b in got for c, _, _ in b)
    n0 = threading.active_count()
    it = ds._host_batches(ds.train_range)
    next(it)
    it.close()                                                                     # consumer walks away
    assert threading.active_count() <= n0
    os.remove(str(root / "v2.mp4"))
    with pytest.raises(IOError, match="v2"):
        list(ds._host_batches(ds.train_range))


def test_dataset_accepts_the_reference_call(tmp_path):
    """the constructor call of r2plus1d_main_universal_attack.py:153-169, argument for argument"""
    root = tmp_path / "videos" / "c"
    root.mkdir(parents=True)
    _write_video(str(root / "v0.mp4"), 12)
    split = tmp_path / "s.txt"
    split.write_text("c/v0,0\n")
    ds = vd.VideoDataset(str(tmp_path / "videos"), seed=None, train_pct=0.75, num_samples=1, sample_length=16,
                         sample_step=1, temporal_jitter=False, temporal_jitter_step=2, random_shift=False, batch_size=1,
                         warning=False, train_split_file=str(split), test_split_file=str(split), video_ext="mp4",
                         train_transforms=vd.get_transforms(train=False), test_transforms=vd.get_transforms(train=False))
    assert (ds.im_scale, ds.input_size) == (128, 112) and len(ds) == 2
    assert vd.VideoDataset(str(tmp_path / "videos"), train_split_file=str(split), test_split_file=str(split),
                           test_transforms=vd.get_transforms(False, {"im_scale": 64, "input_size": 56})).input_size == 56
    with pytest.raises(NotImplementedError):
        vd.get_transforms(train=True)
    with pytest.raises(TypeError):
        vd.VideoDataset(str(tmp_path / "videos"), test_transforms=lambda x: x)


def test_batch_cache_is_transparent(tmp_path):
    """cache=True: a split's batches are produced once; an abandoned first pass caches nothing"""
    import torch
    split = tmp_path / "s.txt"
    split.write_text("")
    made = []

    def make():
        made.append(1)
        for i in range(3):
            yield torch.full((2, 4), i, dtype=torch.uint8), torch.tensor([i, i])

    ds = vd.VideoDataset(str(tmp_path), train_split_file=str(split), test_split_file=str(split), cache=True)
    it = ds._cached("train", make)
    next(it)
    it.close()
    assert "train" not in ds._cache                                  # incomplete pass
    a = [(x.clone(), y.clone()) for x, y in ds._cached("train", make)]
    b = list(ds._cached("train", make))
    assert len(made) == 2 and len(b) == 3 and all(torch.equal(p[0], q[0]) and torch.equal(p[1], q[1]) for p, q in zip(a, b))
    assert ds.cached_bytes() == 3 * (8 + 16)
    ds2 = vd.VideoDataset(str(tmp_path), train_split_file=str(split), test_split_file=str(split))
    list(ds2._cached("train", make)), list(ds2._cached("train", make))
    assert len(made) == 4 and ds2._cache == {}                       # cache off: decoded every epoch


def test_resized_size_equals_torch_interpolate():
    """the output size ResizeVideo produces is torch's own floor(n * scale_factor) — checked against torch on a grid of
    frame sizes (floating-point edge cases of 128 / min(H, W))"""
    import torch
    import torch.nn.functional as F
    for H in list(range(112, 300, 7)) + [320, 360, 480, 540, 720, 1080]:
        for W in list(range(112, 300, 11)) + [320, 340, 426, 480, 640, 854, 1280, 1920]:
            out = F.interpolate(torch.zeros((1, 1, H, W)), scale_factor=128 / min(H, W), mode="bilinear",
                                align_corners=False).shape[-2:]
            assert tuple(out) == vd.resize_geometry(H, W, 128)[:2] == ol.resize_geometry(H, W, 128)[:2], (H, W)


@pytest.mark.parametrize("H,W", [(256, 340), (240, 320), (360, 270), (128, 171), (171, 128), (113, 199), (480, 854)])
def test_oracle_resize_against_torch_interpolate(H, W):
    """the [dep] behind ResizeVideo is torch.nn.functional.interpolate itself: the restatement agrees with it to float32
    rounding on full-size frames (torch's CPU kernel picks its summation order by thread count / memory format)"""
    import torch
    import torch.nn.functional as F
    frames = (torch.rand((2, H, W, 3), generator=torch.Generator().manual_seed(H * 1000 + W)) * 255).round().to(torch.uint8)
    x = frames.float().permute(3, 0, 1, 2) / 255.0                                   # to_tensor: [C,T,H,W]
    r = F.interpolate(x, scale_factor=128 / min(H, W), mode="bilinear", align_corners=False)
    i, j = ol.center_crop_origin(r.shape[-2], r.shape[-1], 112, 112)
    ref = r[..., i:i + 112, j:j + 112].permute(1, 2, 3, 0).numpy()                    # back to [T,h,w,C]
    got = ol.resize_crop(frames.numpy(), 128, 112)
    assert got.shape == ref.shape and float(np.abs(got - ref).max()) <= 3e-7


def test_relative_root_folder_split(tmp_path, monkeypatch):
    """folder-split records carry the root already (dataset.py:372-380): a relative root must not be joined twice"""
    (tmp_path / "vids" / "c").mkdir(parents=True)
    _write_video(str(tmp_path / "vids" / "c" / "v0.mp4"), 10)
    monkeypatch.chdir(tmp_path)
    ds = vd.VideoDataset("vids", train_pct=1.0, sample_length=4)
    assert ds.video_path(ds.video_records[0]) == os.path.join("vids", "c", "v0.mp4")
    assert ds.load_frames(0)[0].shape == (1, 4, 48, 64, 3)
