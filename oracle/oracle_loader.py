"""TEST INFRASTRUCTURE — CPU restatement (numpy, float32, operation by operation) of the torch stack's test-time
clip transform and frame sampling.  Imported by tests/ only; the product path is the CUDA kernel behind
`fav_op_resize_crop` (csrc/loader.cu) and `video_dataset.py`.

Parity: PINNED — `tests/golden/loader_golden.npz` holds outputs of the reference's own transform classes
(utils_cv/action_recognition/references/transforms_video.py) and of `VideoDataset._sample_indices` run in the build
container (generator: tests/golden/make_loader_golden.py); `tests/test_cpu_loader.py` checks this file against them:
frame sampling and the copy path exactly, interpolated values within 2 float32 ulps (torch's CPU bilinear kernel is
itself not bit-reproducible across thread counts: it switches between two summation orders).

Reference lines restated here:
  * to_tensor            references/functional_video.py:65-79   (u8 [T,H,W,C] -> float [C,T,H,W] / 255)
  * ResizeVideo          references/transforms_video.py:23-53   (scale = size / min(H, W) when keep_ratio;
                         interpolate(scale_factor=scale, mode="bilinear", align_corners=False))
  * center_crop          references/functional_video.py:52-62   (i = int(round((h - th) / 2.)))
  * normalize            references/functional_video.py:82-97   ((v - mean) / std)
  * _sample_indices      dataset.py:500-539 (uniform offsets when random_shift is off)
  * _get_frames          dataset.py:541-583 (every sample_step-th frame, last frame repeated past the end)
[dep] torch.nn.functional.interpolate (bilinear, CPU): output size floor(n·scale); source index
ratio·(dst + 0.5) − 0.5 with ratio = float(1/scale), clamped at 0; a dimension whose size does not change is copied.
"""
import math

import numpy as np

DEFAULT_MEAN = (0.43216, 0.394666, 0.37645)      # dataset.py:28
DEFAULT_STD = (0.22803, 0.22145, 0.216989)       # dataset.py:29
F32 = np.float32


def resize_geometry(H, W, size=128, keep_ratio=True):
    """(resized_h, resized_w, ratio_h, ratio_w) of ResizeVideo(size, keep_ratio) for an H×W frame."""
    if keep_ratio:
        scale = size / min(H, W)                                  # transforms_video.py:37
        rh, rw = int(math.floor(H * scale)), int(math.floor(W * scale))
        ratio = F32(1.0 / scale)
        return rh, rw, ratio, ratio
    rh = rw = int(size)                                           # transforms_video.py:39
    return rh, rw, F32(H / rh), F32(W / rw)


def center_crop_origin(rh, rw, th, tw):
    assert rh >= th and rw >= tw, "height and width must be no smaller than crop_size"
    return int(round((rh - th) / 2.0)), int(round((rw - tw) / 2.0))     # functional_video.py:60-61


def _taps(n_in, n_out, ratio):
    dst = np.arange(n_out)
    if n_in == n_out:
        return dst, dst, np.ones(n_out, F32), np.zeros(n_out, F32)
    s = F32(ratio) * (dst.astype(F32) + F32(0.5)) - F32(0.5)
    s = np.where(s < 0, F32(0), s).astype(F32)
    i0 = np.minimum(np.floor(s).astype(np.int64), n_in - 1)
    l1 = np.clip(s - i0.astype(F32), F32(0), F32(1)).astype(F32)
    i1 = i0 + (i0 < n_in - 1)
    return i0, i1, (F32(1) - l1).astype(F32), l1


def resize_crop(frames_u8, size=128, crop=112, keep_ratio=True):
    """frames_u8 [T,H,W,3] uint8 -> float32 [T,crop,crop,3] in [0,1]: to_tensor, ResizeVideo, CenterCropVideo
    (layout kept frames-major; the reference's is [C,T,H,W])."""
    T, H, W, _ = frames_u8.shape
    rh, rw, ratio_h, ratio_w = resize_geometry(H, W, size, keep_ratio)
    ci, cj = center_crop_origin(rh, rw, crop, crop)
    y0, y1, ly0, ly1 = (a[ci:ci + crop] for a in _taps(H, rh, ratio_h))
    x0, x1, lx0, lx1 = (a[cj:cj + crop] for a in _taps(W, rw, ratio_w))
    x = frames_u8.astype(F32) / F32(255.0)
    lx0, lx1 = lx0[None, None, :, None], lx1[None, None, :, None]
    ly0, ly1 = ly0[None, :, None, None], ly1[None, :, None, None]
    top = x[:, y0][:, :, x0] * lx0 + x[:, y0][:, :, x1] * lx1
    bot = x[:, y1][:, :, x0] * lx0 + x[:, y1][:, :, x1] * lx1
    return np.ascontiguousarray((top * ly0 + bot * ly1).astype(F32))


def quantize(v01):
    """nearest uint8 of the resized [0,1] clip (what the engine's apply kernel consumes)"""
    return np.ascontiguousarray(np.clip(np.rint(v01 * F32(255.0)), 0, 255).astype(np.uint8))


def normalize_ncthw(v01, mean=DEFAULT_MEAN, std=DEFAULT_STD):
    """[T,h,w,3] in [0,1] -> the reference's transform output [3,T,h,w]"""
    z = (v01 - np.asarray(mean, F32)) / np.asarray(std, F32)
    return np.ascontiguousarray(z.transpose(3, 0, 1, 2)).astype(F32)


def sample_offsets(num_frames, sample_length, sample_step=1, num_samples=1):
    """uniform clip start offsets (dataset.py:522-531; random_shift off, as the attack drivers' test split runs)"""
    presample = sample_length * sample_step
    if num_frames > presample:
        distance = (num_frames - presample + 1) / num_samples
        return np.array([int(distance / 2.0 + distance * x) for x in range(num_samples)])
    return np.zeros((num_samples,), dtype=int)


def frame_indices(num_frames, offset, sample_length, sample_step=1):
    """indices `_get_frames` reads without temporal jitter (dataset.py:556-583): offset, offset+step, ...; once the
    video ends the last frame read is repeated"""
    idx = [i for i in range(offset, offset + sample_length * sample_step, sample_step) if i < num_frames]
    if not idx:
        raise IndexError(f"offset {offset} past the end of a {num_frames}-frame video")
    return idx + [idx[-1]] * (sample_length - len(idx))
