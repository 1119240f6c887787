"""Generate golden vectors for the torch stack's clip loader by running the REFERENCE's own transform chain
(`get_transforms(train=False)`: ToTensorVideo -> ResizeVideo(128) -> CenterCropVideo(112) -> flip(0) -> NormalizeVideo,
utils_cv/action_recognition/dataset.py:84-123) and `VideoDataset._sample_indices` / `_get_frames`
(dataset.py:500-583) on the CPU of this container.

    python tests/golden/make_loader_golden.py     ->  tests/golden/loader_golden.npz

The reference cannot travel to the GPU box, so the vectors are committed.  `decord`, `IPython` and `matplotlib` are
stubbed as in make_torch_stack_golden.py; `_get_frames` is driven by a small in-memory reader with decord's
seek_accurate / next / skip_frames interface.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_torch_stack_golden import _stub, REF   # noqa: E402

OUT = os.path.join(HERE, "loader_golden.npz")
# (H, W, im_scale, input_size, T): the default 128 / 112 chain once, a 32 / 28 chain (same code path, small fixture) for
# down-scaling landscape / portrait / square frames, a frame whose short side already is im_scale (copy path) and an
# up-scaled one
CASES = [(150, 200, 128, 112, 1), (64, 85, 32, 28, 2), (90, 64, 32, 28, 2), (40, 40, 32, 28, 2), (32, 43, 32, 28, 2),
         (24, 18, 32, 28, 2)]


def synth_frames(T, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((T, H, W, 3), generator=g)
    x[:, : H // 3] = torch.round(x[:, : H // 3])          # saturated block (0 / 255)
    return (x * 255).round().to(torch.uint8)


class _Frame:
    def __init__(self, a):
        self.a = a

    def asnumpy(self):
        return self.a


class _Reader:
    """decord.VideoReader's sequential interface over an array of frame ids"""

    def __init__(self, n):
        self.n, self.pos = n, 0

    def seek_accurate(self, i):
        self.pos = i

    def skip_frames(self, k):
        self.pos += k

    def next(self):
        if self.pos >= self.n:
            raise StopIteration
        self.pos += 1
        return _Frame(np.array(self.pos - 1))


def main():
    for n in ("decord", "IPython", "IPython.display", "matplotlib", "matplotlib.pyplot"):
        _stub(n)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    import utils_cv.action_recognition.dataset as ds   # noqa: E402

    out = {"cases": np.array(CASES), "torch_version": np.array(torch.__version__),
           "num_threads": np.array(torch.get_num_threads())}
    for k, (H, W, im_scale, input_size, T) in enumerate(CASES):
        cfg = ds.get_default_tfms_config(train=False)
        cfg.im_scale, cfg.input_size = im_scale, input_size
        tf = ds.get_transforms(train=False, tfms_config=cfg)
        resize_only = ds.transforms.ResizeVideo(im_scale, True)
        clip = synth_frames(T, H, W, 500 + k)
        out[f"clip{k}"] = clip.numpy()
        out[f"norm{k}"] = tf(clip).numpy()                                        # [3,T,input_size,input_size]
        out[f"resized_shape{k}"] = np.array(resize_only(ds.transforms.ToTensorVideo()(clip)).shape[-2:])

    # frame sampling (test split: random_shift False, temporal_jitter False)
    cases = []
    for num_frames, length, step, samples in [(300, 16, 1, 1), (250, 16, 2, 1), (40, 16, 3, 1), (16, 16, 1, 1),
                                              (10, 16, 1, 1), (100, 8, 4, 3), (33, 16, 2, 2)]:
        d = ds.VideoDataset.__new__(ds.VideoDataset)
        d.presample_length, d.num_samples, d.random_shift, d.warning = length * step, samples, False, False
        d.sample_length, d.sample_step, d.temporal_jitter = length, step, False
        rec = ds.VideoRecord(["x", 0])
        rec._num_frames = num_frames
        offs = np.asarray(d._sample_indices(rec))
        idx = np.array([[int(f) for f in d._get_frames(_Reader(num_frames), int(o))] for o in offs])
        cases.append((num_frames, length, step, samples))
        out[f"offsets{len(cases) - 1}"] = offs
        out[f"indices{len(cases) - 1}"] = idx
    out["sampling_cases"] = np.array(cases)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
