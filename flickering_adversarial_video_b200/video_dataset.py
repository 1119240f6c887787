"""Video-file input of the torch stack: host mirror of the reference's `VideoRecord` /
`VideoDataset` (utils_cv/action_recognition/dataset.py:30-82, 246-625) that feeds `VideoLearnerAdversarial.fit` /
`fit_single_video` with uint8 DEVICE clips.

What differs from the reference, and why:
  * decoding goes through OpenCV's FFmpeg backend instead of `decord` (absent here); the sequential access pattern of
    `_get_frames` (seek, read, skip `sample_step - 1`, repeat the last frame past the end, dataset.py:541-583) is kept;
  * the transform chain of `get_transforms(train=False)` — the one the attack drivers use for BOTH splits
    (r2plus1d_main_statistics_single_video_attack.py:168-169) — runs on the GPU as one kernel over the decoded uint8
    frames (`fav_op_resize_crop`, csrc/loader.cu): ToTensorVideo, ResizeVideo(im_scale, keep_ratio), CenterCropVideo
    (input_size).  NormalizeVideo is not applied here: the engine's apply kernel normalises (row a1), so batches stay
    uint8 — the nearest uint8 of the reference's resized [0,1] clip (|error| <= 0.5/255 per entry; bf16 storage of the
    network input is coarser).  `transform(..., normalized=True)` returns the reference's float tensor instead;
  * batches are assembled by one background thread into pinned host frames (decode of batch i+1 overlaps the attack
    step on batch i); the reference uses DataLoader worker processes (dataset.py:476-494);
  * training-split augmentation (random resized crop, flip, temporal jitter) is not built: the attack never uses it.
There is no CPU fallback for the transform: without libfav.so / a GPU it raises.
"""
import math
import os
import queue
import threading
import warnings
from pathlib import Path

import numpy as np

from . import _lib as L

DEFAULT_MEAN = (0.43216, 0.394666, 0.37645)       # dataset.py:28
DEFAULT_STD = (0.22803, 0.22145, 0.216989)        # dataset.py:29


class VideoRecord:
    """dataset.py:30-82: `[path, label(, label_name)]` rows of a split file."""

    def __init__(self, data):
        assert 2 <= len(data) <= 3
        assert isinstance(data[0], str)
        int(data[1])
        if len(data) == 3:
            assert isinstance(data[2], str)
        self._data = data
        self._num_frames = None

    @property
    def path(self):
        return self._data[0]

    @property
    def num_frames(self):
        if self._num_frames is None:       # frame folders (`img_*` files), as the reference counts them (:68-71)
            self._num_frames = int(len([x for x in Path(self._data[0]).glob("img_*")]) - 1)
        return self._num_frames

    @property
    def label(self):
        return int(self._data[1])

    @property
    def label_name(self):
        return None if len(self._data) <= 2 else self._data[2]


class TransformSpec:
    """What `get_transforms(train=False)` stands for here: the parameters of the fixed chain ToTensorVideo ->
    ResizeVideo(im_scale, keep_ratio) -> CenterCropVideo(input_size) -> NormalizeVideo(mean, std) (dataset.py:84-123,
    212-243); resize + crop run in `fav_op_resize_crop`, the normalisation in the engine's apply kernel."""

    def __init__(self, im_scale=128, input_size=112, mean=DEFAULT_MEAN, std=DEFAULT_STD):
        if tuple(mean) != DEFAULT_MEAN or tuple(std) != DEFAULT_STD:
            raise NotImplementedError("the engine normalises with the Kinetics mean / std of dataset.py:28-29")
        self.im_scale, self.input_size, self.mean, self.std = im_scale, input_size, tuple(mean), tuple(std)

    def __eq__(self, other):
        return isinstance(other, TransformSpec) and vars(self) == vars(other)

    def __repr__(self):
        return f"TransformSpec(im_scale={self.im_scale}, input_size={self.input_size})"


def get_transforms(train=True, tfms_config=None):
    """dataset.py:84-123.  Only the test-time chain exists (flip ratio 0, centre crop): the attack drivers build both
    splits with `get_transforms(train=False)`.  `tfms_config`: an object / dict with `im_scale` and `input_size`."""
    if train:
        raise NotImplementedError("training-split augmentation (random resized crop, flip) is not built")
    if tfms_config is None:
        return TransformSpec()
    get = tfms_config.get if isinstance(tfms_config, dict) else lambda k, d=None: getattr(tfms_config, k, d)
    return TransformSpec(get("im_scale", 128), get("input_size", 112), get("mean", DEFAULT_MEAN), get("std", DEFAULT_STD))


# ---- geometry of ResizeVideo + CenterCropVideo -----------------------------------------------------------------
def resize_geometry(H, W, size=128, keep_ratio=True):
    """(resized_h, resized_w, ratio_h, ratio_w): transforms_video.py:23-53 over torch interpolate's size / ratio rules
    (output = floor(n * scale); the kernel's source-index ratio is float(1 / scale))."""
    if isinstance(size, (tuple, list)):
        if keep_ratio:
            scale = min(size[0] / H, size[1] / W)
        else:
            rh, rw = int(size[0]), int(size[1])
            return rh, rw, np.float32(H / rh), np.float32(W / rw)
    elif keep_ratio:
        scale = size / min(H, W)
    else:
        rh = rw = int(size)
        return rh, rw, np.float32(H / rh), np.float32(W / rw)
    ratio = np.float32(1.0 / scale)
    return int(math.floor(H * scale)), int(math.floor(W * scale)), ratio, ratio


def center_crop_origin(rh, rw, th, tw):
    """functional_video.py:52-62"""
    assert rh >= th and rw >= tw, "height and width must be no smaller than crop_size"
    return int(round((rh - th) / 2.0)), int(round((rw - tw) / 2.0))


def launch_resize_crop(src_ptr, n_frames, H, W, u8_ptr, f32_ptr, im_scale=128, input_size=112, keep_ratio=True,
                       frames_per_clip=None, mean=DEFAULT_MEAN, std=DEFAULT_STD, device=0, stream_ptr=None):
    """Raw-pointer form of `transform` (DEVICE addresses as ints / c_void_p, `stream_ptr` a cudaStream_t or None for
    the default stream): computes the ResizeVideo / CenterCropVideo geometry on the host and launches
    `fav_op_resize_crop`.  Returns (resized_h, resized_w, crop_i, crop_j)."""
    th = tw = int(input_size)
    rh, rw, ratio_h, ratio_w = resize_geometry(H, W, im_scale, keep_ratio)
    ci, cj = center_crop_origin(rh, rw, th, tw)
    nrm = L.NormParams((L.C.c_float * 3)(*mean), (L.C.c_float * 3)(*std), 0.0, 0.0)
    L.check(L.load().fav_op_resize_crop(int(device), src_ptr, int(n_frames), int(H), int(W), rh, rw, float(ratio_h),
                                        float(ratio_w), ci, cj, th, tw, int(frames_per_clip or n_frames),
                                        L.C.byref(nrm), u8_ptr, f32_ptr, stream_ptr), "fav_op_resize_crop")
    return rh, rw, ci, cj


def transform(frames_u8, im_scale=128, input_size=112, keep_ratio=True, frames_per_clip=None, normalized=False,
              mean=DEFAULT_MEAN, std=DEFAULT_STD, out=None, stream=None):
    """frames_u8: uint8 CUDA tensor [N,H,W,3] (frames of one or more clips of the same size).  Returns uint8
    [N,input_size,input_size,3], or with normalized=True the reference's transform output, float32
    [N / frames_per_clip, 3, frames_per_clip, input_size, input_size]."""
    import torch
    if not (frames_u8.is_cuda and frames_u8.dtype == torch.uint8 and frames_u8.dim() == 4 and frames_u8.shape[-1] == 3):
        raise ValueError("transform expects a uint8 CUDA tensor [N,H,W,3]")
    frames_u8 = frames_u8.contiguous()
    N, H, W, _ = frames_u8.shape
    S = int(input_size)
    fpc = int(frames_per_clip or N)
    if fpc <= 0 or N % fpc:
        raise ValueError(f"{N} frames are not a whole number of {fpc}-frame clips")
    dev = frames_u8.device
    if out is None:
        out = (torch.empty((N // fpc, 3, fpc, S, S), dtype=torch.float32, device=dev) if normalized
               else torch.empty((N, S, S, 3), dtype=torch.uint8, device=dev))
    assert out.is_contiguous() and out.device == dev
    assert out.dtype == (torch.float32 if normalized else torch.uint8) and out.numel() == N * S * S * 3
    with torch.cuda.device(dev):
        launch_resize_crop(L.ptr(frames_u8), N, H, W, None if normalized else L.ptr(out),
                           L.ptr(out) if normalized else None, im_scale, input_size, keep_ratio, fpc, mean, std,
                           dev.index or 0, L.stream_ptr(stream))
    return out


# ---- frame sampling ---------------------------------------------------------------------------------------------
def sample_offsets(num_frames, sample_length, sample_step=1, num_samples=1, random_shift=False, rng=None):
    """dataset.py:500-539"""
    presample = sample_length * sample_step
    if num_frames > presample:
        if random_shift:
            rng = rng or np.random
            return np.sort(rng.randint(num_frames - presample + 1, size=num_samples))
        distance = (num_frames - presample + 1) / num_samples
        return np.array([int(distance / 2.0 + distance * x) for x in range(num_samples)])
    return np.zeros((num_samples,), dtype=int)


def read_clip(path, offset, sample_length, sample_step=1, out=None):
    """dataset.py:541-583 over cv2.VideoCapture: frames offset, offset + step, ...; the last decoded frame is
    repeated once the video ends.  Returns uint8 RGB [sample_length,H,W,3] (`out`: a preallocated array to fill)."""
    import cv2
    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        raise IOError(f"cannot open video {path}")
    try:
        if offset > 0:
            # decord's seek_accurate (dataset.py:557): position on exactly frame `offset`.  OpenCV's FFmpeg backend
            # seeks to the preceding key frame and decodes forward; if a container reports another position, fall
            # back to decoding from the start.
            cap.set(cv2.CAP_PROP_POS_FRAMES, int(offset))
            if int(round(cap.get(cv2.CAP_PROP_POS_FRAMES))) != int(offset):
                cap.release()
                cap = cv2.VideoCapture(path)
                for _ in range(int(offset)):
                    if not cap.grab():
                        break
        n = 0
        for k in range(sample_length):
            if k > 0:
                ended = False
                for _ in range(sample_step - 1):
                    if not cap.grab():
                        ended = True
                        break
                if ended:
                    break
            ok, bgr = cap.read()
            if not ok:
                break
            if out is None:
                out = np.empty((sample_length,) + bgr.shape, dtype=np.uint8)
            cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB, dst=out[n])
            n += 1
        if n == 0:
            raise IOError(f"no frame decoded from {path} at offset {offset}")
        out[n:sample_length] = out[n - 1]
        return out[:sample_length]
    finally:
        cap.release()


def video_num_frames(path):
    import cv2
    cap = cv2.VideoCapture(path)
    try:
        if not cap.isOpened():
            raise IOError(f"cannot open video {path}")
        return int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    finally:
        cap.release()


class VideoDataset:
    """dataset.py:246-498.  Same constructor call as the reference's; the transform chain is fixed to the attack's
    test-time transforms (`get_transforms(train=False)` -> `TransformSpec`); `train_batches` / `test_batches` are the
    per-epoch batch sources `VideoLearnerAdversarial.fit` takes (the reference's `train_dl` / `test_dl`)."""

    def __init__(self, root, seed=None, train_pct=0.75, num_samples=1, sample_length=8, sample_step=1,
                 temporal_jitter=False, temporal_jitter_step=2, random_shift=False, batch_size=8, video_ext="mp4",
                 warning=False, train_split_file=None, test_split_file=None, train_transforms=None, test_transforms=None,
                 *, im_scale=None, input_size=None, device=0, prefetch=2, num_workers=None, cache=False):
        """Positional arguments as in the reference (dataset.py:249-266), except that `temporal_jitter` / `random_shift`
        default to False (the attack mains always pass False).  `train_transforms` / `test_transforms` take what
        `get_transforms(train=False)` of this module returns (a `TransformSpec`); both splits use the test-time chain."""
        assert sample_step > 0
        assert num_samples > 0
        spec = test_transforms if test_transforms is not None else train_transforms
        if spec is not None and not isinstance(spec, TransformSpec):
            raise TypeError("train_transforms / test_transforms: pass video_dataset.get_transforms(train=False) "
                            "(the transform chain runs as one CUDA kernel, not as Python callables)")
        if train_transforms is not None and test_transforms is not None and train_transforms != test_transforms:
            raise NotImplementedError("different transforms for the two splits")
        spec = spec or TransformSpec()
        im_scale = spec.im_scale if im_scale is None else im_scale
        input_size = spec.input_size if input_size is None else input_size
        if temporal_jitter or random_shift:
            raise NotImplementedError("training-split augmentation (temporal jitter / random shift) is not built: the "
                                      "attack drivers evaluate both splits with the test-time sampling")
        for f in (train_split_file, test_split_file):
            if f:
                assert Path(f).exists(), f
        if train_split_file or test_split_file:
            assert train_split_file and test_split_file
        self.root, self.seed = root, seed
        self.num_samples, self.sample_length, self.sample_step = num_samples, sample_length, sample_step
        self.presample_length = sample_length * sample_step
        self.batch_size, self.video_ext, self.warning = batch_size, video_ext, warning
        self.im_scale, self.input_size, self.device, self.prefetch = im_scale, input_size, device, prefetch
        # decode threads (OpenCV releases the GIL while decoding); the reference uses DataLoader worker processes
        # (dataset.py:476-494, db_num_workers())
        self.num_workers = min(8, os.cpu_count() or 1) if num_workers is None else max(1, int(num_workers))
        # cache=True keeps the transformed uint8 DEVICE batches of a split after its first complete pass (602 KB per
        # 16x112x112 clip: a 10 000-clip split is 6 GB of the 180 GB HBM).  Test-time sampling and transforms are
        # deterministic, so later epochs see exactly the same batches — without decoding or host->device copies.
        self.cache, self._cache = bool(cache), {}
        if train_split_file:
            self.train_range, self.test_range = self.split_with_file(train_split_file, test_split_file)
        else:
            self.train_range, self.test_range = self.split_by_folder(train_pct)

    def __len__(self):
        return len(self.video_records)

    def set_shard(self, rank, world, device="cpu", group=None):
        """Sharded `fit` (one process per GPU): every rank adopts RANK 0's train / test split (an unseeded
        torch.randperm differs per process) and keeps an equal-count shard of each (every world-th video, remainder
        dropped), so that the ranks attack different clips and run the same number of steps.  The reference trains
        one process over all GPUs (nn.DataParallel, model.py:576-578), i.e. each video once per epoch."""
        from . import dist as fdist
        if world <= 1 or getattr(self, "_shard", None) == (rank, world):
            return
        full_train = getattr(self, "_full_train", None)
        if full_train is None:
            self._full_train, self._full_test = list(self.train_range), list(self.test_range)
        tr = fdist.broadcast_ints(self._full_train, 0, device, group)
        te = fdist.broadcast_ints(self._full_test, 0, device, group)
        self.train_range, self.test_range = fdist.shard_indices(tr, rank, world), fdist.shard_indices(te, rank, world)
        self._shard = (rank, world)
        self._cache = {}

    def split_by_folder(self, train_pct=0.8):
        """dataset.py:337-399: one folder per class; torch.randperm split (seeded when `seed` is set)"""
        import torch
        self.video_records, self.classes = [], []
        dirs = [e for e in os.listdir(self.root) if os.path.isdir(os.path.join(self.root, e))]
        for label, action in enumerate(dirs):
            self.video_records.extend(VideoRecord([os.path.join(self.root, action, vid.split(".")[0]), label, action])
                                      for vid in os.listdir(os.path.join(self.root, action)))
            self.classes.append(action)
        test_num = math.floor(len(self) * (1 - train_pct))
        if self.seed:
            torch.manual_seed(self.seed)
        indices = torch.randperm(len(self)).tolist()
        return indices[test_num:], indices[:test_num]

    def split_with_file(self, train_split_file, test_split_file):
        """dataset.py:401-446: comma-separated `path,label[,name]` rows"""
        self.video_records = [VideoRecord(row.strip().split(",")) for row in open(train_split_file) if row.strip()]
        train_len = len(self.video_records)
        self.video_records.extend(VideoRecord(row.strip().split(",")) for row in open(test_split_file) if row.strip())
        return list(range(train_len)), list(range(train_len, len(self.video_records)))

    def video_path(self, record):
        """dataset.py:592-596: `root/record.path.ext`.  Folder-split records already carry the root (:372-380); joining
        it twice — harmless for the absolute roots the reference uses — is avoided for relative ones."""
        p = record.path
        if not (os.path.isabs(p) or os.path.normpath(p).startswith(os.path.normpath(self.root) + os.sep)):
            p = os.path.join(self.root, p)
        return "{}.{}".format(p, self.video_ext)

    def load_frames(self, idx):
        """decoded uint8 clips [num_samples, T, H, W, 3] (host), label, record path — `__getitem__` (:585-619) before
        the transforms"""
        record = self.video_records[idx]
        path = self.video_path(record)
        record._num_frames = video_num_frames(path)
        if record.num_frames <= self.presample_length and self.warning:
            warnings.warn(f"num_samples and/or sample_length > num_frames in {record.path}")
        offsets = sample_offsets(record.num_frames, self.sample_length, self.sample_step, self.num_samples)
        clips = np.stack([read_clip(path, int(o), self.sample_length, self.sample_step) for o in offsets])
        return clips, record.label, record.path

    def __getitem__(self, idx):
        """(uint8 DEVICE clip [T,S,S,3] (or [num_samples,T,S,S,3]), label, path)"""
        import torch
        clips, label, path = self.load_frames(idx)
        n, T, H, W, _ = clips.shape
        dev = torch.device("cuda", self.device)
        x = transform(torch.from_numpy(clips).reshape(n * T, H, W, 3).to(dev), self.im_scale, self.input_size,
                      frames_per_clip=T)
        x = x.reshape(n, T, self.input_size, self.input_size, 3)
        return (x[0] if self.num_samples == 1 else x), label, path

    # ---- batch sources for VideoLearnerAdversarial.fit --------------------------------------------------------
    def _host_batches(self, indices):
        """lists of `batch_size` decoded items (clips uint8 [num_samples,T,H,W,3], label, path) — full batches only, as
        the engine is planned for a fixed batch; decoding runs up to `prefetch` batches ahead on a background thread,
        which stops when the consumer abandons the iterator and whose exceptions are re-raised in the consumer"""
        B = self.batch_size
        groups = [indices[i:i + B] for i in range(0, len(indices) - B + 1, B)]
        q = queue.Queue(maxsize=max(1, self.prefetch))
        stop = threading.Event()

        def put(item):
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def work():
            from concurrent.futures import ThreadPoolExecutor
            try:
                with ThreadPoolExecutor(max_workers=self.num_workers) as pool:
                    for grp in groups:
                        if not put(list(pool.map(self.load_frames, grp))):      # the videos of a batch decode in parallel
                            return
                put(None)
            except BaseException as e:       # surfaced in the consumer
                put(e)

        th = threading.Thread(target=work, daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            th.join()

    def _batches(self, indices, with_paths=False):
        """yields (uint8 DEVICE clips [B,T,S,S,3], int64 DEVICE labels [B]): pinned upload of the decoded frames and the
        resize / crop kernel per clip (clips of one batch may differ in frame size)"""
        import torch
        dev = torch.device("cuda", self.device)
        B, S = self.batch_size, self.input_size
        for item in self._host_batches(indices):
            out = torch.empty((B, self.sample_length, S, S, 3), dtype=torch.uint8, device=dev)
            for b, (clips, _, _) in enumerate(item):
                fr = torch.from_numpy(clips[0]).pin_memory().to(dev, non_blocking=True)
                transform(fr, self.im_scale, self.input_size, frames_per_clip=self.sample_length, out=out[b])
            labels = torch.tensor([lab for _, lab, _ in item], dtype=torch.int64, device=dev)
            yield (out, labels, [p for _, _, p in item]) if with_paths else (out, labels)

    def _cached(self, key, make):
        if self.cache and key in self._cache:
            yield from self._cache[key]
            return
        items = []
        for item in make():
            if self.cache:
                items.append(item)
            yield item
        if self.cache:                      # reached only when the pass ran to completion
            self._cache[key] = items

    def train_batches(self):
        return self._cached("train", lambda: self._batches(self.train_range))

    def test_batches(self):
        return self._cached("test", lambda: self._batches(self.test_range))

    def cached_bytes(self):
        return sum(t.numel() * t.element_size() for items in self._cache.values() for item in items for t in item[:2])
