"""TFRecord input and TensorBoard scalar output without TensorFlow.

Reference call sites:
  * writer of the clip records: kinetics_to_tf_record_uint8.py:62-95 — one `tf.train.Example` per video with
    `train/label` (int64) and `train/video` (raw bytes of the last 90 uint8 frames, [90,224,224,3]);
  * reader: `tf.data.TFRecordDataset(...).batch(B, drop_remainder=True).map(parse_example_uint8)`
    (i3d_adversarial_main_universal.py:231-248, utils/pre_process_rgb_flow.py:211-236).  The reference casts to
    float and computes u8/128 - 1 on the host; the engine's apply kernel does that on the GPU, so clips stay uint8;
  * scalars: `tf.summary.scalar(tag, ...)` + `SummarySaverHook(save_steps=50)` (universal.py:176-201).

File formats (tensorflow/core/lib/io/record_writer.cc, tensorflow/core/util/event.proto, example.proto, feature.proto):
a record is  u64 length | u32 masked_crc32c(length) | payload | u32 masked_crc32c(payload); payloads are protobuf
messages, hand-encoded here (four message types, wire types 0 / 1 / 2 / 5 only).  CRC-32C and the record index scan
are native (libfavio.so, include/favio.h)."""
import ctypes as C
import os
import socket
import struct
import threading
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_io = None


def _lib():
    global _io
    if _io is None:
        path = os.path.join(_HERE, "libfavio.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(path)
        lib.favio_crc32c.restype = C.c_uint32
        lib.favio_crc32c.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
        lib.favio_masked_crc32c.restype = C.c_uint32
        lib.favio_masked_crc32c.argtypes = [C.c_void_p, C.c_size_t]
        lib.favio_tfrecord_index.restype = C.c_int64
        lib.favio_tfrecord_index.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        _io = lib
    return _io


def _addr(buf):
    """(address, nbytes, keep-alive) of bytes / bytearray / memoryview / contiguous ndarray without copying"""
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf.reshape(-1).view(np.uint8)
    return a.ctypes.data, a.nbytes, a


def crc32c(data, crc=0):
    p, n, keep = _addr(data)
    return int(_lib().favio_crc32c(crc, p, n))


def masked_crc32c(data):
    p, n, keep = _addr(data)
    return int(_lib().favio_masked_crc32c(p, n))


# ---- protobuf wire format (the subset these messages use) ---------------------------------------
def _varint(v):
    v &= (1 << 64) - 1            # negative int64 -> 10-byte two's complement, as protobuf does
    out = bytearray()
    while True:
        b = v & 0x7f
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7f) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _ld(field, payload):          # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + bytes(payload)


def _fields(buf):
    """Iterate (field number, wire type, value) over one message; value is an int (varint / fixed) or a memoryview."""
    buf = memoryview(buf)
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8]); pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4]); pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, v


# ---- tf.train.Example --------------------------------------------------------------------------
def encode_example(features):
    """{name: int | [int] | bytes | [bytes] | float | [float]} -> serialized tf.train.Example
    (Example.features = 1; Features.feature = 1 (map entry: key = 1, value = 2);
     Feature.bytes_list = 1 / float_list = 2 / int64_list = 3, each with repeated `value = 1`)."""
    entries = b""
    for name in sorted(features):                      # protobuf map order is unspecified; sorted = deterministic
        v = features[name]
        vals = v if isinstance(v, (list, tuple)) else [v]
        if all(isinstance(x, (bytes, bytearray, memoryview)) for x in vals):
            feat = _ld(1, b"".join(_ld(1, x) for x in vals))
        elif all(isinstance(x, (int, np.integer)) for x in vals):
            feat = _ld(3, _ld(1, b"".join(_varint(int(x)) for x in vals)))        # packed, like TF writes it
        else:
            feat = _ld(2, _ld(1, struct.pack(f"<{len(vals)}f", *[float(x) for x in vals])))
        entries += _ld(1, _ld(1, name.encode()) + _ld(2, feat))
    return _ld(1, entries)


def decode_example(buf):
    """serialized tf.train.Example -> {name: list of ints / floats / memoryviews (zero-copy bytes)}"""
    out = {}
    for f, wt, feats in _fields(buf):
        if f != 1:
            continue
        for f2, wt2, entry in _fields(feats):
            if f2 != 1:
                continue
            name, feat = None, None
            for f3, wt3, v in _fields(entry):
                if f3 == 1:
                    name = bytes(v).decode()
                elif f3 == 2:
                    feat = v
            vals = []
            for kind, wtk, lst in _fields(feat if feat is not None else b""):
                for f4, wt4, v in _fields(lst):
                    if f4 != 1:
                        continue
                    if kind == 1:
                        vals.append(v)
                    elif kind == 3:
                        if wt4 == 2:                       # packed
                            p = 0
                            while p < len(v):
                                x, p = _read_varint(v, p)
                                vals.append(x - (1 << 64) if x >= 1 << 63 else x)
                        else:
                            vals.append(v - (1 << 64) if v >= 1 << 63 else v)
                    elif kind == 2:
                        if wt4 == 2:
                            vals.extend(struct.unpack(f"<{len(v) // 4}f", bytes(v)))
                        else:
                            vals.append(struct.unpack("<f", v)[0])
            out[name] = vals
    return out


# ---- TFRecord files ----------------------------------------------------------------------------
class TFRecordWriter:
    """tf.python_io.TFRecordWriter(path).write(serialized) (kinetics_to_tf_record_uint8.py:62,94)"""

    def __init__(self, path):
        self._f = open(path, "wb")

    def write(self, payload):
        hdr = struct.pack("<Q", len(payload))
        self._f.write(hdr + struct.pack("<I", masked_crc32c(hdr)))
        self._f.write(payload)
        self._f.write(struct.pack("<I", masked_crc32c(payload)))

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def tfrecord_iterator(path, verify=True):
    """Yield the payload of every record as a zero-copy memoryview into the memory-mapped file.  Raises on a wrong
    length CRC, a truncated file and (verify=True) a wrong payload CRC, like tf.data.TFRecordDataset."""
    if os.path.getsize(path) == 0:
        return
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    lib = _lib()
    n = int(lib.favio_tfrecord_index(mm.ctypes.data, mm.nbytes, 0, None, None, 0))      # count (length CRCs only)
    if n < 0:
        raise IOError(f"{path}: corrupt TFRecord framing at record {-n - 1}")
    off = np.zeros(n, dtype=np.uint64)
    ln = np.zeros(n, dtype=np.uint64)
    r = int(lib.favio_tfrecord_index(mm.ctypes.data, mm.nbytes, int(bool(verify)), off.ctypes.data, ln.ctypes.data, n))
    if r < 0:
        raise IOError(f"{path}: payload CRC mismatch at record {-r - 1}")
    view = memoryview(mm)
    for o, l in zip(off.tolist(), ln.tolist()):
        yield view[o:o + l]


def write_clip_record(writer, frames_u8, label):
    """One record of the reference's conversion script (kinetics_to_tf_record_uint8.py:90-94)."""
    frames_u8 = np.ascontiguousarray(frames_u8, dtype=np.uint8)
    writer.write(encode_example({"train/label": int(label), "train/video": frames_u8.tobytes()}))


def parse_clip_example(payload, height=224, width=224):
    """parse_example_uint8 (utils/pre_process_rgb_flow.py:211-236) for one record: (uint8 [T,H,W,3] view, label).
    The /128 - 1 normalisation of the reference happens in the engine's apply kernel."""
    ex = decode_example(payload)
    video = np.frombuffer(ex["train/video"][0], dtype=np.uint8).reshape(-1, height, width, 3)
    return video, int(ex["train/label"][0])


class ClipRecordDataset:
    """`TFRecordDataset(files).repeat(r).batch(B, drop_remainder=True).map(parse_example_uint8).prefetch(...)`
    (i3d_adversarial_main_universal.py:239-245) yielding (uint8 [B,T,224,224,3], int64 [B]) host arrays; with
    pinned=True they are page-locked torch tensors ready for `FlickerAttack.prefetch`, and a background thread keeps
    `prefetch` batches decoded ahead of the consumer."""

    def __init__(self, filenames, batch_size, frames=None, repeat=1, verify=True, pinned=False, prefetch=2,
                 num_parallel_reads=None):
        """num_parallel_reads=N reproduces the record ORDER of `tf.data.TFRecordDataset(files, num_parallel_reads=N)`
        (the reference passes os.cpu_count(), i3d_adversarial_main_universal.py:239): a deterministic interleave with
        cycle length N and block length 1 — one record from each of N open files in turn; an exhausted file frees its
        slot, which takes the next unopened file when the cycle comes back to it [dep: tf.data InterleaveDataset].
        None reads the files one after the other."""
        self.filenames = [filenames] if isinstance(filenames, str) else list(filenames)
        self.batch_size, self.frames, self.repeat = int(batch_size), frames, int(repeat)
        self.verify, self.pinned, self.prefetch = verify, pinned, int(prefetch)
        self.num_parallel_reads = None if not num_parallel_reads else max(1, int(num_parallel_reads))

    def num_records(self):
        """records of one pass over the files (framing scan only: lengths and length CRCs, no payload is touched)"""
        lib = _lib()
        n = 0
        for path in self.filenames:
            if os.path.getsize(path) == 0:
                continue
            mm = np.memmap(path, dtype=np.uint8, mode="r")
            c = int(lib.favio_tfrecord_index(mm.ctypes.data, mm.nbytes, 0, None, None, 0))
            if c < 0:
                raise IOError(f"{path}: corrupt TFRecord framing")
            n += c
        return n

    def num_batches(self):
        """batches this dataset yields: repeat() precedes batch(drop_remainder=True) in the reference's pipeline"""
        return (self.num_records() * self.repeat) // self.batch_size

    def _payloads(self):
        if self.num_parallel_reads is None or self.num_parallel_reads == 1:
            for path in self.filenames:
                yield from tfrecord_iterator(path, verify=self.verify)
            return
        pending = iter(self.filenames)
        slots = [None] * self.num_parallel_reads
        end_of_input, n_open, i = False, 0, 0
        while not end_of_input or n_open > 0:
            if slots[i] is not None:
                try:
                    yield next(slots[i])
                    i = (i + 1) % len(slots)
                except StopIteration:
                    slots[i], n_open = None, n_open - 1
                    i = (i + 1) % len(slots)
            elif not end_of_input:
                path = next(pending, None)
                if path is None:
                    end_of_input = True
                else:
                    slots[i], n_open = iter(tfrecord_iterator(path, verify=self.verify)), n_open + 1
            else:
                i = (i + 1) % len(slots)

    def _records(self):
        for _ in range(self.repeat):
            for payload in self._payloads():
                yield parse_clip_example(payload)

    def _batches(self):
        vids, labs = [], []
        for v, l in self._records():
            if self.frames is not None:
                v = v[-self.frames:]                  # the drivers attack the last frames, like the conversion script keeps them
            vids.append(v)
            labs.append(l)
            if len(vids) == self.batch_size:
                yield self._pack(vids, labs)
                vids, labs = [], []
        # drop_remainder=True

    def _pack(self, vids, labs):
        if self.pinned:
            import torch
            out = torch.empty((len(vids),) + vids[0].shape, dtype=torch.uint8).pin_memory()
            for i, v in enumerate(vids):
                out[i].copy_(torch.from_numpy(np.array(v, copy=False)))
            return out, torch.tensor(labs, dtype=torch.int64).pin_memory()
        return np.stack(vids), np.asarray(labs, dtype=np.int64)

    def __iter__(self):
        if self.prefetch <= 0:
            yield from self._batches()
            return
        import queue
        q = queue.Queue(maxsize=self.prefetch)
        done, stop = object(), threading.Event()

        def put(item):                     # gives up when the consumer abandoned the iterator
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def work():
            try:
                for b in self._batches():
                    if not put(b):
                        return
                put(done)
            except BaseException as e:     # surface reader errors in the consumer
                put(e)

        th = threading.Thread(target=work, daemon=True)
        th.start()
        try:
            while True:
                b = q.get()
                if b is done:
                    return
                if isinstance(b, BaseException):
                    raise b
                yield b
        finally:
            stop.set()
            th.join()


def read_video_frames(path):
    """all frames of a video file as uint8 RGB [N,H,W,3] (`skvideo.io.vread` in the reference,
    kinetics_to_tf_record_uint8.py:79; here OpenCV's FFmpeg backend)"""
    import cv2
    cap = cv2.VideoCapture(path)
    if not cap.isOpened():
        raise IOError(f"cannot open video {path}")
    frames = []
    try:
        while True:
            ok, bgr = cap.read()
            if not ok:
                break
            frames.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
    finally:
        cap.release()
    if not frames:
        raise IOError(f"no frame decoded from {path}")
    return np.stack(frames)


def videos_to_tfrecords(videos_base_path, class_name, tf_dst_folder, kinetics_classes, n_frames=90, videos_per_file=100,
                        video_ext="mp4"):
    """The reference's conversion script (kinetics_to_tf_record_uint8.py:21-98): for every class folder (or all of
    them when class_name == 'all') the LAST `n_frames` frames of each video, undecimated and unresized (the Kinetics
    crawler already stores 224x224 files), one `tf.train.Example` {train/label, train/video} per video, 100 videos per
    `<dst>/<class>/kinetics_<class>_<k:04>.tfrecords`.  Videos shorter than n_frames are skipped (:84-85).  Unlike the
    reference an undecodable file is skipped, not deleted (:82).  Returns the list of files written."""
    import glob
    classes = sorted(os.listdir(videos_base_path)) if class_name == "all" else [class_name]
    written = []
    for c in classes:
        src = os.path.join(videos_base_path, c)
        if not os.path.isdir(src):
            print("{} not exist".format(src))
            continue
        cls_id = list(kinetics_classes).index(c)
        writer, k, in_file = None, 0, 0
        for v in sorted(glob.glob(os.path.join(src, "*." + video_ext))):
            try:
                frames = read_video_frames(v)
            except IOError:
                continue
            if frames.shape[0] < n_frames:
                continue
            if writer is None or in_file == videos_per_file:
                if writer is not None:
                    writer.close()
                    k += 1
                os.makedirs(os.path.join(tf_dst_folder, c), exist_ok=True)
                path = os.path.join(tf_dst_folder, c, "kinetics_{}_{:04}.tfrecords".format(c, k))
                writer, in_file = TFRecordWriter(path), 0
                written.append(path)
            write_clip_record(writer, frames[-n_frames:], cls_id)
            in_file += 1
        if writer is not None:
            writer.close()
    return written


# ---- TensorBoard scalars -----------------------------------------------------------------------
class SummaryWriter:
    """events.out.tfevents.* with scalar summaries: Event{wall_time = 1 (double), step = 2 (int64), file_version = 3,
    summary = 5}; Summary{value = 1}; Summary.Value{tag = 1, simple_value = 2 (float)} (event.proto, summary.proto).
    Tag names of the reference: i3d_adversarial_main_universal.py:176-196."""

    def __init__(self, logdir, filename_suffix=""):
        os.makedirs(logdir, exist_ok=True)
        name = f"events.out.tfevents.{int(time.time())}.{socket.gethostname()}{filename_suffix}"
        self.path = os.path.join(logdir, name)
        self._w = TFRecordWriter(self.path)
        self._w.write(struct.pack("<Bd", (1 << 3) | 1, time.time()) + _ld(3, b"brain.Event:2"))

    def add_scalar(self, tag, value, step, wall_time=None):
        val = _ld(1, tag.encode()) + struct.pack("<Bf", (2 << 3) | 5, float(value))
        ev = (struct.pack("<Bd", (1 << 3) | 1, time.time() if wall_time is None else wall_time) +
              _varint((2 << 3) | 0) + _varint(int(step)) + _ld(5, _ld(1, val)))
        self._w.write(ev)

    def add_scalars(self, tagged, step):
        for tag, value in tagged.items():
            self.add_scalar(tag, value, step)

    def flush(self):
        self._w._f.flush()

    def close(self):
        self._w.close()


def read_scalars(path):
    """[(step, tag, value)] of an events file (tests; what TensorBoard's event accumulator extracts)."""
    out = []
    for payload in tfrecord_iterator(path):
        step, summary = 0, None
        for f, wt, v in _fields(payload):
            if f == 2:
                step = v
            elif f == 5:
                summary = v
        if summary is None:
            continue
        for f, wt, val in _fields(summary):
            if f != 1:
                continue
            tag, x = None, None
            for f2, wt2, v in _fields(val):
                if f2 == 1:
                    tag = bytes(v).decode()
                elif f2 == 2:
                    x = struct.unpack("<f", v)[0]
            out.append((step, tag, x))
    return out
