"""CPU: TFRecord clip input and TensorBoard scalar output — checksum known answers, the
protobuf encodings against hand-derived wire bytes, framing errors, and the reference's record schema
(kinetics_to_tf_record_uint8.py:90-94 -> utils/pre_process_rgb_flow.py:211-236) as a round trip."""
import struct

import numpy as np
import pytest

from flickering_adversarial_video_b200 import records as R


def test_crc32c_known_answers():
    # RFC 3720 appendix B.4 / the CRC catalogue's check value
    assert R.crc32c(b"123456789") == 0xE3069283
    assert R.crc32c(bytes(32)) == 0x8A9136AA
    assert R.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert R.crc32c(bytes(range(32))) == 0x46DD794E
    # running form and unaligned starts
    data = bytes(range(256)) * 5
    assert R.crc32c(data[100:], R.crc32c(data[:100])) == R.crc32c(data)
    assert R.crc32c(np.frombuffer(data, np.uint8)[3:77]) == R.crc32c(data[3:77])
    # TensorFlow's mask: rotate right 15, add 0xa282ead8
    c = 0xE3069283
    assert R.masked_crc32c(b"123456789") == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_example_wire_bytes():
    # features { feature { key: "train/label" value { int64_list { value: 7 } } } }, packed repeated int64 (proto3)
    want = bytes([0x0A, 0x16, 0x0A, 0x14, 0x0A, 0x0B]) + b"train/label" + bytes([0x12, 0x05, 0x1A, 0x03, 0x0A, 0x01, 0x07])
    assert R.encode_example({"train/label": 7}) == want
    assert R.decode_example(want) == {"train/label": [7]}
    # unpacked int64 (older writers) and negative values decode too
    unpacked = bytes([0x0A, 0x15, 0x0A, 0x13, 0x0A, 0x0B]) + b"train/label" + bytes([0x12, 0x04, 0x1A, 0x02, 0x08, 0x07])
    assert R.decode_example(unpacked) == {"train/label": [7]}
    assert R.decode_example(R.encode_example({"x": [-1, 2 ** 40]}))["x"] == [-1, 2 ** 40]
    ex = R.decode_example(R.encode_example({"f": [0.5, -2.0], "b": [b"ab", b""]}))
    assert ex["f"] == [0.5, -2.0] and [bytes(x) for x in ex["b"]] == [b"ab", b""]


def test_tfrecord_framing_and_corruption(tmp_path):
    path = str(tmp_path / "a.tfrecord")
    payloads = [b"", b"x", bytes(range(200)) * 50]
    with R.TFRecordWriter(path) as w:
        for p in payloads:
            w.write(p)
    raw = open(path, "rb").read()
    # u64 length | masked crc(length) | payload | masked crc(payload)
    assert raw[:8] == struct.pack("<Q", 0) and len(raw) == sum(16 + len(p) for p in payloads)
    assert struct.unpack("<I", raw[8:12])[0] == R.masked_crc32c(struct.pack("<Q", 0))
    assert [bytes(p) for p in R.tfrecord_iterator(path)] == payloads
    bad = bytearray(raw)
    bad[16 + 12 + 0] ^= 1                                  # payload byte of record 1
    open(path, "wb").write(bad)
    with pytest.raises(IOError, match="record 1"):
        list(R.tfrecord_iterator(path))
    assert len(list(R.tfrecord_iterator(path, verify=False))) == 3      # like the reference, payload CRC optional
    open(path, "wb").write(raw[:-3])                       # truncated
    with pytest.raises(IOError):
        list(R.tfrecord_iterator(path))
    open(path, "wb").write(b"")
    assert list(R.tfrecord_iterator(path)) == []


def test_clip_records_round_trip_and_batching(tmp_path):
    rng = np.random.RandomState(0)
    paths, clips, labels = [], [], []
    for f in range(2):
        p = str(tmp_path / f"kinetics_{f}.tfrecords")
        with R.TFRecordWriter(p) as w:
            for i in range(3):
                v = rng.randint(0, 256, size=(6, 224, 224, 3), dtype=np.uint8)
                R.write_clip_record(w, v, 100 * f + i)
                clips.append(v)
                labels.append(100 * f + i)
        paths.append(p)
    ds = R.ClipRecordDataset(paths, batch_size=2, frames=4)
    got = list(ds)
    assert len(got) == 3                                   # 6 records, batch 2, drop_remainder
    for b, (v, l) in enumerate(got):
        assert v.dtype == np.uint8 and v.shape == (2, 4, 224, 224, 3) and l.dtype == np.int64
        for j in range(2):
            assert np.array_equal(v[j], clips[2 * b + j][-4:]) and l[j] == labels[2 * b + j]
    assert len(list(R.ClipRecordDataset(paths, batch_size=4, repeat=2, prefetch=0))) == 3   # repeat(2): 12 records / 4
    ds5 = R.ClipRecordDataset(paths, batch_size=5)
    assert len(list(ds5)) == 1                             # remainder dropped


def test_tensorboard_scalars(tmp_path):
    w = R.SummaryWriter(str(tmp_path))
    tags = {"Loss/total": 1.25, "Perturbation/thickness_%": 0.5, "Probability/prob_to_min": 0.03125}
    for step in (0, 50, 100):
        w.add_scalars({k: v + step for k, v in tags.items()}, step)
    w.close()
    got = R.read_scalars(w.path)
    assert len(got) == 9
    assert got[0] == (0, "Loss/total", 1.25) and got[-1] == (100, "Probability/prob_to_min", 100.03125)
    # first record: file_version "brain.Event:2"
    first = bytes(next(iter(R.tfrecord_iterator(w.path))))
    assert first[0] == 0x09 and first[9:] == bytes([0x1A, 13]) + b"brain.Event:2"
    # one scalar event, hand-derived: step = 50, Summary{Value{tag "a", simple_value 1.0}}
    w2 = R.SummaryWriter(str(tmp_path / "x"))
    w2.add_scalar("a", 1.0, 50, wall_time=0.0)
    w2.close()
    ev = [bytes(p) for p in R.tfrecord_iterator(w2.path)][1]
    assert ev == bytes([0x09]) + bytes(8) + bytes([0x10, 50, 0x2A, 0x0A, 0x0A, 0x08, 0x0A, 0x01]) + b"a" + \
        bytes([0x15]) + struct.pack("<f", 1.0)


def test_videos_to_tfrecords(tmp_path):
    """kinetics_to_tf_record_uint8.py: last n frames of every long-enough video, 100 (here 2) videos per file"""
    cv2 = pytest.importorskip("cv2")
    src = tmp_path / "val"
    lengths = {"hula hooping": [12, 5, 9, 8], "yoga": [8]}
    for cls, ns in lengths.items():
        (src / cls).mkdir(parents=True)
        for i, n in enumerate(ns):
            wr = cv2.VideoWriter(str(src / cls / f"v{i}.mp4"), cv2.VideoWriter_fourcc(*"mp4v"), 25.0, (32, 24))
            if not wr.isOpened():
                pytest.skip("OpenCV cannot encode mp4v here")
            for t in range(n):
                wr.write(np.full((24, 32, 3), 10 * t + 5, np.uint8))
            wr.release()
    (src / "yoga" / "broken.mp4").write_bytes(b"not a video")
    classes = ["abseiling", "hula hooping", "yoga"]
    files = R.videos_to_tfrecords(str(src), "all", str(tmp_path / "rec"), classes, n_frames=8, videos_per_file=2)
    names = [f[len(str(tmp_path / "rec")) + 1:] for f in files]
    assert names == ["hula hooping/kinetics_hula hooping_0000.tfrecords", "hula hooping/kinetics_hula hooping_0001.tfrecords",
                     "yoga/kinetics_yoga_0000.tfrecords"]
    got = [R.parse_clip_example(p, height=24, width=32) for f in files for p in R.tfrecord_iterator(f)]
    assert [lab for _, lab in got] == [1, 1, 1, 2]                       # the 5-frame video and the broken file are skipped
    ids = [[int(round((float(fr.mean()) - 5) / 10)) for fr in v] for v, _ in got]
    assert ids == [list(range(4, 12)), list(range(1, 9)), list(range(0, 8)), list(range(0, 8))]      # the LAST 8 frames
    assert R.videos_to_tfrecords(str(src), "missing class", str(tmp_path / "rec2"), classes + ["missing class"]) == []


def test_prefetch_thread_stops_with_the_consumer(tmp_path):
    import threading
    p = str(tmp_path / "k.tfrecords")
    with R.TFRecordWriter(p) as w:
        for i in range(6):
            R.write_clip_record(w, np.full((2, 224, 224, 3), i, np.uint8), i)
    n0 = threading.active_count()
    it = iter(R.ClipRecordDataset([p], batch_size=1, prefetch=1))
    v, l = next(it)
    assert int(l[0]) == 0 and v.shape == (1, 2, 224, 224, 3)
    it.close()
    assert threading.active_count() <= n0
    assert [int(l[0]) for _, l in R.ClipRecordDataset([p], batch_size=1, prefetch=2)] == list(range(6))
    bad = str(tmp_path / "bad.tfrecords")
    open(bad, "wb").write(open(p, "rb").read()[:-7])
    with pytest.raises(IOError):
        list(R.ClipRecordDataset([bad], batch_size=1, prefetch=2))


def test_interleaved_record_order(tmp_path):
    """TFRecordDataset(files, num_parallel_reads=N): deterministic interleave, cycle N, block 1"""
    sizes = [3, 1, 4, 2]
    paths = []
    for f, n in enumerate(sizes):
        p = str(tmp_path / f"f{f}.tfrecords")
        with R.TFRecordWriter(p) as w:
            for i in range(n):
                R.write_clip_record(w, np.zeros((1, 224, 224, 3), np.uint8), 10 * f + i)
        paths.append(p)

    def order(n):
        return [int(l[0]) for _, l in R.ClipRecordDataset(paths, batch_size=1, prefetch=0, num_parallel_reads=n)]

    assert order(None) == order(1) == [0, 1, 2, 10, 20, 21, 22, 23, 30, 31]
    # two slots (InterleaveDataset's loop by hand): f0 / f1 alternate; f1 is found exhausted on its second turn, which
    # produces nothing and frees the slot; f2 is opened when the cycle returns to it; the same happens to f0 -> f3
    assert order(2) == [0, 10, 1, 2, 20, 21, 30, 22, 31, 23]
    # more slots than files: plain round robin until the short files run out
    assert order(8) == [0, 10, 20, 30, 1, 21, 31, 2, 22, 23]
    assert sorted(order(3)) == sorted(order(None))


def test_record_batches_from_config(tmp_path):
    """the mains' record lists (single_class_gen.py:111-120): sorted globs per path, NUM_OF_* cut, BATCH_SIZE batches"""
    from flickering_adversarial_video_b200.config import AttrDict
    from flickering_adversarial_video_b200.drivers import record_batches_from_config
    for split, per_file in (("train_a", [2, 2]), ("train_b", [2]), ("val", [3, 2])):
        (tmp_path / split).mkdir()
        for f, n in enumerate(per_file):
            with R.TFRecordWriter(str(tmp_path / split / f"kinetics_x_{f:04}.tfrecords")) as w:
                for i in range(n):
                    R.write_clip_record(w, np.zeros((3, 224, 224, 3), np.uint8), {"train_a": 0, "train_b": 50, "val": 90}[split] + 10 * f + i)
    cfg = AttrDict(TF_RECORDS_TRAIN_PATH=[str(tmp_path / "train_a"), str(tmp_path / "train_b")],
                   TF_RECORDS_VAL_PATH=[str(tmp_path / "val")], NUM_OF_TRAIN_TF_RECORDS=2, NUM_OF_VAL_TF_RECORDS=5,
                   BATCH_SIZE=2)
    train, val = record_batches_from_config(cfg, frames=2)
    assert [l.tolist() for _, l in train()] == [[0, 1], [10, 11]]            # train_b is cut off by NUM_OF_TRAIN_TF_RECORDS
    assert [l.tolist() for _, l in train()] == [[0, 1], [10, 11]]            # a fresh iterator per call
    got = list(val())
    assert [l.tolist() for _, l in got] == [[90, 91], [92, 100]] and got[0][0].shape == (2, 2, 224, 224, 3)     # remainder dropped
    cfg.NUM_OF_TRAIN_TF_RECORDS = 3
    r1, v1 = record_batches_from_config(cfg, rank=1, world=2)
    assert [l.tolist() for _, l in r1()] == [[10, 11]]                       # rank 1 of 2 takes every second file
    assert [l.tolist() for _, l in v1()] == [[100, 101]]
    cfg.TF_RECORDS_VAL_PATH = [str(tmp_path / "nothing")]
    with pytest.raises(FileNotFoundError):
        record_batches_from_config(cfg)
