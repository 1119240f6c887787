"""CPU: the TensorFlow-free record / event / Example code (records.py, rows f1 / f4) against INDEPENDENT implementations
shipped in this image — TensorBoard's own record writer and reader (tensorboard.summary.writer.record_writer,
tensorboard.compat.tensorflow_stub.pywrap_tensorflow: the TF team's TFRecord framing and masked CRC-32C), its Event /
Summary protobuf classes, and Google's protobuf runtime for tf.train.Example (message types declared here from
tensorflow/core/example/{example,feature}.proto).  Skipped where TensorBoard / protobuf are not installed."""
import os
import struct

import numpy as np
import pytest

from flickering_adversarial_video_b200 import records as R

pytest.importorskip("tensorboard")
pytest.importorskip("google.protobuf")


def _tb_read(path):
    from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import PyRecordReader_New
    from tensorboard.compat.tensorflow_stub import errors
    rd = PyRecordReader_New(path)
    out = []
    while True:
        try:
            rd.GetNext()
        except errors.OutOfRangeError:
            return out
        out.append(bytes(rd.record()))


def test_our_tfrecords_are_read_by_tensorboards_reader(tmp_path):
    path = str(tmp_path / "ours.tfrecord")
    payloads = [b"", b"flicker", bytes(range(256)) * 40, os.urandom(4097)]
    with R.TFRecordWriter(path) as w:
        for p in payloads:
            w.write(p)
    assert _tb_read(path) == payloads          # its reader verifies both CRCs and raises DataLossError otherwise


def test_tensorboards_records_are_read_by_ours(tmp_path):
    from tensorboard.summary.writer.record_writer import RecordWriter
    path = str(tmp_path / "theirs.tfrecord")
    payloads = [b"x" * n for n in (0, 1, 15, 16, 17, 70000)]
    with open(path, "wb") as f:
        w = RecordWriter(f)
        for p in payloads:
            w.write(p)
        w.flush()
    assert [bytes(p) for p in R.tfrecord_iterator(path, verify=True)] == payloads
    # byte-identical framing, too
    ours = str(tmp_path / "ours.tfrecord")
    with R.TFRecordWriter(ours) as w:
        for p in payloads:
            w.write(p)
    assert open(ours, "rb").read() == open(path, "rb").read()


def test_masked_crc_matches_tensorboards():
    from tensorboard.summary.writer.record_writer import masked_crc32c
    rng = np.random.RandomState(1)
    for n in (0, 1, 7, 8, 9, 63, 64, 65, 4096, 100003):
        data = rng.randint(0, 256, n).astype(np.uint8).tobytes()
        assert R.masked_crc32c(data) == masked_crc32c(data), n


def test_our_event_files_load_in_tensorboard(tmp_path):
    from tensorboard.backend.event_processing.event_file_loader import LegacyEventFileLoader
    w = R.SummaryWriter(str(tmp_path))
    # tag names of the reference's summaries (i3d_adversarial_main_universal.py:176-201)
    vals = {"Loss/total": 0.75, "Loss/adversarial": 0.5, "Perturbation/thickness_%": 1.5, "Perturbation/roughness_%": 0.25}
    for step in (0, 50):
        w.add_scalars({k: v * (1 + step) for k, v in vals.items()}, step)
    w.close()
    events = list(LegacyEventFileLoader(w.path).Load())
    assert events[0].file_version == "brain.Event:2"
    got = [(e.step, v.tag, v.simple_value) for e in events[1:] for v in e.summary.value]
    assert got == [(s, k, np.float32(v * (1 + s))) for s in (0, 50) for k, v in vals.items()]
    assert all(e.wall_time > 1e9 for e in events)
    assert got == [(s, t, np.float32(x)) for s, t, x in R.read_scalars(w.path)]


def test_tensorboards_event_files_are_read_by_ours(tmp_path):
    from tensorboard.compat.proto import event_pb2, summary_pb2
    from tensorboard.summary.writer.event_file_writer import EventFileWriter
    w = EventFileWriter(str(tmp_path))
    for step, val in ((3, 0.125), (2 ** 33, -7.5)):
        s = summary_pb2.Summary(value=[summary_pb2.Summary.Value(tag="Loss/total", simple_value=val),
                                       summary_pb2.Summary.Value(tag="Probability/prob_to_min", simple_value=val / 2)])
        w.add_event(event_pb2.Event(wall_time=12.5, step=step, summary=s))
    w.close()
    path = [os.path.join(str(tmp_path), f) for f in os.listdir(str(tmp_path)) if "tfevents" in f][0]
    assert R.read_scalars(path) == [(3, "Loss/total", 0.125), (3, "Probability/prob_to_min", 0.0625),
                                    (2 ** 33, "Loss/total", -7.5), (2 ** 33, "Probability/prob_to_min", -3.75)]


def _example_classes():
    """tf.train.Example and its parts (tensorflow/core/example/example.proto, feature.proto), declared at run time"""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    f = descriptor_pb2.FileDescriptorProto(name="fav_test_example.proto", package="favtest", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name, *fields):
        m = f.message_type.add(name=name)
        for fname, num, ftype, label, tname, extra in fields:
            fd = m.field.add(name=fname, number=num, type=ftype, label=label)
            if tname:
                fd.type_name = tname
            if extra == "packed":
                fd.options.packed = True
            if extra == "oneof":
                fd.oneof_index = 0
        return m

    msg("BytesList", ("value", 1, T.TYPE_BYTES, T.LABEL_REPEATED, None, None))
    msg("FloatList", ("value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED, None, "packed"))
    msg("Int64List", ("value", 1, T.TYPE_INT64, T.LABEL_REPEATED, None, "packed"))
    feat = msg("Feature", ("bytes_list", 1, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".favtest.BytesList", "oneof"),
               ("float_list", 2, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".favtest.FloatList", "oneof"),
               ("int64_list", 3, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".favtest.Int64List", "oneof"))
    feat.oneof_decl.add(name="kind")
    feats = msg("Features", ("feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, ".favtest.Features.FeatureEntry", None))
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.field.add(name="key", number=1, type=T.TYPE_STRING, label=T.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=T.TYPE_MESSAGE, label=T.LABEL_OPTIONAL, type_name=".favtest.Feature")
    entry.options.map_entry = True
    msg("Example", ("features", 1, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".favtest.Features", None))
    pool = descriptor_pool.DescriptorPool()
    pool.Add(f)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("favtest.Example"))


def test_example_encoding_against_protobuf_runtime():
    Example = _example_classes()
    video = np.random.RandomState(2).randint(0, 256, (3, 8, 8, 3), dtype=np.uint8)
    # the reference's schema (kinetics_to_tf_record_uint8.py:90-94): train/label int64, train/video raw bytes
    ours = R.encode_example({"train/label": 217, "train/video": video.tobytes()})
    ex = Example.FromString(ours)
    assert list(ex.features.feature["train/label"].int64_list.value) == [217]
    assert ex.features.feature["train/video"].bytes_list.value[0] == video.tobytes()
    # and the other direction, including negative / 64-bit ints and floats
    theirs = Example()
    theirs.features.feature["train/label"].int64_list.value.extend([5, -3, 2 ** 50])
    theirs.features.feature["train/video"].bytes_list.value.append(video.tobytes())
    theirs.features.feature["score"].float_list.value.extend([0.25, -1.5])
    got = R.decode_example(theirs.SerializeToString())
    assert got["train/label"] == [5, -3, 2 ** 50] and got["score"] == [0.25, -1.5]
    assert bytes(got["train/video"][0]) == video.tobytes()
    # a record file written by protobuf + TensorBoard's writer parses as a clip
    frames, label = R.parse_clip_example(theirs.SerializeToString(), height=8, width=8)
    assert label == 5 and np.array_equal(frames, video)


def _bundle_classes():
    """BundleHeaderProto / BundleEntryProto / TensorShapeProto (tensorflow/core/protobuf/tensor_bundle.proto,
    framework/tensor_shape.proto, framework/versions.proto), declared at run time for Google's protobuf runtime"""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    f = descriptor_pb2.FileDescriptorProto(name="fav_test_bundle.proto", package="favbundle", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto
    shape = f.message_type.add(name="TensorShapeProto")
    dim = shape.nested_type.add(name="Dim")
    dim.field.add(name="size", number=1, type=T.TYPE_INT64, label=T.LABEL_OPTIONAL)
    dim.field.add(name="name", number=2, type=T.TYPE_STRING, label=T.LABEL_OPTIONAL)
    shape.field.add(name="dim", number=2, type=T.TYPE_MESSAGE, label=T.LABEL_REPEATED,
                    type_name=".favbundle.TensorShapeProto.Dim")
    shape.field.add(name="unknown_rank", number=3, type=T.TYPE_BOOL, label=T.LABEL_OPTIONAL)
    ver = f.message_type.add(name="VersionDef")
    ver.field.add(name="producer", number=1, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    ver.field.add(name="min_consumer", number=2, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    hdr = f.message_type.add(name="BundleHeaderProto")
    hdr.field.add(name="num_shards", number=1, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    hdr.field.add(name="endianness", number=2, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    hdr.field.add(name="version", number=3, type=T.TYPE_MESSAGE, label=T.LABEL_OPTIONAL, type_name=".favbundle.VersionDef")
    ent = f.message_type.add(name="BundleEntryProto")
    ent.field.add(name="dtype", number=1, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    ent.field.add(name="shape", number=2, type=T.TYPE_MESSAGE, label=T.LABEL_OPTIONAL,
                  type_name=".favbundle.TensorShapeProto")
    ent.field.add(name="shard_id", number=3, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    ent.field.add(name="offset", number=4, type=T.TYPE_INT64, label=T.LABEL_OPTIONAL)
    ent.field.add(name="size", number=5, type=T.TYPE_INT64, label=T.LABEL_OPTIONAL)
    ent.field.add(name="crc32c", number=6, type=T.TYPE_FIXED32, label=T.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(f)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("favbundle." + n))
    return get("BundleHeaderProto"), get("BundleEntryProto")


def test_checkpoint_index_protos_against_protobuf_runtime(tmp_path):
    """Row f2: the BundleHeaderProto / BundleEntryProto values inside the index our writer produces parse with Google's
    protobuf runtime to the right dtype / shape / extent, their CRCs agree with TensorBoard's masked CRC-32C of the
    tensor bytes, and entries SERIALISED by the protobuf runtime are understood by the reader.  (The LevelDB table
    around them has no second implementation in this image: that part of f2 stays unpinned.)"""
    from tensorboard.summary.writer.record_writer import masked_crc32c
    from flickering_adversarial_video_b200 import ckpt
    Header, Entry = _bundle_classes()
    rng = np.random.RandomState(4)
    tensors = {"RGB/inception_i3d/Conv3d_1a_7x7/conv_3d/w": rng.randn(7, 7, 7, 3, 64).astype(np.float32),
               "RGB/inception_i3d/Conv3d_1a_7x7/batch_norm/beta": rng.randn(1, 1, 1, 1, 64).astype(np.float32),
               "RGB/inception_i3d/Logits/Conv3d_0c_1x1/conv_3d/b": rng.randn(400).astype(np.float32),
               "global_step": np.array(12345, dtype=np.int64)}
    prefix = str(tmp_path / "model.ckpt")
    ckpt.write_tf_checkpoint(prefix, tensors)
    index = memoryview(open(prefix + ".index", "rb").read())
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    footer = index[-48:]
    _, _, p = ckpt._handle(footer, 0)
    ioff, isize, _ = ckpt._handle(footer, p)
    seen = {}
    for _, hv in ckpt._block_entries(ckpt._read_block(index, ioff, isize)):
        doff, dsize, _ = ckpt._handle(hv, 0)
        for key, val in ckpt._block_entries(ckpt._read_block(index, doff, dsize)):
            if key == b"":
                h = Header.FromString(bytes(val))
                assert h.num_shards == 1 and h.endianness == 0
                continue
            e = Entry.FromString(bytes(val))
            t = tensors[key.decode()]
            assert e.dtype == (1 if t.dtype == np.float32 else 9)            # DT_FLOAT / DT_INT64
            assert tuple(d.size for d in e.shape.dim) == t.shape and e.shard_id == 0 and e.size == t.nbytes
            raw = data[e.offset:e.offset + e.size]
            assert raw == t.tobytes() and e.crc32c == masked_crc32c(raw)
            seen[key.decode()] = e
    assert sorted(seen) == sorted(tensors)
    # the reader's entry parser on bytes serialised by the protobuf runtime
    e = Entry(dtype=1, shard_id=0, offset=2 ** 33, size=4 * 5 * 7, crc32c=0xDEADBEEF)
    e.shape.dim.add(size=5)
    e.shape.dim.add(size=7)
    blk = ckpt._build_block([(b"", Header(num_shards=1).SerializeToString()), (b"v", e.SerializeToString())])
    got = dict(ckpt._block_entries(blk))
    fields = {f: v for f, wt, v in ckpt._fields(got[b"v"])}
    assert fields[1] == 1 and fields[4] == 2 ** 33 and fields[5] == 140
    assert struct.unpack("<I", fields[6])[0] == 0xDEADBEEF and ckpt._shape_of(fields[2]) == (5, 7)


def test_example_round_trips_property_based():
    """random feature maps through both encoders / decoders (ours and the protobuf runtime's) — varint edges, negative
    and 64-bit ints, empty lists, non-ASCII keys, large byte strings"""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")
    Example = _example_classes()
    ints = st.lists(st.integers(min_value=-2 ** 63, max_value=2 ** 63 - 1), max_size=6)
    floats = st.lists(st.floats(width=32, allow_nan=False), max_size=6)
    blobs = st.lists(st.binary(max_size=300), max_size=3)
    feats = st.dictionaries(st.text(min_size=1, max_size=12), st.one_of(ints, floats, blobs), max_size=5)

    def kind(v):
        return "bytes" if v and isinstance(v[0], bytes) else "float" if v and isinstance(v[0], float) else "int"

    @hyp.settings(max_examples=150, deadline=None)
    @hyp.given(feats)
    def check(features):
        features = {k: v for k, v in features.items() if v}               # an empty list carries no type
        ours = R.encode_example(features)
        ex = Example.FromString(ours)                                        # protobuf parses what we wrote
        assert set(ex.features.feature) == set(features)
        theirs = Example()
        for k, v in features.items():
            f = ex.features.feature[k]
            got = {"bytes": list(f.bytes_list.value), "float": list(f.float_list.value), "int": list(f.int64_list.value)}[kind(v)]
            assert got == v, (k, v, got)
            dst = theirs.features.feature[k]
            {"bytes": dst.bytes_list, "float": dst.float_list, "int": dst.int64_list}[kind(v)].value.extend(v)
        back = R.decode_example(theirs.SerializeToString())                  # and we parse what protobuf wrote
        assert {k: [bytes(x) if kind(features[k]) == "bytes" else x for x in vals] for k, vals in back.items()} == features

    check()
