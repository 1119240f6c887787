"""CPU: TF1 checkpoint (tensor bundle) ingest, SURVEY §8 row f2 — format known answers and a round trip through the
writer with the I3D variable names the reference's Saver restores (utils/kinetics_i3d_utils.py:41-62).
PARITY UNPINNED against a TensorFlow-written file (none available offline): see ckpt.py."""
import struct

import numpy as np
import pytest

from flickering_adversarial_video_b200 import ckpt, records


def _weights():
    rng = np.random.RandomState(0)
    w = {}
    for unit, shape in (("Conv3d_1a_7x7", (7, 7, 7, 3, 64)), ("Conv3d_2b_1x1", (1, 1, 1, 64, 64)),
                        ("Mixed_3b/Branch_1/Conv3d_0b_3x3", (3, 3, 3, 96, 128))):
        base = f"RGB/inception_i3d/{unit}"
        w[f"{base}/conv_3d/w"] = rng.randn(*shape).astype(np.float32)
        for v in ("beta", "moving_mean", "moving_variance"):
            w[f"{base}/batch_norm/{v}"] = rng.rand(1, 1, 1, 1, shape[-1]).astype(np.float32)
    w["RGB/inception_i3d/Logits/Conv3d_0c_1x1/conv_3d/w"] = rng.randn(1, 1, 1, 1024, 400).astype(np.float32)
    w["RGB/inception_i3d/Logits/Conv3d_0c_1x1/conv_3d/b"] = rng.randn(400).astype(np.float32)
    w["global_step"] = np.array(1234, dtype=np.int64)
    return w


@pytest.mark.parametrize("block_size", [4096, 64])
def test_round_trip_and_structure(tmp_path, block_size):
    prefix = str(tmp_path / "model.ckpt")
    w = _weights()
    ckpt.write_tf_checkpoint(prefix, w, block_size=block_size)
    raw = open(prefix + ".index", "rb").read()
    # LevelDB table footer: 40 bytes of handles + padding, then the magic number
    assert struct.unpack("<Q", raw[-8:])[0] == 0xdb4775248b80fb57 and len(raw) >= 48
    # the first data block starts with the header entry: key "" (shared 0, non_shared 0)
    assert raw[0] == 0 and raw[1] == 0
    entries, shards = ckpt.list_variables(prefix)
    assert shards == 1 and set(entries) == set(w)
    assert entries["RGB/inception_i3d/Conv3d_1a_7x7/conv_3d/w"]["shape"] == (7, 7, 7, 3, 64)
    assert entries["global_step"]["dtype"] == 9 and entries["global_step"]["shape"] == ()
    got = ckpt.read_tf_checkpoint(prefix)
    for k, v in w.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    only = ckpt.read_tf_checkpoint(prefix, names=lambda n: n.endswith("/w"))
    assert set(only) == {k for k in w if k.endswith("/w")}


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "model.ckpt")
    ckpt.write_tf_checkpoint(prefix, _weights())
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[1000] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(data)
    with pytest.raises(IOError, match="CRC"):
        ckpt.read_tf_checkpoint(prefix)
    assert ckpt.read_tf_checkpoint(prefix, verify=False)            # readable when asked not to check
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[5] ^= 1
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(IOError, match="CRC"):
        ckpt.list_variables(prefix)
    open(prefix + ".index", "wb").write(idx[:-1])
    with pytest.raises(IOError, match="magic"):
        ckpt.list_variables(prefix)


def test_crc_mask_inverse():
    c = records.crc32c(b"tensor bytes")
    assert ckpt._unmask(records.masked_crc32c(b"tensor bytes")) == c


def test_kinetics_loader_accepts_checkpoint_prefix(tmp_path):
    from flickering_adversarial_video_b200.kinetics_i3d import load_weights
    prefix = str(tmp_path / "rgb_imagenet" / "model.ckpt")
    w = _weights()
    ckpt.write_tf_checkpoint(prefix, w)
    got = load_weights(prefix)
    assert "global_step" not in got and len(got) == len(w) - 1
    assert np.array_equal(got["RGB/inception_i3d/Conv3d_2b_1x1/conv_3d/w"], w["RGB/inception_i3d/Conv3d_2b_1x1/conv_3d/w"])
