"""TF1 checkpoint ("tensor bundle") ingest without TensorFlow.

The reference restores the Kinetics I3D weights with `tf.train.Saver(var_list=rgb_variable_map).restore(sess,
'data/checkpoints/rgb_imagenet/model.ckpt')` (utils/kinetics_i3d_utils.py:41-62, run_config.yml:6-7); the variable
names are `RGB/inception_i3d/<unit>/conv_3d/w`, `.../batch_norm/{beta,moving_mean,moving_variance}` and
`RGB/inception_i3d/Logits/Conv3d_0c_1x1/conv_3d/{w,b}` — exactly the names `fav_load_weights` consumes.

Format (tensorflow/core/util/tensor_bundle/tensor_bundle.cc, tensorflow/core/lib/io/{table,block,format}.cc — the
LevelDB table format):
  <prefix>.index                 sorted string table: data blocks, a meta-index block, an index block and a 48-byte
                                 footer (two varint block handles, zero padding, magic 0xdb4775248b80fb57).  A block is
                                 prefix-compressed entries (shared, non_shared, value_len varints + key suffix + value),
                                 a restart array and its length, followed on disk by a type byte (0 = raw, 1 = snappy)
                                 and the masked CRC-32C of block + type.  Key "" holds BundleHeaderProto, every other
                                 key a BundleEntryProto {dtype = 1, shape = 2, shard_id = 3, offset = 4, size = 5,
                                 crc32c = 6 (fixed32, masked)}.
  <prefix>.data-NNNNN-of-MMMMM   the raw little-endian tensor bytes.

PARITY UNPINNED: no TensorFlow and no checkpoint file are available in the build container, so this reader is
checked against the published format (structure known answers: magic, footer layout, masked CRCs) and against the
writer below, not against a file TensorFlow wrote; the BundleHeaderProto / BundleEntryProto encodings and the tensor
CRCs are cross-checked against Google's protobuf runtime and TensorBoard's masked CRC-32C
(tests/test_cpu_records_crosscheck.py), the LevelDB table framing is not.  Snappy-compressed index blocks (not what BundleWriter emits) are
rejected with an error rather than guessed at."""
import os
import struct

import numpy as np

from .records import _fields, _ld, _read_varint, _varint, crc32c, masked_crc32c

_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           14: None, 17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}     # 14 = bfloat16 (as uint16 bits)
_DT_OF = {np.dtype(v): k for k, v in _DTYPES.items() if v is not None}


def _unmask(m):
    r = (m - 0xa282ead8) & 0xffffffff
    return ((r >> 17) | (r << 15)) & 0xffffffff


def _read_block(buf, offset, size, verify=True):
    raw = buf[offset:offset + size]
    btype = buf[offset + size]
    if verify:
        want = struct.unpack("<I", buf[offset + size + 1:offset + size + 5])[0]
        if masked_crc32c(buf[offset:offset + size + 1]) != want:
            raise IOError(f"checkpoint index: block at {offset} fails its CRC")
    if btype != 0:
        raise NotImplementedError("checkpoint index block is snappy-compressed (BundleWriter writes raw blocks)")
    return raw


def _block_entries(block):
    """(key, value) pairs of one table block."""
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _handle(buf, pos):
    off, pos = _read_varint(buf, pos)
    size, pos = _read_varint(buf, pos)
    return off, size, pos


def _shape_of(msg):
    dims = []
    for f, wt, v in _fields(msg):
        if f == 2:                                  # TensorShapeProto.dim
            size = 0
            for f2, wt2, v2 in _fields(v):
                if f2 == 1:
                    size = v2
            dims.append(size)
    return tuple(dims)


def list_variables(prefix, verify=True):
    """{name: (dtype code, shape, shard, offset, size, masked crc)} and the bundle header's shard count"""
    index = memoryview(open(prefix + ".index", "rb").read())
    if len(index) < 48 or struct.unpack("<Q", index[-8:])[0] != _MAGIC:
        raise IOError(f"{prefix}.index is not a TensorFlow checkpoint index (bad magic)")
    footer = index[-48:]
    _, _, p = _handle(footer, 0)                    # meta-index (empty in bundles)
    ioff, isize, _ = _handle(footer, p)
    entries, num_shards = {}, 1
    for _, hv in _block_entries(_read_block(index, ioff, isize, verify)):
        doff, dsize, _ = _handle(hv, 0)
        for key, val in _block_entries(_read_block(index, doff, dsize, verify)):
            if key == b"":
                for f, wt, v in _fields(val):
                    if f == 1:
                        num_shards = v
                    elif f == 2 and v != 0:
                        raise NotImplementedError("big-endian checkpoint")
                continue
            e = dict(dtype=0, shape=(), shard=0, offset=0, size=0, crc=None)
            for f, wt, v in _fields(val):
                if f == 1:
                    e["dtype"] = v
                elif f == 2:
                    e["shape"] = _shape_of(v)
                elif f == 3:
                    e["shard"] = v
                elif f == 4:
                    e["offset"] = v
                elif f == 5:
                    e["size"] = v
                elif f == 6:
                    e["crc"] = struct.unpack("<I", v)[0]
                elif f == 7:
                    raise NotImplementedError(f"{key.decode()}: sliced (partitioned) variables are not supported")
            entries[key.decode()] = e
    return entries, num_shards


def read_tf_checkpoint(prefix, names=None, verify=True):
    """{variable name: ndarray} of a TF1 checkpoint prefix (e.g. 'data/checkpoints/rgb_imagenet/model.ckpt').
    `names`: optional filter (iterable or predicate); verify checks every block and tensor CRC."""
    entries, num_shards = list_variables(prefix, verify)
    keep = (lambda n: True) if names is None else (names if callable(names) else set(names).__contains__)
    shards, out = {}, {}
    for name, e in entries.items():
        if not keep(name):
            continue
        sid = e["shard"]
        if sid not in shards:
            shards[sid] = np.memmap(f"{prefix}.data-{sid:05d}-of-{num_shards:05d}", dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        if verify and e["crc"] is not None and crc32c(raw) != _unmask(e["crc"]):
            raise IOError(f"{name}: tensor bytes fail their CRC")
        if e["dtype"] not in _DTYPES:
            raise NotImplementedError(f"{name}: dtype code {e['dtype']}")
        dt = _DTYPES[e["dtype"]]
        if dt is None:                              # bfloat16 -> float32
            arr = (np.frombuffer(raw, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)
        else:
            arr = np.frombuffer(raw, dtype=dt)
        out[name] = np.array(arr.reshape(e["shape"]))
    return out


# ---- writer (tests, and converting .npz weights into the layout the reference's Saver restores) ---------------
def _build_block(pairs, restart_interval=16):
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(pairs):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _varint(shared) + _varint(len(k) - shared) + _varint(len(v)) + k[shared:] + bytes(v)
        last = k
    if not restarts:
        restarts = [0]
    out += b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))
    return bytes(out)


def write_tf_checkpoint(prefix, tensors, block_size=4096):
    """Write {name: ndarray} as a single-shard tensor bundle (BundleWriter layout: sorted keys, raw blocks)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    header = _varint(1 << 3) + _varint(1) + _ld(3, _varint(1 << 3) + _varint(1))     # num_shards = 1, version.producer = 1
    items = [(b"", header)]
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as data:
        for name in sorted(tensors):
            a = np.asarray(tensors[name])          # (ascontiguousarray would turn a scalar into shape (1,))
            if not a.flags.c_contiguous:
                a = a.copy(order="C")
            raw = a.tobytes()
            shape = b"".join(_ld(2, _varint(1 << 3) + _varint(int(d))) for d in a.shape)
            entry = (_varint(1 << 3) + _varint(_DT_OF[a.dtype]) + _ld(2, shape) +
                     (_varint(4 << 3) + _varint(offset) if offset else b"") + _varint(5 << 3) + _varint(len(raw)) +
                     struct.pack("<B", (6 << 3) | 5) + struct.pack("<I", masked_crc32c(raw)))
            items.append((name.encode(), entry))
            data.write(raw)
            offset += len(raw)
    out = bytearray()
    index_pairs = []

    def emit(block):
        off = len(out)
        out.extend(block)
        out.append(0)                                                              # kNoCompression
        out.extend(struct.pack("<I", masked_crc32c(bytes(block) + b"\x00")))
        return off, len(block)

    cur, cur_bytes = [], 0
    for k, v in items:
        cur.append((k, v))
        cur_bytes += len(k) + len(v) + 3
        if cur_bytes >= block_size:
            off, size = emit(_build_block(cur))
            index_pairs.append((cur[-1][0], _varint(off) + _varint(size)))
            cur, cur_bytes = [], 0
    if cur:
        off, size = emit(_build_block(cur))
        index_pairs.append((cur[-1][0], _varint(off) + _varint(size)))
    moff, msize = emit(_build_block([]))
    ioff, isize = emit(_build_block(index_pairs, restart_interval=1))
    footer = _varint(moff) + _varint(msize) + _varint(ioff) + _varint(isize)
    out.extend(footer + bytes(40 - len(footer)) + struct.pack("<Q", _MAGIC))
    with open(prefix + ".index", "wb") as f:
        f.write(out)
