"""CPU: pins the [dep] semantics the I3D oracle restates (TF `SAME` padding of strided Conv3D / max_pool3d with -inf
padding, BatchNorm inference with eps 1e-3 and no scale, the VALID average pool of the head) against an INDEPENDENT
implementation: OpenCV's DNN module running ONNX graphs hand-encoded here (`auto_pad = SAME_UPPER` is ONNX's name for
TensorFlow's SAME rule: out = ceil(n / s), the odd padding element goes at the end).  TensorFlow 1.15 itself cannot be
installed; this is the strongest pin available offline.  The full-network test builds InceptionI3d
(i3d.py:144-479) as one ONNX graph from the same variable dictionary the engine loads."""
import numpy as np
import pytest
import torch

from oracle import oracle_i3d as O

cv2 = pytest.importorskip("cv2")
if not hasattr(cv2, "dnn"):
    pytest.skip("OpenCV without the dnn module", allow_module_level=True)


# ---- a minimal ONNX writer (onnx.proto3: ModelProto / GraphProto / NodeProto / AttributeProto / TensorProto) ----
def _varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _vf(field, v):
    return _varint(field << 3) + _varint(v)


def _ld(field, payload):
    return _varint(field << 3 | 2) + _varint(len(payload)) + payload


def _s(field, text):
    return _ld(field, text.encode())


def a_ints(name, vals):
    return _s(1, name) + b"".join(_vf(8, v) for v in vals) + _vf(20, 7)


def a_int(name, v):
    return _s(1, name) + _vf(3, v) + _vf(20, 2)


def a_float(name, v):
    return _s(1, name) + _varint(2 << 3 | 5) + np.float32(v).tobytes() + _vf(20, 1)


def a_str(name, v):
    return _s(1, name) + _ld(4, v.encode()) + _vf(20, 3)


class Graph:
    def __init__(self):
        self.nodes, self.inits, self.n = [], [], 0

    def tensor(self, arr):
        arr = np.ascontiguousarray(arr, np.float32)
        name = f"w{len(self.inits)}"
        self.inits.append(b"".join(_vf(1, d) for d in arr.shape) + _vf(2, 1) + _s(8, name) + _ld(9, arr.tobytes()))
        return name

    def node(self, op, inputs, attrs=()):
        out = f"t{self.n}"
        self.n += 1
        self.nodes.append(b"".join(_s(1, i) for i in inputs) + _s(2, out) + _s(3, "n_" + out) + _s(4, op) +
                          b"".join(_ld(5, a) for a in attrs))
        return out

    @staticmethod
    def _info(name, shape):
        dims = b"".join(_ld(1, _vf(1, d)) for d in shape)
        return _s(1, name) + _ld(2, _ld(1, _vf(1, 1) + _ld(2, dims)))

    def run(self, x, out_name, out_shape):
        g = b"".join(_ld(1, n) for n in self.nodes) + _s(2, "g") + b"".join(_ld(5, t) for t in self.inits) + \
            _ld(11, self._info("x", x.shape)) + _ld(12, self._info(out_name, out_shape))
        model = _vf(1, 6) + _ld(8, _s(1, "") + _vf(2, 11)) + _ld(7, g)
        net = cv2.dnn.readNetFromONNX(np.frombuffer(model, np.uint8))
        net.setInput(np.ascontiguousarray(x, np.float32), "x")
        return net.forward(out_name)

    # layers
    def conv(self, x, w_tf, stride=(1, 1, 1), bias=None):
        w = np.transpose(np.asarray(w_tf, np.float32), (4, 3, 0, 1, 2))           # [kt,kh,kw,Cin,Cout] -> OIDHW
        ins = [x, self.tensor(w)] + ([self.tensor(bias)] if bias is not None else [])
        return self.node("Conv", ins, [a_ints("kernel_shape", w.shape[2:]), a_ints("strides", stride),
                                       a_str("auto_pad", "SAME_UPPER")])

    def maxpool(self, x, k, s):
        return self.node("MaxPool", [x], [a_ints("kernel_shape", k), a_ints("strides", s), a_str("auto_pad", "SAME_UPPER")])

    def unit(self, x, weights, scope, stride=(1, 1, 1)):
        """Unit3D (i3d.py:51-71): Conv3D SAME no bias -> BatchNorm(inference, eps 1e-3, no scale) -> ReLU"""
        y = self.conv(x, weights[O.ROOT + scope + "/conv_3d/w"], stride)
        bn = [np.asarray(weights[O.ROOT + scope + "/batch_norm/" + n], np.float32).reshape(-1)
              for n in ("beta", "moving_mean", "moving_variance")]
        y = self.node("BatchNormalization", [y, self.tensor(np.ones_like(bn[0])), self.tensor(bn[0]), self.tensor(bn[1]),
                                             self.tensor(bn[2])], [a_float("epsilon", 1e-3)])
        return self.node("Relu", [y])


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("shape,k,s", [
    ((1, 3, 9, 11, 10), (7, 7, 7), (2, 2, 2)),        # the I3D stem on odd / even sizes
    ((1, 3, 16, 20, 20), (7, 7, 7), (2, 2, 2)),
    ((2, 5, 6, 9, 8), (3, 3, 3), (1, 1, 1)),
    ((1, 4, 5, 8, 8), (1, 1, 1), (1, 1, 1)),
    ((1, 3, 8, 15, 12), (3, 3, 3), (2, 2, 2)),
])
def test_conv3d_same_matches_opencv(shape, k, s):
    rng = np.random.RandomState(0)
    x = rng.randn(*shape).astype(np.float32)
    w_tf = rng.randn(*k, shape[1], 6).astype(np.float32)
    ref = O.conv3d_same(torch.from_numpy(x), torch.from_numpy(w_tf), s).numpy()
    g = Graph()
    got = g.run(x, g.conv("x", w_tf, s), ref.shape)
    assert got.shape == ref.shape and _rel(got, ref) < 1e-5
    if s != (1, 1, 1):      # the symmetric padding a torch-style port would use gives a different answer
        pad = tuple((kk - 1) // 2 for kk in k)
        sym = torch.nn.functional.conv3d(torch.from_numpy(x), torch.from_numpy(w_tf).permute(4, 3, 0, 1, 2), stride=s,
                                         padding=pad).numpy()
        assert sym.shape != ref.shape or _rel(sym, ref) > 1e-2


@pytest.mark.parametrize("k,s", [((1, 3, 3), (1, 2, 2)), ((3, 3, 3), (2, 2, 2)), ((2, 2, 2), (2, 2, 2)),
                                 ((3, 3, 3), (1, 1, 1))])          # every pool of i3d.py:174,189,252,398 + Branch_3
@pytest.mark.parametrize("shape", [(1, 3, 9, 11, 10), (1, 2, 8, 14, 14), (1, 2, 5, 7, 7)])
def test_maxpool3d_same_matches_opencv(k, s, shape):
    x = np.random.RandomState(3).randn(*shape).astype(np.float32) - 3.0       # mostly negative: zero padding would win
    ref = O.maxpool3d_same(torch.from_numpy(x), k, s).numpy()
    g = Graph()
    got = g.run(x, g.maxpool("x", k, s), ref.shape)
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_unit3d_matches_opencv():
    from flickering_adversarial_video_b200 import synthetic
    weights = synthetic.i3d_weights(seed=0)
    x = np.random.RandomState(1).randn(1, 64, 4, 9, 9).astype(np.float32)
    model = O.OracleI3D(weights)
    ref = model.unit(torch.from_numpy(x), "Conv3d_2b_1x1").numpy()
    g = Graph()
    got = g.run(x, g.unit("x", weights, "Conv3d_2b_1x1"), ref.shape)
    assert _rel(got, ref) < 1e-5 and (ref == 0).mean() > 0.05          # the ReLU is active


def test_full_i3d_forward_matches_opencv():
    """InceptionI3d(final_endpoint='Logits') on a 17-frame clip (odd T: asymmetric SAME padding in time at every
    strided stage; T' = 2 at the head so the mean over T' is exercised): oracle vs OpenCV's DNN running the same
    graph."""
    from flickering_adversarial_video_b200 import synthetic
    weights = synthetic.i3d_weights(seed=0)
    clip = synthetic.clips_u8(1, 17, seed=1009)
    x = O.normalize_u8(clip)                                             # [1,T,224,224,3]
    with torch.no_grad():
        ref = O.OracleI3D(weights).forward(x).numpy()

    g = Graph()
    net = g.unit("x", weights, "Conv3d_1a_7x7", (2, 2, 2))
    net = g.maxpool(net, (1, 3, 3), (1, 2, 2))
    net = g.unit(net, weights, "Conv3d_2b_1x1")
    net = g.unit(net, weights, "Conv3d_2c_3x3")
    net = g.maxpool(net, (1, 3, 3), (1, 2, 2))
    for name, *_ in O.BLOCKS:
        if name == "Mixed_4b":
            net = g.maxpool(net, (3, 3, 3), (2, 2, 2))
        if name == "Mixed_5b":
            net = g.maxpool(net, (2, 2, 2), (2, 2, 2))
        b2b = "Conv3d_0a_3x3" if name == "Mixed_5b" else "Conv3d_0b_3x3"        # i3d.py:418 naming quirk
        b0 = g.unit(net, weights, f"{name}/Branch_0/Conv3d_0a_1x1")
        b1 = g.unit(g.unit(net, weights, f"{name}/Branch_1/Conv3d_0a_1x1"), weights, f"{name}/Branch_1/Conv3d_0b_3x3")
        b2 = g.unit(g.unit(net, weights, f"{name}/Branch_2/Conv3d_0a_1x1"), weights, f"{name}/Branch_2/{b2b}")
        b3 = g.unit(g.maxpool(net, (3, 3, 3), (1, 1, 1)), weights, f"{name}/Branch_3/Conv3d_0b_1x1")
        net = g.node("Concat", [b0, b1, b2, b3], [a_int("axis", 1)])
    net = g.node("AveragePool", [net], [a_ints("kernel_shape", (2, 7, 7)), a_ints("strides", (1, 1, 1))])
    out = g.conv(net, weights[O.ROOT + "Logits/Conv3d_0c_1x1/conv_3d/w"],
                 bias=np.asarray(weights[O.ROOT + "Logits/Conv3d_0c_1x1/conv_3d/b"], np.float32).reshape(-1))
    got = g.run(x.permute(0, 4, 1, 2, 3).numpy(), out, (1, 400, 2, 1, 1))
    assert got.shape == (1, 400, 2, 1, 1)
    logits = got.reshape(1, 400, 2).mean(-1)                              # tf.reduce_mean over T' (i3d.py:472)
    assert _rel(logits, ref) < 1e-4 and int(logits.argmax()) == int(ref.argmax())
