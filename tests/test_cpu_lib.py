"""not gpu: the C-ABI shared library loads and exports every symbol include/fav.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fav.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fav_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    from flickering_adversarial_video_b200 import _lib
    declared = _declared_symbols()
    assert set(declared) == set(_lib.SIGNATURES), (declared, sorted(_lib.SIGNATURES))
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libfav.so does not export {name}"
    assert _lib.load().fav_build_info().decode().startswith("sm_100a")


def test_no_gpu_is_a_loud_error():
    """no CPU fallback: creating an engine without a CUDA device must fail with a clear message"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from flickering_adversarial_video_b200 import _lib
    from flickering_adversarial_video_b200.engine import FlickerEngine
    with pytest.raises(_lib.FavError):
        FlickerEngine(1, 16)
    lib = _lib.load()
    h = ctypes.c_void_p()
    desc = _lib.NetDesc(0, 1, 16, 224, 224, 400)
    st = lib.fav_create(ctypes.byref(h), 0, ctypes.byref(desc))
    assert st == -4 and b"no CUDA device" in lib.fav_last_error()


def test_favio_header_symbols_are_exported():
    """include/favio.h (host I/O helpers of the f1 / f4 rows): every declared symbol is exported by libfavio.so"""
    text = open(os.path.join(ROOT, "include", "favio.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(favio_[a-z0-9_]+)\s*\(", text)))
    assert declared == ["favio_crc32c", "favio_masked_crc32c", "favio_tfrecord_index"]
    path = os.path.join(ROOT, "flickering_adversarial_video_b200", "libfavio.so")
    if not os.path.exists(path):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), f"libfavio.so does not export {name}"


def test_forward_weight_conversion_is_ieee_half_rne():
    """fav_load_weights packs the forward operands as fp16 on the host: the conversion must be numpy's (IEEE
    round-to-nearest-even incl. subnormals) everywhere below the saturation point, and saturate instead of inf."""
    import numpy as np
    from flickering_adversarial_video_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.standard_normal(20000).astype(np.float32) * np.float32(0.05),            # weight-like
        (rng.standard_normal(5000) * 10.0 ** rng.uniform(-9, 4.5, 5000)).astype(np.float32),   # wide dynamic range
        np.array([0.0, -0.0, 1.0, -1.0, 65504.0, -65504.0, 65519.9, 6.1035156e-05, 6.0975552e-05, 5.9604645e-08,
                  2.9802322e-08, 2.98e-08, 3.0e-08, 8.9406967e-08, 1.0009765625, 1.00048828125, 1.00146484375],
                 dtype=np.float32),
    ])
    got = np.array([lib.fav_debug_f32_to_f16(float(v)) for v in vals], dtype=np.uint16)
    with np.errstate(over="ignore"):
        ref = vals.astype(np.float16).view(np.uint16)
    finite = np.abs(vals) < 65520.0
    assert np.array_equal(got[finite], ref[finite]), np.flatnonzero(got[finite] != ref[finite])[:10]
    big = np.array([70000.0, -1e9, 65520.0], dtype=np.float32)
    sat = [lib.fav_debug_f32_to_f16(float(v)) for v in big]
    assert sat == [0x7bff, 0xfbff, 0x7bff]
