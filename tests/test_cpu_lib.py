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
