"""ctypes binding of libfav.so (the C-ABI declared in include/fav.h).

There is no CPU fallback: if the shared library is missing the import of the product path fails
loudly (run `python -c "import __graft_entry__ as g; g.build()"` or `make -C
flickering_adversarial_video_b200/csrc`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfav.so")

FAV_OK = 0
FAV_U8, FAV_F32 = 0, 1
FAV_STACK_TF, FAV_STACK_TORCH = 0, 1
FAV_NET_I3D, FAV_NET_R3D_18, FAV_NET_MC3_18, FAV_NET_R2PLUS1D_18 = 0, 1, 2, 3
ARCHS = {"i3d": FAV_NET_I3D, "r3d_18": FAV_NET_R3D_18, "mc3_18": FAV_NET_MC3_18, "r2plus1d_18": FAV_NET_R2PLUS1D_18}

# indices into the device scalar block (include/fav.h)
S_ADV_LOSS, S_FOOLED, S_SUM_P_MIN, S_SUM_P_MAX = 0, 1, 2, 3
S_NORM_REG, S_DIFF_REG, S_LAP_REG, S_THICKNESS, S_ROUGHNESS, S_TOTAL_LOSS, S_SAT_COUNT = 4, 5, 6, 7, 8, 9, 10
S_COUNT = 16
PROF_KINDS = ["apply", "stem_conv", "conv_halo", "conv_tap", "pool_fwd", "pool_bwd", "head_loss", "stem_grad",
              "delta_update", "other"]


class FavError(RuntimeError):
    pass


class NetDesc(C.Structure):
    _fields_ = [("arch", C.c_int32), ("batch", C.c_int32), ("frames", C.c_int32),
                ("height", C.c_int32), ("width", C.c_int32), ("num_classes", C.c_int32)]


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("ndim", C.c_int32),
                ("dims", C.c_int64 * 5)]


class NormParams(C.Structure):
    _fields_ = [("mean", C.c_float * 3), ("std", C.c_float * 3), ("lo", C.c_float), ("hi", C.c_float)]


class LossParams(C.Structure):
    _fields_ = [("improve_loss", C.c_int32), ("targeted", C.c_int32), ("use_logits", C.c_int32),
                ("margin", C.c_float), ("grad_scale", C.c_float), ("global_batch", C.c_int32),
                ("stack", C.c_int32)]


class RegParams(C.Structure):
    _fields_ = [("beta0", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("beta3", C.c_float),
                ("delta_clip", C.c_float)]


class AdamParams(C.Structure):
    _fields_ = [("lr", C.c_float), ("b1", C.c_float), ("b2", C.c_float), ("eps", C.c_float),
                ("stack", C.c_int32)]


# every symbol include/fav.h declares, with its signature
_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
SIGNATURES = {
    "fav_create": (_i, [C.POINTER(_vp), _i, C.POINTER(NetDesc)]),
    "fav_create_eval": (_i, [C.POINTER(_vp), _i, C.POINTER(NetDesc)]),
    "fav_eval_batch": (_i, [_vp, _vp, _vp, _i, _vp, _f, _vp, _i, _i, _i64, _i, C.POINTER(LossParams), _vp, _vp, _vp, _vp]),
    "fav_destroy": (_i, [_vp]),
    "fav_last_error": (C.c_char_p, []),
    "fav_device_bytes": (_i64, [_vp]),
    "fav_load_weights": (_i, [_vp, C.POINTER(Tensor), _i]),
    "fav_apply_flicker": (_i, [_vp, _vp, _i, _vp, _f, _f, _vp, _vp, _vp]),
    "fav_forward": (_i, [_vp, _vp, _vp]),
    "fav_loss": (_i, [_vp, _vp, C.POINTER(LossParams), _vp, _vp, _vp]),
    "fav_backward_delta": (_i, [_vp, _vp, _vp]),
    "fav_delta_update": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(RegParams), C.POINTER(AdamParams), _vp, _vp]),
    "fav_pixels_enable": (_i, [_vp]),
    "fav_apply_pixels": (_i, [_vp, _vp, _vp, _f, _f, _vp, _vp]),
    "fav_backward_pixels": (_i, [_vp, _vp, _vp]),
    "fav_pixels_update": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _f, _f, C.POINTER(AdamParams), _vp, _vp]),
    "fav_op_conv3d": (_i, [_i, _vp, _i64, _i64, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i64, _i64,
                           _i, _i, _i, _i, _i, _i, _vp, _i64, _i64, _vp]),
    "fav_op_maxpool3d": (_i, [_i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fav_op_maxpool3d_bwd": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "fav_op_loss": (_i, [_i, _vp, _vp, C.POINTER(LossParams), _i, _i, _vp, _vp, _vp, _vp]),
    "fav_op_delta_update": (_i, [_i, _vp, _vp, _vp, _vp, _vp, C.POINTER(RegParams), C.POINTER(AdamParams), _f, _vp, _i, _vp]),
    "fav_op_resize_crop": (_i, [_i, _vp, _i, _i, _i, _i, _i, _f, _f, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "fav_debug_read": (_i64, [_vp, C.c_char_p, _vp, _i64, _vp]),
    "fav_profile_begin": (_i, []),
    "fav_profile_end": (_i, [C.POINTER(C.c_double), _i]),
    "fav_launch_count": (_i64, []),
    "fav_debug_f32_to_f16": (C.c_uint16, [_f]),
    "fav_debug_stem_ts_schedule": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp]),
    "fav_build_info": (C.c_char_p, []),
}

_lib = None


def load():
    """Load libfav.so and bind every declared symbol. Raises FavError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FavError(
            f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU fallback. "
            "Build it with `make -C flickering_adversarial_video_b200/csrc` (needs nvcc, sm_100a).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().fav_last_error().decode("utf-8", "replace")


def check(status, what=""):
    if status != FAV_OK:
        raise FavError(f"{what} failed with status {status}: {last_error()}")
    return status


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def stream_ptr(stream=None, device=None):
    """raw cudaStream_t of `stream`, or of torch's current stream ON `device` (the engine's device, which need not
    be the process's current device)"""
    import torch
    s = stream if stream is not None else torch.cuda.current_stream(device)
    return C.c_void_p(s.cuda_stream)
