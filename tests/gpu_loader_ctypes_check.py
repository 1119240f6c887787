"""Torch-free GPU check of the loader's Python binding (run by hand on a GPU box: `python tests/gpu_loader_ctypes_check.py`;
not collected by pytest).  numpy + ctypes only, so it starts in a second: device buffers come from libcudart, the
launch goes through video_dataset.launch_resize_crop -> _lib.SIGNATURES -> libfav.so, the result is compared with
oracle/oracle_loader.py bit for bit."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_loader as ol                                     # noqa: E402
from flickering_adversarial_video_b200 import video_dataset as vd          # noqa: E402


def main():
    rt = C.CDLL("libcudart.so")
    rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rt.cudaMemset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]

    def dev_alloc(nbytes):
        p = C.c_void_p()
        assert rt.cudaMalloc(C.byref(p), nbytes) == 0
        return p

    rc = 0
    rng = np.random.RandomState(3)
    for (T, H, W, im_scale, size) in [(16, 256, 340, 128, 112), (8, 320, 240, 128, 112), (4, 128, 171, 128, 112),
                                      (64, 240, 320, 128, 112)]:
        clip = rng.randint(0, 256, (T, H, W, 3)).astype(np.uint8)
        clip[:, : H // 4][rng.rand(T, H // 4, W, 3) < 0.5] = 255
        v = ol.resize_crop(clip, im_scale, size)
        ref_u8, ref_f32 = ol.quantize(v), ol.normalize_ncthw(v)
        d_src, d_u8, d_f32 = dev_alloc(clip.nbytes), dev_alloc(ref_u8.nbytes), dev_alloc(ref_f32.nbytes)
        assert rt.cudaMemcpy(d_src, clip.ctypes.data, clip.nbytes, 1) == 0
        got_u8, got_f32 = np.empty_like(ref_u8), np.empty_like(ref_f32)

        def fetch_u8(tag):
            assert rt.cudaDeviceSynchronize() == 0, tag
            assert rt.cudaMemcpy(got_u8.ctypes.data, d_u8, got_u8.nbytes, 2) == 0
            d = np.abs(got_u8.astype(np.int32) - ref_u8.astype(np.int32))
            print(f"  [{tag}] uint8 mismatches {int((d > 0).sum())}/{d.size}, max |diff| {int(d.max())}, "
                  f"bytes still 0x5A {int((got_u8 == 0x5A).sum())}, head got {got_u8.ravel()[:9].tolist()} "
                  f"ref {ref_u8.ravel()[:9].tolist()}")
            return int((d > 0).sum())

        rt.cudaMemset(d_u8, 0x5A, ref_u8.nbytes)
        rt.cudaDeviceSynchronize()
        vd.launch_resize_crop(d_src, T, H, W, d_u8, None, im_scale, size, frames_per_clip=T)
        bad8 = fetch_u8("uint8 only")
        rt.cudaMemset(d_u8, 0x5A, ref_u8.nbytes)
        rt.cudaDeviceSynchronize()
        vd.launch_resize_crop(d_src, T, H, W, d_u8, d_f32, im_scale, size, frames_per_clip=T)
        bad8 += fetch_u8("uint8 + fp32")
        assert rt.cudaMemcpy(got_f32.ctypes.data, d_f32, got_f32.nbytes, 2) == 0
        badf = int((got_f32 != ref_f32).sum())
        assert ref_u8.flags.c_contiguous and ref_f32.flags.c_contiguous      # raw memcpy targets below
        print(f"launch_resize_crop {clip.shape} -> {got_u8.shape}: uint8 mismatches {bad8}/{got_u8.size}, "
              f"fp32 mismatches {badf}/{got_f32.size} (max abs {float(np.abs(got_f32 - ref_f32).max()):.3e})")
        rc |= int(bad8 != 0 or badf != 0)
        for p in (d_src, d_u8, d_f32):
            rt.cudaFree(p)
    print("CTYPES CHECK PASS" if rc == 0 else "CTYPES CHECK FAIL")
    return rc


if __name__ == "__main__":
    sys.exit(main())
