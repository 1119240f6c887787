"""Thin host-side wrapper of the libfav C-ABI: one `FlickerEngine` per GPU (per rank).

PyTorch is used for device memory and streams only; every kernel on the hot path lives in
libfav.so.  Replaces the TF session + graph of utils/kinetics_i3d_utils.py:76-208.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


class FlickerEngine:
    _EVAL = False   # EvalEngine: forward-only handle (fav_create_eval)

    def __init__(self, batch, frames, height=None, width=None, num_classes=400, device=0, arch="i3d"):
        """arch: "i3d" (TF stack, 224x224) or "r3d_18" / "mc3_18" / "r2plus1d_18" (torch stack, 112x112)."""
        if not torch.cuda.is_available():
            raise L.FavError("FlickerEngine needs a CUDA device (sm_100a); there is no CPU fallback")
        if arch not in L.ARCHS:
            raise ValueError(f"unknown arch {arch!r}; expected one of {sorted(L.ARCHS)}")
        self.lib = L.load()
        self.device = torch.device("cuda", device)
        self.arch = arch
        self.torch_stack = arch != "i3d"
        side = 112 if self.torch_stack else 224
        height = side if height is None else height
        width = side if width is None else width
        self.B, self.T, self.H, self.W, self.K = batch, frames, height, width, num_classes
        desc = L.NetDesc(L.ARCHS[arch], batch, frames, height, width, num_classes)
        h = C.c_void_p()
        create = self.lib.fav_create_eval if self._EVAL else self.lib.fav_create
        with torch.cuda.device(self.device):
            L.check(create(C.byref(h), device, C.byref(desc)), "fav_create_eval" if self._EVAL else "fav_create")
        self.h = h
        self.scalars = torch.zeros(L.S_COUNT, dtype=torch.float32, device=self.device)
        if self._EVAL:
            return
        self.logits = torch.zeros((batch, num_classes), dtype=torch.float32, device=self.device)
        self.probs = torch.zeros((batch, num_classes), dtype=torch.float32, device=self.device)
        self.grad = torch.zeros((frames, 3), dtype=torch.float32, device=self.device)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.fav_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_bytes(self):
        return int(self.lib.fav_device_bytes(self.h))

    def load_weights(self, weights):
        """weights: {tf variable name: float32 ndarray} (reference ckpt naming/layout) for I3D, or a
        torchvision `state_dict()` (names and [Cout,Cin,kt,kh,kw] layout as torch stores them) for the
        video ResNets; integer entries (num_batches_tracked) are ignored."""
        weights = {k: v for k, v in weights.items() if not k.endswith("num_batches_tracked")}
        n = len(weights)
        arr = (L.Tensor * n)()
        keep = []
        for i, (name, a) in enumerate(weights.items()):
            if isinstance(a, torch.Tensor):
                a = a.detach().cpu().numpy()
            a = np.ascontiguousarray(a, dtype=np.float32)
            keep.append(a)
            arr[i].name = name.encode()
            arr[i].data = a.ctypes.data_as(C.POINTER(C.c_float))
            arr[i].ndim = min(a.ndim, 5)
            dims = list(a.shape)[:5] + [1] * (5 - min(a.ndim, 5))
            for j in range(5):
                arr[i].dims[j] = dims[j]
        with torch.cuda.device(self.device):
            L.check(self.lib.fav_load_weights(self.h, arr, n), "fav_load_weights")

    # ---- hot path ------------------------------------------------------------------------
    def apply(self, clip, delta, adv_flag=1.0, delta_clip=0.4, adv_u8=None, adv_f32=None, stream=None):
        """clip [B,T,H,W,3] uint8 (or float32 for I3D); delta [T,3].  Torch-stack archs: delta_clip is the
        Perturbation's dynamic_max_norm and adv_f32, when given, is NCTHW [B,3,T,H,W] (normalised)."""
        assert clip.is_cuda and clip.is_contiguous() and tuple(clip.shape) == (self.B, self.T, self.H, self.W, 3)
        assert delta.is_cuda and delta.dtype == torch.float32 and delta.numel() == self.T * 3
        dt = L.FAV_U8 if clip.dtype == torch.uint8 else L.FAV_F32
        if dt == L.FAV_F32:
            assert clip.dtype == torch.float32
        L.check(self.lib.fav_apply_flicker(self.h, L.ptr(clip), dt, L.ptr(delta), adv_flag, delta_clip,
                                           L.ptr(adv_u8), L.ptr(adv_f32), L.stream_ptr(stream, self.device)), "fav_apply_flicker")

    def forward(self, stream=None):
        L.check(self.lib.fav_forward(self.h, L.ptr(self.logits), L.stream_ptr(stream, self.device)), "fav_forward")
        return self.logits

    def loss(self, labels, improve_loss=True, targeted=False, use_logits=False, margin=0.05, grad_scale=1.0,
             global_batch=0, stack=L.FAV_STACK_TF, stream=None):
        assert labels.is_cuda and labels.dtype == torch.int64 and labels.numel() == self.B
        p = L.LossParams(int(improve_loss), int(targeted), int(use_logits), margin, grad_scale, global_batch, stack)
        L.check(self.lib.fav_loss(self.h, L.ptr(labels), C.byref(p), L.ptr(self.probs), L.ptr(self.scalars),
                                  L.stream_ptr(stream, self.device)), "fav_loss")
        return self.scalars

    def backward(self, stream=None):
        L.check(self.lib.fav_backward_delta(self.h, L.ptr(self.grad), L.stream_ptr(stream, self.device)), "fav_backward_delta")
        return self.grad

    def update(self, delta, grad, m, v, step, beta0, beta1, beta2, beta3, lr=1e-3, delta_clip=0.4,
               b1=0.9, b2=0.999, eps=1e-8, stack=L.FAV_STACK_TF, stream=None):
        reg = L.RegParams(beta0, beta1, beta2, beta3, delta_clip)
        adam = L.AdamParams(lr, b1, b2, eps, stack)
        L.check(self.lib.fav_delta_update(self.h, L.ptr(delta), L.ptr(grad), L.ptr(m), L.ptr(v), L.ptr(step),
                                          C.byref(reg), C.byref(adam), L.ptr(self.scalars), L.stream_ptr(stream, self.device)),
                "fav_delta_update")
        return self.scalars

    # ---- sparse per-pixel attack (kinetics_i3d_L12 / torch attack_type "L12") ------------------
    def pixels_enable(self):
        with torch.cuda.device(self.device):
            L.check(self.lib.fav_pixels_enable(self.h), "fav_pixels_enable")

    def apply_pixels(self, clip_u8, delta_px, adv_flag=1.0, delta_clip=0.0, adv_f32=None, stream=None):
        """delta_px [T,H,W,3] float32 (the torch stack's [3,T,H,W] permuted); delta_clip <= 0: no clamp."""
        assert clip_u8.is_cuda and clip_u8.dtype == torch.uint8 and clip_u8.is_contiguous()
        assert tuple(clip_u8.shape) == (self.B, self.T, self.H, self.W, 3)
        assert delta_px.is_cuda and delta_px.dtype == torch.float32 and tuple(delta_px.shape) == (self.T, self.H, self.W, 3)
        L.check(self.lib.fav_apply_pixels(self.h, L.ptr(clip_u8), L.ptr(delta_px), adv_flag, delta_clip, L.ptr(adv_f32),
                                          L.stream_ptr(stream, self.device)), "fav_apply_pixels")

    def backward_pixels(self, grad_px, stream=None):
        assert grad_px.is_cuda and grad_px.dtype == torch.float32 and tuple(grad_px.shape) == (self.T, self.H, self.W, 3)
        L.check(self.lib.fav_backward_pixels(self.h, L.ptr(grad_px), L.stream_ptr(stream, self.device)), "fav_backward_pixels")
        return grad_px

    def update_pixels(self, delta_px, grad_px, m, v, step, reg_weight, delta_clip=0.0, lr=1e-3, b1=0.9, b2=0.999,
                      eps=1e-8, stack=L.FAV_STACK_TF, stream=None):
        adam = L.AdamParams(lr, b1, b2, eps, stack)
        L.check(self.lib.fav_pixels_update(self.h, L.ptr(delta_px), L.ptr(grad_px), L.ptr(m), L.ptr(v), L.ptr(step),
                                           reg_weight, delta_clip, C.byref(adam), L.ptr(self.scalars),
                                           L.stream_ptr(stream, self.device)), "fav_pixels_update")
        return self.scalars

    def read(self, name, shape):
        """Debug read of an internal activation / gradient buffer as fp32 [B,T,H,W,C]."""
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        n = self.lib.fav_debug_read(self.h, name.encode(), L.ptr(out), out.numel(), L.stream_ptr(None, self.device))
        if n < 0:
            raise L.FavError(f"fav_debug_read({name}): {L.last_error()}")
        assert n == out.numel(), (name, n, out.numel())
        return out


class EvalEngine(FlickerEngine):
    """Forward-only evaluation handle (include/fav.h `fav_create_eval` / `fav_eval_batch`): the
    clean and the perturbed version of every validation batch run through the network in ONE pass and the fooling-ratio
    counters stay on the device.  Replaces the two `sess.run` per batch of `kinetics_i3d.evaluate`
    (utils/kinetics_i3d_utils.py:217-250) and the two `model(...)` calls of the torch stack's validation phase
    (utils_cv/action_recognition/model.py:697-713).  No gradient buffers: ~40 % of a training handle's memory per clip."""
    _EVAL = True

    def __init__(self, batch, frames, height=None, width=None, num_classes=400, device=0, arch="i3d"):
        super().__init__(batch, frames, height, width, num_classes, device, arch)
        self.counts = torch.zeros(2, dtype=torch.int64, device=self.device)          # [miss, valid], accumulates
        self.probs = torch.zeros((2 * batch, num_classes), dtype=torch.float32, device=self.device)

    def reset_counts(self):
        self.counts.zero_()

    def eval_batch(self, clips, delta, labels, clips_adv=None, n_clips=None, delta_clip=0.4, targeted=False,
                   target_class=0, exclude_misclassify=True, loss=None, want_probs=False, stream=None):
        """clips [B,T,H,W,3] uint8 (float32 allowed for I3D); clips_adv: the clips the perturbation is added to (None =
        the same); labels int64 [B]; n_clips: how many clips of a ragged last batch count; loss: None or the keyword
        dict of `FlickerEngine.loss` (the loss scalars of the perturbed rows land in `self.scalars`).  Asynchronous:
        `self.counts` += [miss, valid]; with want_probs `self.probs[:B]` / `[B:]` hold the clean / perturbed softmax."""
        shape = (self.B, self.T, self.H, self.W, 3)
        for c in (clips, clips_adv):
            assert c is None or (c.is_cuda and c.is_contiguous() and tuple(c.shape) == shape)
        assert clips_adv is None or clips_adv.dtype == clips.dtype
        assert delta.is_cuda and delta.dtype == torch.float32 and delta.numel() == self.T * 3 and delta.is_contiguous()
        assert labels.is_cuda and labels.dtype == torch.int64 and labels.numel() == self.B
        dt = L.FAV_U8 if clips.dtype == torch.uint8 else L.FAV_F32
        assert dt == L.FAV_U8 or clips.dtype == torch.float32
        lp = None
        if loss is not None:
            lp = L.LossParams(int(loss.get("improve_loss", True)), int(loss.get("targeted", False)),
                              int(loss.get("use_logits", False)), float(loss.get("margin", 0.05)),
                              float(loss.get("grad_scale", 1.0)), int(loss.get("global_batch", 0)),
                              int(loss.get("stack", L.FAV_STACK_TF)))
        L.check(self.lib.fav_eval_batch(
            self.h, L.ptr(clips), L.ptr(clips_adv), dt, L.ptr(delta), delta_clip, L.ptr(labels),
            self.B if n_clips is None else int(n_clips), int(targeted), int(target_class or 0), int(exclude_misclassify),
            None if lp is None else C.byref(lp), L.ptr(self.counts), L.ptr(self.probs) if want_probs else None,
            L.ptr(self.scalars) if lp is not None else None, L.stream_ptr(stream, self.device)), "fav_eval_batch")
        return self.counts


# ---- op-level wrappers used by the parity tests ---------------------------------------------
def op_conv3d(x16, w_tf, bias=None, relu=False, dgrad=False, relu_src=None, y=None, y_coff=0, x_coff=0,
              cin=None, cout=None):
    """x16 [B,T,H,W,Cs] cuda; w_tf [kt,kh,kw,cin,cout] float32 (TF layout).  Storage formats follow the engine's:
    the forward conv reads and writes float16 (activations), the data gradient (dgrad=True) reads and writes
    bfloat16 (gradients) and takes its ReLU mask `relu_src` from a float16 activation."""
    lib = L.load()
    x_bf16 = x16
    want = torch.bfloat16 if dgrad else torch.float16
    assert x16.dtype == want, f"op_conv3d(dgrad={dgrad}) takes {want} input, got {x16.dtype}"
    assert relu_src is None or relu_src.dtype == torch.float16
    B, T, H, W, xcs = x_bf16.shape
    kt, kh, kw, wcin, wcout = w_tf.shape
    cin = wcin if cin is None else cin
    cout = wcout if cout is None else cout
    n_out = cin if dgrad else cout
    if y is None:
        y = torch.zeros((B, T, H, W, n_out), dtype=want, device=x_bf16.device)
    assert y.dtype == want
    w_host = np.ascontiguousarray(w_tf.detach().cpu().numpy(), dtype=np.float32)
    b_host = None if bias is None else np.ascontiguousarray(bias.detach().cpu().numpy(), dtype=np.float32)
    st = lib.fav_op_conv3d(
        x_bf16.device.index or 0, L.ptr(x_bf16), xcs, x_coff,
        w_host.ctypes.data_as(C.c_void_p), None if b_host is None else b_host.ctypes.data_as(C.c_void_p),
        kt, kh, kw, cin, cout, L.ptr(y), y.shape[-1], y_coff, B, T, H, W, int(relu), int(dgrad),
        L.ptr(relu_src), 0 if relu_src is None else relu_src.shape[-1], 0, L.stream_ptr())
    L.check(st, "fav_op_conv3d")
    return y


def op_maxpool3d(x_f16, k, s):
    """forward pool on float16 activations -> (y float16, arg-max codes uint8)"""
    lib = L.load()
    x_bf16 = x_f16
    assert x_f16.dtype == torch.float16
    B, T, H, W, Cc = x_bf16.shape
    To, Ho, Wo = -(-T // s[0]), -(-H // s[1]), -(-W // s[2])
    y = torch.empty((B, To, Ho, Wo, Cc), dtype=torch.float16, device=x_bf16.device)
    idx = torch.empty((B, To, Ho, Wo, Cc), dtype=torch.uint8, device=x_bf16.device)
    L.check(lib.fav_op_maxpool3d(x_bf16.device.index or 0, L.ptr(x_bf16), L.ptr(y), L.ptr(idx), B, T, H, W, Cc,
                                 k[0], k[1], k[2], s[0], s[1], s[2], L.stream_ptr()), "fav_op_maxpool3d")
    return y, idx


def op_maxpool3d_bwd(dy, idx, in_shape, k, s, add=None, relu_src=None):
    """dy / add / result bfloat16 (gradients); relu_src float16 (the pool input's producer activation)"""
    lib = L.load()
    assert dy.dtype == torch.bfloat16 and (add is None or add.dtype == torch.bfloat16)
    assert relu_src is None or relu_src.dtype == torch.float16
    B, T, H, W, Cc = in_shape
    dx = torch.empty(in_shape, dtype=torch.bfloat16, device=dy.device)
    L.check(lib.fav_op_maxpool3d_bwd(dy.device.index or 0, L.ptr(dy), L.ptr(idx), L.ptr(add), L.ptr(relu_src),
                                     L.ptr(dx), B, T, H, W, Cc, k[0], k[1], k[2], s[0], s[1], s[2],
                                     L.stream_ptr()), "fav_op_maxpool3d_bwd")
    return dx


def op_loss(logits, labels, improve_loss=True, targeted=False, use_logits=False, margin=0.05, stack=L.FAV_STACK_TF,
            global_batch=0):
    """returns (probs, dlogits, scalars) — the loss kernel on caller tensors"""
    lib = L.load()
    B, K = logits.shape
    probs = torch.empty_like(logits)
    dlogits = torch.empty_like(logits)
    scalars = torch.zeros(L.S_COUNT, dtype=torch.float32, device=logits.device)
    p = L.LossParams(int(improve_loss), int(targeted), int(use_logits), margin, 1.0, global_batch, stack)
    L.check(lib.fav_op_loss(logits.device.index or 0, L.ptr(logits), L.ptr(labels), C.byref(p), B, K, L.ptr(probs),
                            L.ptr(dlogits), L.ptr(scalars), L.stream_ptr()), "fav_op_loss")
    return probs, dlogits, scalars


def op_delta_update(delta, grad, m, v, step, beta0, beta1, beta2, beta3, lr=1e-3, delta_clip=0.4, stack=L.FAV_STACK_TF,
                    adv_flag=1.0, scalars=None):
    lib = L.load()
    if scalars is None:
        scalars = torch.zeros(L.S_COUNT, dtype=torch.float32, device=delta.device)
    reg = L.RegParams(beta0, beta1, beta2, beta3, delta_clip)
    adam = L.AdamParams(lr, 0.9, 0.999, 1e-8, stack)
    L.check(lib.fav_op_delta_update(delta.device.index or 0, L.ptr(delta), L.ptr(grad), L.ptr(m), L.ptr(v), L.ptr(step),
                                    C.byref(reg), C.byref(adam), adv_flag, L.ptr(scalars), delta.shape[0],
                                    L.stream_ptr()), "fav_op_delta_update")
    return scalars
