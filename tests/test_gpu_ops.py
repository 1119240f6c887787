"""-m gpu: layer-wise parity of the CUDA kernels (through the C-ABI op entry points) against a torch
fp32 reference of the same op on the same 16-bit-rounded inputs (forward ops: float16 activations and weights;
data gradients: bfloat16 gradients and weights, ReLU mask from a float16 activation).  Tolerances: the kernels
accumulate in fp32 and round the result to the storage format, so |err| <= 2^-8 * |ref| (bf16) or 2^-10 * |ref|
(fp16) + small absolute slack."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _same_pad(x, k, s, value=0.0):
    from oracle.oracle_i3d import same_pads
    T, H, W = x.shape[2:]
    pt, ph, pw = same_pads(T, k[0], s[0]), same_pads(H, k[1], s[1]), same_pads(W, k[2], s[2])
    return F.pad(x, (pw[0], pw[1], ph[0], ph[1], pt[0], pt[1]), value=value)


def _ref_conv(x_ndhwc, w_tf, bias=None, relu=False):
    x = x_ndhwc.float().permute(0, 4, 1, 2, 3)
    w = w_tf.float().permute(4, 3, 0, 1, 2).contiguous()
    y = F.conv3d(_same_pad(x, w.shape[2:], (1, 1, 1)), w)
    if bias is not None:
        y = y + bias.view(1, -1, 1, 1, 1)
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 4, 1).contiguous()


def _check(got, ref, what, rtol=2 ** -7, atol=2e-2, scale=None):
    """|got - ref| <= rtol * scale + atol element-wise; scale defaults to |ref|"""
    got = got.float()
    err = (got - ref).abs()
    tol = rtol * (ref.abs() if scale is None else scale) + atol
    bad = (err > tol)
    nbad = int(bad.sum())
    rel = float(err.max() / (ref.abs().max() + 1e-12))
    print(f"{what}: max|err|={float(err.max()):.4g} max|ref|={float(ref.abs().max()):.4g} rel={rel:.3g} bad={nbad}/{bad.numel()}")
    if nbad:
        idx = torch.nonzero(bad)[:8]
        for i in idx:
            i = tuple(int(v) for v in i)
            print("   at", i, "got", float(got[i]), "ref", float(ref[i]))
    assert nbad == 0, f"{what}: {nbad} elements out of tolerance (rel {rel:.3g})"


CONV_CASES = [
    # (B, T, H, W, k, cin, cout)
    (1, 2, 8, 8, 1, 64, 64),        # flat 1x1x1, M == 128
    (1, 3, 7, 7, 1, 192, 96),       # flat, ragged M
    (2, 4, 14, 14, 1, 832, 384),    # flat, two N tiles, 13 k-blocks
    (1, 2, 8, 8, 3, 64, 64),        # 3x3x3, exact box
    (1, 4, 16, 16, 3, 64, 192),
    (2, 5, 14, 14, 3, 96, 208),     # odd extents, partial k-block (96), N=208
    (1, 6, 7, 7, 3, 16, 48),        # tiny cin (one 16-wide k-step)
    (1, 3, 28, 28, 3, 32, 96),
    (1, 4, 14, 14, 3, 160, 320),    # two N tiles of 160
    (1, 3, 7, 7, 3, 192, 384),
    (1, 3, 56, 56, 3, 64, 192),     # halo path: 2 rows per tile
    (2, 4, 28, 28, 3, 96, 128),     # halo path: 4 rows per tile, partial k-block
    (1, 5, 14, 14, 3, 112, 224),    # halo path: 8 + 6 rows
    (1, 3, 28, 28, 3, 16, 32),      # halo path, tiny cin
]


@pytest.mark.parametrize("B,T,H,W,k,cin,cout", CONV_CASES)
def test_conv3d_forward(B, T, H, W, k, cin, cout):
    from flickering_adversarial_video_b200.engine import op_conv3d
    g = torch.Generator(device="cuda").manual_seed(1234 + cin + cout + k)
    x = torch.randn((B, T, H, W, cin), generator=g, device="cuda").to(torch.float16)
    w = (torch.randn((k, k, k, cin, cout), generator=g, device="cuda") * (2.0 / (k ** 3 * cin)) ** 0.5)
    w = w.to(torch.float16).float()
    bias = torch.randn((cout,), generator=g, device="cuda") * 0.1
    y = op_conv3d(x, w, bias=bias, relu=True)
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, bias, relu=True)
    _check(y, ref, f"conv fwd {B}x{T}x{H}x{W} k{k} {cin}->{cout}", rtol=2 ** -9, atol=5e-3)


@pytest.mark.parametrize("B,T,H,W,k,cin,cout", CONV_CASES)
def test_conv3d_dgrad(B, T, H, W, k, cin, cout):
    from flickering_adversarial_video_b200.engine import op_conv3d
    if cin % 16 or cout % 16:
        pytest.skip("dgrad GEMM-K needs cout % 16 == 0")
    g = torch.Generator(device="cuda").manual_seed(4321 + cin + cout + k)
    xin = torch.randn((B, T, H, W, cin), generator=g, device="cuda").to(torch.float16)    # relu source (an activation)
    dy = torch.randn((B, T, H, W, cout), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((k, k, k, cin, cout), generator=g, device="cuda") * (2.0 / (k ** 3 * cout)) ** 0.5)
    w = w.to(torch.bfloat16).float()
    dx = op_conv3d(dy, w, dgrad=True, relu_src=xin)
    torch.cuda.synchronize()
    # reference: autograd of the forward conv, then the ReLU mask of the conv input
    xr = xin.float().clone().requires_grad_(True)
    yr = _ref_conv(xr, w)
    (gx,) = torch.autograd.grad(yr, xr, dy.float())
    ref = gx * (xin.float() > 0)
    _check(dx, ref, f"conv dgrad {B}x{T}x{H}x{W} k{k} {cin}<-{cout}")


def test_conv3d_channel_slices():
    """branch epilogues write a channel slice of a wider tensor and read a slice as input"""
    from flickering_adversarial_video_b200.engine import op_conv3d
    g = torch.Generator(device="cuda").manual_seed(99)
    B, T, H, W = 1, 3, 14, 14
    xw = torch.randn((B, T, H, W, 256), generator=g, device="cuda").to(torch.float16)
    w = (torch.randn((3, 3, 3, 64, 32), generator=g, device="cuda") * 0.05).to(torch.float16).float()
    y = torch.full((B, T, H, W, 128), 7.0, dtype=torch.float16, device="cuda")
    op_conv3d(xw, w, relu=False, y=y, y_coff=64, x_coff=128, cin=64, cout=32)
    torch.cuda.synchronize()
    ref = _ref_conv(xw[..., 128:192], w)
    _check(y[..., 64:96], ref, "conv slice")
    assert float((y[..., :64].float() - 7.0).abs().max()) == 0.0
    assert float((y[..., 96:].float() - 7.0).abs().max()) == 0.0


POOL_CASES = [
    # (B,T,H,W,C,k,s)
    (1, 5, 16, 16, 64, (1, 3, 3), (1, 2, 2)),
    (2, 5, 14, 14, 64, (3, 3, 3), (1, 1, 1)),
    (1, 9, 14, 14, 32, (3, 3, 3), (2, 2, 2)),
    (1, 7, 14, 14, 32, (2, 2, 2), (2, 2, 2)),
    (1, 4, 28, 28, 192, (3, 3, 3), (1, 1, 1)),
    # separable streaming path (pool3.cu): T spanning several segments, odd channel-group divisors, 7x7 planes
    (2, 19, 28, 28, 32, (3, 3, 3), (1, 1, 1)),
    (1, 16, 14, 14, 528, (3, 3, 3), (1, 1, 1)),
    (2, 8, 7, 7, 832, (3, 3, 3), (1, 1, 1)),
    (1, 11, 7, 7, 48, (3, 3, 3), (1, 1, 1)),
    (1, 3, 9, 5, 24, (3, 3, 3), (1, 1, 1)),
    # patch-per-thread stride-2 backward: odd sizes (pad_before = 1) and partial patches
    (1, 5, 15, 15, 16, (1, 3, 3), (1, 2, 2)),
    (2, 6, 15, 13, 24, (3, 3, 3), (2, 2, 2)),
    (1, 45, 8, 8, 8, (3, 3, 3), (2, 2, 2)),
]


@pytest.mark.parametrize("B,T,H,W,Cc,k,s", POOL_CASES)
def test_maxpool_fwd_bwd(B, T, H, W, Cc, k, s):
    from flickering_adversarial_video_b200.engine import op_maxpool3d, op_maxpool3d_bwd
    g = torch.Generator(device="cuda").manual_seed(5)
    # distinct values avoid fp16 ties, which the kernel resolves first-match like torch
    x = torch.randn((B, T, H, W, Cc), generator=g, device="cuda").to(torch.float16)
    y, idx = op_maxpool3d(x, k, s)
    xr = x.float().permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    yr = F.max_pool3d(_same_pad(xr, k, s, float("-inf")), k, s)
    assert torch.equal(y.float(), yr.permute(0, 2, 3, 4, 1)), "maxpool forward mismatch"
    dy = torch.randn(y.shape, generator=g, device="cuda").to(torch.bfloat16)
    add = torch.randn(x.shape, generator=g, device="cuda").to(torch.bfloat16)
    dx = op_maxpool3d_bwd(dy, idx, tuple(x.shape), k, s, add=add, relu_src=x)
    (gx,) = torch.autograd.grad(yr, xr, dy.float().permute(0, 4, 1, 2, 3), retain_graph=True)
    (gabs,) = torch.autograd.grad(yr, xr, dy.float().abs().permute(0, 4, 1, 2, 3))
    ref = (gx.permute(0, 2, 3, 4, 1) + add.float()) * (x.float() > 0)
    # The kernels sum the windows an element wins in packed bf16 (one rounding per add, at most three adds per
    # separable stage): the error is bounded by 2^-7 of the MAGNITUDES routed to the element, not of their (possibly
    # cancelling) sum.  With random inputs an element wins up to 27 windows; in the network it rarely wins more than one.
    mag = gabs.permute(0, 2, 3, 4, 1) + add.float().abs()
    _check(dx, ref, f"maxpool bwd {k}/{s}", rtol=2 ** -7, atol=3e-2, scale=mag)


def test_maxpool3_ties_route_like_torch():
    """post-ReLU inputs are full of exact ties (zeros and repeated fp16 values): the separable
    first-wins stages must route each window's gradient to the same element torch's scan picks."""
    from flickering_adversarial_video_b200.engine import op_maxpool3d, op_maxpool3d_bwd
    g = torch.Generator(device="cuda").manual_seed(11)
    k, s = (3, 3, 3), (1, 1, 1)
    # coarse quantisation -> many equal positive values inside every window
    x = (torch.randn((1, 10, 14, 14, 64), generator=g, device="cuda").clamp_min(0) * 4).round().div(4).to(torch.float16)
    y, idx = op_maxpool3d(x, k, s)
    xr = x.float().permute(0, 4, 1, 2, 3).clone().requires_grad_(True)
    yr = F.max_pool3d(_same_pad(xr, k, s, float("-inf")), k, s)
    assert torch.equal(y.float(), yr.permute(0, 2, 3, 4, 1))
    dy = torch.randn(y.shape, generator=g, device="cuda").to(torch.bfloat16)
    dx = op_maxpool3d_bwd(dy, idx, tuple(x.shape), k, s, add=None, relu_src=None)
    (gx,) = torch.autograd.grad(yr, xr, dy.float().permute(0, 4, 1, 2, 3), retain_graph=True)
    (gabs,) = torch.autograd.grad(yr, xr, dy.float().abs().permute(0, 4, 1, 2, 3))
    _check(dx, gx.permute(0, 2, 3, 4, 1), "maxpool3 ties", rtol=2 ** -7, atol=3e-2, scale=gabs.permute(0, 2, 3, 4, 1))
