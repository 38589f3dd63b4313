"""tcgen05 implicit-GEMM convolution (csrc/rcf_conv64.cu) against ATen in fp64 (GPU box only)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {3: 3e-5, 2: 4e-3, 1: 8e-3}     # rel-L2 against the fp64 result, per number of bf16 products


@pytest.fixture(scope="module")
def c64():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import rcf_unsupvideoseg_b200 as pkg
    pkg.load_library()
    from rcf_unsupvideoseg_b200 import conv64
    return conv64


def _rel(a, b):
    return float((a.double() - b).norm() / b.norm())


@pytest.mark.parametrize("N,H,W", [(1, 8, 8), (2, 19, 23), (3, 48, 48), (2, 96, 96), (1, 70, 130), (1, 3, 200), (5, 33, 7)])
@pytest.mark.parametrize("nprod", [3, 2, 1])
def test_conv64_forward_and_data_gradient_vs_fp64(c64, N, H, W, nprod):
    from rcf_unsupvideoseg_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + H * 10 + W)
    x = torch.randn(N, 64, H, W, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24).cuda()
    gy = torch.randn(N, 64, H, W, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    y = c64.conv64_raw(x, c64.pack_weights(w, False), nprod)
    dx = c64.conv64_raw(gy, c64.pack_weights(w, True), nprod)
    torch.cuda.synchronize()
    assert _lib.load_library().rcf_debug_conv64_status() == 0, "barrier time-out inside the tcgen05 kernel"
    xd = x.double().requires_grad_(True)
    yd = F.conv2d(xd, w.double(), None, 1, 1)
    (dxd,) = torch.autograd.grad(yd, xd, gy.double())
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert _rel(y, yd) <= TOL[nprod], (_rel(y, yd), nprod)
    assert _rel(dx, dxd) <= TOL[nprod], (_rel(dx, dxd), nprod)


WTOL = {3: 5e-5, 2: 5e-3, 1: 1e-2}


@pytest.mark.parametrize("N,H,W", [(1, 8, 8), (2, 19, 23), (3, 48, 48), (2, 96, 96), (1, 70, 130), (1, 3, 200), (5, 33, 7)])
@pytest.mark.parametrize("nprod", [3, 2, 1])
def test_conv64_weight_gradient_vs_fp64(c64, N, H, W, nprod):
    from rcf_unsupvideoseg_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(N * 999 + H * 10 + W)
    x = torch.randn(N, 64, H, W, generator=g).cuda()
    gy = torch.randn(N, 64, H, W, generator=g).cuda()
    dw = c64.conv64_wgrad_raw(x, gy, nprod)
    torch.cuda.synchronize()
    assert _lib.load_library().rcf_debug_conv64_status() == 0, "barrier time-out inside the tcgen05 kernel"
    wd = torch.zeros(64, 64, 3, 3, dtype=torch.float64, device="cuda", requires_grad=True)
    (dwd,) = torch.autograd.grad(F.conv2d(x.double(), wd, None, 1, 1), wd, gy.double())
    assert _rel(dw, dwd) <= WTOL[nprod], (_rel(dw, dwd), nprod)
    assert torch.equal(dw, c64.conv64_wgrad_raw(x, gy, nprod))          # fixed-order reduction


def test_conv64_autograd_function(c64):
    g = torch.Generator(device="cpu").manual_seed(7)
    x = torch.randn(2, 64, 21, 35, generator=g).cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(64, 64, 3, 3, generator=g) / 24).cuda().requires_grad_(True)
    gy = torch.randn(2, 64, 21, 35, generator=g).cuda()
    y = c64.conv64(x, w, 3)
    dx, dw = torch.autograd.grad(y, (x, w), gy)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    dxd, dwd = torch.autograd.grad(F.conv2d(xd, wd, None, 1, 1), (xd, wd), gy.double())
    assert _rel(y, F.conv2d(xd, wd, None, 1, 1).detach()) <= TOL[3]
    assert _rel(dx, dxd) <= TOL[3]
    assert _rel(dw, dwd) <= WTOL[3]


def test_conv64_one_cta_and_cta_pair_kernels_agree(c64):
    """RCF_OPT_CONV64_PAIR = 0 selects the one-CTA kernel (same MMA sequence per tile, other pipeline): identical results."""
    from rcf_unsupvideoseg_b200 import _lib
    lib = _lib.load_library()
    x = torch.randn(3, 64, 37, 61, device="cuda").contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, 64, 3, 3, device="cuda") / 24
    wp = c64.pack_weights(w, False)
    try:
        for nprod in (1, 2, 3):
            lib.rcf_debug_set_option(7, 1)
            a = c64.conv64_raw(x, wp, nprod)
            lib.rcf_debug_set_option(7, 0)
            b = c64.conv64_raw(x, wp, nprod)
            torch.cuda.synchronize()
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-6), nprod      # fp32 accumulation order inside the MMAs may differ between M = 128 and M = 256
    finally:
        lib.rcf_debug_set_option(7, 1)
    assert lib.rcf_debug_conv64_status() == 0


@pytest.mark.parametrize("N,H,W", [(1, 8, 8), (2, 19, 23), (2, 96, 96), (1, 70, 130)])
def test_conv64_fp16_operand_formats_vs_fp64(c64, N, H, W):
    """The TF32-class single-product mode of the head (rcf_head_* with nprod = 2): IEEE fp16 operands (11-bit significands)
    on both sides of every product; the gradient carries a power-of-two scale that the caller divides out again."""
    torch.manual_seed(N * 1000 + H)
    x = (torch.randn(N, 64, H, W, device="cuda") * 3).contiguous(memory_format=torch.channels_last)
    g = (torch.randn(N, 64, H, W, device="cuda") * 1e-7).contiguous(memory_format=torch.channels_last)     # loss-gradient magnitudes
    w = torch.randn(64, 64, 3, 3, device="cuda") / 24
    s = 2.0 ** 36                                                   # what rcf_grad_scale would pick for max |g| ~ 4e-7
    wf, wb = c64.pack_weights(w, False, f16=True), c64.pack_weights(w, True, f16=True)
    y = c64.conv64_pair(x.half(), None, wf, 1, a_f16=True, w_f16=True)
    dx = c64.conv64_pair((g * s).half(), None, wb, 1, a_f16=True, w_f16=True) / s
    dw = c64.conv64_wgrad_pair(x.half(), None, (g * s).half(), None, 1, f16=True) / s
    torch.cuda.synchronize()
    y_ref = F.conv2d(x.double(), w.double(), padding=1)
    dx_ref = torch.nn.grad.conv2d_input(x.shape, w.double(), g.double(), padding=1)
    dw_ref = torch.nn.grad.conv2d_weight(x.double(), w.shape, g.double(), padding=1)
    e = (_rel(y, y_ref), _rel(dx, dx_ref), _rel(dw, dw_ref))
    assert max(e) < 1e-3, e                                         # 2^-11 per operand
    y1 = c64.conv64_pair(x.bfloat16(), None, c64.pack_weights(w, False), 1)
    assert e[0] < 0.5 * _rel(y1, y_ref)                             # and well below the plain bf16 product
    from rcf_unsupvideoseg_b200 import _lib
    assert _lib.load_library().rcf_debug_conv64_status() == 0


def test_conv64_pack_both_orientations_in_one_launch(c64):
    """transpose_flip = 2 writes the forward image followed by the data-gradient image: byte-identical to the two single packs."""
    w = torch.randn(64, 64, 3, 3, device="cuda") / 24
    f, b = c64.pack_weights_both(w)
    torch.cuda.synchronize()
    assert torch.equal(f, c64.pack_weights(w, False)) and torch.equal(b, c64.pack_weights(w, True))


def test_stem_tf32_mode_vs_fp64():
    """nprod < 3 on the stem entry points = one round-to-nearest TF32 product (the allow_tf32 / autocast policy): errors at
    TF32 level (2^-11 per operand), against 1e-6 for the 3xTF32 mode."""
    from rcf_unsupvideoseg_b200 import fused_head, stem
    torch.manual_seed(3)
    B, H, W = 2, 37, 53
    fl = [torch.randn(B, 2, H, W, device="cuda") * 6 for _ in range(2)]
    w = torch.randn(64, 2, 3, 3, device="cuda") / 4
    b = torch.randn(64, device="cuda") / 4
    x = torch.cat([f.clamp(-20, 20) for f in fl], 0).double()
    pre = torch.nn.functional.conv2d(x, w.double(), b.double(), padding=1)
    ref = torch.nn.functional.leaky_relu(pre, 0.1)
    g = torch.randn(2 * B, 64, H, W, device="cuda").contiguous(memory_format=torch.channels_last)
    errs = {}
    for nprod in (3, 1):
        hi, lo, sign = fused_head.stem_forward_pair(fl, w, b, 20.0, 0.1, want_lo=True, nprod=nprod)
        act = hi.float() + lo.float()
        dw, db = stem.stem_backward_raw(fl, tuple(w.shape), 20.0, 0.1, None, sign, g, nprod=nprod)
        torch.cuda.synchronize()
        # the backward is checked for the LeakyReLU mask its own forward produced: pre-activations within the forward's
        # rounding error of zero (3e-4 of the elements in TF32) take the other branch, in cuDNN's TF32 kernels as well
        gpre = g.double() * torch.where(act > 0, 1.0, 0.1)
        dw_ref = torch.nn.grad.conv2d_weight(x, w.shape, gpre, padding=1)
        db_ref = gpre.sum((0, 2, 3))
        errs[nprod] = (_rel(act, ref), _rel(dw, dw_ref), _rel(db, db_ref))
    assert max(errs[3]) < 2e-5, errs          # activation limited by the bf16 pair (2^-17), gradients fp32-grade
    assert max(errs[1]) < 2e-3 and errs[1][1] > errs[3][1], errs


def test_conv64_is_bit_reproducible(c64):
    x = torch.randn(2, 64, 40, 52, device="cuda").contiguous(memory_format=torch.channels_last)
    w = torch.randn(64, 64, 3, 3, device="cuda") / 24
    wp = c64.pack_weights(w, False)
    a = c64.conv64_raw(x, wp, 2)
    b = c64.conv64_raw(x, wp, 2)
    assert torch.equal(a, b)
