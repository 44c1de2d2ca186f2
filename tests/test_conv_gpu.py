"""tcgen05 implicit-GEMM conv (forward + dgrad) vs a torch fp32 reference of the same op.

The kernel is a floating-point kernel, so the checker is torch's fp32 conv2d evaluated on
the SAME bf16-rounded operands; tolerance 2e-3 relative L2 (bf16 output rounding ~ 2^-9)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from yolov5m_b200 import _lib

pytestmark = pytest.mark.gpu
TOL = 2e-3


def rel(a, b):
    a = a.double(); b = b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def pack_fwd(w):  # OIHW -> [Cout][kh*kw][Cin]
    return w.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def pack_dgrad(w):  # OIHW -> [Cin][kh*kw][Cout]
    return w.permute(1, 2, 3, 0).contiguous().to(torch.bfloat16)


def conv_fwd(x_nhwc, wp, ks, stride, cout, out_kind=0, scale=None, shift=None, act=0, addend=None, stats=False,
             pitch_extra=0):
    L = _lib.lib()
    N, H, W, Cin = x_nhwc.shape
    Ho, Wo = H // stride, W // stride
    if out_kind == 0:
        ybuf = torch.full((N, Ho, Wo, cout + pitch_extra), 7.0, device="cuda", dtype=torch.bfloat16)
        y = ybuf[..., :cout]
        pitch = cout + pitch_extra
    elif out_kind == 2:
        ybuf = torch.zeros((N, Ho, Wo, cout), device="cuda", dtype=torch.float32); y = ybuf; pitch = cout
    else:
        ybuf = torch.zeros((N, 3, Ho, Wo, cout // 3), device="cuda", dtype=torch.float32); y = ybuf; pitch = cout
    st = None
    rows = ctypes.c_int(0)
    if stats:
        st = torch.zeros(L.yb_conv_max_partials(), 2, cout, device="cuda", dtype=torch.float32)
    _lib.check(L.yb_conv2d_fwd(_lib.ptr(x_nhwc), N, H, W, Cin, _lib.c_i64(x_nhwc.stride(2)), _lib.ptr(wp), cout, ks,
                               stride, _lib.ptr(ybuf), _lib.c_i64(pitch), out_kind, _lib.ptr(scale), _lib.ptr(shift),
                               act, _lib.ptr(addend), _lib.c_i64(addend.stride(2) if addend is not None else 0),
                               _lib.ptr(st), ctypes.byref(rows), 3, cout // 3 if out_kind == 1 else 85,
                               _lib.stream()))
    torch.cuda.synchronize()
    if stats:
        return y, ybuf, st[: rows.value].sum(0)
    return y, ybuf


CASES = [
    # N, H, W, Cin, Cout, ks, stride
    (2, 16, 16, 64, 64, 1, 1),
    (2, 16, 16, 64, 64, 3, 1),
    (1, 8, 16, 192, 192, 3, 1),
    (3, 20, 20, 384, 384, 3, 1),    # 20x20: multi-image patch
    (2, 40, 40, 192, 96, 1, 1),
    (2, 24, 40, 96, 96, 3, 1),      # KC=32 (SWIZZLE_64B)
    (2, 32, 32, 48, 96, 3, 2),      # KC=16 (SWIZZLE_32B), stride 2
    (2, 32, 32, 48, 48, 1, 1),
    (1, 32, 32, 16, 48, 3, 1),      # stem after space-to-depth
    (2, 16, 16, 384, 768, 3, 2),    # N split into 3 x 256
    (5, 22, 22, 192, 384, 3, 2),    # odd output grid 11x11, partial tiles
    (2, 10, 14, 768, 384, 1, 1),
    (1, 20, 20, 1536, 768, 1, 1),
    (70, 4, 4, 64, 32, 3, 1),       # many images per tile, N tail
]


@pytest.mark.parametrize("case", CASES)
def test_conv_fwd_raw(case):
    N, H, W, Cin, Cout, ks, s = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, ks, ks, generator=g) / (Cin * ks * ks) ** 0.5).to(torch.bfloat16)
    ref = F.conv2d(x.float(), w.float(), None, s, ks // 2)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    y, ybuf, st = conv_fwd(xd, pack_fwd(w.float()).cuda(), ks, s, Cout, stats=True, pitch_extra=8)
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert rel(got, ref) < TOL
    assert torch.all(ybuf[..., Cout:].float() == 7.0)  # slice write must not touch neighbours
    m = ref.numel() // Cout
    assert rel(st[0].cpu() / m, ref.mean((0, 2, 3))) < 1e-3 or (st[0].cpu() / m - ref.mean((0, 2, 3))).abs().max() < 1e-4
    assert rel(st[1].cpu() / m, (ref * ref).mean((0, 2, 3))) < 1e-3


def test_conv_fwd_fused_epilogue():
    N, H, W, Cin, Cout = 2, 16, 24, 96, 96
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(torch.bfloat16)
    sc = torch.rand(Cout, generator=g) + 0.5
    sh = torch.randn(Cout, generator=g) * 0.1
    res = torch.randn(N, Cout, H, W, generator=g).to(torch.bfloat16)
    ref = F.silu(F.conv2d(x.float(), w.float(), None, 1, 1) * sc[None, :, None, None] + sh[None, :, None, None]) + res.float()
    y, _ = conv_fwd(x.permute(0, 2, 3, 1).contiguous().cuda(), pack_fwd(w.float()).cuda(), 3, 1, Cout,
                    scale=sc.cuda(), shift=sh.cuda(), act=1, addend=res.permute(0, 2, 3, 1).contiguous().cuda())
    assert rel(y.float().cpu().permute(0, 3, 1, 2), ref) < TOL


@pytest.mark.parametrize("cin,hw", [(192, (16, 24)), (768, (4, 6))])
def test_conv_head_layout(cin, hw):
    N, (H, W) = 2, hw
    g = torch.Generator().manual_seed(5)
    x = torch.randn(N, cin, H, W, generator=g).to(torch.bfloat16)
    w = (torch.randn(255, cin, 1, 1, generator=g) / cin ** 0.5).to(torch.bfloat16)
    b = torch.randn(255, generator=g)
    ref = F.conv2d(x.float(), w.float(), b).view(N, 3, 85, H, W).permute(0, 1, 3, 4, 2).contiguous()
    y, _ = conv_fwd(x.permute(0, 2, 3, 1).contiguous().cuda(), pack_fwd(w.float()).cuda(), 1, 1, 255, out_kind=1,
                    shift=b.cuda())
    assert y.shape == ref.shape
    assert rel(y.cpu(), ref) < 1e-4  # fp32 output: only accumulation-order differences


DG = [
    (2, 16, 16, 64, 64, 1, 1),
    (2, 16, 16, 64, 128, 3, 1),
    (2, 24, 40, 96, 48, 3, 1),
    (3, 20, 20, 384, 384, 3, 1),
    (2, 32, 32, 48, 96, 3, 2),
    (2, 16, 16, 384, 768, 3, 2),
    (5, 22, 22, 192, 384, 3, 2),
]


@pytest.mark.parametrize("case", DG)
def test_conv_dgrad(case):
    N, H, W, Cin, Cout, ks, s = case
    g = torch.Generator().manual_seed(hash(case) % 1000 + 1)
    w = (torch.randn(Cout, Cin, ks, ks, generator=g) / (Cout * ks * ks) ** 0.5).to(torch.bfloat16)
    dy = torch.randn(N, Cout, H // s, W // s, generator=g).to(torch.bfloat16)
    ref = torch.nn.grad.conv2d_input((N, Cin, H, W), w.float(), dy.float(), s, ks // 2)
    L = _lib.lib()
    dyd = dy.permute(0, 2, 3, 1).contiguous().cuda()
    dx = torch.full((N, H, W, Cin), 3.0, device="cuda", dtype=torch.bfloat16)
    add = None
    if s == 1:
        add_c = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16)
        add = add_c.permute(0, 2, 3, 1).contiguous().cuda()
        ref = ref + add_c.float()
    _lib.check(L.yb_conv2d_dgrad(_lib.ptr(dyd), N, H, W, Cout, _lib.c_i64(Cout), _lib.ptr(pack_dgrad(w.float()).cuda()),
                                 Cin, ks, s, _lib.ptr(dx), _lib.c_i64(Cin), _lib.ptr(add), _lib.c_i64(Cin),
                                 0, _lib.stream()))
    torch.cuda.synchronize()
    assert rel(dx.float().cpu().permute(0, 3, 1, 2), ref) < TOL


# ---- halo-patch kernel (csrc/conv_patch.cu) forced on: same checks, every super-tile shape / stride / epilogue option
PATCH_FWD = [
    # N, H, W, Cin, Cout, ks, stride
    (2, 32, 32, 48, 48, 3, 1),     # BLOCK_N 48: 2x2 super-tile, one padded K chunk
    (1, 64, 48, 16, 48, 3, 1),     # stem-like: 16 real channels in a 64-wide chunk
    (2, 32, 40, 96, 96, 3, 1),     # BLOCK_N 96: two tiles per weight stage, 2 K chunks (second half padded)
    (2, 24, 40, 192, 192, 3, 1),   # single tile, H not a multiple of 16
    (1, 20, 20, 384, 384, 3, 1),   # partial tiles in both directions, 2 N tiles
    (2, 64, 64, 48, 96, 3, 2),     # stride 2: four parity patches
    (3, 22, 22, 192, 384, 3, 2),   # odd 11x11 output grid
    (2, 32, 32, 48, 48, 1, 1),     # 1x1 through the patch kernel: four tiles per weight stage
    (2, 32, 48, 192, 96, 1, 1),
    (1, 40, 40, 384, 192, 1, 1),
]


@pytest.fixture
def patch_mode():
    L = _lib.lib()
    L.yb_set_conv_patch_mode(2)
    yield
    L.yb_set_conv_patch_mode(0)


@pytest.mark.parametrize("case", PATCH_FWD)
def test_conv_patch_fwd(case, patch_mode):
    test_conv_fwd_raw(case)


def test_conv_patch_fused_epilogue(patch_mode):
    test_conv_fwd_fused_epilogue()


PATCH_DG = [
    (2, 32, 32, 48, 48, 1, 1),
    (2, 32, 48, 96, 192, 1, 1),
    (2, 32, 32, 48, 48, 3, 1),
    (2, 32, 40, 96, 48, 3, 1),
    (2, 24, 40, 192, 192, 3, 1),
    (1, 20, 20, 384, 384, 3, 1),
    (2, 64, 64, 48, 96, 3, 2),
    (2, 32, 32, 96, 192, 3, 2),
    (3, 22, 22, 192, 384, 3, 2),
]


@pytest.mark.parametrize("case", PATCH_DG)
def test_conv_patch_dgrad(case, patch_mode):
    test_conv_dgrad(case)


# ---- ks code 31: 3x1 kernel (three vertical taps, stride 1) -- the stem after horizontal tap gathering
@pytest.mark.parametrize("N,H,W,Cin,Cout,patch", [(2, 32, 32, 48, 48, 0), (2, 32, 32, 48, 48, 2), (1, 24, 20, 48, 96, 0)])
def test_conv_fwd_3x1(N, H, W, Cin, Cout, patch):
    L = _lib.lib()
    L.yb_set_conv_patch_mode(patch)
    try:
        g = torch.Generator().manual_seed(21)
        x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16)
        w = (torch.randn(Cout, Cin, 3, 1, generator=g) / (Cin * 3) ** 0.5).to(torch.bfloat16)
        ref = F.conv2d(x.float(), w.float(), None, 1, (1, 0))
        y, _, st = conv_fwd(x.permute(0, 2, 3, 1).contiguous().cuda(), pack_fwd(w.float()).cuda(), 31, 1, Cout, stats=True)
        assert rel(y.float().cpu().permute(0, 3, 1, 2), ref) < TOL
        m = ref.numel() // Cout
        assert rel(st[1].cpu() / m, (ref * ref).mean((0, 2, 3))) < 1e-3
    finally:
        L.yb_set_conv_patch_mode(0)


@pytest.mark.parametrize("N,H,W,Cout,patch", [(2, 32, 32, 48, 0), (2, 32, 32, 48, 2), (1, 24, 20, 96, 0), (2, 64, 40, 48, 0)])
def test_conv_fwd_3x1_stem_view(N, H, W, Cout, patch):
    """ks 31 with x_pitch (16) < Cin (48): the row-padded 16-channel staging of yb_prep_input read through an
    overlapping-window tensor map -- pixel w's channels are padded columns w, w+1, w+2 (include/yolov5m_b200.h)."""
    L = _lib.lib()
    L.yb_set_conv_patch_mode(patch)
    try:
        g = torch.Generator().manual_seed(23)
        xs = torch.zeros(N, H, W + 2, 16)
        xs[:, :, 1:-1] = torch.randn(N, H, W, 16, generator=g)
        xs = xs.to(torch.bfloat16)
        view = xs.flatten(2).unfold(2, 48, 16)                    # (N, H, W, 48), pixel stride 16
        assert view.shape == (N, H, W, 48) and view.stride(2) == 16
        w = (torch.randn(Cout, 48, 3, 1, generator=g) / 12.0).to(torch.bfloat16)
        ref = F.conv2d(view.float().permute(0, 3, 1, 2), w.float(), None, 1, (1, 0))
        y, _, st = conv_fwd(xs.cuda().flatten(2).unfold(2, 48, 16), pack_fwd(w.float()).cuda(), 31, 1, Cout, stats=True)
        assert rel(y.float().cpu().permute(0, 3, 1, 2), ref) < TOL
        m = ref.numel() // Cout
        assert rel(st[1].cpu() / m, (ref * ref).mean((0, 2, 3))) < 1e-3
    finally:
        L.yb_set_conv_patch_mode(0)
