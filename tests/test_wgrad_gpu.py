"""tcgen05 weight-gradient conv vs torch fp32 conv2d_weight on the same bf16-rounded operands.
fp32 output; tolerance 2e-3 relative L2 (fp32 accumulation-order differences only are ~1e-6;
the bound leaves room for the split-K partial ordering)."""
import ctypes

import pytest
import torch

from yolov5m_b200 import _lib

pytestmark = pytest.mark.gpu
TOL = 2e-3


def rel(a, b):
    a = a.double(); b = b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


CASES = [
    # N, H, W, Cin, Cout, ks, stride
    (2, 16, 16, 64, 64, 1, 1),       # 128B swizzle on both operands
    (2, 16, 16, 64, 128, 3, 1),
    (2, 24, 40, 96, 96, 3, 1),       # 64B swizzle (32-channel boxes)
    (2, 32, 32, 48, 48, 3, 1),       # 32B swizzle, M tile mostly empty
    (2, 32, 32, 48, 96, 3, 2),       # stride 2: parity maps
    (3, 20, 20, 384, 384, 3, 1),     # N split 2 x 192, M 3 x 128
    (2, 16, 16, 384, 768, 3, 2),
    (5, 22, 22, 192, 384, 3, 2),     # odd 11x11 output grid
    (1, 20, 20, 1536, 768, 1, 1),    # N split 6 x 256
    (2, 32, 32, 16, 48, 3, 1),       # stem after space-to-depth
    (2, 40, 40, 192, 256, 1, 1),     # head: 255 -> 256 padded rows
    (70, 4, 4, 64, 32, 3, 1),
]


@pytest.mark.parametrize("case", CASES)
def test_conv_wgrad(case):
    N, H, W, Cin, Cout, ks, s = case
    g = torch.Generator().manual_seed(hash(case) % 1000 + 7)
    x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16)
    dy = torch.randn(N, Cout, H // s, W // s, generator=g).to(torch.bfloat16)
    ref = torch.nn.grad.conv2d_weight(x.float(), (Cout, Cin, ks, ks), dy.float(), s, ks // 2)
    L = _lib.lib()
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    dyd = dy.permute(0, 2, 3, 1).contiguous().cuda()
    ws = torch.empty(8 << 20, device="cuda", dtype=torch.float32)
    out_rows = Cout - 1 if Cout == 256 else Cout
    dw = torch.full((Cout, ks * ks, Cin), 3.0, device="cuda", dtype=torch.float32)
    for acc in (0, 1):
        _lib.check(L.yb_conv2d_wgrad(_lib.ptr(xd), N, H, W, Cin, _lib.c_i64(Cin), _lib.ptr(dyd), Cout, _lib.c_i64(Cout),
                                     ks, s, _lib.ptr(ws), _lib.c_i64(ws.numel()), 0, _lib.ptr(dw), out_rows, None, acc,
                                     _lib.stream()))
    torch.cuda.synchronize()
    got = dw.cpu().view(Cout, ks, ks, Cin).permute(0, 3, 1, 2)
    assert rel(got[:out_rows], 2 * ref[:out_rows]) < TOL  # second call accumulated
    if out_rows < Cout:
        assert torch.all(got[out_rows:] == 3.0)


def test_wgrad_index_map_and_pitch():
    """x and dy as channel slices of wider buffers; output scattered through an index map."""
    N, H, W, Cin, Cout, ks = 2, 16, 16, 32, 64, 3
    g = torch.Generator().manual_seed(1)
    xb = torch.randn(N, H, W, Cin + 32, generator=g).to(torch.bfloat16).cuda()
    dyb = torch.randn(N, H, W, Cout + 64, generator=g).to(torch.bfloat16).cuda()
    x = xb[..., 32:]; dy = dyb[..., 64:]
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, ks, ks), dy.float().permute(0, 3, 1, 2),
                                      1, 1).cpu()
    n = Cout * ks * ks * Cin
    perm = torch.randperm(n, generator=g).int()
    perm[::7] = -1
    L = _lib.lib()
    ws = torch.empty(4 << 20, device="cuda", dtype=torch.float32)
    dw = torch.zeros(n, device="cuda", dtype=torch.float32)
    _lib.check(L.yb_conv2d_wgrad(_lib.ptr(x), N, H, W, Cin, _lib.c_i64(Cin + 32), _lib.ptr(dy), Cout,
                                 _lib.c_i64(Cout + 64), ks, 1, _lib.ptr(ws), _lib.c_i64(ws.numel()), 0, _lib.ptr(dw),
                                 Cout, _lib.ptr(perm.cuda()), 0, _lib.stream()))
    torch.cuda.synchronize()
    packed = ref.permute(0, 2, 3, 1).reshape(-1)
    want = torch.zeros(n)
    m = perm >= 0
    want[perm[m].long()] = packed[m]
    assert rel(dw.cpu(), want) < TOL


# ---- halo-patch variant (csrc/conv_wgrad_patch.cu) forced on for every eligible shape
@pytest.fixture
def wpatch_mode():
    L = _lib.lib()
    L.yb_set_wgrad_patch_mode(1)
    yield
    L.yb_set_wgrad_patch_mode(0)


@pytest.mark.parametrize("case", [c for c in CASES if c[5] == 3 and c[6] == 1] + [(2, 32, 48, 192, 192, 3, 1), (1, 40, 40, 96, 256, 3, 1)])
def test_conv_wgrad_patch(case, wpatch_mode):
    test_conv_wgrad(case)


def test_wgrad_patch_index_map_and_pitch(wpatch_mode):
    test_wgrad_index_map_and_pitch()


def test_conv_wgrad_3x1():
    """ks code 31: 3x1 kernel (stem after horizontal tap gathering): dw [Cout][3][Cin]."""
    N, H, W, Cin, Cout = 2, 32, 32, 48, 48
    g = torch.Generator().manual_seed(9)
    x = torch.randn(N, Cin, H, W, generator=g).to(torch.bfloat16)
    dy = torch.randn(N, Cout, H, W, generator=g).to(torch.bfloat16)
    ref = torch.nn.grad.conv2d_weight(x.float(), (Cout, Cin, 3, 1), dy.float(), 1, (1, 0))
    L = _lib.lib()
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    dyd = dy.permute(0, 2, 3, 1).contiguous().cuda()
    ws = torch.empty(4 << 20, device="cuda", dtype=torch.float32)
    dw = torch.zeros(Cout, 3, Cin, device="cuda", dtype=torch.float32)
    _lib.check(L.yb_conv2d_wgrad(_lib.ptr(xd), N, H, W, Cin, _lib.c_i64(Cin), _lib.ptr(dyd), Cout, _lib.c_i64(Cout), 31, 1,
                                 _lib.ptr(ws), _lib.c_i64(ws.numel()), 0, _lib.ptr(dw), Cout, None, 0, _lib.stream()))
    torch.cuda.synchronize()
    assert rel(dw.cpu().view(Cout, 3, 1, Cin).permute(0, 3, 1, 2), ref) < TOL


def test_conv_wgrad_3x1_stem_view():
    """ks 31 with x_pitch (16) < Cin (48): x is the row-padded 16-channel staging read as an overlapping window."""
    N, H, W, Cout = 2, 32, 40, 48
    g = torch.Generator().manual_seed(10)
    xs = torch.zeros(N, H, W + 2, 16)
    xs[:, :, 1:-1] = torch.randn(N, H, W, 16, generator=g)
    xs = xs.to(torch.bfloat16)
    view = xs.flatten(2).unfold(2, 48, 16)
    dy = torch.randn(N, Cout, H, W, generator=g).to(torch.bfloat16)
    ref = torch.nn.grad.conv2d_weight(view.float().permute(0, 3, 1, 2), (Cout, 48, 3, 1), dy.float(), 1, (1, 0))
    L = _lib.lib()
    xd = xs.cuda()
    dyd = dy.permute(0, 2, 3, 1).contiguous().cuda()
    ws = torch.empty(4 << 20, device="cuda", dtype=torch.float32)
    dw = torch.zeros(Cout, 3, 48, device="cuda", dtype=torch.float32)
    _lib.check(L.yb_conv2d_wgrad(_lib.ptr(xd), N, H, W, 48, _lib.c_i64(16), _lib.ptr(dyd), Cout, _lib.c_i64(Cout), 31, 1,
                                 _lib.ptr(ws), _lib.c_i64(ws.numel()), 0, _lib.ptr(dw), Cout, None, 0, _lib.stream()))
    torch.cuda.synchronize()
    assert rel(dw.cpu().view(Cout, 3, 1, 48).permute(0, 3, 1, 2), ref) < TOL
