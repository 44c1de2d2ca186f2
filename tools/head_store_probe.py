"""head conv store cost: the (B,na,H,W,no) fp32 layout (out_kind 1) against a pixel-major fp32 tensor with a 256-float
pitch (out_kind 2) on the P3/P4/P5 head shapes of the detect config (bs 32 = a quarter of the batch)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov5m_b200 import _lib
L = _lib.lib(); st = _lib.stream()
B = int(os.environ.get("BS", "32"))
for (H, Cin) in ((160, 192), (80, 384), (40, 768)):
    x = torch.randn(B, H, H, Cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(256, Cin, device="cuda") / Cin ** 0.5).to(torch.bfloat16)
    w[255:] = 0
    bias = torch.randn(256, device="cuda")
    for kind, cout, shape in ((1, 255, (B, 3, H, H, 85)), (2, 256, (B, H, H, 256))):
        y = torch.empty(shape, device="cuda", dtype=torch.float32)
        rows = ctypes.c_int(0)
        def run():
            _lib.check(L.yb_conv2d_fwd(x.data_ptr(), B, H, H, Cin, _lib.c_i64(Cin), w.data_ptr(), cout, 1, 1, y.data_ptr(),
                                       _lib.c_i64(cout), kind, None, bias.data_ptr(), 0, None, _lib.c_i64(0), None,
                                       ctypes.byref(rows), 3, 85, st))
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        gb = (x.numel() * 2 + y.numel() * 4) / 1e9
        print(f"H={H} Cin={Cin} kind={kind}: {us:8.1f} us  {gb / us * 1e6:7.0f} GB/s")
