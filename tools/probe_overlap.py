"""Probe: can a tensor-bound weight-gradient kernel (512 threads x 48 registers) run concurrently with the HBM-bound
BN-backward passes on the same SMs?  Launches `nw` wgrad kernels on one stream and reduce + apply passes on another and
compares the elapsed time with the same work on ONE stream.  YB_WGRAD_SMEM_KB caps the wgrad operand ring (plan time).

    [YB_WGRAD_SMEM_KB=190] python tools/probe_overlap.py
"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov5m_b200 import _lib  # noqa: E402


def main():
    L = _lib.lib()
    B = 64
    # wgrad 192 -> 192 3x3 @40x40 (tensor-bound, ~60 us)
    cin = cout = 192
    x = torch.randn(B, 40, 40, cin, device="cuda").to(torch.bfloat16)
    dyw = torch.randn(B, 40, 40, cout, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(cout, 9, cin, device="cuda")
    ws = torch.empty(64 << 20, device="cuda", dtype=torch.float32)
    wp = _lib.checkp(L.yb_conv_wgrad_plan(x.data_ptr(), B, 40, 40, cin, cin, dyw.data_ptr(), cout, cout, 3, 1, ws.data_ptr(),
                                          ws.numel(), 0))
    # BN backward on 96 channels @160x160 (HBM-bound, ~125 + ~170 us)
    C, npix = 96, B * 160 * 160
    da = torch.randn(npix, C, device="cuda").to(torch.bfloat16)
    y = torch.randn(npix, C, device="cuda").to(torch.bfloat16)
    dy = torch.empty_like(y)
    par = [torch.rand(C, device="cuda") + 0.5 for _ in range(4)]
    coef = torch.zeros(2, C, device="cuda")
    part = torch.zeros(L.yb_bwd_reduce_max_rows(), 2, C, device="cuda")
    rows = ctypes.c_int(0)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def wgrads(st, n):
        for _ in range(n):
            _lib.check(L.yb_wgrad_plan_run(wp, dw.data_ptr(), cout, None, 0, st.cuda_stream))

    def bn_bwd(st):
        _lib.check(L.yb_bn_act_bwd_reduce(da.data_ptr(), C, y.data_ptr(), C, npix, C, par[0].data_ptr(), par[1].data_ptr(),
                                          par[2].data_ptr(), par[3].data_ptr(), part.data_ptr(), ctypes.byref(rows),
                                          st.cuda_stream))
        _lib.check(L.yb_bn_act_bwd_apply(da.data_ptr(), C, y.data_ptr(), C, npix, C, par[0].data_ptr(), par[1].data_ptr(),
                                         par[2].data_ptr(), par[3].data_ptr(), coef.data_ptr(), dy.data_ptr(), C,
                                         st.cuda_stream))

    def timed(fn, iters=20):
        cur = torch.cuda.current_stream()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for _ in range(iters):
            s1.wait_stream(cur); s2.wait_stream(cur)
            fn()
            cur.wait_stream(s1); cur.wait_stream(s2)
        e1.record(cur)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    nw = 4
    out = {"smem_kb": os.environ.get("YB_WGRAD_SMEM_KB", "227"),
           "wgrad_only_us": timed(lambda: wgrads(s1, nw)),
           "bn_bwd_only_us": timed(lambda: bn_bwd(s1)),
           "serial_us": timed(lambda: (wgrads(s1, nw), bn_bwd(s1))),
           "two_streams_wgrad_first_us": timed(lambda: (wgrads(s1, nw), bn_bwd(s2))),
           "two_streams_bn_first_us": timed(lambda: (bn_bwd(s2), wgrads(s1, nw)))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
