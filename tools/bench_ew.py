"""Achieved HBM bandwidth of the BN/SiLU passes (elementwise.cu) on the activation shapes of YOLOV5m at 640x640, bs=64.

    python tools/bench_ew.py [--bs 64]
Algorithmic bytes: fwd = read y + write a (2 x 2 B/elem); bwd_reduce = read da + y; bwd_apply = read da + y, write dy.
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov5m_b200 import _lib  # noqa: E402

SHAPES = [(48, 320, 1), (96, 160, 2), (48, 160, 6), (192, 80, 3), (96, 80, 16), (384, 40, 5), (192, 40, 27),
          (768, 20, 5), (384, 20, 16)]  # C, H, count (approximate layer census)


def time_it(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bs", type=int, default=64)
    a = ap.parse_args()
    L = _lib.lib()
    st = _lib.stream()
    tot = {"fwd": [0, 0], "reduce": [0, 0], "apply": [0, 0]}
    for C, H, cnt in SHAPES:
        B = a.bs
        npix = B * H * H
        y = torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16)
        out = torch.empty_like(y)
        da = torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16)
        dy = torch.empty_like(y)
        sc = torch.rand(C, device="cuda") + 0.5
        sh = torch.randn(C, device="cuda") * 0.1
        mu = torch.randn(C, device="cuda") * 0.1
        iv = torch.rand(C, device="cuda") + 0.5
        coef = torch.randn(2 * C, device="cuda") * 0.01
        part = torch.zeros(L.yb_bwd_reduce_max_rows() * 2 * C, device="cuda")
        rows = ctypes.c_int(0)
        n = npix * C
        t_f = time_it(lambda: L.yb_bn_act_fwd(y.data_ptr(), C, B, H, H, C, sc.data_ptr(), sh.data_ptr(), None, 0,
                                              out.data_ptr(), C, None, 0, st))
        t_r = time_it(lambda: L.yb_bn_act_bwd_reduce(da.data_ptr(), C, y.data_ptr(), C, npix, C, sc.data_ptr(), sh.data_ptr(),
                                                     mu.data_ptr(), iv.data_ptr(), part.data_ptr(), ctypes.byref(rows), st))
        t_a = time_it(lambda: L.yb_bn_act_bwd_apply(da.data_ptr(), C, y.data_ptr(), C, npix, C, sc.data_ptr(), sh.data_ptr(),
                                                    mu.data_ptr(), iv.data_ptr(), coef.data_ptr(), dy.data_ptr(), C, st))
        r = dict(C=C, H=H, count=cnt, MB=n * 2 / 1e6, fwd_us=t_f * 1e6, fwd_GBs=4 * n / t_f / 1e9, reduce_us=t_r * 1e6,
                 reduce_GBs=4 * n / t_r / 1e9, apply_us=t_a * 1e6, apply_GBs=6 * n / t_a / 1e9)
        print(json.dumps({k: round(v, 1) if isinstance(v, float) else v for k, v in r.items()}), flush=True)
        for k, t, b in (("fwd", t_f, 4 * n), ("reduce", t_r, 4 * n), ("apply", t_a, 6 * n)):
            tot[k][0] += t * cnt
            tot[k][1] += b * cnt
    print(json.dumps({k: dict(ms=v[0] * 1e3, GBs=v[1] / v[0] / 1e9) for k, v in tot.items()}))


if __name__ == "__main__":
    main()
