"""Per-layer timing of the tcgen05 conv kernels (fwd with BN-stat partials, dgrad, wgrad) on the distinct
conv shapes of YOLOV5m at 640x640.  CUDA events, inputs larger than L2 at bs=64.

    python tools/bench_conv.py [--bs 64] [--iters 10] [--only fwd,dgrad,wgrad]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolov5m_b200 import _lib  # noqa: E402

SHAPES = [  # Cin, Cout, k, stride, Hout, count;  k = 31: the stem, a 3x1 conv over the window view of the row-padded
            # 16-channel space-to-depth staging (Cin = 48 per pixel, pitch 16; include/yolov5m_b200.h: yb_prep_input)
    (48, 48, 31, 1, 320, 1), (48, 96, 3, 2, 160, 1), (96, 48, 1, 1, 160, 2), (48, 48, 1, 1, 160, 2),
    (48, 48, 3, 1, 160, 2), (96, 96, 1, 1, 160, 1), (96, 192, 3, 2, 80, 1), (192, 96, 1, 1, 80, 2),
    (96, 96, 1, 1, 80, 6), (96, 96, 3, 1, 80, 6), (192, 192, 1, 1, 80, 2), (384, 96, 1, 1, 80, 2),
    (192, 384, 3, 2, 40, 1), (384, 192, 1, 1, 40, 5), (192, 192, 1, 1, 40, 10), (192, 192, 3, 1, 40, 10),
    (384, 384, 1, 1, 40, 3), (768, 192, 1, 1, 40, 2), (192, 192, 3, 2, 40, 1), (384, 768, 3, 2, 20, 1),
    (768, 384, 1, 1, 20, 6), (384, 384, 1, 1, 20, 4), (384, 384, 3, 1, 20, 4), (768, 768, 1, 1, 20, 2),
    (1536, 768, 1, 1, 20, 1), (384, 384, 3, 2, 20, 1), (192, 256, 1, 1, 80, 1), (384, 256, 1, 1, 40, 1),
    (768, 256, 1, 1, 20, 1),
]


def _peaks():
    try:
        d = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        return float(d.get("bf16_tflops_sustained", 1383.4)), float(d.get("hbm_gbs", 6584.8))
    except Exception:
        return 1383.4, 6584.8


PEAK_TF, PEAK_HBM = _peaks()


def time_it(fn, iters):
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
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="fwd,dgrad,wgrad")
    ap.add_argument("--out", default=None)
    ap.add_argument("--shapes", default=None, help="comma-separated indices into SHAPES (default: all)")
    ap.add_argument("--patch", type=int, default=None, help="yb_set_conv_patch_mode (-1 generic only, 0 auto, 1 force)")
    a = ap.parse_args()
    L = _lib.lib()
    if a.patch is not None:
        L.yb_set_conv_patch_mode(a.patch)
    st = _lib.stream()
    ws = torch.empty(64 << 20, device="cuda", dtype=torch.float32)
    rows = []
    tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    totf = 0.0
    sel = SHAPES if a.shapes is None else [SHAPES[int(i)] for i in a.shapes.split(",")]
    for (cin, cout, k, s, ho, cnt) in sel:
        B, hin = a.bs, ho * s
        stem = k == 31
        taps = 3 if stem else k * k
        xpitch = 16 if stem else cin
        x = (torch.randn(B, hin, hin + 2, 16, device="cuda") if stem else torch.randn(B, hin, hin, cin, device="cuda")).to(torch.bfloat16)
        y = torch.empty(B, ho, ho, cout, device="cuda", dtype=torch.bfloat16)
        w = (torch.randn(cout, taps, cin, device="cuda") * 0.05).to(torch.bfloat16)
        wt = (torch.randn(cin, taps, cout, device="cuda") * 0.05).to(torch.bfloat16)
        dw = torch.zeros(cout, taps, cin, device="cuda")
        stats = torch.zeros(L.yb_conv_max_partials(), 2, cout, device="cuda")
        # the stem's algorithmic work is the reference's 6x6 conv over 3 channels (model.py:184)
        flops = 2.0 * B * ho * ho * cout * (108 if stem else cin * k * k)
        r = dict(shape=f"3->{cout} k6 s2 (3x1 window view) @{ho}" if stem else f"{cin}->{cout} k{k} s{s} @{ho}", count=cnt,
                 gflop=flops / 1e9)
        # the layer's own floor: max(FLOPs / measured bf16 peak, activation bytes in + out / measured copy bandwidth)
        nbytes = 2.0 * (B * hin * hin * (16 if stem else cin) + B * ho * ho * cout)
        r["floor_us"] = max(flops / (PEAK_TF * 1e12), nbytes / (PEAK_HBM * 1e9)) * 1e6
        r["bound"] = "tensor" if flops / (PEAK_TF * 1e12) > nbytes / (PEAK_HBM * 1e9) else "hbm"
        nrows = ctypes.c_int(0)
        if "fwd" in a.only:
            p = _lib.checkp(L.yb_conv_fwd_plan(x.data_ptr(), B, hin, hin, cin, xpitch, w.data_ptr(), cout, k, s, y.data_ptr(),
                                               cout, 0, None, None, 0, None, 0, stats.data_ptr(), ctypes.byref(nrows), 3, 85))
            t = time_it(lambda: L.yb_plan_run(p, st), a.iters)
            r["fwd_us"] = t * 1e6; r["fwd_tf"] = flops / t / 1e12; tot["fwd"] += t * cnt
            L.yb_plan_destroy(p)
        if "dgrad" in a.only and not stem:
            p = _lib.checkp(L.yb_conv_dgrad_plan(y.data_ptr(), B, hin, hin, cout, cout, wt.data_ptr(), cin, k, s,
                                                 x.data_ptr(), cin, None, 0, 0))
            t = time_it(lambda: L.yb_plan_run(p, st), a.iters)
            r["dgrad_us"] = t * 1e6; r["dgrad_tf"] = flops / t / 1e12; tot["dgrad"] += t * cnt
            L.yb_plan_destroy(p)
        if "wgrad" in a.only:
            p = _lib.checkp(L.yb_conv_wgrad_plan(x.data_ptr(), B, hin, hin, cin, xpitch, y.data_ptr(), cout, cout, k, s,
                                                 ws.data_ptr(), ws.numel(), 0))
            t = time_it(lambda: L.yb_wgrad_plan_run(p, dw.data_ptr(), cout, None, 0, st), a.iters)
            r["wgrad_us"] = t * 1e6; r["wgrad_tf"] = flops / t / 1e12; tot["wgrad"] += t * cnt
            L.yb_plan_destroy(p)
        totf += flops * cnt
        rows.append(r)
        print(json.dumps({k_: (round(v, 1) if isinstance(v, float) else v) for k_, v in r.items()}), flush=True)
        del x, y, w, wt, dw
    floor_ms = sum(r["floor_us"] * r["count"] for r in rows) * 1e-3
    summ = {k_: dict(ms=v * 1e3, tflops=totf / v / 1e12 if v else None, floor_ms=floor_ms, frac_of_floor=floor_ms / (v * 1e3) if v else None)
            for k_, v in tot.items()}
    print(json.dumps(dict(bs=a.bs, total_gflop_per_pass=totf / 1e9, summary=summ)))
    if a.out:
        json.dump(dict(rows=rows, summary=summ, bs=a.bs), open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
