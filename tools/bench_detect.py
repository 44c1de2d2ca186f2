"""Detect path (BASELINE.json configs[4]): eval forward at 1280x1280 -> cells_to_bboxes -> non_max_suppression
(conf 0.25, iou 0.45, max_det 300) on synthetic images, bs=128 by default.  Two score distributions:
D1 "realistic": head objectness bias -5 (a few % of the 100,800 cells pass); D2 "adversarial": untrained head, every
cell passes.  Reports img/s per stage (CUDA events) and checks the NMS keep sets of a few images against the oracle.

    python tools/bench_detect.py [--bs 128] [--size 1280] [--iters 5] [--check 2]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolov5m_b200 as yb  # noqa: E402
from yolov5m_b200.boxes import nms_device  # noqa: E402


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bs", type=int, default=128)
    ap.add_argument("--size", type=int, default=1280)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--check", type=int, default=2, help="images whose keep set is compared with the CPU oracle")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = yb.YOLOV5m(first_out=yb.FIRST_OUT, nc=80, anchors=yb.ANCHORS, ch=(192, 384, 768)).to(dev).eval()
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 256, (a.bs, 3, a.size, a.size), dtype=torch.uint8, generator=g).to(dev)
    res = {"bs": a.bs, "size": a.size}
    for tag, bias in (("D2_untrained_all_pass", None), ("D1_obj_bias_-5", -5.0)):
        if bias is not None:
            with torch.no_grad():
                for conv in m.head.out_convs:
                    conv.bias.view(3, -1)[:, 4] = bias
        with torch.no_grad():
            t_f, out = timed(lambda: m(x), a.iters)
            t_d, dec = timed(lambda: yb.cells_to_bboxes(out, m.head.anchors, m.head.stride, is_pred=True, to_list=False), a.iters)
            t_n, (rows, counts) = timed(lambda: nms_device(dec, 0.45, 0.25, 300), a.iters)
        cand = int((dec[..., 1] > 0.25).sum().item())
        r = {"fwd_img_s": a.bs / t_f, "decode_img_s": a.bs / t_d, "nms_img_s": a.bs / t_n,
             "end_to_end_img_s": a.bs / (t_f + t_d + t_n), "ms": {"fwd": t_f * 1e3, "decode": t_d * 1e3, "nms": t_n * 1e3},
             "candidates_per_image": cand / a.bs, "kept_per_image": float(counts.float().mean().item())}
        if a.check:
            from oracle import nms_ref
            sub = dec[: a.check].cpu()
            ref, _ = nms_ref.non_max_suppression(sub, 0.45, 0.25, 300)
            ok = all(int(counts[i]) == len(ref[i]) and np.array_equal(rows[i, : len(ref[i])].cpu().numpy(), ref[i].astype(np.float32))
                     for i in range(a.check))
            r["keep_sets_bit_exact_vs_oracle"] = bool(ok)
        res[tag] = r
        print(json.dumps({tag: r}), flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
