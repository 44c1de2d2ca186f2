"""One launch of every hot-path kernel at a representative BASELINE shape, inside a cudaProfilerStart/Stop range, for

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_kernels \
        python tools/profile_kernels.py
    ncu -i gpurun_out/prof_kernels.ncu-rep --page raw --csv > gpurun_out/prof_kernels.raw.csv
    python tools/summarize_ncu.py gpurun_out/prof_kernels.raw.csv profiles/ncu_kernels_<tag>.json

Each kernel runs once un-profiled (warm-up: plans, allocator, L2 state irrelevant -- tensors exceed L2) and once profiled.
Shapes: bs=64, 640x640 training tensors; bs=128, 1280x1280 detect tensors (BASELINE.json configs[2] / [4]).
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import yolov5m_b200 as yb  # noqa: E402
from yolov5m_b200 import _lib  # noqa: E402
from yolov5m_b200.boxes import nms_device  # noqa: E402
from yolov5m_b200.trainer import Adam  # noqa: E402

L = _lib.lib()
JOBS = []


def job(fn):
    JOBS.append(fn)
    return fn


def main():
    dev = torch.device("cuda", 0)
    st = _lib.stream()
    g = torch.Generator(device="cuda").manual_seed(0)
    keep = []

    # ---- BN / SiLU passes (elementwise.cu) on (64, 80, 80, 192): 78.6 M elements
    B, H, C = 64, 80, 192
    npix = B * H * H
    y = torch.randn(B, H, H, C, device=dev, generator=g).to(torch.bfloat16)
    da = torch.randn(B, H, H, C, device=dev, generator=g).to(torch.bfloat16)
    out, dy = torch.empty_like(y), torch.empty_like(y)
    sc, sh = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
    mu, iv = torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev) + 0.5
    coef = torch.randn(2 * C, device=dev) * 0.01
    part = torch.zeros(L.yb_bwd_reduce_max_rows() * 2 * C, device=dev)
    rows = ctypes.c_int(0)
    job(lambda: L.yb_bn_act_fwd(y.data_ptr(), C, B, H, H, C, sc.data_ptr(), sh.data_ptr(), None, 0, out.data_ptr(), C, None, 0, st))
    job(lambda: L.yb_bn_act_bwd_reduce(da.data_ptr(), C, y.data_ptr(), C, npix, C, sc.data_ptr(), sh.data_ptr(), mu.data_ptr(),
                                       iv.data_ptr(), part.data_ptr(), ctypes.byref(rows), st))
    job(lambda: L.yb_bn_act_bwd_apply(da.data_ptr(), C, y.data_ptr(), C, npix, C, sc.data_ptr(), sh.data_ptr(), mu.data_ptr(),
                                      iv.data_ptr(), coef.data_ptr(), dy.data_ptr(), C, st))
    # ---- SPPF pooling on (64, 20, 20, 384), input staging of a 64 x 3 x 640 x 640 uint8 batch
    xp = torch.randn(64, 20, 20, 384, device=dev, generator=g).to(torch.bfloat16)
    yp, am = torch.empty_like(xp), torch.empty(64, 20, 20, 384, device=dev, dtype=torch.uint8)
    job(lambda: L.yb_maxpool5_fwd(xp.data_ptr(), 384, 64, 20, 20, 384, yp.data_ptr(), 384, am.data_ptr(), st))
    job(lambda: L.yb_maxpool5_bwd(yp.data_ptr(), 384, am.data_ptr(), 64, 20, 20, 384, xp.data_ptr(), 384, 0, st))
    img = torch.randint(0, 256, (64, 3, 640, 640), device=dev, dtype=torch.uint8)
    x16 = torch.empty(64, 320, 322, 16, device=dev, dtype=torch.bfloat16)
    job(lambda: L.yb_prep_input(img.data_ptr(), 1, 64, 640, 640, x16.data_ptr(), st))

    # ---- ComputeLoss kernels (loss.cu): bs=64 head tensors at 640x640, 512 targets, forward + backward
    class _Head:
        nc, nl, naxs, stride = 80, 3, 3, [8, 16, 32]
        anchors = (torch.tensor(yb.ANCHORS).float().view(3, 3, 2) / torch.tensor([8., 16., 32.]).view(3, 1, 1)).to(dev)

    class _M:
        head = _Head()

        def parameters(self):
            return iter([torch.nn.Parameter(torch.zeros(1, device=dev))])
    import recipes
    p = [t.to(dev).requires_grad_(True) for t in recipes.head_outputs(1, 64, 640, 640)]
    tg = recipes.targets(2, 64, 512)
    loss_fn = yb.ComputeLoss(_M())

    def loss_job():
        for t in p:
            t.grad = None
        loss_fn(p, tg, None).backward()
    job(loss_job)

    # ---- detect path (nms.cu): bs=128, 1280x1280 head tensors, ~2 % of the cells above conf 0.25
    pd = [torch.randn(128, 3, 1280 // s, 1280 // s, 85, device=dev, generator=g) for s in (8, 16, 32)]
    for t in pd:
        t[..., 4] = t[..., 4] * 1.0 - 3.15   # P(logit > log(1/3)) ~ 2 %
    anchors = _Head.anchors
    state = {}

    def decode_job():
        state["dec"] = yb.cells_to_bboxes(pd, anchors, [8, 16, 32], is_pred=True, to_list=False)
    job(decode_job)
    job(lambda: nms_device(state["dec"], 0.45, 0.25, 300))

    # ---- optimiser tail (optim.cu) on the 21.19 M-parameter flat buffers
    torch.manual_seed(0)
    m = yb.YOLOV5m(first_out=48, nc=80, anchors=yb.ANCHORS, ch=(192, 384, 768)).to(dev)
    opt = Adam(m)
    m.flat_grads.normal_(0, 1e-3)
    job(lambda: opt.step(grad_scale=1.0, max_norm=10.0))

    # ---- representative tcgen05 convs (bs=64): 96->96 3x3 @80 (halo-patch kernel), 192->192 3x3 @40 (generic kernel)
    ws = torch.empty(32 << 20, device=dev, dtype=torch.float32)
    for cin, cout, k, hh in ((96, 96, 3, 80), (192, 192, 3, 40), (96, 96, 1, 80)):
        x = torch.randn(64, hh, hh, cin, device=dev, generator=g).to(torch.bfloat16)
        yy = torch.empty(64, hh, hh, cout, device=dev, dtype=torch.bfloat16)
        w = (torch.randn(cout, k * k * cin, device=dev, generator=g) / (cin * k * k) ** 0.5).to(torch.bfloat16)
        wt = (torch.randn(cin, k * k * cout, device=dev, generator=g) / (cin * k * k) ** 0.5).to(torch.bfloat16)
        stats = torch.zeros(L.yb_conv_max_partials() * 2 * cout, device=dev)
        dw = torch.zeros(cout * k * k * cin, device=dev)
        nrows = ctypes.c_int(0)
        pf = _lib.checkp(L.yb_conv_fwd_plan(x.data_ptr(), 64, hh, hh, cin, cin, w.data_ptr(), cout, k, 1, yy.data_ptr(), cout, 0,
                                            None, None, 0, None, 0, stats.data_ptr(), ctypes.byref(nrows), 3, 85))
        pdg = _lib.checkp(L.yb_conv_dgrad_plan(yy.data_ptr(), 64, hh, hh, cout, cout, wt.data_ptr(), cin, k, 1, x.data_ptr(), cin,
                                               None, 0, 0))
        pw = _lib.checkp(L.yb_conv_wgrad_plan(x.data_ptr(), 64, hh, hh, cin, cin, yy.data_ptr(), cout, cout, k, 1, ws.data_ptr(),
                                              ws.numel(), 0))
        keep += [x, yy, w, wt, stats, dw]
        job(lambda pf=pf: L.yb_plan_run(pf, st))
        job(lambda pdg=pdg: L.yb_plan_run(pdg, st))
        job(lambda pw=pw, dw=dw, cout=cout: L.yb_wgrad_plan_run(pw, dw.data_ptr(), cout, None, 0, st))

    if len(sys.argv) > 1 and sys.argv[1].startswith("conv:"):   # e.g. conv:96,96,1,80,fwd -> only that conv launch
        cin, cout, k, hh, which = sys.argv[1][5:].split(",")
        sel = {(96, 96, 3, 80): 0, (192, 192, 3, 40): 1, (96, 96, 1, 80): 2}[(int(cin), int(cout), int(k), int(hh))]
        base = len(JOBS) - 9 + 3 * sel + {"fwd": 0, "dgrad": 1, "wgrad": 2}[which]
        JOBS[:] = [JOBS[base]]
    for fn in JOBS:   # warm-up
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for fn in JOBS:
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", len(JOBS), "jobs")


if __name__ == "__main__":
    main()
