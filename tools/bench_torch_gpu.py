"""Practical bar to beat (last row): the SAME network through stock PyTorch / cuDNN on the same B200.

A plain nn.Module restatement of the reference architecture (model.py:12-239: Conv2d + BatchNorm2d(eps 1e-3, mom 0.03) +
SiLU blocks, C3 / SPPF / PANet wiring, three 1x1 heads), channels_last, bf16 autocast, fused Adam, clip_grad_norm_(10).
The loss is a SURROGATE (mean of squares of the three head outputs): the reference ComputeLoss is ~1 % of the step and
its ~120 tiny launches would only slow this arm down, so the number printed here is an upper bound of what the stock
path can do.  Not part of the product path and not a bench.py arm: a measurement tool, output under profiles/.

    python tools/bench_torch_gpu.py [--bs 64] [--size 640] [--steps 20] [--warmup 5] [--eval]
"""
import argparse
import json

import torch
import torch.nn as nn
import torch.nn.functional as F


class CBL(nn.Sequential):
    def __init__(self, cin, cout, k, s):
        super().__init__(nn.Conv2d(cin, cout, k, s, 2 if k == 6 else k // 2, bias=False),
                         nn.BatchNorm2d(cout, eps=1e-3, momentum=0.03), nn.SiLU(inplace=True))


class C3(nn.Module):
    def __init__(self, cin, cout, width, depth, residual):
        super().__init__()
        c_ = int(cin * width)
        self.residual = residual
        self.c1, self.skip, self.out = CBL(cin, c_, 1, 1), CBL(cin, c_, 1, 1), CBL(2 * c_, cout, 1, 1)
        self.seq = nn.ModuleList([nn.Sequential(CBL(c_, c_, 1, 1), CBL(c_, c_, 3, 1)) for _ in range(depth)])

    def forward(self, x):
        a = self.c1(x)
        for blk in self.seq:
            a = blk(a) + a if self.residual else blk(a)
        return self.out(torch.cat([a, self.skip(x)], 1))


class SPPF(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.c1, self.out = CBL(c, c // 2, 1, 1), CBL(2 * c, c, 1, 1)

    def forward(self, x):
        x = self.c1(x)
        p1 = F.max_pool2d(x, 5, 1, 2)
        p2 = F.max_pool2d(p1, 5, 1, 2)
        return self.out(torch.cat([x, p1, p2, F.max_pool2d(p2, 5, 1, 2)], 1))


class Net(nn.Module):
    def __init__(self, c=48, nc=80):
        super().__init__()
        self.b = nn.ModuleList([CBL(3, c, 6, 2), CBL(c, 2 * c, 3, 2), C3(2 * c, 2 * c, .5, 2, True), CBL(2 * c, 4 * c, 3, 2),
                                C3(4 * c, 4 * c, .5, 4, True), CBL(4 * c, 8 * c, 3, 2), C3(8 * c, 8 * c, .5, 6, True),
                                CBL(8 * c, 16 * c, 3, 2), C3(16 * c, 16 * c, .5, 2, True), SPPF(16 * c)])
        self.n = nn.ModuleList([CBL(16 * c, 8 * c, 1, 1), C3(16 * c, 8 * c, .25, 2, False), CBL(8 * c, 4 * c, 1, 1),
                                C3(8 * c, 4 * c, .25, 2, False), CBL(4 * c, 4 * c, 3, 2), C3(8 * c, 8 * c, .5, 2, False),
                                CBL(8 * c, 8 * c, 3, 2), C3(16 * c, 16 * c, .5, 2, False)])
        self.h = nn.ModuleList([nn.Conv2d(ch, 3 * (5 + nc), 1) for ch in (4 * c, 8 * c, 16 * c)])
        self.no = 5 + nc

    def forward(self, x):
        taps = []
        for i, m in enumerate(self.b):
            x = m(x)
            if i in (4, 6):
                taps.append(x)
        n = self.n
        x = n0 = n[0](x)
        x = n[1](torch.cat([F.interpolate(x, scale_factor=2, mode="nearest"), taps[1]], 1))
        x = n2 = n[2](x)
        x = p3 = n[3](torch.cat([F.interpolate(x, scale_factor=2, mode="nearest"), taps[0]], 1))
        x = p4 = n[5](torch.cat([n[4](x), n2], 1))
        p5 = n[7](torch.cat([n[6](x), n0], 1))
        outs = []
        for f, h in zip((p3, p4, p5), self.h):
            o = h(f)
            B, _, H, W = o.shape
            outs.append(o.view(B, 3, self.no, H, W).permute(0, 1, 3, 4, 2).contiguous())
        return outs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bs", type=int, default=64)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--eval", action="store_true", help="inference forward only (detect path)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    net = Net().to(dev).to(memory_format=torch.channels_last)
    x = torch.rand(a.bs, 3, a.size, a.size, device=dev).contiguous(memory_format=torch.channels_last)
    opt = torch.optim.Adam(net.parameters(), lr=5e-4, weight_decay=5e-4, fused=True)

    def train_step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = net(x)
            loss = sum(o.float().pow(2).mean() for o in outs)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)
        opt.step()
        opt.zero_grad(set_to_none=True)

    def eval_step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            net(x)

    net.train(not a.eval)
    step = eval_step if a.eval else train_step
    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"impl": "torch %s + cuDNN %s, channels_last, bf16 autocast, fused Adam, surrogate loss" %
                      (torch.__version__, torch.backends.cudnn.version()),
                      "mode": "eval forward" if a.eval else "train step", "bs": a.bs, "size": a.size, "steps": a.steps,
                      "warmup": a.warmup, "ms_per_step": ms, "img_per_s": a.bs / ms * 1e3,
                      "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
