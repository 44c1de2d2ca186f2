"""Oracle: functional fp32 restatement of the reference network.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/model.py:
  CBL        model.py:12-28    conv(bias=False) -> BN(eps=1e-3, momentum=0.03) -> SiLU
  Bottleneck model.py:32-50    c2(c1(x)) + x
  C3         model.py:54-92    c_out(cat[seq(c1(x)), c_skipped(x)])
  SPPF       model.py:96-112   c_out(cat[x, p(x), p(p(x)), p(p(p(x)))]), p = maxpool 5/1/2
  HEADS      model.py:143-175  1x1 conv + bias, view(B,3,85,H,W).permute(0,1,3,4,2)
  YOLOV5m    model.py:178-239  backbone / neck wiring

The network is evaluated from a flat ``state_dict``-style mapping that uses the
reference's parameter names, so the same weights drive the reference, the
oracle and the CUDA path.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

ANCHORS = [  # config.py:33-37
    [(10, 13), (16, 30), (33, 23)],
    [(30, 61), (62, 45), (59, 119)],
    [(116, 90), (156, 198), (373, 326)],
]
STRIDES = [8, 16, 32]  # model.py:153
BN_EPS = 1e-3  # model.py:17
BN_MOMENTUM = 0.03


def _q(t, quant):
    """Optional bf16 quantisation point (emulates the CUDA path's storage type)."""
    return t.to(torch.bfloat16).to(torch.float32) if quant else t


class Net:
    """Functional evaluator. ``sd`` maps reference state_dict names -> tensors.

    quant=False : pure fp32 (the reference's arithmetic).
    quant=True  : weights and stored activations are rounded to bf16 at the
                  points where the CUDA bf16 path stores them (conv inputs,
                  raw conv outputs), arithmetic stays fp32.  Used to separate
                  "kernel is wrong" from "bf16 storage rounding".
    """

    def __init__(self, sd, first_out=48, nc=80, train=True, quant=False, update_stats=True):
        self.sd = sd
        self.c = first_out
        self.nc = nc
        self.train = train
        self.quant = quant
        self.update_stats = update_stats

    # -- blocks ------------------------------------------------------------
    def cbl(self, x, name, k, s, p):
        sd = self.sd
        w = _q(sd[name + ".cbl.0.weight"], self.quant)
        y = F.conv2d(_q(x, self.quant), w, None, s, p)
        g, b = sd[name + ".cbl.1.weight"], sd[name + ".cbl.1.bias"]
        rm, rv = sd[name + ".cbl.1.running_mean"], sd[name + ".cbl.1.running_var"]
        if self.train:
            # batch statistics are taken from the fp32 accumulator (before the
            # bf16 store) -- this is what the CUDA epilogue does.
            mean = y.mean(dim=(0, 2, 3))
            var = y.var(dim=(0, 2, 3), unbiased=False)
            if self.update_stats:
                n = y.numel() // y.shape[1]
                with torch.no_grad():
                    rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
                    rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.detach() * n / max(n - 1, 1))
                    sd[name + ".cbl.1.num_batches_tracked"] += 1
        else:
            mean, var = rm, rv
        yq = _q(y, self.quant)
        scale = g / torch.sqrt(var + BN_EPS)
        shift = b - mean * scale
        z = yq * scale[None, :, None, None] + shift[None, :, None, None]
        return F.silu(z)

    def c3(self, x, name, depth, backbone):
        a = self.cbl(x, name + ".c1", 1, 1, 0)
        for j in range(depth):
            if backbone:
                h = self.cbl(a, f"{name}.seq.{j}.c1", 1, 1, 0)
                h = self.cbl(h, f"{name}.seq.{j}.c2", 3, 1, 1)
                a = h + _q(a, self.quant)
            else:
                h = self.cbl(a, f"{name}.seq.{j}.0", 1, 1, 0)
                a = self.cbl(h, f"{name}.seq.{j}.1", 3, 1, 1)
        s = self.cbl(x, name + ".c_skipped", 1, 1, 0)
        return self.cbl(torch.cat([a, s], 1), name + ".c_out", 1, 1, 0)

    def sppf(self, x, name):
        x = self.cbl(x, name + ".c1", 1, 1, 0)
        x = _q(x, self.quant)
        p1 = F.max_pool2d(x, 5, 1, 2)
        p2 = F.max_pool2d(p1, 5, 1, 2)
        p3 = F.max_pool2d(p2, 5, 1, 2)
        return self.cbl(torch.cat([x, p1, p2, p3], 1), name + ".c_out", 1, 1, 0)

    # -- network -----------------------------------------------------------
    def forward(self, x):
        assert x.shape[2] % 32 == 0 and x.shape[3] % 32 == 0, "Width and Height aren't divisible by 32!"
        x = self.cbl(x, "backbone.0", 6, 2, 2)
        x = self.cbl(x, "backbone.1", 3, 2, 1)
        x = self.c3(x, "backbone.2", 2, True)
        x = self.cbl(x, "backbone.3", 3, 2, 1)
        x = tap_a = self.c3(x, "backbone.4", 4, True)
        x = self.cbl(x, "backbone.5", 3, 2, 1)
        x = tap_b = self.c3(x, "backbone.6", 6, True)
        x = self.cbl(x, "backbone.7", 3, 2, 1)
        x = self.c3(x, "backbone.8", 2, True)
        x = self.sppf(x, "backbone.9")

        x = n0 = self.cbl(x, "neck.0", 1, 1, 0)
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = self.c3(torch.cat([x, tap_b], 1), "neck.1", 2, False)
        x = n2 = self.cbl(x, "neck.2", 1, 1, 0)
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = p3 = self.c3(torch.cat([x, tap_a], 1), "neck.3", 2, False)
        x = self.cbl(x, "neck.4", 3, 2, 1)
        x = p4 = self.c3(torch.cat([x, n2], 1), "neck.5", 2, False)
        x = self.cbl(x, "neck.6", 3, 2, 1)
        p5 = self.c3(torch.cat([x, n0], 1), "neck.7", 2, False)

        outs = []
        for i, f in enumerate((p3, p4, p5)):
            w = _q(self.sd[f"head.out_convs.{i}.weight"], self.quant)
            o = F.conv2d(_q(f, self.quant), w, self.sd[f"head.out_convs.{i}.bias"])
            bs, _, gy, gx = o.shape
            outs.append(o.view(bs, 3, 5 + self.nc, gy, gx).permute(0, 1, 3, 4, 2).contiguous())
        return outs


def param_specs(first_out=48, nc=80):
    """Ordered (name, shape, kind) list reproducing the reference state_dict
    (481 entries for first_out=48, nc=80).  kind in {conv, bn_w, bn_b, bn_rm,
    bn_rv, bn_nbt, anchors, head_w, head_b}."""
    c = first_out
    specs = []

    def cbl(name, cin, cout, k):
        specs.append((name + ".cbl.0.weight", (cout, cin, k, k), "conv"))
        specs.append((name + ".cbl.1.weight", (cout,), "bn_w"))
        specs.append((name + ".cbl.1.bias", (cout,), "bn_b"))
        specs.append((name + ".cbl.1.running_mean", (cout,), "bn_rm"))
        specs.append((name + ".cbl.1.running_var", (cout,), "bn_rv"))
        specs.append((name + ".cbl.1.num_batches_tracked", (), "bn_nbt"))

    def c3(name, cin, cout, width, depth, backbone):
        c_ = int(width * cin)
        cbl(name + ".c1", cin, c_, 1)
        cbl(name + ".c_skipped", cin, c_, 1)
        for j in range(depth):
            a, b = ("c1", "c2") if backbone else ("0", "1")
            cbl(f"{name}.seq.{j}.{a}", c_, c_, 1)
            cbl(f"{name}.seq.{j}.{b}", c_, c_, 3)
        cbl(name + ".c_out", 2 * c_, cout, 1)

    cbl("backbone.0", 3, c, 6)
    cbl("backbone.1", c, 2 * c, 3)
    c3("backbone.2", 2 * c, 2 * c, 0.5, 2, True)
    cbl("backbone.3", 2 * c, 4 * c, 3)
    c3("backbone.4", 4 * c, 4 * c, 0.5, 4, True)
    cbl("backbone.5", 4 * c, 8 * c, 3)
    c3("backbone.6", 8 * c, 8 * c, 0.5, 6, True)
    cbl("backbone.7", 8 * c, 16 * c, 3)
    c3("backbone.8", 16 * c, 16 * c, 0.5, 2, True)
    cbl("backbone.9.c1", 16 * c, 8 * c, 1)
    cbl("backbone.9.c_out", 32 * c, 16 * c, 1)
    cbl("neck.0", 16 * c, 8 * c, 1)
    c3("neck.1", 16 * c, 8 * c, 0.25, 2, False)
    cbl("neck.2", 8 * c, 4 * c, 1)
    c3("neck.3", 8 * c, 4 * c, 0.25, 2, False)
    cbl("neck.4", 4 * c, 4 * c, 3)
    c3("neck.5", 8 * c, 8 * c, 0.5, 2, False)
    cbl("neck.6", 8 * c, 8 * c, 3)
    c3("neck.7", 16 * c, 16 * c, 0.5, 2, False)
    specs.append(("head.anchors", (3, 3, 2), "anchors"))
    for i, ch in enumerate((4 * c, 8 * c, 16 * c)):
        specs.append((f"head.out_convs.{i}.weight", ((5 + nc) * 3, ch, 1, 1), "head_w"))
        specs.append((f"head.out_convs.{i}.bias", ((5 + nc) * 3,), "head_b"))
    return specs


def head_anchors():
    """model.py:156-157: anchors divided by the level stride."""
    a = torch.tensor(ANCHORS).float().view(3, -1, 2)
    s = torch.tensor(STRIDES).repeat(6, 1).T.reshape(3, 3, 2)
    return a / s


def make_state_dict(seed=0, first_out=48, nc=80, bn_noise=True):
    """Deterministic synthetic weights (CPU generator => identical on every box).

    Conv weights ~ U(-b, b), b = 1/sqrt(fan_in) (the nn.Conv2d default bound);
    BN gamma/beta/running stats get mild noise when ``bn_noise`` so that eval
    mode is not the identity transform."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape, kind in param_specs(first_out, nc):
        if kind in ("conv", "head_w"):
            fan_in = shape[1] * shape[2] * shape[3]
            b = 1.0 / math.sqrt(fan_in)
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind == "head_b":
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind == "bn_w":
            sd[name] = 1 + 0.2 * (torch.rand(shape, generator=g) - 0.5) if bn_noise else torch.ones(shape)
        elif kind == "bn_b":
            sd[name] = 0.2 * (torch.rand(shape, generator=g) - 0.5) if bn_noise else torch.zeros(shape)
        elif kind == "bn_rm":
            sd[name] = 0.1 * (torch.rand(shape, generator=g) - 0.5) if bn_noise else torch.zeros(shape)
        elif kind == "bn_rv":
            sd[name] = 0.5 + torch.rand(shape, generator=g) if bn_noise else torch.ones(shape)
        elif kind == "bn_nbt":
            sd[name] = torch.tensor(0, dtype=torch.long)
        elif kind == "anchors":
            sd[name] = head_anchors()
    return sd


def forward(sd, x, train=True, quant=False, first_out=48, nc=80, update_stats=True):
    return Net(sd, first_out, nc, train, quant, update_stats).forward(x)
