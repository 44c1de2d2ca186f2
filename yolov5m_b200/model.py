"""YOLOV5m: drop-in for the reference network (reference model.py:178-239) running on hand-written sm_100a kernels.

Same constructor, module tree / state_dict keys (481 entries for first_out=48, nc=80), ``forward()`` contract
(list of three ``(B, 3, H/s, W/s, 5+nc)`` raw-logit tensors, autograd-connected in train mode) and head attributes
(``model.head.nc/nl/naxs/anchors/stride``) as the reference, so it drops in under train.py / detect.py.

What is different underneath (B200-first, not a translation):
  * activations are NHWC bf16; every ``torch.cat`` of the reference (model.py:91,112,226,230) is a channel slice of a
    shared buffer that the producing kernels write directly (zero-copy concat);
  * Conv2d fprop / dgrad / wgrad are tcgen05 implicit GEMMs (csrc/conv_igemm.cu, csrc/conv_wgrad.cu); the conv epilogue
    emits the BatchNorm batch-statistic partials, BN-apply + SiLU (+ residual, + nearest-2x upsample) is one pass;
  * eval mode folds BN into the conv epilogue (scale/shift + SiLU in registers, model.py:17-23);
  * the 6x6/s2 stem runs as a 3x3/s1 conv over a space-to-depth staging of the image;
  * parameters live in ONE flat fp32 buffer (the tensors in ``state_dict()`` are views of it, conv weights in
    channels-last memory order = the kernels' [Cout][tap][Cin] packing); gradients land in one flat fp32 buffer that is
    also the NCCL all-reduce bucket and the input of the fused clip+Adam kernel.
There is no CPU / PyTorch fallback: calling forward without the CUDA library or on CPU tensors raises.
"""
import ctypes
import math
import os
import weakref

import numpy as np
import torch
import torch.nn as nn

from . import _lib

BN_EPS = 1e-3      # reference model.py:17
BN_MOMENTUM = 0.03
HEAD_PAD = 256     # head conv Cout = 3*(5+nc) padded to a multiple of 16 for the bf16 gradient tensor


# ------------------------------------------------------------------------------------------------ module tree
_OWNERS = weakref.WeakValueDictionary()  # id(block) -> the YOLOV5m that owns it (blocks called on their own, run_submodule)


class _Conv(nn.Module):
    """Parameter holder for an nn.Conv2d (reference model.py:16,162).  The arithmetic runs in the engine."""

    def __init__(self, cin, cout, k, stride, bias=False):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size, self.stride = cin, cout, k, stride
        ref = nn.Conv2d(cin, cout, k, stride, k // 2 if k != 6 else 2, bias=bias)  # same init / RNG draws as the reference
        self.weight = nn.Parameter(ref.weight.data)
        self.bias = nn.Parameter(ref.bias.data) if bias else None

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, k={self.kernel_size}, s={self.stride}, bias={self.bias is not None}"

    def forward(self, x):
        """Sub-modules hold parameters; the arithmetic runs in the owning network's engine.  CBL / Bottleneck / C3 / SPPF
        blocks can still be called on their own for inspection (``model.backbone[0](img)``): inference semantics (running
        BN statistics), no autograd -- see YOLOV5m.run_submodule."""
        net = _OWNERS.get(id(self))
        if net is None or not isinstance(self, (CBL, Bottleneck, C3, SPPF)) or not any(m is self for m in net.modules()):
            raise _lib.YBError("this sub-module holds parameters only; call YOLOV5m.forward (fused sm_100a engine)")
        return net.run_submodule(self, x)


class _BN(nn.Module):
    """Parameter / buffer holder for nn.BatchNorm2d(eps=1e-3, momentum=0.03) (reference model.py:17)."""

    def __init__(self, c):
        super().__init__()
        self.num_features, self.eps, self.momentum = c, BN_EPS, BN_MOMENTUM
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    forward = _Conv.forward


class CBL(nn.Module):  # reference model.py:12-28
    def __init__(self, cin, cout, k, stride, padding=None):
        super().__init__()
        self.cbl = nn.Sequential(_Conv(cin, cout, k, stride), _BN(cout), nn.SiLU(inplace=True))

    forward = _Conv.forward


class Bottleneck(nn.Module):  # reference model.py:32-50
    def __init__(self, cin, cout, width_multiple=1):
        super().__init__()
        c_ = int(width_multiple * cin)
        self.c1 = CBL(cin, c_, 1, 1)
        self.c2 = CBL(c_, cout, 3, 1)

    forward = _Conv.forward


class C3(nn.Module):  # reference model.py:54-92
    def __init__(self, cin, cout, width_multiple=1, depth=1, backbone=True):
        super().__init__()
        c_ = int(width_multiple * cin)
        self.is_backbone, self.depth, self.hidden = backbone, depth, c_
        self.c1 = CBL(cin, c_, 1, 1)
        self.c_skipped = CBL(cin, c_, 1, 1)
        if backbone:
            self.seq = nn.Sequential(*[Bottleneck(c_, c_, 1) for _ in range(depth)])
        else:
            self.seq = nn.Sequential(*[nn.Sequential(CBL(c_, c_, 1, 1), CBL(c_, c_, 3, 1)) for _ in range(depth)])
        self.c_out = CBL(c_ * 2, cout, 1, 1)

    forward = _Conv.forward


class SPPF(nn.Module):  # reference model.py:96-112
    def __init__(self, cin, cout):
        super().__init__()
        c_ = cin // 2
        self.c1 = CBL(cin, c_, 1, 1)
        self.pool = nn.MaxPool2d(kernel_size=5, stride=1, padding=2)
        self.c_out = CBL(c_ * 4, cout, 1, 1)

    forward = _Conv.forward


class HEADS(nn.Module):  # reference model.py:143-175
    def __init__(self, nc=80, anchors=(), ch=()):
        super().__init__()
        self.nc = nc
        self.nl = len(anchors)
        self.naxs = len(anchors[0])
        self.stride = [8, 16, 32]
        anchors_ = torch.tensor(anchors).float().view(self.nl, -1, 2) / torch.tensor(self.stride).repeat(6, 1).T.reshape(3, 3, 2)
        self.register_buffer("anchors", anchors_)
        self.out_convs = nn.ModuleList([_Conv(c, (5 + nc) * self.naxs, 1, 1, bias=True) for c in ch])

    forward = _Conv.forward


# ------------------------------------------------------------------------------------------------ engine pieces
PARITY_PASSES = ((0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0))  # (activation plane, weight plane) per conv pass


class _Work(float):
    """algorithmic FLOPs of one conv launch, carrying its algorithmic bytes (activation read once + activation written
    once + weights), so that the timed wrappers can report the launch's roofline floor max(flops / tensor, bytes / HBM)"""

    def __new__(cls, flops, nbytes):
        o = float.__new__(cls, flops)
        o.nbytes = float(nbytes)
        return o


class _Buf:
    """NHWC bf16 allocation (+ same-shaped gradient when training).  Parity mode (csrc/parity.cu): the tensor and its
    gradient are fp32 and `pl` holds the three bf16 planes of the 3-way split that the tcgen05 convs read."""

    def __init__(self, N, H, W, C, grad, dev, parity=False, planes=True):
        self.N, self.H, self.W, self.C = N, H, W, C
        dt = torch.float32 if parity else torch.bfloat16
        self.esz = 4 if parity else 2
        self.t = torch.empty(N, H, W, C, device=dev, dtype=dt)
        self.g = torch.empty(N, H, W, C, device=dev, dtype=dt) if grad else None
        self.pl = torch.empty(3, N, H, W, C, device=dev, dtype=torch.bfloat16) if (parity and planes) else None
        self.plane_stride = N * H * W * C
        self.gw = np.zeros(C, bool)  # (backward-list construction) which gradient channels already hold a value
        self.pending = []            # identity gradient contributions (c0, C, src_view) not yet materialised

    def v(self, c0=0, C=None):
        return _View(self, c0, self.C - c0 if C is None else C)


class _View:
    def __init__(self, buf, c0, C):
        self.buf, self.c0, self.C = buf, c0, C
        self.N, self.H, self.W, self.pitch = buf.N, buf.H, buf.W, buf.C
        self.npix = buf.N * buf.H * buf.W

    @property
    def ptr(self):
        return self.buf.t.data_ptr() + self.buf.esz * self.c0

    @property
    def gptr(self):
        return self.buf.g.data_ptr() + self.buf.esz * self.c0

    def plptr(self, k=0):
        """parity mode: bf16 plane k of the 3-way split of this view"""
        return self.buf.pl.data_ptr() + 2 * (k * self.buf.plane_stride + self.c0)

    @property
    def plane_stride(self):
        return self.buf.plane_stride

    def pitch_c(self):
        """channels per pixel a conv over this view actually moves (the stem's window view re-reads its 16 stored ones)"""
        return min(self.C, self.pitch)

    def tensor(self):
        return self.buf.t[..., self.c0:self.c0 + self.C]

    def gtensor(self):
        return self.buf.g[..., self.c0:self.c0 + self.C]


class _StemView(_View):
    """The stem's input: the row-padded 16-channel space-to-depth staging (N, H, W+2, 16) read as (N, H, W, 48) with
    pitch 16 -- pixel w's 48 channels are padded columns w..w+2, i.e. the three horizontal taps (csrc: stem_view)."""

    def __init__(self, buf):
        super().__init__(buf, 0, 48)
        self.W = buf.W - 2
        self.npix = buf.N * buf.H * self.W


class _LayerRec:
    """Flat-buffer bookkeeping of one conv (+BN)."""
    __slots__ = ("name", "conv", "bn", "cin", "cout", "k", "stride", "w_off", "g_off", "b_off", "wt_off", "bias_off",
                 "rm", "rv", "nbt", "is_stem", "is_head", "kin", "kk", "cout_pad")


class _Engine:
    """Static launch plan for one (batch, height, width, mode): buffers, TMA plans, forward / backward op lists."""

    def __init__(self, net, B, H, W, train, parity=False):
        self.net, self.B, self.H, self.W, self.train, self.parity = net, B, H, W, train, parity
        self.dev = net._pflat.device
        self.L = _lib.lib()
        self.fwd_ops, self.bwd_ops, self.tape = [], [], []
        self.bwd_writes = {}      # backward op index -> [(offset, length)] slices of the flat gradient bucket it writes
        self.on_grad_chunks = None  # (schedule {op index: [chunk]}, callback(chunk, gflat)) set by trainer.GradSync
        self.prof = None
        self.conv_flops = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}  # algorithmic FLOPs of the reference convs per step
        self._fold, self._fold_eps, self._fold_dev = [], None, None  # inference: BatchNorms folded by ONE launch per forward
        self.plans = []
        self.bufs = []
        self.nbytes = 0
        L = self.L
        self.max_rows = L.yb_conv_max_partials()
        self.red_rows = L.yb_bwd_reduce_max_rows()
        # per-layer fp32 scratch: stats partials + scale/shift/mean/invstd
        tot_c = sum(r.cout for r in net._recs if not r.is_head)
        self.stat_pool = torch.zeros(tot_c * (2 * self.max_rows + 4), device=self.dev, dtype=torch.float32)
        self._stat_off = 0
        if train:
            cmax = max(HEAD_PAD, max(r.cout for r in net._recs))
            self.red_partial = torch.zeros(self.red_rows * 2 * cmax, device=self.dev, dtype=torch.float32)
            self.coef = torch.zeros(2 * cmax, device=self.dev, dtype=torch.float32)
            self.wg_ws = torch.empty(32 << 20, device=self.dev, dtype=torch.float32)
            # The split-K reduction of a weight gradient (wgrad_reduce_kernel: small, reads L2-resident partials) runs on a
            # second stream beside the next layer's backward kernels instead of serialising 74 short launches into the main
            # stream.  Two workspaces alternate, so a wgrad kernel never waits for the reduction that read "its" buffer two
            # layers ago; run_backward joins the stream before the gradient bucket is used.  YB_WGRAD_REDUCE_STREAM=0: off.
            self.rstream = (torch.cuda.Stream(device=self.dev)
                            if os.environ.get("YB_WGRAD_REDUCE_STREAM", "1") != "0" and not parity else None)
            self.wg_ws2 = torch.empty_like(self.wg_ws) if self.rstream is not None else None
            self._ws_toggle, self._plan_ws, self._rs_on = 0, {}, False
            self._ws_mma_ev = [torch.cuda.Event(), torch.cuda.Event()]
            self._ws_red_ev = [torch.cuda.Event(), torch.cuda.Event()]
            self._ws_pending = [False, False]
            self.dy = None  # allocated after the forward build (max conv output size)
            self._dy_elems = 0
            # YB_WGRAD_STREAM=1 (experiment, off by default): weight-gradient kernels on a second stream, meant to overlap
            # the tensor-bound wgrad of layer L with the HBM-bound BN/SiLU backward of layer L-1.  Measured on B200
            # (profiles/ab_wgrad_side_stream_r1.json): no gain -- the elementwise passes are launched as one full resident
            # wave and hold the register file, so a 512-thread wgrad CTA cannot become co-resident until they drain.
            self.side = (torch.cuda.Stream(device=self.dev)
                         if os.environ.get("YB_WGRAD_STREAM", "0") == "1" and not parity else None)
            self._bwd_events = []
        self._build()

    # -- timed launch wrappers (bench.py roofline: CUDA events around every tensor-core kernel launch when prof is a list)
    def _conv(self, plan, st, flops, kind):
        if self.prof is None:
            _lib.check(self.L.yb_plan_run(plan, st))
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(self.L.yb_plan_run(plan, st))
        e1.record()
        self.prof.append((kind, flops, e0, e1, getattr(flops, "nbytes", 0.0)))

    def _ew(self, kind, nbytes, fn, *args):
        """HBM-bound pass (BN/SiLU forward, backward reduce, backward apply): timed like the convs when prof is a list; the
        second field of the record is the pass's ALGORITHMIC bytes (every tensor read / written once)"""
        if self.prof is None:
            _lib.check(fn(*args))
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(fn(*args))
        e1.record()
        self.prof.append((kind, nbytes, e0, e1, nbytes))

    def _next_ws(self):
        """workspace of the next weight-gradient plan: (pointer, floats, buffer index)"""
        if self.rstream is None:
            return self.wg_ws.data_ptr(), self.wg_ws.numel(), 0
        i = self._ws_toggle
        self._ws_toggle ^= 1
        t = self.wg_ws if i == 0 else self.wg_ws2
        return t.data_ptr(), t.numel(), i

    def _wgrad(self, plan, st, flops, *args):
        if self.prof is None:
            b = self._plan_ws.get(plan) if self._rs_on else None
            if b is None:
                _lib.check(self.L.yb_wgrad_plan_run(plan, *args, st))
                return
            if self._ws_pending[b]:
                self._main.wait_event(self._ws_red_ev[b])  # the reduction that last read this workspace has finished
            _lib.check(self.L.yb_wgrad_plan_run_phase(plan, *args, 1, st))
            self._ws_mma_ev[b].record(self._main)
            self.rstream.wait_event(self._ws_mma_ev[b])
            _lib.check(self.L.yb_wgrad_plan_run_phase(plan, *args, 2, self.rstream.cuda_stream))
            self._ws_red_ev[b].record(self.rstream)
            self._ws_pending[b] = True
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(self.L.yb_wgrad_plan_run(plan, *args, st))
        e1.record()
        self.prof.append(("wgrad", flops, e0, e1, getattr(flops, "nbytes", 0.0)))

    def _layer_events(self):
        if self.side is None:
            return None
        evs = (torch.cuda.Event(), torch.cuda.Event())  # (dy ready + dgrad issued on the main stream, wgrad done on the side)
        self._bwd_events.append(evs)
        return evs

    def _wgrad_async(self, evs, plan, st, flops, *args):
        """Weight gradient of one layer.  With the side stream enabled it is ordered AFTER the layer's dgrad (so the dgrad
        -- which the next layer's backward waits for -- gets the SMs first) and then runs concurrently with the HBM-bound
        BN/SiLU-backward passes of the next layer; nothing on the main stream waits for it until its dy buffer is reused
        (two layers later) or the step ends (run_backward)."""
        if not self._side_on:
            self._wgrad(plan, st, flops, *args)
            return
        evs[0].record(self._main)
        self.side.wait_event(evs[0])
        _lib.check(self.L.yb_wgrad_plan_run(plan, *args, self.side.cuda_stream))
        evs[1].record(self.side)

    # -- allocation helpers
    def buf(self, N, H, W, C, grad=None, planes=True):
        b = _Buf(N, H, W, C, self.train if grad is None else grad, self.dev, self.parity, planes)
        self.bufs.append(b)
        self.nbytes += b.t.numel() * b.esz * (2 if b.g is not None else 1) + (b.pl.numel() * 2 if b.pl is not None else 0)
        return b

    def _stat(self, n):
        t = self.stat_pool[self._stat_off:self._stat_off + n]
        self._stat_off += n
        return t

    # -- forward construction
    def pair_setup(self, c3mod, xin):
        """C3's c1 and c_skipped (model.py:90-92) read the same input: ONE 1x1 GEMM with N = 2 * c_ (their weights are
        adjacent in the flat buffer).  Returns the shared state the two cbl() calls use, or None (parity mode / no pair)."""
        net, L = self.net, self.L
        if self.parity or c3mod not in net._pairs or os.environ.get("YB_C3_FUSE", "1") == "0":
            return None
        ra, rb, wt2 = net._pairs[c3mod]
        C, N, H, W = ra.cout, xin.N, xin.H, xin.W
        pr = {"C": C, "wt2": wt2, "ra": ra, "rb": rb, "flops": _Work(2.0 * xin.npix * 2 * C * ra.cin, 2.0 * (xin.npix * (xin.C + 2 * C) + 2 * C * ra.cin)), "xin": xin}
        w_ptr = net._wfwd.data_ptr() + 2 * ra.w_off
        if self.train:
            pr["y"] = self.buf(N, H, W, 2 * C, grad=False)
            pr["stats"] = self._stat(2 * self.max_rows * 2 * C)
            rows = ctypes.c_int(0)
            pr["plan"] = _lib.checkp(L.yb_conv_fwd_plan(xin.ptr, N, H, W, xin.C, xin.pitch, w_ptr, 2 * C, 1, 1, pr["y"].t.data_ptr(),
                                                        2 * C, 0, None, None, 0, None, 0, pr["stats"].data_ptr(),
                                                        ctypes.byref(rows), 3, 85))
            pr["nrows"] = rows.value
            pr["dy"] = torch.empty(xin.npix * 2 * C, device=self.dev, dtype=torch.bfloat16)
            self.nbytes += pr["dy"].numel() * 2
        else:
            pr["scale"], pr["shift"] = self._stat(2 * C), self._stat(2 * C)
        if pr.get("plan") is not None:
            self.plans.append(pr["plan"])
        return pr

    def cbl(self, mod, xin, out, res=None, up=None, pair=None, half=0):
        net, L = self.net, self.L
        r = net._rec_of[mod.cbl[0]]
        Ho, Wo = xin.H // (1 if r.is_stem else r.stride), xin.W // (1 if r.is_stem else r.stride)
        assert (out.H, out.W, out.C) == (Ho, Wo, r.cout), (r.name, out.H, out.W, out.C)
        C = r.cout
        flops = _Work(2.0 * xin.N * Ho * Wo * C * r.k * r.k * r.cin,  # the reference conv (stem: 6x6 over 3 channels)
                      2.0 * (xin.npix * xin.pitch_c() + xin.N * Ho * Wo * C + C * r.k * r.k * r.cin))
        self.conv_flops["fwd"] += flops
        if pair is not None:
            return self._cbl_pair(r, xin, out, pair, half, flops)
        scale, shift = self._stat(C), self._stat(C)
        w_ptr = net._wstem.data_ptr() if r.is_stem else net._wfwd.data_ptr() + 2 * r.w_off
        gam, bet = net._pflat.data_ptr() + 4 * r.g_off, net._pflat.data_ptr() + 4 * r.b_off
        rm, rv, nbt = r.rm.data_ptr(), r.rv.data_ptr(), r.nbt.data_ptr()
        k, s = (31, 1) if r.is_stem else (r.k, r.stride)  # stem: 3x1 over the tap-gathered staging (yb_prep_input)
        if self.parity:
            return self._cbl_parity(r, xin, out, res, up, flops, scale, shift, gam, bet, rm, rv, nbt, k, s)
        if not self.train:
            # eval: BN folded into the conv epilogue (running statistics), SiLU + residual in registers
            plan = _lib.checkp(L.yb_conv_fwd_plan(xin.ptr, xin.N, xin.H, xin.W, xin.C, xin.pitch, w_ptr, C, k, s, out.ptr,
                                                  out.pitch, 0, scale.data_ptr(), shift.data_ptr(), 1,
                                                  res.ptr if res is not None else None, res.pitch if res is not None else 0,
                                                  None, None, 3, 85))
            self.plans.append(plan)
            self._fold.append((gam, bet, rm, rv, scale.data_ptr(), shift.data_ptr(), C, r.bn))
            self.fwd_ops.append(lambda st, plan=plan: self._conv(plan, st, flops, "fwd"))
            if up is not None:
                self.fwd_ops.append(lambda st: _lib.check(L.yb_upsample2x_fwd(out.ptr, out.pitch, out.N, out.H, out.W, C,
                                                                               up.ptr, up.pitch, st)))
            return
        mean, invstd = self._stat(C), self._stat(C)
        stats = self._stat(2 * self.max_rows * C)
        y = self.buf(xin.N, Ho, Wo, C, grad=False).v()
        rows = ctypes.c_int(0)
        plan = _lib.checkp(L.yb_conv_fwd_plan(xin.ptr, xin.N, xin.H, xin.W, xin.C, xin.pitch, w_ptr, C, k, s, y.ptr, C, 0,
                                              None, None, 0, None, 0, stats.data_ptr(), ctypes.byref(rows), 3, 85))
        self.plans.append(plan)
        nrows, count = rows.value, float(y.npix)
        ptrs = (stats.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr())
        resp, resl = (res.ptr, res.pitch) if res is not None else (None, 0)
        upp, upl = (up.ptr, up.pitch) if up is not None else (None, 0)
        # algorithmic bytes of the BN+SiLU pass: read y, write a (+ read the residual, + 4 upsampled copies), 2 B / element
        ew_bytes = 2.0 * y.npix * C * (2 + (1 if res is not None else 0) + (4 if up is not None else 0))
        bn = r.bn  # eps / momentum are read at launch time, like nn.BatchNorm2d reads its attributes

        def op(st):
            self._conv(plan, st, flops, "fwd")
            _lib.check(L.yb_bn_finalize(ptrs[0], nrows, C, count, gam, bet, bn.eps, bn.momentum, rm, rv, nbt, ptrs[1], ptrs[2],
                                        ptrs[3], ptrs[4], 1, st))
            self._ew("bn_fwd", ew_bytes, L.yb_bn_act_fwd, y.ptr, C, y.N, y.H, y.W, C, ptrs[1], ptrs[2], resp, resl, out.ptr,
                     out.pitch, upp, upl, st)
        self.fwd_ops.append(op)
        self._dy_elems = max(self._dy_elems, y.npix * C)
        self.tape.append(("cbl", r, xin, out, y, res, up, ptrs, flops))

    def _cbl_pair(self, r, xin, out, pr, half, flops):
        """one half (0 = c1, 1 = c_skipped) of a fused C3 input pair; half 0 launches the shared GEMM"""
        net, L, C = self.net, self.L, pr["C"]
        gam, bet = net._pflat.data_ptr() + 4 * r.g_off, net._pflat.data_ptr() + 4 * r.b_off
        rm, rv, nbt = r.rm.data_ptr(), r.rv.data_ptr(), r.nbt.data_ptr()
        if not self.train:
            # eval: both halves leave the ONE conv epilogue (folded BN + SiLU) straight into the concat buffer: c1's output
            # occupies the slot that the last bottleneck overwrites later (it is dead by then: depth >= 2)
            sp, hp = pr["scale"].data_ptr() + 4 * half * C, pr["shift"].data_ptr() + 4 * half * C
            if half == 0:
                assert out.c0 == 0 and out.pitch == 2 * C, "fused C3 pair: c1 must write slot 0 of the concat buffer"
                plan = _lib.checkp(L.yb_conv_fwd_plan(xin.ptr, xin.N, xin.H, xin.W, xin.C, xin.pitch,
                                                      net._wfwd.data_ptr() + 2 * pr["ra"].w_off, 2 * C, 1, 1, out.ptr, out.pitch, 0,
                                                      pr["scale"].data_ptr(), pr["shift"].data_ptr(), 1, None, 0, None, None,
                                                      3, 85))
                self.plans.append(plan)
                rb = pr["rb"]
                gb, bb = net._pflat.data_ptr() + 4 * rb.g_off, net._pflat.data_ptr() + 4 * rb.b_off
                sp1, hp1 = pr["scale"].data_ptr() + 4 * C, pr["shift"].data_ptr() + 4 * C
                pflops = pr["flops"]

                self._fold.append((gam, bet, rm, rv, sp, hp, C, r.bn))
                self._fold.append((gb, bb, rb.rm.data_ptr(), rb.rv.data_ptr(), sp1, hp1, C, rb.bn))
                self.fwd_ops.append(lambda st: self._conv(plan, st, pflops, "fwd"))
            return
        scale, shift, mean, invstd = self._stat(C), self._stat(C), self._stat(C), self._stat(C)
        y = pr["y"].v(half * C, C)
        stats_ptr = pr["stats"].data_ptr() + 4 * half * C
        ptrs = (stats_ptr, scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr())
        nrows, count, plan, pflops, bn = pr["nrows"], float(y.npix), pr["plan"], pr["flops"], r.bn
        ew_bytes = 4.0 * y.npix * C

        def op(st):
            if half == 0:
                self._conv(plan, st, pflops, "fwd")
            _lib.check(L.yb_bn_finalize_ld(ptrs[0], nrows, C, 2 * C, count, gam, bet, bn.eps, bn.momentum, rm, rv, nbt, ptrs[1],
                                           ptrs[2], ptrs[3], ptrs[4], 1, st))
            self._ew("bn_fwd", ew_bytes, L.yb_bn_act_fwd, y.ptr, y.pitch, y.N, y.H, y.W, C, ptrs[1], ptrs[2], None, 0, out.ptr,
                     out.pitch, None, 0, st)
        self.fwd_ops.append(op)
        self.tape.append(("cbl", r, xin, out, y, None, None, ptrs, flops, (pr, half)))

    # -- parity mode (fp32 activations, six bf16-split passes of the same tcgen05 kernels per conv; csrc/parity.cu)
    def _parity_conv_plans(self, xin, r, k, s, out_ptr, out_pitch, head=False, bias_ptr=None):
        net, L = self.net, self.L
        plans = []
        for n_, (i, j) in enumerate(PARITY_PASSES):
            w_ptr = net._wstem_pl[j].data_ptr() if r.is_stem else net._wfwd_pl[j].data_ptr() + 2 * r.w_off
            kind = (1 if n_ == 0 else 4) if head else (2 if n_ == 0 else 3)
            plans.append(_lib.checkp(L.yb_conv_fwd_plan(xin.plptr(i), xin.N, xin.H, xin.W, xin.C, xin.pitch, w_ptr, r.cout, k, s,
                                                        out_ptr, out_pitch, kind, None, bias_ptr if n_ == 0 else None, 0,
                                                        None, 0, None, None, self.net.head.naxs, 5 + self.net.head.nc)))
        self.plans += plans
        return plans

    def _cbl_parity(self, r, xin, out, res, up, flops, scale, shift, gam, bet, rm, rv, nbt, k, s):
        L, C = self.L, r.cout
        mean, invstd = self._stat(C), self._stat(C)
        stats = self._stat(2 * self.max_rows * C)
        y = self.buf(xin.N, out.H, out.W, C, grad=False, planes=False).v()
        plans = self._parity_conv_plans(xin, r, k, s, y.ptr, C)
        count = float(y.npix)
        ptrs = (stats.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), invstd.data_ptr())
        resp, resl = (res.ptr, res.pitch) if res is not None else (None, 0)
        upa = (up.ptr, up.plptr(0), up.pitch, up.plane_stride) if up is not None else (None, None, 0, 0)
        train, max_rows, bn = self.train, self.max_rows, r.bn

        def op(st):
            for pl in plans:
                _lib.check(L.yb_plan_run(pl, st))
            if train:
                rows = ctypes.c_int(0)
                _lib.check(L.yb_p32_bn_stats(y.ptr, C, y.npix, C, ptrs[0], max_rows, ctypes.byref(rows), st))
                _lib.check(L.yb_bn_finalize(ptrs[0], rows.value, C, count, gam, bet, bn.eps, bn.momentum, rm, rv, nbt, ptrs[1],
                                            ptrs[2], ptrs[3], ptrs[4], 1, st))
            else:
                _lib.check(L.yb_bn_finalize(None, 0, C, 1.0, gam, bet, BN_EPS, BN_MOMENTUM, rm, rv, None, ptrs[1], ptrs[2], None,
                                            None, 0, st))
            _lib.check(L.yb_p32_bn_act_fwd(y.ptr, C, y.N, y.H, y.W, C, ptrs[1], ptrs[2], resp, resl, out.ptr, out.plptr(0),
                                           out.pitch, out.plane_stride, upa[0], upa[1], upa[2], upa[3], st))
        self.fwd_ops.append(op)
        if self.train:
            self._dy_elems = max(self._dy_elems, y.npix * C)
            self.tape.append(("cbl", r, xin, out, y, res, up, ptrs, flops))

    def c3(self, mod, xin, out):
        c_ = mod.hidden
        N, H, W = xin.N, xin.H, xin.W
        cat = self.buf(N, H, W, 2 * c_)
        pair = self.pair_setup(mod, xin) if mod.depth >= 2 else None
        # eval + fused pair: c1's output lives in slot 0 of the concat buffer until the last bottleneck overwrites it
        a = cat.v(0, c_) if (pair is not None and not self.train) else self.buf(N, H, W, c_).v()
        self.cbl(mod.c1, xin, a, pair=pair, half=0)
        for j in range(mod.depth):
            blk = mod.seq[j]
            first, second = (blk.c1, blk.c2) if mod.is_backbone else (blk[0], blk[1])
            h = self.buf(N, H, W, c_).v()
            self.cbl(first, a, h)
            dst = cat.v(0, c_) if j == mod.depth - 1 else self.buf(N, H, W, c_).v()
            self.cbl(second, h, dst, res=a if mod.is_backbone else None)
            a = dst
        self.cbl(mod.c_skipped, xin, cat.v(c_, c_), pair=pair, half=1)
        self.cbl(mod.c_out, cat.v(), out)

    def sppf(self, mod, xin, out):
        L = self.L
        c_ = xin.C // 2
        N, H, W = xin.N, xin.H, xin.W
        cat = self.buf(N, H, W, 4 * c_)
        self.cbl(mod.c1, xin, cat.v(0, c_))
        views = [cat.v(i * c_, c_) for i in range(4)]
        ams = [torch.empty(N, H, W, c_, device=self.dev, dtype=torch.uint8) if self.train else None for _ in range(3)]
        amp = [a.data_ptr() if a is not None else None for a in ams]
        # the three chained 5x5 pools of model.py:108-110 as ONE launch when the map fits a CTA's shared memory
        # (csrc/elementwise.cu: sppf_pool3_*); otherwise (and in parity mode) three maxpool launches
        fused_fwd = not self.parity and c_ % 16 == 0 and H * W * (80 if self.train else 64) <= 200 * 1024
        fused_bwd = fused_fwd and H * W * 144 <= 200 * 1024
        if fused_fwd:
            v = views
            self.fwd_ops.append(lambda st: _lib.check(L.yb_sppf_pool3_fwd(v[0].ptr, v[0].pitch, N, H, W, c_, v[1].ptr, v[2].ptr,
                                                                          v[3].ptr, v[1].pitch, amp[0], amp[1], amp[2], st)))
        for i in range(3):
            src, dst = views[i], views[i + 1]
            if self.parity:
                self.fwd_ops.append(lambda st, src=src, dst=dst, a=amp[i]: _lib.check(
                    L.yb_p32_maxpool5_fwd(src.ptr, src.pitch, N, H, W, c_, dst.ptr, dst.plptr(0), dst.pitch, dst.plane_stride,
                                          a, st)))
            elif not fused_fwd:
                self.fwd_ops.append(lambda st, src=src, dst=dst, a=amp[i]: _lib.check(
                    L.yb_maxpool5_fwd(src.ptr, src.pitch, N, H, W, c_, dst.ptr, dst.pitch, a, st)))
            if self.train and not fused_bwd:
                self.tape.append(("pool", src, dst, ams[i]))
        if self.train and fused_bwd:
            self.tape.append(("pool3", views, ams))
        self.cbl(mod.c_out, cat.v(), out)

    def head(self, i, xin):
        net, L = self.net, self.L
        r = net._rec_of[net.head.out_convs[i]]
        na, no = net.head.naxs, 5 + net.head.nc
        out = torch.empty(xin.N, na, xin.H, xin.W, no, device=self.dev, dtype=torch.float32)
        flops = _Work(2.0 * xin.npix * r.cout * r.cin, xin.npix * (2.0 * r.cin + 4.0 * r.cout) + 2.0 * r.cout * r.cin)
        self.conv_flops["fwd"] += flops
        if self.parity:
            plans = self._parity_conv_plans(xin, r, 1, 1, out.data_ptr(), r.cout, head=True,
                                            bias_ptr=net._pflat.data_ptr() + 4 * r.bias_off)

            def op(st):
                for pl in plans:
                    _lib.check(L.yb_plan_run(pl, st))
            self.fwd_ops.append(op)
            self.outs.append(out)
            if self.train:
                dyh = torch.zeros(xin.N, xin.H, xin.W, HEAD_PAD, device=self.dev, dtype=torch.float32)
                dyp = torch.zeros(3, xin.N, xin.H, xin.W, HEAD_PAD, device=self.dev, dtype=torch.bfloat16)
                self.head_dy.append(dyh)
                self.head_dy_pl.append(dyp)
                self.tape.append(("head", r, xin, (dyh, dyp), flops))
            return
        plan = _lib.checkp(L.yb_conv_fwd_plan(xin.ptr, xin.N, xin.H, xin.W, xin.C, xin.pitch,
                                              net._wfwd.data_ptr() + 2 * r.w_off, r.cout, 1, 1, out.data_ptr(), r.cout, 1,
                                              None, net._pflat.data_ptr() + 4 * r.bias_off, 0, None, 0, None, None, na, no))
        self.plans.append(plan)
        self.fwd_ops.append(lambda st: self._conv(plan, st, flops, "fwd"))
        self.outs.append(out)
        if self.train:
            dyh = torch.zeros(xin.N, xin.H, xin.W, HEAD_PAD, device=self.dev, dtype=torch.bfloat16)
            self.head_dy.append(dyh)
            self.tape.append(("head", r, xin, dyh, flops))

    def _build(self):
        net, B, H, W = self.net, self.B, self.H, self.W
        bb, nk = net.backbone, net.neck
        c = net._first_out
        self.outs, self.head_dy, self.head_dy_pl = [], [], []
        self.x16 = self.buf(B, H // 2, W // 2 + 2, 16, grad=False)  # stem staging: space-to-depth, one zero pixel either side
        b0 = self.buf(B, H // 2, W // 2, c); self.cbl(bb[0], _StemView(self.x16), b0.v())
        b1 = self.buf(B, H // 4, W // 4, 2 * c); self.cbl(bb[1], b0.v(), b1.v())
        b2 = self.buf(B, H // 4, W // 4, 2 * c); self.c3(bb[2], b1.v(), b2.v())
        b3 = self.buf(B, H // 8, W // 8, 4 * c); self.cbl(bb[3], b2.v(), b3.v())
        cat3 = self.buf(B, H // 8, W // 8, 8 * c)      # [up(N2) | tapA]   (model.py:226, second pass)
        tapA = cat3.v(4 * c, 4 * c); self.c3(bb[4], b3.v(), tapA)
        b5 = self.buf(B, H // 16, W // 16, 8 * c); self.cbl(bb[5], tapA, b5.v())
        cat1 = self.buf(B, H // 16, W // 16, 16 * c)   # [up(N0) | tapB]   (model.py:226, first pass)
        tapB = cat1.v(8 * c, 8 * c); self.c3(bb[6], b5.v(), tapB)
        b7 = self.buf(B, H // 32, W // 32, 16 * c); self.cbl(bb[7], tapB, b7.v())
        b8 = self.buf(B, H // 32, W // 32, 16 * c); self.c3(bb[8], b7.v(), b8.v())
        b9 = self.buf(B, H // 32, W // 32, 16 * c); self.sppf(bb[9], b8.v(), b9.v())
        cat7 = self.buf(B, H // 32, W // 32, 16 * c)   # [neck.6 | N0]     (model.py:230)
        N0 = cat7.v(8 * c, 8 * c); self.cbl(nk[0], b9.v(), N0, up=cat1.v(0, 8 * c))
        n1 = self.buf(B, H // 16, W // 16, 8 * c); self.c3(nk[1], cat1.v(), n1.v())
        cat5 = self.buf(B, H // 16, W // 16, 8 * c)    # [neck.4 | N2]
        N2 = cat5.v(4 * c, 4 * c); self.cbl(nk[2], n1.v(), N2, up=cat3.v(0, 4 * c))
        P3 = self.buf(B, H // 8, W // 8, 4 * c); self.c3(nk[3], cat3.v(), P3.v())
        self.cbl(nk[4], P3.v(), cat5.v(0, 4 * c))
        P4 = self.buf(B, H // 16, W // 16, 8 * c); self.c3(nk[5], cat5.v(), P4.v())
        self.cbl(nk[6], P4.v(), cat7.v(0, 8 * c))
        P5 = self.buf(B, H // 32, W // 32, 16 * c); self.c3(nk[7], cat7.v(), P5.v())
        for i, p in enumerate((P3, P4, P5)):
            self.head(i, p.v())
        if self.train:
            ndy = 2 if self.side is not None else 1  # double-buffered when wgrad(L) overlaps the BN backward of L-1
            npl = 3 if self.parity else 1            # parity: the three bf16 planes of dy, _dy_elems apart
            self.dy = [torch.empty(npl * self._dy_elems, device=self.dev, dtype=torch.bfloat16) for _ in range(ndy)]
            self.nbytes += self._dy_elems * 2 * ndy * npl
            self._build_backward()

    def _add_bwd_op(self, op, writes):
        """backward op that writes the parameter-gradient slices `writes` = [(offset, length), ...] of the flat bucket"""
        self.bwd_writes[len(self.bwd_ops)] = writes
        self.bwd_ops.append(op)

    def grad_chunk_schedule(self, bounds):
        """For the all-reduce chunks [bounds[c], bounds[c+1]) of the flat gradient bucket: {backward op index: [chunks that
        are complete once that op has been issued]} -- a chunk is complete after the LAST op (in execution order) that
        writes a slice intersecting it.  (Data-parallel overlap: yolov5m_b200.trainer.GradSync reduces a chunk while the
        rest of the backward pass still runs.)"""
        from .trainer import chunk_ready_after
        return chunk_ready_after(self.bwd_writes, bounds)

    # -- backward construction (reverse tape); gradient fan-in is resolved statically
    def _contrib_state(self, view):
        w = view.buf.gw[view.c0:view.c0 + view.C]
        if w.all():
            return True
        assert not w.any(), "partially written gradient slice"
        return False

    def _flush_pending(self, view):
        """materialise identity contributions that target `view` (called before its gradient is consumed)."""
        L = self.L
        keep = []
        for (c0, C, src) in view.buf.pending:
            if c0 >= view.c0 and c0 + C <= view.c0 + view.C:
                tgt = view.buf.v(c0, C)
                acc = 1 if self._contrib_state(tgt) else 0
                add_into = L.yb_p32_add_into if self.parity else L.yb_add_into
                self.bwd_ops.append(lambda st, g, src=src, tgt=tgt, acc=acc, add_into=add_into: _lib.check(
                    add_into(src.gptr, src.pitch, tgt.gptr, tgt.pitch, tgt.npix, tgt.C, acc, st)))
                view.buf.gw[c0:c0 + C] = True
            else:
                keep.append((c0, C, src))
        view.buf.pending = keep

    def _take_pending_exact(self, view):
        for i, (c0, C, src) in enumerate(view.buf.pending):
            if c0 == view.c0 and C == view.C:
                del view.buf.pending[i]
                return src
        return None

    def _parity_bwd_plans(self, r, xin, dy_ptr, dy_ps, Cdy, k, s, acc, ws, wsn):
        """parity mode: the six dgrad passes (dy plane i x weight plane j -> fp32 xin.g, accumulating) and the six wgrad
        passes (x plane i x dy plane j) of one conv; dy planes are dy_ps elements apart"""
        net, L = self.net, self.L
        dplans, wplans = [], []
        for n_, (i, j) in enumerate(PARITY_PASSES):
            if not r.is_stem:
                dplans.append(_lib.checkp(L.yb_conv_dgrad_plan(dy_ptr + 2 * i * dy_ps, xin.N, xin.H, xin.W, Cdy, Cdy,
                                                               net._wdg_pl[j].data_ptr() + 2 * r.wt_off, xin.C, k, s,
                                                               xin.gptr, xin.pitch, None, 0, 3 if (acc or n_) else 2)))
            wplans.append(_lib.checkp(L.yb_conv_wgrad_plan(xin.plptr(i), xin.N, xin.H, xin.W, xin.C, xin.pitch,
                                                           dy_ptr + 2 * j * dy_ps, Cdy, Cdy, k, s, ws, wsn, 0)))
        self.plans += dplans + wplans
        return dplans, wplans

    def _build_backward(self):
        net, L = self.net, self.L
        redp, coefp = self.red_partial.data_ptr(), self.coef.data_ptr()
        ws, wsn = self.wg_ws.data_ptr(), self.wg_ws.numel()
        nlayer, done_by_slot = 0, {}
        for rec in reversed(self.tape):
            kind = rec[0]
            if kind == "head" and self.parity:
                _, r, xin, (dyh, dypl), flops = rec
                npix, ps = xin.npix, dyh.numel()
                acc = 1 if self._contrib_state(xin) else 0
                xin.buf.gw[xin.c0:xin.c0 + xin.C] = True
                dplans, wplans = self._parity_bwd_plans(r, xin, dypl.data_ptr(), ps, HEAD_PAD, 1, 1, acc, ws, wsn)

                def op(st, g, r=r, dyh=dyh, npix=npix, dplans=dplans, wplans=wplans):
                    rows = ctypes.c_int(0)
                    _lib.check(L.yb_p32_bn_stats(dyh.data_ptr(), HEAD_PAD, npix, HEAD_PAD, redp, self.red_rows,
                                                 ctypes.byref(rows), st))
                    _lib.check(L.yb_reduce_rows(redp, rows.value, 2 * HEAD_PAD, r.cout, g + 4 * r.bias_off, 0, st))
                    for pl in dplans:
                        _lib.check(L.yb_plan_run(pl, st))
                    for n_, pl in enumerate(wplans):
                        _lib.check(L.yb_wgrad_plan_run(pl, g + 4 * r.w_off, r.cout, None, 1 if n_ else 0, st))
                self._add_bwd_op(op, [(r.bias_off, r.cout), (r.w_off, r.cout * r.cin)])
            elif kind == "head":
                _, r, xin, dyh, flops = rec
                self.conv_flops["dgrad"] += flops
                self.conv_flops["wgrad"] += flops
                dyp = dyh.data_ptr()
                npix = xin.npix
                ws, wsn, wsi = self._next_ws()
                wplan = _lib.checkp(L.yb_conv_wgrad_plan(xin.ptr, xin.N, xin.H, xin.W, xin.C, xin.pitch, dyp, HEAD_PAD,
                                                         HEAD_PAD, 1, 1, ws, wsn, 0))
                self._plan_ws[wplan] = wsi
                acc = 1 if self._contrib_state(xin) else 0
                dplan = _lib.checkp(L.yb_conv_dgrad_plan(dyp, xin.N, xin.H, xin.W, HEAD_PAD, HEAD_PAD,
                                                         net._wdg.data_ptr() + 2 * r.wt_off, xin.C, 1, 1, xin.gptr,
                                                         xin.pitch, xin.gptr if acc else None, xin.pitch if acc else 0, 0))
                self.plans += [wplan, dplan]
                xin.buf.gw[xin.c0:xin.c0 + xin.C] = True
                rows = ctypes.c_int(0)

                evs = self._layer_events()

                def op(st, g, r=r, dyp=dyp, npix=npix, wplan=wplan, dplan=dplan, rows=rows, flops=flops, evs=evs):
                    _lib.check(L.yb_colsum(dyp, HEAD_PAD, npix, HEAD_PAD, redp, ctypes.byref(rows), st))
                    _lib.check(L.yb_reduce_rows(redp, rows.value, 2 * HEAD_PAD, r.cout, g + 4 * r.bias_off, 0, st))
                    self._conv(dplan, st, flops, "dgrad")
                    self._wgrad_async(evs, wplan, st, flops, g + 4 * r.w_off, r.cout, None, 0)
                self._add_bwd_op(op, [(r.bias_off, r.cout), (r.w_off, r.cout * r.cin)])
            elif kind == "pool3":
                _, v, ams = rec
                for d in v[1:]:
                    self._flush_pending(d)
                    assert self._contrib_state(d), "SPPF: pooled-slice gradient never produced"
                acc = 1 if self._contrib_state(v[0]) else 0
                v[0].buf.gw[v[0].c0:v[0].c0 + v[0].C] = True
                self.bwd_ops.append(lambda st, g, v=v, ams=ams, acc=acc: _lib.check(
                    L.yb_sppf_pool3_bwd(v[1].gptr, v[2].gptr, v[3].gptr, v[1].pitch, ams[0].data_ptr(), ams[1].data_ptr(),
                                        ams[2].data_ptr(), v[0].N, v[0].H, v[0].W, v[0].C, v[0].gptr, v[0].pitch, acc, st)))
            elif kind == "pool":
                _, src, dst, am = rec
                self._flush_pending(dst)
                acc = 1 if self._contrib_state(src) else 0
                src.buf.gw[src.c0:src.c0 + src.C] = True
                pool_bwd = L.yb_p32_maxpool5_bwd if self.parity else L.yb_maxpool5_bwd
                self.bwd_ops.append(lambda st, g, src=src, dst=dst, am=am, acc=acc, pool_bwd=pool_bwd: _lib.check(
                    pool_bwd(dst.gptr, dst.pitch, am.data_ptr(), src.N, src.H, src.W, src.C, src.gptr, src.pitch, acc, st)))
            elif len(rec) == 10:
                # one half of a fused C3 input pair (c1 || c_skipped as ONE GEMM): its BN/SiLU backward writes its half of the
                # pair's dy buffer; the half that comes LAST in the backward order (c1) then runs the ONE dgrad (K = 2 c_) and
                # the ONE wgrad (N = 2 c_, whose output is the two adjacent weight-gradient slices of the flat bucket)
                _, r, xin, out, y, _res, _up, ptrs, flops, (pr, half) = rec
                self.conv_flops["wgrad"] += flops
                self.conv_flops["dgrad"] += flops
                C, npix = r.cout, y.npix
                dy2 = pr["dy"].data_ptr()
                self._flush_pending(out)
                assert self._contrib_state(out), f"{r.name}: output gradient never produced"
                dplan = wplan = None
                if half == 0:
                    src = self._take_pending_exact(xin)
                    written = self._contrib_state(xin)
                    if src is not None and written:
                        self.bwd_ops.append(lambda st, g, src=src, xin=xin: _lib.check(
                            L.yb_add_into(src.gptr, src.pitch, xin.gptr, xin.pitch, xin.npix, xin.C, 1, st)))
                        src = None
                    addp, addl = (src.gptr, src.pitch) if src is not None else ((xin.gptr, xin.pitch) if written else (None, 0))
                    dplan = _lib.checkp(L.yb_conv_dgrad_plan(dy2, xin.N, xin.H, xin.W, 2 * C, 2 * C,
                                                             net._wdg.data_ptr() + 2 * pr["wt2"], xin.C, 1, 1, xin.gptr, xin.pitch,
                                                             addp, addl, 0))
                    ws, wsn, wsi = self._next_ws()
                    wplan = _lib.checkp(L.yb_conv_wgrad_plan(xin.ptr, xin.N, xin.H, xin.W, xin.C, xin.pitch, dy2, 2 * C, 2 * C, 1, 1,
                                                             ws, wsn, 0))
                    self._plan_ws[wplan] = wsi
                    self.plans += [dplan, wplan]
                    xin.buf.gw[xin.c0:xin.c0 + xin.C] = True
                rows = ctypes.c_int(0)
                count, pflops, w0 = float(npix), pr["flops"], pr["ra"].w_off

                def op(st, g, r=r, out=out, y=y, ptrs=ptrs, C=C, npix=npix, wplan=wplan, dplan=dplan, rows=rows, count=count,
                       pflops=pflops, dy2=dy2, half=half, w0=w0):
                    self._ew("bn_bwd_reduce", 4.0 * npix * C, L.yb_bn_act_bwd_reduce, out.gptr, out.pitch, y.ptr, y.pitch, npix, C,
                             ptrs[1], ptrs[2], ptrs[3], ptrs[4], redp, ctypes.byref(rows), st)
                    _lib.check(L.yb_bn_bwd_finalize(redp, rows.value, C, count, g + 4 * r.g_off, g + 4 * r.b_off, coefp, 0, st))
                    self._ew("bn_bwd_apply", 6.0 * npix * C, L.yb_bn_act_bwd_apply, out.gptr, out.pitch, y.ptr, y.pitch, npix, C,
                             ptrs[1], ptrs[2], ptrs[3], ptrs[4], coefp, dy2 + 2 * half * C, 2 * C, st)
                    if dplan is not None:
                        self._conv(dplan, st, pflops, "dgrad")
                        self._wgrad(wplan, st, pflops, g + 4 * w0, 2 * C, None, 0)
                writes = [(r.g_off, C), (r.b_off, C)]
                if half == 0:  # the fused wgrad writes both (adjacent) weight-gradient slices
                    writes += [(pr["ra"].w_off, C * r.cin), (pr["rb"].w_off, C * r.cin)]
                self._add_bwd_op(op, writes)
            else:
                _, r, xin, out, y, res, up, ptrs, flops = rec
                self.conv_flops["wgrad"] += flops
                if not r.is_stem:
                    self.conv_flops["dgrad"] += flops
                C, npix = r.cout, y.npix
                k, s = (31, 1) if r.is_stem else (r.k, r.stride)
                # dy scratch of this layer: alternates between two buffers when the weight gradients run on the side
                # stream (wgrad of layer i reads dy[i % 2] while the BN backward of layer i+1 already writes the other one)
                slot = nlayer % len(self.dy)
                dy_ptr = self.dy[slot].data_ptr()
                evs = self._layer_events()
                dy_free = done_by_slot.get(slot)  # wgrad-done event of the previous user of this dy buffer
                done_by_slot[slot] = evs[1] if evs is not None else None
                nlayer += 1
                self._flush_pending(out)
                assert self._contrib_state(out), f"{r.name}: output gradient never produced"
                if up is not None:
                    assert self._contrib_state(up), f"{r.name}: upsampled gradient never produced"
                    up_bwd = L.yb_p32_upsample2x_bwd if self.parity else L.yb_upsample2x_bwd
                    self.bwd_ops.append(lambda st, g, up=up, out=out, up_bwd=up_bwd: _lib.check(
                        up_bwd(up.gptr, up.pitch, out.N, out.H, out.W, out.C, out.gptr, out.pitch, 1, st)))
                if res is not None:
                    res.buf.pending.append((res.c0, res.C, out))
                if self.parity:
                    acc = 0
                    if not r.is_stem:
                        src = self._take_pending_exact(xin)
                        written = self._contrib_state(xin)
                        if src is not None:  # identity (residual) contribution: copy / add it first, then accumulate the dgrad
                            self.bwd_ops.append(lambda st, g, src=src, xin=xin, a=1 if written else 0: _lib.check(
                                L.yb_p32_add_into(src.gptr, src.pitch, xin.gptr, xin.pitch, xin.npix, xin.C, a, st)))
                        acc = 1 if (src is not None or written) else 0
                        xin.buf.gw[xin.c0:xin.c0 + xin.C] = True
                    dplans, wplans = self._parity_bwd_plans(r, xin, dy_ptr, self._dy_elems, C, k, s, acc, ws, wsn)
                    mapp = net._stem_map.data_ptr() if r.is_stem else None
                    count = float(npix)

                    def op(st, g, r=r, out=out, y=y, ptrs=ptrs, C=C, npix=npix, dplans=dplans, wplans=wplans, mapp=mapp,
                           count=count, dy_ptr=dy_ptr):
                        rows = ctypes.c_int(0)
                        _lib.check(L.yb_p32_bn_act_bwd_reduce(out.gptr, out.pitch, y.ptr, C, npix, C, ptrs[1], ptrs[2], ptrs[3],
                                                              ptrs[4], redp, self.red_rows, ctypes.byref(rows), st))
                        _lib.check(L.yb_bn_bwd_finalize(redp, rows.value, C, count, g + 4 * r.g_off, g + 4 * r.b_off, coefp, 0, st))
                        _lib.check(L.yb_p32_bn_act_bwd_apply(out.gptr, out.pitch, y.ptr, C, npix, C, ptrs[1], ptrs[2], ptrs[3],
                                                             ptrs[4], coefp, None, dy_ptr, C, self._dy_elems, st))
                        for pl in dplans:
                            _lib.check(L.yb_plan_run(pl, st))
                        for n_, pl in enumerate(wplans):
                            _lib.check(L.yb_wgrad_plan_run(pl, g + 4 * r.w_off, C, mapp, 1 if n_ else 0, st))
                    self._add_bwd_op(op, [(r.g_off, C), (r.b_off, C), (r.w_off, r.conv.weight.numel())])
                    continue
                ws, wsn, wsi = self._next_ws()
                wplan = _lib.checkp(L.yb_conv_wgrad_plan(xin.ptr, xin.N, xin.H, xin.W, xin.C, xin.pitch, dy_ptr, C, C, k, s,
                                                         ws, wsn, 0))
                self._plan_ws[wplan] = wsi
                self.plans.append(wplan)
                dplan = None
                if not r.is_stem:
                    src = self._take_pending_exact(xin)
                    written = self._contrib_state(xin)
                    if src is not None and written:  # both an identity source and an earlier value: fold the identity first
                        self.bwd_ops.append(lambda st, g, src=src, xin=xin: _lib.check(
                            L.yb_add_into(src.gptr, src.pitch, xin.gptr, xin.pitch, xin.npix, xin.C, 1, st)))
                        src = None
                    if src is not None:
                        addp, addl = src.gptr, src.pitch
                    elif written:
                        addp, addl = xin.gptr, xin.pitch
                    else:
                        addp, addl = None, 0
                    dplan = _lib.checkp(L.yb_conv_dgrad_plan(dy_ptr, xin.N, xin.H, xin.W, C, C,
                                                             net._wdg.data_ptr() + 2 * r.wt_off, xin.C, k, s, xin.gptr,
                                                             xin.pitch, addp, addl, 0))
                    self.plans.append(dplan)
                    xin.buf.gw[xin.c0:xin.c0 + xin.C] = True
                mapp = net._stem_map.data_ptr() if r.is_stem else None
                rows = ctypes.c_int(0)
                count = float(npix)

                def op(st, g, r=r, out=out, y=y, ptrs=ptrs, C=C, npix=npix, wplan=wplan, dplan=dplan, mapp=mapp, rows=rows,
                       count=count, flops=flops, dy_ptr=dy_ptr, evs=evs, dy_free=dy_free):
                    self._ew("bn_bwd_reduce", 4.0 * npix * C, L.yb_bn_act_bwd_reduce, out.gptr, out.pitch, y.ptr, C, npix, C,
                             ptrs[1], ptrs[2], ptrs[3], ptrs[4], redp, ctypes.byref(rows), st)
                    _lib.check(L.yb_bn_bwd_finalize(redp, rows.value, C, count, g + 4 * r.g_off, g + 4 * r.b_off, coefp, 0, st))
                    if dy_free is not None and self._side_on:
                        self._main.wait_event(dy_free)  # the wgrad that still reads this dy buffer (two layers back)
                    self._ew("bn_bwd_apply", 6.0 * npix * C, L.yb_bn_act_bwd_apply, out.gptr, out.pitch, y.ptr, C, npix, C,
                             ptrs[1], ptrs[2], ptrs[3], ptrs[4], coefp, dy_ptr, C, st)
                    if dplan is not None:
                        self._conv(dplan, st, flops, "dgrad")
                    self._wgrad_async(evs, wplan, st, flops, g + 4 * r.w_off, C, mapp, 0)
                self._add_bwd_op(op, [(r.g_off, C), (r.b_off, C), (r.w_off, r.conv.weight.numel())])

    # -- execution
    def run_forward(self, x):
        L = self.L
        st = _lib.stream()
        dt = 0 if x.dtype == torch.float32 else 1
        Hs, Ws = x.shape[2], x.shape[3]
        if self.parity:
            # three fp32 NCHW planes of the image (each bf16-exact), each staged like the production input
            if (Hs, Ws) != (self.H, self.W):
                raise _lib.YBError("parity mode: multi-scale resampling is not available (resize the batch first)")
            xp = torch.empty((3,) + tuple(x.shape), device=self.dev, dtype=torch.float32)
            _lib.check(L.yb_p32_split_flat(x.data_ptr(), dt, x.numel(), xp[0].data_ptr(), xp[1].data_ptr(), xp[2].data_ptr(), st))
            for k in range(3):
                _lib.check(L.yb_prep_input(xp[k].data_ptr(), 0, self.B, self.H, self.W, self.x16.v().plptr(k), st))
        elif (Hs, Ws) == (self.H, self.W):
            _lib.check(L.yb_prep_input(x.data_ptr(), dt, self.B, self.H, self.W, self.x16.t.data_ptr(), st))
        else:  # multi-scale training: bilinear resample fused into the stem staging (training_utils.py:11-28)
            _lib.check(L.yb_prep_input_resized(x.data_ptr(), dt, self.B, Hs, Ws, self.H, self.W, self.x16.t.data_ptr(), st))
        self._run_fwd_ops(st)
        return self.outs

    def _run_fwd_ops(self, st):
        L = self.L
        if self._fold:
            # running statistics -> (scale, shift) of every conv epilogue, re-read on every forward like nn.BatchNorm2d does
            eps = tuple(float(f[7].eps) for f in self._fold)
            if eps != self._fold_eps:
                tab = np.zeros((len(self._fold), 8), np.int64)
                for i, f in enumerate(self._fold):
                    tab[i, :7] = f[:7]
                    tab[i, 7] = int(np.float32(eps[i]).view(np.int32))
                self._fold_dev, self._fold_eps = torch.from_numpy(tab).to(self.dev), eps
            _lib.check(L.yb_bn_fold_batch(self._fold_dev.data_ptr(), len(self._fold), st))
        for op in self.fwd_ops:
            op(st)

    def run_backward(self, gflat):
        st = _lib.stream()
        g = gflat.data_ptr()
        self._main = torch.cuda.current_stream(self.dev)
        self._side_on = self.side is not None and self.prof is None  # the per-kernel timing pass stays on one stream
        hook = self.on_grad_chunks if not self._side_on and self.prof is None else None
        # reductions on the second stream: not under the per-kernel timing pass, not with the whole wgrad already on a side
        # stream, and not when chunks of the bucket are handed to an overlapped all-reduce as soon as the main stream wrote them
        self._rs_on = self.rstream is not None and self.prof is None and not self._side_on and hook is None
        self._ws_pending = [False, False]
        for i, op in enumerate(self.bwd_ops):
            op(st, g)
            if hook is not None and i in hook[0]:
                for c in hook[0][i]:
                    hook[1](c, gflat)  # this chunk of the bucket is final: its all-reduce overlaps the remaining ops
        if self._side_on:
            self._main.wait_stream(self.side)  # every weight gradient is in the bucket before all-reduce / optimiser
        if self._rs_on:
            for b in range(2):
                if self._ws_pending[b]:
                    self._main.wait_event(self._ws_red_ev[b])
            self._ws_pending = [False, False]

    def __del__(self):
        try:
            for p in self.plans:
                self.L.yb_plan_destroy(p)
        except Exception:
            pass


class _SubEngine(_Engine):
    """Inference plan of ONE block (CBL / Bottleneck / C3 / SPPF) called on its own (reference model.py:27,49,89,106):
    NCHW float input -> NHWC bf16 -> the block's fused kernels with folded BatchNorm -> NCHW float32."""

    def __init__(self, net, mod, B, C, H, W):
        self.mod, self.cin = mod, C
        super().__init__(net, B, H, W, False, False)

    def _build(self):
        net, mod, B, H, W, C = self.net, self.mod, self.B, self.H, self.W, self.cin
        self.outs, self.x16, self.xin = [], None, None
        if isinstance(mod, CBL):
            r = net._rec_of[mod.cbl[0]]
            if r.is_stem:  # the 6x6/s2 stem reads the image through the space-to-depth staging (yb_prep_input)
                if C != 3 or H % 2 or W % 2:
                    raise ValueError(f"stem CBL: expected (B, 3, even H, even W), got C={C} H={H} W={W}")
                self.x16 = self.buf(B, H // 2, W // 2 + 2, 16, grad=False)
                xin, Ho, Wo = _StemView(self.x16), H // 2, W // 2
            else:
                if C != r.cin or H % r.stride or W % r.stride:
                    raise ValueError(f"CBL: expected {r.cin} input channels and a size divisible by {r.stride}")
                self.xin = self.buf(B, H, W, C, grad=False)
                xin, Ho, Wo = self.xin.v(), H // r.stride, W // r.stride
            self.out = self.buf(B, Ho, Wo, r.cout, grad=False)
            self.cbl(mod, xin, self.out.v())
            return
        cin = net._rec_of[mod.c1.cbl[0]].cin
        if C != cin:
            raise ValueError(f"{type(mod).__name__}: expected {cin} input channels, got {C}")
        self.xin = self.buf(B, H, W, C, grad=False)
        if isinstance(mod, Bottleneck):  # model.py:49: c2(c1(x)) + x
            cout = net._rec_of[mod.c2.cbl[0]].cout
            h = self.buf(B, H, W, net._rec_of[mod.c1.cbl[0]].cout, grad=False)
            self.out = self.buf(B, H, W, cout, grad=False)
            self.cbl(mod.c1, self.xin.v(), h.v())
            self.cbl(mod.c2, h.v(), self.out.v(), res=self.xin.v())
        else:
            self.out = self.buf(B, H, W, net._rec_of[mod.c_out.cbl[0]].cout, grad=False)
            (self.c3 if isinstance(mod, C3) else self.sppf)(mod, self.xin.v(), self.out.v())

    def run_forward(self, x):
        L, st = self.L, _lib.stream()
        if self.x16 is not None:
            dt = 0 if x.dtype == torch.float32 else 1
            _lib.check(L.yb_prep_input(x.data_ptr(), dt, self.B, self.H, self.W, self.x16.t.data_ptr(), st))
        else:
            self.xin.t.copy_(x.permute(0, 2, 3, 1))  # layout + dtype plumbing (fp32 NCHW -> bf16 NHWC)
        self._run_fwd_ops(st)
        return self.out.t.permute(0, 3, 1, 2).float()


class _NetFn(torch.autograd.Function):
    """Autograd boundary: forward / backward of the whole network are the engine's launch lists."""

    @staticmethod
    def forward(ctx, net, eng, x, *params):
        outs = eng.run_forward(x)
        ctx.net, ctx.eng = net, eng
        eng.head_ready = False
        res = tuple(o.view_as(o) for o in outs)  # fresh autograd outputs over the engine's head tensors
        return res

    @staticmethod
    def backward(ctx, *gouts):
        net, eng = ctx.net, ctx.eng
        L = _lib.lib()
        st = _lib.stream()
        # fast path (eng.head_ready): ComputeLoss.backward already wrote dL/dp into eng.head_dy and returned stride-0 zero
        # sentinels.  Anything else that reaches an output (an auxiliary term, a second loss on the same outputs) arrives
        # here as a dense gradient -- possibly summed with a sentinel by autograd -- and is ADDED to head_dy.
        fast = eng.head_ready
        for i, go in enumerate(gouts):
            dyh = eng.head_dy[i]
            sentinel = go is not None and go.numel() > 1 and all(s == 0 for s in go.stride())
            if go is None or sentinel:
                if not fast:
                    dyh.zero_()
                    if eng.parity:
                        eng.head_dy_pl[i].zero_()
                continue
            go = go.contiguous().float()
            B, na, H, W, no = go.shape
            if eng.parity:
                _lib.check(L.yb_p32_head_grad_pack(go.data_ptr(), B, na, H, W, no, dyh.data_ptr(), eng.head_dy_pl[i].data_ptr(),
                                                   HEAD_PAD, dyh.numel(), 0, st))
                continue
            _lib.check(L.yb_head_grad_pack(go.data_ptr(), B, na, H, W, no, dyh.data_ptr(), HEAD_PAD, 1 if fast else 0, st))
        eng.head_ready = False
        if net._accumulate_grads and not net.expose_param_grads:
            # gradient accumulation on the fused-optimiser path (trainer.TrainStep(accumulate=k)): this micro-batch's
            # gradients go to the second flat bucket and are added into the first
            if net._gflat[1] is None:
                net._gflat[1] = torch.zeros_like(net._pflat)
            gflat = net._gflat[1]
            eng.run_backward(gflat)
            _lib.check(L.yb_accumulate_f32(net.flat_grads.data_ptr(), gflat.data_ptr(), gflat.numel(), st))
            return (None, None, None) + (None,) * len(net._poffs)
        gflat = net._grad_target()
        eng.run_backward(gflat)
        if not net.expose_param_grads:  # fused optimiser path: the flat bucket is consumed directly (trainer.py)
            return (None, None, None) + (None,) * len(net._poffs)
        grads = tuple(net._grad_views(gflat))
        return (None, None, None) + grads


# ------------------------------------------------------------------------------------------------ the network
class YOLOV5m(nn.Module):
    """Same signature as the reference: YOLOV5m(first_out, nc=80, anchors=(), ch=(), inference=False)  (model.py:178-180)."""

    def __init__(self, first_out, nc=80, anchors=(), ch=(), inference=False):
        super().__init__()
        c = first_out
        self.inference = inference
        self._first_out = c
        self.backbone = nn.ModuleList([
            CBL(3, c, 6, 2), CBL(c, c * 2, 3, 2), C3(c * 2, c * 2, 0.5, 2), CBL(c * 2, c * 4, 3, 2),
            C3(c * 4, c * 4, 0.5, 4), CBL(c * 4, c * 8, 3, 2), C3(c * 8, c * 8, 0.5, 6), CBL(c * 8, c * 16, 3, 2),
            C3(c * 16, c * 16, 0.5, 2), SPPF(c * 16, c * 16)])
        self.neck = nn.ModuleList([
            CBL(c * 16, c * 8, 1, 1), C3(c * 16, c * 8, 0.25, 2, backbone=False), CBL(c * 8, c * 4, 1, 1),
            C3(c * 8, c * 4, 0.25, 2, backbone=False), CBL(c * 4, c * 4, 3, 2), C3(c * 8, c * 8, 0.5, 2, backbone=False),
            CBL(c * 8, c * 8, 3, 2), C3(c * 16, c * 16, 0.5, 2, backbone=False)])
        self.head = HEADS(nc=nc, anchors=anchors, ch=ch)
        assert tuple(ch) == (c * 4, c * 8, c * 16), "ch must be (4, 8, 16) x first_out (reference call sites)"
        assert (5 + nc) * self.head.naxs <= HEAD_PAD, "head width exceeds the padded gradient width"
        assert c % 16 == 0, "first_out must be a multiple of 16 (bf16 NHWC channel alignment)"
        self._engines = {}
        self._packed_sig = None
        self._gflat = [None, None]
        # True: backward publishes p.grad views of the flat gradient bucket (stock torch optimisers work);
        # False: gradients stay only in `flat_grads` (yolov5m_b200.trainer.Adam reads the bucket) -- saves 243 view objects
        self.expose_param_grads = True
        # fp32 parity mode (csrc/parity.cu): fp32 activations / gradients, every conv = six bf16-split passes of the same
        # tcgen05 kernels.  A numerics instrument for BASELINE config 2 ("fp32 vs reference", 1e-3), not the fast path.
        self.parity = os.environ.get("YB_PARITY", "0") == "1"
        self._accumulate_grads = False  # set by trainer.TrainStep around the backward of an accumulated micro-batch
        for m in self.modules():  # blocks called on their own route to run_submodule (registry: nothing added to the modules)
            if isinstance(m, (CBL, Bottleneck, C3, SPPF)):
                _OWNERS[id(m)] = self
        self._flatten()

    # -- flat parameter storage -----------------------------------------------------------------
    def _flatten(self):
        """(re)build the flat fp32 master buffer and point every Parameter at its slice."""
        params = list(self.parameters())
        dev = params[0].device
        recs, rec_of = [], {}
        off = 0
        offs = {}
        # storage order = parameter order, except that a C3's c_skipped conv weight sits right behind its c1 conv weight:
        # the two 1x1 convs read the same input (model.py:90-92) and run as ONE GEMM over the [2c_][Cin] matrix they then form
        # (forward operand, weight gradient and Adam state all become contiguous; state_dict() order is unaffected)
        order = list(params)
        for m in self.modules():
            if isinstance(m, C3):
                a, b = m.c1.cbl[0].weight, m.c_skipped.cbl[0].weight
                ib = next(i for i, q in enumerate(order) if q is b)
                order.pop(ib)
                ia = next(i for i, q in enumerate(order) if q is a)
                order.insert(ia + 1, b)
        for p in order:
            offs[id(p)] = off
            off += (p.numel() + 3) // 4 * 4  # 16-byte aligned slices
        flat = torch.zeros(off, device=dev, dtype=torch.float32)
        for p in params:
            o, n = offs[id(p)], p.numel()
            if p.dim() == 4:
                co, ci, kh, kw = p.shape
                view = flat[o:o + n].view(co, kh, kw, ci).permute(0, 3, 1, 2)  # channels-last: memory = [Cout][tap][Cin]
            else:
                view = flat[o:o + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self._pflat = flat
        self._poffs = [(offs[id(p)], p.numel()) for p in params]
        wt_off = 0
        for name, m in self.named_modules():
            if isinstance(m, CBL):
                conv, bn = m.cbl[0], m.cbl[1]
                r = _LayerRec()
                r.name, r.conv, r.bn = name, conv, bn
                r.cin, r.cout, r.k, r.stride = conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride
                r.is_stem, r.is_head = conv.kernel_size == 6, False
                r.w_off, r.g_off, r.b_off = offs[id(conv.weight)], offs[id(bn.weight)], offs[id(bn.bias)]
                r.bias_off = -1
                r.rm, r.rv, r.nbt = bn.running_mean, bn.running_var, bn.num_batches_tracked
                r.cout_pad = r.cout
            elif isinstance(m, _Conv) and m.bias is not None:
                r = _LayerRec()
                r.name, r.conv, r.bn = name, m, None
                r.cin, r.cout, r.k, r.stride = m.in_channels, m.out_channels, 1, 1
                r.is_stem, r.is_head = False, True
                r.w_off, r.bias_off, r.g_off, r.b_off = offs[id(m.weight)], offs[id(m.bias)], -1, -1
                r.rm = r.rv = r.nbt = None
                r.cout_pad = HEAD_PAD
            else:
                continue
            if r.is_stem:
                r.wt_off = -1
            else:
                r.wt_off = wt_off
                wt_off += r.cin * r.k * r.k * r.cout_pad
            recs.append(r)
            rec_of[r.conv] = r
        # fused c1 || c_skipped pairs: one more dgrad operand [Cin][2c_] per C3
        self._pairs = {}
        for m in self.modules():
            if isinstance(m, C3):
                ra, rb = rec_of[m.c1.cbl[0]], rec_of[m.c_skipped.cbl[0]]
                if rb.w_off == ra.w_off + ra.cout * ra.cin and (ra.cout * ra.cin) % 4 == 0:
                    self._pairs[m] = (ra, rb, wt_off)
                    wt_off += ra.cin * 2 * ra.cout
        self._recs, self._rec_of, self._wdg_elems = recs, rec_of, wt_off
        self._engines = {}
        self._packed_sig = None
        self._gflat = [None, None]
        if dev.type == "cuda":
            self._alloc_device_side()

    def _alloc_device_side(self):
        dev = self._pflat.device
        self._wfwd = torch.zeros(self._pflat.numel(), device=dev, dtype=torch.bfloat16)
        self._wdg = torch.zeros(self._wdg_elems, device=dev, dtype=torch.bfloat16)
        stem = next(r for r in self._recs if r.is_stem)
        self._wstem = torch.zeros(stem.cout * 9 * 16, device=dev, dtype=torch.bfloat16)
        rows, cum = [], 0
        for r in self._recs:
            if r.is_stem:
                continue
            n = r.cin * r.k * r.k * r.cout_pad
            rows.append([r.w_off, r.wt_off, r.cout, r.k * r.k, r.cin, r.cout_pad, cum, cum + n])
            cum += n
        for (ra, rb, wt2) in self._pairs.values():
            n = ra.cin * 2 * ra.cout
            rows.append([ra.w_off, wt2, 2 * ra.cout, 1, ra.cin, 2 * ra.cout, cum, cum + n])
            cum += n
        self._dg_table = torch.tensor(rows, dtype=torch.int64, device=dev)
        # stem wgrad: packed 3x3/16ch gradient index -> offset inside the [Cout][6][6][3] master slice
        co, tap, ch = np.meshgrid(np.arange(stem.cout), np.arange(9), np.arange(16), indexing="ij")
        c, rs = ch % 3, ch // 3
        rr, ss, a, b = rs >> 1, rs & 1, tap // 3, tap % 3
        m = ((co * 6 + 2 * a + rr) * 6 + 2 * b + ss) * 3 + c
        m[ch >= 12] = -1
        self._stem_map = torch.from_numpy(m.reshape(-1).astype(np.int32)).to(dev)
        self._stem_rec = stem

    def _apply(self, fn, recurse=True):
        super()._apply(fn)
        p0 = next(self.parameters())
        if p0.dtype != torch.float32:
            raise TypeError("YOLOV5m (B200): master weights are fp32 and compute is bf16; dtype casts of the module are not supported")
        self._flatten()
        return self

    def _param_signature(self):
        return (sum(p._version for p in self.parameters()), self.parity)

    def refresh_packed(self, force=False):
        """bf16 tensor-core operands derived from the fp32 master weights (re-run after any parameter update)."""
        sig = self._param_signature()
        if not force and sig == self._packed_sig:
            return
        L, st = _lib.lib(), _lib.stream()
        _lib.check(L.yb_cast_bf16(self._pflat.data_ptr(), self._wfwd.data_ptr(), self._pflat.numel(), st))
        self._repack_derived(st)
        if self.parity:
            self._pack_parity(st)
        self._packed_sig = sig

    def _pack_parity(self, st):
        """parity mode: the three bf16 planes of every tensor-core operand (forward, dgrad and stem layouts), produced by
        the SAME packing kernels from the three bf16-exact fp32 planes of the master weights"""
        L, n = _lib.lib(), self._pflat.numel()
        dev = self._pflat.device
        if getattr(self, "_wfwd_pl", None) is None or self._wfwd_pl[0].device != dev:
            self._wfwd_pl = [torch.zeros_like(self._wfwd) for _ in range(3)]
            self._wdg_pl = [torch.zeros_like(self._wdg) for _ in range(3)]
            self._wstem_pl = [torch.zeros_like(self._wstem) for _ in range(3)]
        pf = torch.empty(3, n, device=dev, dtype=torch.float32)
        _lib.check(L.yb_p32_split_flat(self._pflat.data_ptr(), 0, n, pf[0].data_ptr(), pf[1].data_ptr(), pf[2].data_ptr(), st))
        s = self._stem_rec
        for k in range(3):
            _lib.check(L.yb_cast_bf16(pf[k].data_ptr(), self._wfwd_pl[k].data_ptr(), n, st))
            _lib.check(L.yb_repack_dgrad(pf[k].data_ptr(), self._wdg_pl[k].data_ptr(), self._dg_table.data_ptr(),
                                         self._dg_table.shape[0], self._wdg_elems, st))
            _lib.check(L.yb_repack_stem(pf[k].data_ptr() + 4 * s.w_off, self._wstem_pl[k].data_ptr(), s.cout, st))

    def _repack_derived(self, st):
        L = _lib.lib()
        _lib.check(L.yb_repack_dgrad(self._pflat.data_ptr(), self._wdg.data_ptr(), self._dg_table.data_ptr(),
                                     self._dg_table.shape[0], self._wdg_elems, st))
        s = self._stem_rec
        _lib.check(L.yb_repack_stem(self._pflat.data_ptr() + 4 * s.w_off, self._wstem.data_ptr(), s.cout, st))

    # -- gradients ---------------------------------------------------------------------------------
    def _grad_target(self):
        """flat fp32 gradient buffer the next backward writes: buffer 0 unless live .grad tensors alias it
        (gradient accumulation across backward calls), then buffer 1."""
        if self._gflat[0] is None:
            self._gflat[0] = torch.zeros_like(self._pflat)
        base = self._gflat[0].untyped_storage().data_ptr()
        for p in self.parameters():
            if p.grad is not None and p.grad.untyped_storage().data_ptr() == base:
                if self._gflat[1] is None:
                    self._gflat[1] = torch.zeros_like(self._pflat)
                return self._gflat[1]
        return self._gflat[0]

    def _grad_views(self, gflat):
        out = []
        for p, (o, n) in zip(self.parameters(), self._poffs):
            if p.dim() == 4:
                co, ci, kh, kw = p.shape
                out.append(gflat[o:o + n].view(co, kh, kw, ci).permute(0, 3, 1, 2))
            else:
                out.append(gflat[o:o + n].view(p.shape))
        return out

    @property
    def flat_params(self):
        return self._pflat

    @property
    def flat_grads(self):
        if self._gflat[0] is None:
            self._gflat[0] = torch.zeros_like(self._pflat)
        return self._gflat[0]

    # -- forward -------------------------------------------------------------------------------------
    # engines (buffers + launch plans) are cached per (batch, height, width, mode).  Multi-scale training visits up to 11
    # square sizes (320..640 step 32): at bs=64 all of them together hold ~120 GB of activations, so the cache is bounded
    # by bytes (YB_ENGINE_CACHE_GB, default 100 of the 180 GB), oldest first.
    _ENGINE_CACHE_BYTES = int(float(os.environ.get("YB_ENGINE_CACHE_GB", "100")) * (1 << 30))

    def engine(self, B, H, W, train):
        key = (B, H, W, bool(train), bool(self.parity))
        e = self._engines.get(key)
        if e is None:
            e = _Engine(self, B, H, W, train, bool(self.parity))
            while self._engines and sum(v.nbytes for v in self._engines.values()) + e.nbytes > self._ENGINE_CACHE_BYTES:
                self._engines.pop(next(iter(self._engines)))
            self._engines[key] = e
        else:
            self._engines[key] = self._engines.pop(key)  # most recently used last
        return e

    def run_submodule(self, mod, x):
        """``mod(x)`` for a CBL / Bottleneck / C3 / SPPF block of this network (reference model.py:27,49,89,106), e.g.
        ``model.backbone[0](img)`` for feature inspection.  Inference only: BatchNorm uses its running statistics and the
        result carries no autograd graph -- a block in training mode under grad raises instead of returning a tensor that
        silently would not train (train through YOLOV5m.forward).  bf16 arithmetic like the full forward."""
        if mod.training and torch.is_grad_enabled():
            raise _lib.YBError("sub-module calls are inference-only on the B200 engine: use model.eval() / torch.no_grad(), "
                               "or train through YOLOV5m.forward")
        if not (torch.is_tensor(x) and x.is_cuda and x.dim() == 4 and x.device == self._pflat.device):
            raise _lib.YBError("sub-module input must be a (B, C, H, W) CUDA tensor on the model's device (no CPU fallback)")
        if x.dtype not in (torch.float32, torch.uint8):
            x = x.float()
        x = x.contiguous()
        B, C, H, W = (int(v) for v in x.shape)
        with torch.cuda.device(self._pflat.device), torch.no_grad():
            self.refresh_packed()
            key = ("sub", id(mod), B, C, H, W)
            eng = self._engines.get(key)
            if eng is None:
                eng = _SubEngine(self, mod, B, C, H, W)
                while self._engines and sum(v.nbytes for v in self._engines.values()) + eng.nbytes > self._ENGINE_CACHE_BYTES:
                    self._engines.pop(next(iter(self._engines)))
                self._engines[key] = eng
            else:
                self._engines[key] = self._engines.pop(key)
            return eng.run_forward(x)

    def forward(self, x, size=None):
        """x: (B,3,H,W) float32 in [0,1] or uint8.  ``size=(h, w)`` (extension, both % 32 == 0): run the network on the
        image bilinearly resampled to (h, w) -- the reference's multi_scale() (training_utils.py:11-28, :100) without
        materialising the resized float image (the resample is fused into the stem's input staging)."""
        H, W = (int(size[0]), int(size[1])) if size is not None else (x.shape[2], x.shape[3])
        assert H % 32 == 0 and W % 32 == 0, "Width and Height aren't divisible by 32!"  # model.py:211
        if not (torch.is_tensor(x) and x.is_cuda and self._pflat.is_cuda):
            raise _lib.YBError("YOLOV5m (B200): model and input must be on a CUDA device (no CPU fallback)")
        if x.dtype not in (torch.float32, torch.uint8):
            x = x.float()
        if x.device != self._pflat.device:
            raise _lib.YBError(f"YOLOV5m (B200): input on {x.device} but the model lives on {self._pflat.device}")
        x = x.contiguous()
        B = x.shape[0]
        # every launch goes to the current stream of the MODEL's device, whatever the process's current device is
        with torch.cuda.device(self._pflat.device):
            self.refresh_packed()
            train = self.training
            eng = self.engine(B, H, W, train)
            if train and torch.is_grad_enabled():
                outs = _NetFn.apply(self, eng, x, *self.parameters())
                outs = list(outs)
                for o in outs:
                    o._yb_engine = eng
                return outs
            with torch.no_grad():
                outs = eng.run_forward(x)
            return [o.clone() for o in outs] if not train else list(outs)
