"""ComputeLoss: drop-in for the reference loss (reference ultralytics_loss.py:17-311) on hand-written sm_100a kernels.

Same constructor ``ComputeLoss(model, save_logs=False, filename=None, resume=False)``, same call
``loss_fn(p, targets, pred_size, batch_idx=None, epoch=None) -> Tensor of shape (1,)`` (differentiable w.r.t. ``p``),
same ``build_targets(p, targets) -> (tcls, tbox, indices, anchors)``, same CSV side effect when ``save_logs``.

Underneath (csrc/loss.cu): one ordered-compaction kernel builds the target rows of all three levels in the reference's
row order (bit-exact indices), one warp-per-row kernel does gather + decode + GIoU + class BCE, one dense pass does the
objectness BCE with "last write wins" tobj, and the backward pass writes dL/dp either as fp32 tensors (generic autograd)
or -- when ``p`` are the live outputs of a :class:`yolov5m_b200.model.YOLOV5m` -- straight into the head convolutions'
bf16 gradient operand, so no dense fp32 gradient ever exists.  There is no CPU / PyTorch fallback.
"""
import csv
import ctypes
import os

import torch

from . import _lib

IMAGE_SIZE = 640  # reference config.py:24 (scales lambda_obj, ultralytics_loss.py:32)


class _Level(ctypes.Structure):  # mirrors yb_loss_level in include/yolov5m_b200.h
    _fields_ = [("p", ctypes.c_void_p), ("H", ctypes.c_int32), ("W", ctypes.c_int32), ("balance", ctypes.c_float),
                ("pad_", ctypes.c_int32), ("idx", ctypes.c_void_p), ("tbox", ctypes.c_void_p), ("anch", ctypes.c_void_p),
                ("tcls", ctypes.c_void_p), ("row_val", ctypes.c_void_p), ("row_grad", ctypes.c_void_p),
                ("row_prev", ctypes.c_void_p), ("cell_head", ctypes.c_void_p), ("obj_partial", ctypes.c_void_p),
                ("grad_f32", ctypes.c_void_p), ("grad_bf16", ctypes.c_void_p)]


class _Workspace:
    """device scratch of one loss evaluation (kept alive until its backward has run)."""

    def __init__(self, shapes, cap, no, dev, obj_rows):
        self.shapes, self.cap, self.busy = shapes, cap, False
        nl = len(shapes)
        self.counts = torch.zeros(nl, dtype=torch.int32, device=dev)
        self.out4 = torch.zeros(4, dtype=torch.float32, device=dev)
        self.idx = [torch.zeros(4, cap, dtype=torch.int64, device=dev) for _ in range(nl)]
        self.tbox = [torch.zeros(cap, 4, dtype=torch.float32, device=dev) for _ in range(nl)]
        self.anch = [torch.zeros(cap, 2, dtype=torch.float32, device=dev) for _ in range(nl)]
        self.tcls = [torch.zeros(cap, dtype=torch.int64, device=dev) for _ in range(nl)]
        self.row_val = [torch.zeros(3, cap, dtype=torch.float32, device=dev) for _ in range(nl)]
        self.row_grad = [torch.zeros(cap, no, dtype=torch.float32, device=dev) for _ in range(nl)]
        self.row_prev = [torch.zeros(cap, dtype=torch.int32, device=dev) for _ in range(nl)]
        self.cell_head = [torch.zeros(B * na * H * W, dtype=torch.int32, device=dev) for (B, na, H, W, _) in shapes]
        self.obj_partial = [torch.zeros(obj_rows, dtype=torch.float32, device=dev) for _ in range(nl)]
        self.levels = (_Level * nl)()
        for i, (B, na, H, W, _) in enumerate(shapes):
            lv = self.levels[i]
            lv.H, lv.W = H, W
            lv.idx, lv.tbox, lv.anch, lv.tcls = (self.idx[i].data_ptr(), self.tbox[i].data_ptr(), self.anch[i].data_ptr(),
                                                 self.tcls[i].data_ptr())
            lv.row_val, lv.row_grad, lv.row_prev = self.row_val[i].data_ptr(), self.row_grad[i].data_ptr(), self.row_prev[i].data_ptr()
            lv.cell_head, lv.obj_partial = self.cell_head[i].data_ptr(), self.obj_partial[i].data_ptr()


class _Lease:
    """Marks a workspace busy for as long as a backward pass can still need it.  Held by the autograd node (ctx): released by
    backward, or -- when the graph is dropped without a backward (validation loss, logging) -- by the node's destruction, so
    forward-only evaluations never strand a workspace in the pool."""

    def __init__(self, ws):
        self.ws = ws
        ws.busy = True

    def release(self):
        if self.ws is not None:
            self.ws.busy = False
            self.ws = None

    __del__ = release


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, owner, targets, *p):
        ws, fast = owner._run_forward(p, targets)
        ctx.owner, ctx.ws, ctx.fast = owner, ws, fast
        ctx.shapes = [tuple(t.shape) for t in p]
        ctx.p = p  # the logits are re-read by the backward kernel (objectness sigmoid)
        # only a differentiable evaluation keeps the workspace (under no_grad / detached inputs nothing can call backward)
        ctx.lease = _Lease(ws) if any(ctx.needs_input_grad[2:]) else None
        owner.last_parts = ws.out4[1:4].clone()
        return ws.out4[0:1].clone()

    @staticmethod
    def backward(ctx, gout):
        owner, ws = ctx.owner, ctx.ws
        L, st = _lib.lib(), _lib.stream()
        gout = gout.to(torch.float32).contiguous()
        grads = []
        eng = ctx.fast
        if eng is not None and eng.head_ready:
            eng = None  # a second loss on the same outputs: hand autograd a dense gradient, _NetFn.backward adds it in
        for i, pi in enumerate(ctx.p):
            lv = ws.levels[i]
            lv.p = pi.data_ptr()
            if eng is not None:
                lv.grad_f32, lv.grad_bf16 = None, eng.head_dy[i].data_ptr()
                grads.append(torch.zeros((), dtype=pi.dtype, device=pi.device).expand(pi.shape))  # sentinel: see model._NetFn
            else:
                g = torch.empty_like(pi)
                lv.grad_f32, lv.grad_bf16 = g.data_ptr(), None
                grads.append(g)
        B, na, _, _, no = ctx.shapes[0]
        cpad = eng.head_dy[0].shape[-1] if eng is not None else 0
        _lib.check(L.yb_loss_bwd(ws.levels, len(ctx.p), B, na, no, ws.cap, ws.counts.data_ptr(), owner._nobj_ptr(ws), owner.lambda_box,
                                 owner.lambda_obj, owner.lambda_class, gout.data_ptr(), cpad, st))
        if eng is not None:
            eng.head_ready = True
        if ctx.lease is not None:
            ctx.lease.release()
        return (None, None) + tuple(grads)


class ComputeLoss:
    sort_obj_iou = False
    nan_on_empty = False   # ComputeLoss: a level without targets contributes nothing (ultralytics_loss.py:76)
    rows_per_target = None  # row capacity per target row: 5 * na (ultralytics_loss.py:248), set in _workspace

    def _nobj_ptr(self, ws):
        return None  # every row is an object row: the means divide by counts

    def __init__(self, model, save_logs=False, filename=None, resume=False, image_size=IMAGE_SIZE):
        device = next(model.parameters()).device
        # hyper-parameters: ultralytics_loss.py:31-41
        self.lambda_class = 0.5 * (model.head.nc / 80 * 3 / model.head.nl)
        self.lambda_obj = 1 * ((image_size / 640) ** 2 * 3 / model.head.nl)
        self.lambda_box = 0.05 * (3 / model.head.nl)
        self.anchor_t = 4.0
        self.balance = [4.0, 1.0, 0.4]
        self.na, self.nc, self.nl = model.head.naxs, model.head.nc, model.head.nl
        self.anchors = model.head.anchors
        self.device = device
        self.save_logs, self.filename = save_logs, filename
        self.last_parts = None
        self._ws = {}
        if self.save_logs and not resume:  # ultralytics_loss.py:46-58
            folder = os.path.join("train_eval_metrics", filename)
            os.makedirs(folder, exist_ok=True)
            with open(os.path.join(folder, "loss.csv"), "w") as f:
                csv.writer(f).writerow(["epoch", "batch_idx", "box_loss", "object_loss", "class_loss"])

    # -- plumbing ---------------------------------------------------------------------------------------------------
    def _prep(self, p, targets):
        if len(p) != self.nl:
            raise ValueError(f"ComputeLoss: expected {self.nl} prediction levels, got {len(p)}")
        for t in p:
            if not (torch.is_tensor(t) and t.is_cuda):
                raise _lib.YBError("ComputeLoss (B200): predictions must be CUDA tensors (no CPU fallback)")
        dev = p[0].device
        targets = torch.as_tensor(targets).to(device=dev, dtype=torch.float32, non_blocking=True).reshape(-1, 6).contiguous()
        return dev, targets

    def _workspace(self, shapes, nt, dev, need_free):
        nt_cap = 64
        while nt_cap < nt:
            nt_cap *= 2
        cap = 5 * self.na * nt_cap
        key = (tuple(shapes), cap)
        pool = self._ws.setdefault(key, [])
        for ws in pool:
            if not (need_free and ws.busy):
                return ws
        ws = _Workspace(tuple(shapes), cap, shapes[0][4], dev, _lib.lib().yb_loss_obj_rows())
        pool.append(ws)
        return ws

    def _anchors_dev(self, dev):
        a = self.anchors
        if a.device != dev or a.dtype != torch.float32 or not a.is_contiguous():
            a = a.to(device=dev, dtype=torch.float32).contiguous()
        return a

    def _build(self, ws, p, targets, dev):
        L, st = _lib.lib(), _lib.stream()
        for i, pi in enumerate(p):
            ws.levels[i].p = pi.data_ptr()
            ws.levels[i].balance = self.balance[i]
        anchors = self._anchors_dev(dev)
        _lib.check(L.yb_build_targets(targets.data_ptr(), targets.shape[0], anchors.data_ptr(), ws.levels, self.nl, self.na,
                                      self.anchor_t, ws.cap, ws.counts.data_ptr(), st))
        ws._keep = (targets, anchors)

    def _run_forward(self, p, targets):
        dev = p[0].device
        shapes = [tuple(t.shape) for t in p]
        ws = self._workspace(shapes, targets.shape[0], dev, need_free=True)
        self._build(ws, p, targets, dev)
        L, st = _lib.lib(), _lib.stream()
        B, na, _, _, no = shapes[0]
        _lib.check(L.yb_loss_fwd(ws.levels, self.nl, B, na, no, ws.cap, ws.counts.data_ptr(), self._nobj_ptr(ws),
                                 1 if self.nan_on_empty else 0, self.lambda_box, self.lambda_obj,
                                 self.lambda_class, ws.out4.data_ptr(), st))
        # fast path: p are the live head tensors of a YOLOV5m engine -> backward writes its bf16 gradient operand
        eng = getattr(p[0], "_yb_engine", None)
        if eng is not None:
            ok = eng.train and not eng.parity and all(getattr(t, "_yb_engine", None) is eng and t.data_ptr() == o.data_ptr()
                                   for t, o in zip(p, eng.outs))
            eng = eng if ok else None
        return ws, eng

    # -- reference API ------------------------------------------------------------------------------------------------
    def __call__(self, p, targets, pred_size=None, batch_idx=None, epoch=None):
        dev, targets = self._prep(p, targets)
        pc = []
        for t in p:
            t2 = t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()
            if t2 is not t and hasattr(t, "_yb_engine"):
                t2._yb_engine = None
            pc.append(t2)
        with torch.cuda.device(dev):  # launches go to the predictions' device, whatever the current device is
            loss = _LossFn.apply(self, targets, *pc)
        if self.save_logs and batch_idx is not None and batch_idx % 100 == 0:  # ultralytics_loss.py:108-116
            lbox, lobj, lcls = (float(v) for v in self.last_parts.tolist())
            with open(os.path.join("train_eval_metrics", self.filename, "loss.csv"), "a") as f:
                csv.writer(f).writerow([epoch, batch_idx, lbox, lobj, lcls])
        return loss

    def build_targets(self, p, targets):
        """ultralytics_loss.py:122-311 -> (tcls, tbox, indices, anchors), each a list over levels."""
        dev, targets = self._prep(p, targets)
        shapes = [tuple(t.shape) for t in p]
        with torch.cuda.device(dev):
            ws = self._workspace(shapes, targets.shape[0], dev, need_free=True)
            self._build(ws, p, targets, dev)
        counts = ws.counts.tolist()
        tcls, tbox, indices, anch = [], [], [], []
        for i, n in enumerate(counts):
            idx = ws.idx[i][:, :n].clone()
            indices.append((idx[0], idx[1], idx[2], idx[3]))
            tbox.append(ws.tbox[i][:n].clone())
            anch.append(ws.anch[i][:n].clone())
            tcls.append(ws.tcls[i][:n].clone())
        return tcls, tbox, indices, anch
