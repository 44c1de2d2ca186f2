"""YOLO_LOSS: drop-in for the reference's default loss (reference loss.py:20-246; train.py:102-106 picks it whenever
``--ultralytics_loss`` is absent) on the sm_100a kernels of csrc/yolo_loss.cu + csrc/loss.cu.

Same constructor ``YOLO_LOSS(model, rect_training, save_logs=False, filename=None, resume=False)``, same call
``loss_fn(preds, targets, pred_size, batch_idx=None, epoch=None)`` with ``targets`` = one ``(n_i, 5)`` array
``[class, x, y, w, h]`` per image (what ``Training_Dataset.collate_fn`` yields, dataset.py:199-202), same
``build_targets(input_tensor, bboxes, pred_size)`` -> list of three ``(3, H, W, 6)`` target tensors.

What replaces what:
  * the per-box Python loop of ``build_targets`` (loss.py:101-192: nine-anchor IoU ranking, sequential "cell taken" /
    "scale has anchor" / "ignore" rules) runs as ONE kernel for the whole batch, one thread per image (boxes of an image are
    sequential by definition), float64 arithmetic like the reference (np.loadtxt labels);
  * ``compute_loss`` (loss.py:195-246) has ComputeLoss's form (GIoU box term, dense objectness BCE with
    ``tobj = GIoU.clamp(0)``, class BCE, per-level balance, ``* bs``), so it runs on the same fused kernels; the only new
    ingredient is the "ignore" rows (objectness target -1, which the reference feeds to the BCE as is, loss.py:190,:224).

Reference quirk: ``iou_width_height`` divides the loss's anchor tensor by 640 IN PLACE on every call
(utils/bboxes_utils.py:18), i.e. once per box, for the lifetime of the loss object -- the k-th box ever seen is matched
against ``anchors / 640**k`` and from the 17th box on against all-zero anchors (every IoU 0 => anchors visited in index
order).  ``mirror_anchor_decay=True`` (default) reproduces exactly that, call count included, so results are bit-compatible
with the reference; ``mirror_anchor_decay=False`` normalises the anchors once (what the code evidently meant).
"""
import csv
import os

import numpy as np
import torch

from . import _lib
from .loss import ComputeLoss, IMAGE_SIZE

_DECAY_ROWS = 24  # anchors / 640**k underflows to exactly zero in fp32 at k = 17; rows beyond stay zero


class YOLO_LOSS(ComputeLoss):
    nan_on_empty = True  # (1 - iou).mean() of an empty selection is NaN in the reference (loss.py:211)

    def __init__(self, model, rect_training=False, save_logs=False, filename=None, resume=False, image_size=IMAGE_SIZE,
                 mirror_anchor_decay=True):
        super().__init__(model, save_logs=save_logs, filename=filename, resume=resume, image_size=image_size)
        self.rect_training = rect_training
        self.S = model.head.stride
        self.ignore_iou_thresh = 0.5
        self.num_anchors_per_scale = self.na
        self.mirror_anchor_decay = mirror_anchor_decay
        self.iou_calls = 0  # how many times the reference would have divided its anchors by 640 so far
        # rows k = 0.. of the decay table, computed with the reference's own fp32 operations (bboxes_utils.py:18-20)
        a = model.head.anchors.detach().clone().to("cpu", torch.float32)
        stride = torch.tensor(self.S).repeat(6, 1).T.reshape(9, 2)
        rows = []
        for _ in range(_DECAY_ROWS):
            rows.append(a.reshape(9, 2) * stride)
            a = a / 640
        self._table_cpu = torch.stack(rows).contiguous()
        self._table = {}

    # -- plumbing: the row lists come from yb_yolo_build_targets, everything downstream is ComputeLoss's ---------------
    def _nobj_ptr(self, ws):
        return ws.nobj.data_ptr()

    def _labels(self, targets):
        """tuple of per-image (n_i,5) arrays -> (float64 [nt,5], int32 offsets [B+1])"""
        rows, off = [], [0]
        for t in targets:
            t = np.asarray(t.cpu() if torch.is_tensor(t) else t, dtype=np.float64).reshape(-1, 5) if len(t) else np.zeros((0, 5))
            rows.append(t)
            off.append(off[-1] + t.shape[0])
        lab = np.concatenate(rows, 0) if rows else np.zeros((0, 5))
        return np.ascontiguousarray(lab, dtype=np.float64), np.asarray(off, dtype=np.int32)

    def _prep(self, p, targets):
        if len(p) != self.nl:
            raise ValueError(f"YOLO_LOSS: expected {self.nl} prediction levels, got {len(p)}")
        for t in p:
            if not (torch.is_tensor(t) and t.is_cuda):
                raise _lib.YBError("YOLO_LOSS (B200): predictions must be CUDA tensors (no CPU fallback)")
        if len(targets) != p[0].shape[0]:
            raise ValueError(f"YOLO_LOSS: {len(targets)} label arrays for a batch of {p[0].shape[0]} images")
        return p[0].device, targets

    def _workspace(self, shapes, nt, dev, need_free):
        ws = super()._workspace(shapes, nt, dev, need_free)  # cap = 5 * na * nt_cap >= 3 * nt
        if not hasattr(ws, "nobj"):
            cells = sum(B * na * H * W for (B, na, H, W, _) in shapes)
            ws.nobj = torch.zeros(len(shapes), dtype=torch.int32, device=dev)
            ws.state = torch.zeros(cells, dtype=torch.int8, device=dev)
        return ws

    def _build(self, ws, p, targets, dev, dense=None):
        L, st = _lib.lib(), _lib.stream()
        for i, pi in enumerate(p):
            ws.levels[i].p = pi.data_ptr()
            ws.levels[i].balance = self.balance[i]
        lab, off = self._labels(targets)
        nt, B = lab.shape[0], len(off) - 1
        lab_d = torch.from_numpy(lab).to(dev)
        off_d = torch.from_numpy(off).to(dev)
        table = self._table.get(dev)
        if table is None:
            table = self._table[dev] = self._table_cpu.to(dev)
        anchors = self._anchors_dev(dev)
        stride = 1 if self.mirror_anchor_decay else 0
        _lib.check(L.yb_yolo_build_targets(lab_d.data_ptr(), off_d.data_ptr(), B, nt, table.data_ptr(), _DECAY_ROWS,
                                           self.iou_calls if self.mirror_anchor_decay else 0, stride, anchors.data_ptr(),
                                           ws.levels, self.nl, self.na, self.ignore_iou_thresh, ws.cap, ws.state.data_ptr(),
                                           dense.data_ptr() if dense is not None else None, ws.counts.data_ptr(),
                                           ws.nobj.data_ptr(), st))
        if self.mirror_anchor_decay:
            self.iou_calls += nt
        ws._keep = (lab_d, off_d, anchors, table)

    # -- reference API ------------------------------------------------------------------------------------------------
    def __call__(self, preds, targets, pred_size=None, batch_idx=None, epoch=None):
        from .loss import _LossFn
        dev, targets = self._prep(preds, targets)
        pc = []
        for t in preds:
            t2 = t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()
            if t2 is not t and hasattr(t, "_yb_engine"):
                t2._yb_engine = None
            pc.append(t2)
        with torch.cuda.device(dev):
            loss = _LossFn.apply(self, _Targets(targets), *pc)
        if self.save_logs and batch_idx is not None and batch_idx % 100 == 0:  # loss.py:80-88: mean over the three levels
            lbox, lobj, lcls = (float(v) / 3.0 for v in self.last_parts.tolist())
            with open(os.path.join("train_eval_metrics", self.filename, "loss.csv"), "a") as f:
                csv.writer(f).writerow([epoch, batch_idx, lbox, lobj, lcls])
        return loss

    def build_targets(self, input_tensor, bboxes, pred_size=None):
        """loss.py:101-192 for ONE image: list of three (3, H, W, 6) CPU tensors
        [x_cell, y_cell, w_cell, h_cell, objectness (1 / -1 ignore / 0), class]; like the reference, the call advances the
        anchor decay by one step per box when ``mirror_anchor_decay``."""
        shapes = [(1, self.na, int(t.shape[2]), int(t.shape[3]), 5 + self.nc) for t in input_tensor]
        dev = input_tensor[0].device if input_tensor[0].is_cuda else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            ws = self._workspace(shapes, len(bboxes), dev, need_free=True)
            dense = torch.zeros(sum(s[1] * s[2] * s[3] * 6 for s in shapes), dtype=torch.float32, device=dev)
            fake = [torch.empty(0, device=dev) for _ in shapes]
            self._build(ws, fake, (bboxes,), dev, dense=dense)
        out, o = [], 0
        for (_, na, H, W, _) in shapes:
            n = na * H * W * 6
            out.append(dense[o:o + n].view(na, H, W, 6).cpu())
            o += n
        return out


class _Targets:
    """opaque wrapper so that autograd passes the per-image label arrays through untouched; ``shape[0]`` = label rows"""

    def __init__(self, targets):
        self.targets = targets
        self.shape = (sum(len(t) for t in targets),)

    def __iter__(self):
        return iter(self.targets)

    def __len__(self):
        return len(self.targets)
