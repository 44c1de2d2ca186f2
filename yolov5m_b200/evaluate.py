"""YOLO_EVAL: drop-in for the reference's evaluation glue (reference utils/validation_utils.py:11-143) around the on-device
decode + NMS of this package.

Same constructor ``YOLO_EVAL(save_logs, conf_threshold, nms_iou_thresh, map_iou_thresh, device, filename, resume)`` and
the same two methods:
  * ``check_class_accuracy(model, loader)``   (:44-83)  class / "obj" accuracy over the labelled cells; the per-cell arg-max
    and threshold run in one kernel per level (csrc/nms.cu: class_accuracy_kernel), three counters come back per call;
  * ``map_pr_rec(model, loader, anchors, epoch)`` (:85-143)  forward -> cells_to_bboxes of the predictions AND of the label
    tensors -> non_max_suppression of both (tolist=False: ONE concatenated tensor per batch, reference quirk App. B4) ->
    the ``preds`` / ``targets`` dict lists -> mean average precision.  Everything up to the dict lists stays on the GPU
    (the reference's own GPU path fails there: make_grids builds CPU grids, App. B2).
The reference scores the lists with ``torchmetrics.detection.mean_ap.MeanAveragePrecision``; that is used when importable,
otherwise :class:`MeanAP` below (COCO-style 101-point AP at IoU 0.50 and 0.75, the two numbers the reference logs).
Quirks mirrored: images are divided by 255 (uint8 input takes the fused path of the stem staging); the "obj accuracy"
thresholds channel 0 of the logits (:67), not the objectness; the CSV row layout of eval.csv.
"""
import csv
import os

import torch

from . import _lib
from .boxes import cells_to_bboxes, non_max_suppression


def _box_iou(a, b):
    """(n,4) x (m,4) xyxy -> (n,m)"""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter)


class MeanAP:
    """Minimal stand-in for torchmetrics' MeanAveragePrecision (same update / compute protocol, keys ``map_50`` / ``map_75``):
    per class, detections of all samples ranked by score, greedily matched to the unmatched ground-truth box of the same
    sample and class with the highest IoU >= threshold; AP = mean of the interpolated precision at 101 recall points (COCO);
    mean over the classes that have ground truth; at most ``max_det`` detections per sample."""

    def __init__(self, iou_thresholds=(0.5, 0.75), max_det=100):
        self.iou_thresholds, self.max_det = tuple(iou_thresholds), max_det
        self.preds, self.targets = [], []

    def update(self, preds, targets):
        self.preds += list(preds)
        self.targets += list(targets)

    def _ap(self, thr):
        classes = sorted({int(c) for t in self.targets for c in t["labels"].tolist()})
        aps = []
        for c in classes:
            scores, hits, ngt = [], [], 0
            for p, t in zip(self.preds, self.targets):
                gt = t["boxes"][t["labels"] == c].float().cpu()
                ngt += gt.shape[0]
                sel = p["labels"] == c
                sc = p["scores"][sel].float().cpu()
                bx = p["boxes"][sel].float().cpu()
                order = torch.argsort(sc, descending=True, stable=True)[: self.max_det]
                sc, bx = sc[order], bx[order]
                used = torch.zeros(gt.shape[0], dtype=torch.bool)
                iou = _box_iou(bx, gt) if gt.shape[0] and bx.shape[0] else torch.zeros(bx.shape[0], gt.shape[0])
                for i in range(bx.shape[0]):
                    hit = False
                    if gt.shape[0]:
                        cand = iou[i].clone()
                        cand[used] = -1
                        j = int(torch.argmax(cand))
                        if cand[j] >= thr:
                            used[j] = True
                            hit = True
                    scores.append(float(sc[i]))
                    hits.append(hit)
            if ngt == 0:
                continue
            if not scores:
                aps.append(0.0)
                continue
            order = sorted(range(len(scores)), key=lambda i: -scores[i])
            tp = torch.tensor([hits[i] for i in order], dtype=torch.float64)
            ctp, cfp = torch.cumsum(tp, 0), torch.cumsum(1 - tp, 0)
            rec, prec = ctp / ngt, ctp / (ctp + cfp)
            for i in range(prec.numel() - 2, -1, -1):  # monotone precision envelope
                prec[i] = torch.maximum(prec[i], prec[i + 1])
            rpts = torch.linspace(0, 1, 101, dtype=torch.float64)
            idx = torch.searchsorted(rec, rpts, right=False)
            pr = torch.where(idx < prec.numel(), prec[idx.clamp(max=prec.numel() - 1)], torch.zeros_like(rpts))
            aps.append(float(pr.mean()))
        return torch.tensor(sum(aps) / len(aps) if aps else -1.0)

    def compute(self):
        out = {}
        for thr in self.iou_thresholds:
            out["map_%d" % round(thr * 100)] = self._ap(thr)
        return out


def _metric():
    try:
        from torchmetrics.detection.mean_ap import MeanAveragePrecision
        return MeanAveragePrecision()
    except Exception:
        return MeanAP()


class YOLO_EVAL:
    def __init__(self, save_logs, conf_threshold, nms_iou_thresh, map_iou_thresh, device, filename, resume):
        self.save_logs = save_logs
        self.conf_threshold = conf_threshold
        self.nms_iou_thresh = nms_iou_thresh
        self.map_iou_threshold = map_iou_thresh
        self.device = device
        self.filename = filename
        if self.save_logs and not resume:  # validation_utils.py:23-36
            folder = os.path.join("train_eval_metrics", filename)
            os.makedirs(folder, exist_ok=True)
            with open(os.path.join(folder, "eval.csv"), "w") as f:
                csv.writer(f).writerow(["epoch", "class_accuracy", "obj_accuracy", "map50", "map75"])
        self.class_accuracy = None
        self.noobj_accuracy = None
        self.obj_accuracy = None
        self.last_preds, self.last_targets = None, None  # the lists handed to the metric (inspection / tests)

    def _dev(self):
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise _lib.YBError("YOLO_EVAL (B200): device must be a CUDA device (no CPU fallback)")
        return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())

    def check_class_accuracy(self, model, loader):
        model.eval()
        dev = self._dev()
        L = _lib.lib()
        counters = torch.zeros(3, dtype=torch.int64, device=dev)
        for images, y in loader:
            images = images.to(dev)  # uint8 stays uint8: /255 is fused into the stem staging (validation_utils.py:54)
            with torch.no_grad():
                out = model(images)
            with torch.cuda.device(dev):
                st = _lib.stream()
                for i in range(3):
                    yi = y[i].to(dev, torch.float32).contiguous()
                    oi = out[i].to(dev, torch.float32).contiguous()
                    _lib.check(L.yb_class_accuracy(oi.data_ptr(), yi.data_ptr(), yi.numel() // yi.shape[-1], oi.shape[-1],
                                                   yi.shape[-1], float(self.conf_threshold), counters.data_ptr(), st))
        tot, cc, co = (float(v) for v in counters.tolist())
        class_accuracy = cc / (tot + 1e-16)
        obj_accuracy = co / (tot + 1e-16)
        if self.save_logs:
            self.class_accuracy = round(float(class_accuracy), 3)
            self.obj_accuracy = round(float(obj_accuracy), 3)
        print("Class accuracy: {:.2f}%".format(class_accuracy * 100))
        print("Obj accuracy: {:.2f}%".format(obj_accuracy * 100))
        model.train()
        return class_accuracy, obj_accuracy

    def map_pr_rec(self, model, loader, anchors, epoch):
        model.eval()
        dev = self._dev()
        preds, targets = [], []
        for images, labels in loader:
            images = images.to(dev)
            with torch.no_grad():
                predictions = model(images)
            pred_boxes = cells_to_bboxes(predictions, anchors, strides=model.head.stride, is_pred=True, to_list=False)
            true_boxes = cells_to_bboxes([t.to(dev) for t in labels], anchors, strides=model.head.stride, is_pred=False,
                                         to_list=False)
            pred_boxes = non_max_suppression(pred_boxes, iou_threshold=self.nms_iou_thresh, threshold=self.conf_threshold,
                                             tolist=False, max_detections=300)
            true_boxes = non_max_suppression(true_boxes, iou_threshold=self.nms_iou_thresh, threshold=self.conf_threshold,
                                             tolist=False, max_detections=300)
            preds.append(dict(boxes=pred_boxes[..., 2:], scores=pred_boxes[..., 1], labels=pred_boxes[..., 0]))
            targets.append(dict(boxes=true_boxes[..., 2:], labels=true_boxes[..., 0]))
        self.last_preds, self.last_targets = preds, targets
        metric = _metric()
        metric.update(preds, targets)
        metrics = metric.compute()
        map50, map75 = metrics["map_50"], metrics["map_75"]
        print(f"MAP50: {map50}, \nMAP75: {map75}")
        if self.save_logs:
            with open(os.path.join("train_eval_metrics", self.filename, "eval.csv"), "a") as f:
                csv.writer(f).writerow([epoch, self.class_accuracy, self.obj_accuracy, float(map50), float(map75)])
        return float(map50), float(map75)
