"""Oracle: restatement of cells_to_bboxes + non_max_suppression.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

cells_to_bboxes       utils/plot_utils.py:10-54 (is_pred=True branch)
non_max_suppression   utils/bboxes_utils.py:175-209
greedy_nms            torchvision.ops.nms (third party, torchvision==0.12.0 pinned by
                      requirements.txt:11; csrc/ops/cpu/nms_kernel.cpp): stable
                      descending sort by score; box i kept unless suppressed; box j
                      suppressed when inter/(area_i+area_j-inter) > thr (strict).
                      All arithmetic fp32, no FMA contraction.
"""
import numpy as np
import torch

F32 = np.float32


def cells_to_bboxes(preds, anchors, strides):
    """preds list of (B,3,H,W,5+nc) logits -> (B, sum(3*H*W), 6) [cls, obj, cx, cy, w, h] px."""
    outs = []
    for i, p in enumerate(preds):
        bs, na, ny, nx, _ = p.shape
        s = p.sigmoid()
        ys, xs = torch.meshgrid(torch.arange(ny), torch.arange(nx), indexing="ij")
        grid = torch.stack([xs, ys], -1).view(1, 1, ny, nx, 2)  # int64, like make_grids :42-54
        ag = (anchors[i] * strides[i]).view(1, na, 1, 1, 2)
        obj = s[..., 4:5]
        xy = (2 * s[..., 0:2] + grid - 0.5) * strides[i]  # :25
        wh = ((2 * s[..., 2:4]) ** 2) * ag  # :26
        best = torch.argmax(s[..., 5:], dim=-1).unsqueeze(-1)  # :27
        outs.append(torch.cat((best, obj, xy, wh), dim=-1).reshape(bs, -1, 6))
    return torch.cat(outs, dim=1)


def greedy_nms(boxes, scores, thr):
    """boxes (n,4) x1y1x2y2 float32 numpy, scores (n,) -> kept indices (int64), score order."""
    boxes = np.asarray(boxes, F32); scores = np.asarray(scores, F32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros(0, np.int64)
    order = np.argsort(-scores, kind="stable")
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    supp = np.zeros(n, bool)
    keep = []
    thr = F32(thr)
    for _i in range(n):
        i = order[_i]
        if supp[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest]); yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest]); yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(F32(0), xx2 - xx1); h = np.maximum(F32(0), yy2 - yy1)
        inter = w * h
        ovr = inter / (areas[i] + areas[rest] - inter)
        supp[rest[ovr > thr]] = True
    return np.asarray(keep, np.int64)


def non_max_suppression(batch_bboxes, iou_threshold, threshold, max_detections=300):
    """batch_bboxes (B,N,6) [cls, score, cx, cy, w, h] -> list of (k,6) float32 arrays
    [cls, score, x1, y1, x2, y2] and the kept candidate indices (into the original N)."""
    bb = batch_bboxes.detach().cpu().numpy().astype(F32) if torch.is_tensor(batch_bboxes) else np.asarray(batch_bboxes, F32)
    outs, idxs = [], []
    for boxes in bb:
        sel = np.nonzero(boxes[:, 1] > F32(threshold))[0]
        c = boxes[sel].copy()
        c[:, 2] = c[:, 2] - c[:, 4] / F32(2)  # :190
        c[:, 3] = c[:, 3] - c[:, 5] / F32(2)  # :191
        c[:, 5] = c[:, 5] + c[:, 3]  # :192
        c[:, 4] = c[:, 4] + c[:, 2]  # :193
        keep = greedy_nms(c[:, 2:] + c[:, 0:1], c[:, 1], iou_threshold)  # :195 class offset = +cls
        keep = keep[:max_detections]  # :202-203
        outs.append(c[keep]); idxs.append(sel[keep])
    return outs, idxs
