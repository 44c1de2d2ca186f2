"""Oracle: CPU restatement of the reference's default loss YOLO_LOSS -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/loss.py:
  build_targets  loss.py:101-192  per box: IoU with the nine anchors (utils/bboxes_utils.py:6-29; float32 by torch's type
                                  promotion although the labels are float64), anchors in descending IoU, one object cell per scale, `ignore`
                                  (-1) for further free cells with IoU > 0.5
  compute_loss   loss.py:195-246  GIoU box term, BCE objectness with target GIoU.clamp(0) / -1 / 0, class BCE, * bs
including the in-place `anchors /= 640` of iou_width_height (bboxes_utils.py:18): `YoloLossRef` keeps the decaying anchor
tensor exactly like the reference object does.  Pinned by tests/test_oracle_golden.py against tests/golden/yolo_loss.npz
(outputs of the real reference).
"""
import numpy as np
import torch

from .loss_ref import bce_logits, giou_midpoint

BALANCE = [4.0, 1.0, 0.4]
STRIDE = torch.tensor([8, 16, 32]).repeat(6, 1).T.reshape(9, 2)


class YoloLossRef:
    def __init__(self, head_anchors, nc=80, image_size=640, mirror_anchor_decay=True):
        self.nc = nc
        self.anchors_d = head_anchors.clone().float()       # loss.py:41 (stays intact)
        self.anchors = head_anchors.clone().float()         # loss.py:42 (decays)
        self.mirror = mirror_anchor_decay
        self.lam_cls = 0.5 * (nc / 80 * 3 / 3)
        self.lam_obj = 1.0 * ((image_size / 640) ** 2 * 3 / 3)
        self.lam_box = 0.05 * (3 / 3)

    def _iou_anchors(self, wh):
        """utils/bboxes_utils.py:6-29 with its in-place division"""
        if self.mirror:
            self.anchors /= 640                             # :18
            a = self.anchors
        else:
            a = self.anchors_d / 640
        # torch type promotion: the float64 box values are 0-dim tensors, the anchors a dimensioned float32 tensor, so every
        # mixed operation runs in float32 (only w*h, 0-dim x 0-dim, is a float64 product, rounded when it meets the anchors)
        a = (a.reshape(9, 2) * STRIDE).numpy()              # :19-20 (float32)
        f32 = np.float32
        inter = np.minimum(f32(wh[0]), a[:, 0]) * np.minimum(f32(wh[1]), a[:, 1])   # :22-24
        union = (f32(wh[0] * wh[1]) + a[:, 0] * a[:, 1]) - inter                    # :25-27
        with np.errstate(invalid="ignore", divide="ignore"):
            return (inter / union).astype(np.float32)

    def build_targets(self, shapes, bboxes):
        """shapes: [(H, W)] * 3; bboxes (n,5) float64 [class, x, y, w, h] -> list of (3,H,W,6) float32 arrays"""
        tg = [np.zeros((3, h, w, 6), np.float32) for (h, w) in shapes]
        bboxes = np.asarray(bboxes, np.float64).reshape(-1, 5)
        for row in bboxes:
            cls, x, y, w, h = row
            iou = self._iou_anchors(np.array([w, h]))
            order = torch.from_numpy(iou).argsort(descending=True, dim=0).numpy()  # loss.py:119
            has = [False] * 3
            for a in order:
                s, aos = int(a) // 3, int(a) % 3
                sy, sx = shapes[s]
                i, j = int(sy * y), int(sx * x)                                    # :148
                taken = tg[s][aos, i, j, 4]
                if not taken and not has[s]:                                       # :162
                    tg[s][aos, i, j, 4] = 1
                    tg[s][aos, i, j, 0:4] = [sx * x - j, sy * y - i, w * sx, h * sy]
                    tg[s][aos, i, j, 5] = int(cls)
                    has[s] = True
                elif not taken and iou[a] > np.float32(0.5):                       # :189
                    tg[s][aos, i, j, 4] = -1
        return tg

    def __call__(self, preds, targets):
        """preds: list of 3 (B,3,H,W,85) tensors; targets: per-image label arrays -> loss (1,), parts"""
        shapes = [(int(p.shape[2]), int(p.shape[3])) for p in preds]
        per_img = [self.build_targets(shapes, t) for t in targets]                # loss.py:70
        total = torch.zeros(1)
        parts = []
        for lvl, p in enumerate(preds):
            t = torch.from_numpy(np.stack([pi[lvl] for pi in per_img], 0))
            bs = p.shape[0]
            anchors = self.anchors_d[lvl].reshape(1, 3, 1, 1, 2)
            obj = t[..., 4] == 1
            pxy = p[..., 0:2].sigmoid() * 2 - 0.5
            pwh = (p[..., 2:4].sigmoid() * 2) ** 2 * anchors
            pbox = torch.cat((pxy[obj], pwh[obj]), -1)
            tbox = t[..., 0:4][obj]
            giou = giou_midpoint(pbox, tbox).reshape(-1)
            lbox = (1.0 - giou).mean()
            tobj = t[..., 4].clone()
            tobj[obj] = tobj[obj] * giou.detach().clamp(0)                        # :217-218
            lobj = bce_logits(p[..., 4], tobj).mean() * BALANCE[lvl]
            pc = p[..., 5:][obj]
            tc = torch.zeros_like(pc)
            tc[torch.arange(tc.shape[0]), t[..., 5][obj].long()] = 1.0
            lcls = bce_logits(pc, tc).mean()
            total = total + (self.lam_box * lbox + self.lam_obj * lobj + self.lam_cls * lcls) * bs
            parts.append((self.lam_box * lbox, self.lam_obj * lobj, self.lam_cls * lcls))
        return total, parts
