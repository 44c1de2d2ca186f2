"""The UNMODIFIED reference's hot path timed on the host cores -- MEASUREMENT INFRASTRUCTURE (bench.py reference arm).

Train step = the body of the reference train_loop (utils/training_utils.py:97-122) at accumulate = 1 (what it is at the
metric's bs=64): images.float()/255 -> model(images) -> ComputeLoss -> backward -> clip_grad_norm_(10) -> Adam.step ->
zero_grad, built from the reference's own YOLOV5m (model.py:178), ComputeLoss (ultralytics_loss.py:17) and
torch.optim.Adam(lr, weight_decay) (train.py:61), on CPU in fp32 (GradScaler / autocast are no-ops on CPU).
Detect = model.eval() forward -> cells_to_bboxes -> non_max_suppression (detect.py:50-54).
"""
import os
import time

import torch

from . import refshim


def synthetic_batch(seed, bs, size=640):
    """SURVEY.md 8(d) recipe (uint8 images as the reference's loader yields them, nt = 8 boxes per image)"""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 256, (bs, 3, size, size), dtype=torch.uint8, generator=g)
    nt = 8 * bs
    t = torch.cat([torch.randint(0, bs, (nt, 1), generator=g).float(), torch.randint(0, 80, (nt, 1), generator=g).float(),
                   torch.rand(nt, 2, generator=g), torch.rand(nt, 2, generator=g) * 0.5 + 0.005], 1)
    return x, t


class RefTrainer:
    kind = "reference"

    def __init__(self, threads=None, seed=0):
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.ref = refshim.import_reference("cpu")
        cfg = self.ref.config
        torch.manual_seed(seed)
        fo = cfg.FIRST_OUT
        self.model = self.ref.model.YOLOV5m(first_out=fo, nc=80, anchors=cfg.ANCHORS, ch=(fo * 4, fo * 8, fo * 16)).train()
        self.loss_fn = self.ref.ultralytics_loss.ComputeLoss(self.model, save_logs=False)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=5e-4, weight_decay=5e-4)

    def step(self, images_u8, targets):
        images = images_u8.float() / 255                                          # training_utils.py:98
        out = self.model(images)                                                  # :107
        loss = self.loss_fn(out, targets, pred_size=images.shape[2:4], batch_idx=None, epoch=None)  # :108
        loss.backward()                                                           # :114
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=10.0)    # :118
        self.opt.step()                                                           # :119
        self.opt.zero_grad(set_to_none=True)                                      # :121
        return float(loss.detach())

    def time_steps(self, bs, steps, warmup, size=640, seed=1):
        x, t = synthetic_batch(seed, bs, size)
        for _ in range(warmup):
            self.step(x, t)
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step(x, t)
        dt = time.perf_counter() - t0
        return bs * steps / dt, dt / steps


class RefDetector:
    kind = "reference"

    def __init__(self, threads=None, seed=0, obj_bias=None):
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.ref = refshim.import_reference("cpu")
        cfg = self.ref.config
        torch.manual_seed(seed)
        fo = cfg.FIRST_OUT
        self.model = self.ref.model.YOLOV5m(first_out=fo, nc=80, anchors=cfg.ANCHORS, ch=(fo * 4, fo * 8, fo * 16)).eval()
        if obj_bias is not None:  # SURVEY.md 8(d) distribution D1: objectness prior so that a few % of the cells pass 0.25
            with torch.no_grad():
                for conv in self.model.head.out_convs:
                    conv.bias.view(3, -1)[:, 4] = obj_bias

    @torch.no_grad()
    def detect(self, images, conf=0.25, iou=0.45, max_det=300):
        out = self.model(images)                                                                      # detect.py:51
        boxes = self.ref.plot_utils.cells_to_bboxes(out, self.model.head.anchors, self.model.head.stride, is_pred=True,
                                                    to_list=False)                                    # :53
        return self.ref.bboxes_utils.non_max_suppression(boxes, iou_threshold=iou, threshold=conf,
                                                         max_detections=max_det, tolist=True)        # :54

    def time_detect(self, bs, steps, warmup, size=1280, seed=2):
        g = torch.Generator().manual_seed(seed)
        x = torch.rand(bs, 3, size, size, generator=g)
        for _ in range(warmup):
            self.detect(x)
        t0 = time.perf_counter()
        for _ in range(steps):
            self.detect(x)
        dt = time.perf_counter() - t0
        return bs * steps / dt, dt / steps
