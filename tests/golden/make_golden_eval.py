"""Golden fixtures for the eval glue YOLO_EVAL.map_pr_rec / check_class_accuracy (utils/validation_utils.py:44-143), generated
by running the REAL reference on CPU with a stub model that returns fixed head tensors (so the fixture pins the glue --
decode of predictions AND of label tensors, NMS of both, the dict lists handed to MeanAveragePrecision -- independently of
the network's precision).  torchmetrics is not in this image: MeanAveragePrecision is a recorder that captures update().

    python tests/golden/make_golden_eval.py        -> tests/golden/eval.npz
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("YOLO_REF", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
for m in ["albumentations", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "imagesize",
          "torchmetrics", "torchmetrics.detection", "torchmetrics.detection.mean_ap"]:
    sys.modules[m] = MagicMock()
sys.path.insert(0, REF)
import config  # noqa: E402  (reference)
config.DEVICE = "cpu"
import utils.validation_utils as vu  # noqa: E402  (reference)

import recipes  # noqa: E402
from oracle import model_ref  # noqa: E402

CAPTURED = {}


class Recorder:
    def update(self, preds, targets):
        CAPTURED["preds"], CAPTURED["targets"] = preds, targets

    def compute(self):
        return {"map_50": torch.tensor(0.0), "map_75": torch.tensor(0.0)}


vu.MeanAveragePrecision = Recorder
vu.time.sleep = lambda s: None


def main():
    batches = recipes.eval_batches()
    model = recipes.StubModel([b[2] for b in batches])
    loader = [(b[0], [t.clone() for t in b[1]]) for b in batches]
    ev = vu.YOLO_EVAL(save_logs=False, conf_threshold=0.3, nms_iou_thresh=0.6, map_iou_thresh=0.5, device="cpu",
                      filename=None, resume=False)
    ev.map_pr_rec(model, loader, anchors=model.head.anchors, epoch=1)
    out = {}
    for k, (p, t) in enumerate(zip(CAPTURED["preds"], CAPTURED["targets"])):
        out[f"pred{k}_boxes"] = p["boxes"].numpy(); out[f"pred{k}_scores"] = p["scores"].numpy()
        out[f"pred{k}_labels"] = p["labels"].numpy()
        out[f"true{k}_boxes"] = t["boxes"].numpy(); out[f"true{k}_labels"] = t["labels"].numpy()
    # check_class_accuracy (validation_utils.py:44-83): quirks included (it thresholds channel 0, not the objectness)
    model2 = recipes.StubModel([b[2] for b in batches])
    ev2 = vu.YOLO_EVAL(save_logs=True, conf_threshold=0.3, nms_iou_thresh=0.6, map_iou_thresh=0.5, device="cpu",
                       filename="_golden_tmp", resume=True)
    ev2.check_class_accuracy(model2, [(b[0], [t.clone() for t in b[1]]) for b in batches])
    out["class_accuracy"] = np.array(ev2.class_accuracy)
    out["obj_accuracy"] = np.array(ev2.obj_accuracy)
    np.savez_compressed(os.path.join(HERE, "eval.npz"), **out)
    print({k: v.shape for k, v in out.items()}, out["class_accuracy"], out["obj_accuracy"])


if __name__ == "__main__":
    main()
