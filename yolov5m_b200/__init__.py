"""yolov5m_b200: B200-native (sm_100a) hot path of AlessandroMondin/YOLOV5m.

Drop-in surface (same names / signatures as the reference):
    YOLOV5m(first_out, nc=80, anchors=(), ch=(), inference=False)      model.py:178
    ComputeLoss(model, save_logs=False, filename=None, resume=False)   ultralytics_loss.py:17
    cells_to_bboxes(predictions, anchors, strides, is_pred, to_list)   utils/plot_utils.py:10
    non_max_suppression(batch_bboxes, iou_threshold, threshold, ...)   utils/bboxes_utils.py:175
    intersection_over_union(...)                                       utils/bboxes_utils.py:33
    YOLO_LOSS(model, rect_training, save_logs, filename, resume)       loss.py:20   (train.py's default loss)
    YOLO_EVAL(save_logs, conf_threshold, nms_iou_thresh, ...)          utils/validation_utils.py:11
"""
from . import _lib  # noqa: F401
from .model import YOLOV5m  # noqa: F401
from .loss import ComputeLoss  # noqa: F401
from .boxes import cells_to_bboxes, non_max_suppression, intersection_over_union  # noqa: F401
from .yolo_loss import YOLO_LOSS  # noqa: F401
from .evaluate import YOLO_EVAL  # noqa: F401

ANCHORS = [  # reference config.py:33-37
    [(10, 13), (16, 30), (33, 23)],
    [(30, 61), (62, 45), (59, 119)],
    [(116, 90), (156, 198), (373, 326)],
]
FIRST_OUT = 48  # reference config.py:15

__all__ = ["YOLOV5m", "ComputeLoss", "YOLO_LOSS", "YOLO_EVAL", "cells_to_bboxes", "non_max_suppression",
           "intersection_over_union", "ANCHORS", "FIRST_OUT"]
