"""yolov5m_b200: B200-native (sm_100a) hot path of AlessandroMondin/YOLOV5m.

Drop-in surface (same names / signatures as the reference):
    YOLOV5m(first_out, nc=80, anchors=(), ch=(), inference=False)      model.py:178
    ComputeLoss(model, save_logs=False, filename=None, resume=False)   ultralytics_loss.py:17
    cells_to_bboxes(predictions, anchors, strides, is_pred, to_list)   utils/plot_utils.py:10
    non_max_suppression(batch_bboxes, iou_threshold, threshold, ...)   utils/bboxes_utils.py:175
    intersection_over_union(...)                                       utils/bboxes_utils.py:33
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
