"""cells_to_bboxes / non_max_suppression / intersection_over_union: drop-ins for the reference detect path
(reference utils/plot_utils.py:10-54, utils/bboxes_utils.py:33-87 and :175-209) on hand-written sm_100a kernels
(csrc/nms.cu).  Same argument names and return types; the misspelt keyword aliases that the reference's own call
sites use (``list_output=`` in plot_utils.py:77, ``to_list=`` in detect.py:54) are accepted too.
There is no CPU / PyTorch fallback: inputs are moved to the GPU, the arithmetic runs in the CUDA library.
"""
import torch

from . import _lib


def _dev():
    if not torch.cuda.is_available():
        raise _lib.YBError("yolov5m_b200: a CUDA device is required (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def cells_to_bboxes(predictions, anchors, strides, is_pred=False, to_list=True, list_output=None):
    """plot_utils.py:10-40.  predictions: list of (B, na, H, W, 5+nc) tensors (is_pred=True: raw logits; False: target
    tensors with 6 channels).  Returns (B, sum(na*H*W), 6) rows [class, objectness, cx, cy, w, h] in pixels, as a nested
    list when ``to_list`` (the reference default) else as a tensor."""
    if list_output is not None:
        to_list = list_output
    src_dev = predictions[0].device
    dev = src_dev if src_dev.type == "cuda" else _dev()
    with torch.cuda.device(dev):  # launches follow the tensors' device, whatever the process's current device is
        out = _decode(predictions, anchors, strides, is_pred, dev)
    if to_list:
        return out.tolist()
    return out if src_dev.type == "cuda" else out.to(src_dev)


def _decode(predictions, anchors, strides, is_pred, dev):
    L, st = _lib.lib(), _lib.stream()
    preds = [p.to(device=dev, dtype=torch.float32).contiguous() for p in predictions]
    anchors = torch.as_tensor(anchors).to(device=dev, dtype=torch.float32)
    B = preds[0].shape[0]
    rows = [p.shape[1] * p.shape[2] * p.shape[3] for p in preds]
    total = sum(rows)
    out = torch.empty(B, total, 6, device=dev, dtype=torch.float32)
    off = 0
    keep = []
    for i, p in enumerate(preds):
        _, na, H, W, no = p.shape
        apx = (anchors[i] * strides[i]).contiguous()  # make_grids: anchor_grid = anchors[i]*stride, plot_utils.py:52
        keep.append(apx)
        _lib.check(L.yb_decode_level(p.data_ptr(), B, na, H, W, no, float(strides[i]), apx.data_ptr(), 1 if is_pred else 0,
                                     out.data_ptr(), total, off, st))
        off += rows[i]
    return out


class _NmsScratch:
    buf = None

    @classmethod
    def get(cls, nbytes, dev):
        if cls.buf is None or cls.buf.numel() < nbytes or cls.buf.device != dev:
            cls.buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return cls.buf


def nms_device(batch_bboxes, iou_threshold, threshold, max_detections=300, want_index=False):
    """(B,N,6) CUDA tensor -> (rows (B,max_det,6), counts (B,) int32[, index (B,max_det) int32]) all on the GPU, no sync."""
    bb = batch_bboxes
    with torch.cuda.device(bb.device):
        return _nms_device(bb, iou_threshold, threshold, max_detections, want_index)


def _nms_device(bb, iou_threshold, threshold, max_detections, want_index):
    L, st = _lib.lib(), _lib.stream()
    B, N, _ = bb.shape
    dev = bb.device
    out = torch.zeros(B, max_detections, 6, device=dev, dtype=torch.float32)
    counts = torch.zeros(B, device=dev, dtype=torch.int32)
    index = torch.zeros(B, max_detections, device=dev, dtype=torch.int32) if want_index else None
    scratch = _NmsScratch.get(max(16, int(L.yb_nms_scratch_bytes(B, N))), dev)
    _lib.check(L.yb_nms_batched(bb.data_ptr(), B, N, float(iou_threshold), float(threshold), int(max_detections),
                                scratch.data_ptr(), out.data_ptr(), counts.data_ptr(),
                                index.data_ptr() if want_index else None, None, st))
    return (out, counts, index) if want_index else (out, counts)


def non_max_suppression(batch_bboxes, iou_threshold, threshold, max_detections=300, tolist=True, to_list=None):
    """bboxes_utils.py:175-209.  batch_bboxes (B,N,6) rows [class, score, cx, cy, w, h] (tensor or nested list; never
    mutated).  Per image: rows with score > threshold, converted to [class, score, x1, y1, x2, y2], NMS'ed with the
    class index as box offset, sorted by score, at most ``max_detections``.  Returns a list (per image) of row lists when
    ``tolist`` else ONE concatenated tensor, exactly like the reference."""
    if to_list is not None:
        tolist = to_list
    src = batch_bboxes
    src_dev = src.device if torch.is_tensor(src) else torch.device("cpu")
    dev = src_dev if src_dev.type == "cuda" else _dev()
    bb = torch.as_tensor(src, dtype=torch.float32).to(dev)
    if bb.dim() == 2:
        bb = bb.unsqueeze(0)
    bb = bb.contiguous()
    out, counts = nms_device(bb, iou_threshold, threshold, max_detections)
    counts_h = counts.tolist()
    if tolist:
        out_h = out.cpu()
        return [out_h[i, :n].tolist() for i, n in enumerate(counts_h)]
    res = torch.cat([out[i, :n] for i, n in enumerate(counts_h)], dim=0) if counts_h else out.reshape(0, 6)
    return res if src_dev.type == "cuda" else res.to(src_dev)


def intersection_over_union(boxes_preds, boxes_labels, box_format="midpoint", GIoU=False, eps=1e-7):
    """bboxes_utils.py:33-87.  (..., 4) x (..., 4) -> (..., 1) IoU (or GIoU) of corresponding boxes; forward only (the
    differentiable use inside the loss is fused in csrc/loss.cu)."""
    src_dev = boxes_preds.device
    dev = src_dev if src_dev.type == "cuda" else _dev()
    a = boxes_preds.detach().to(device=dev, dtype=torch.float32)
    b = boxes_labels.detach().to(device=dev, dtype=torch.float32)
    a, b = torch.broadcast_tensors(a, b)
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty(a.shape[:-1] + (1,), device=dev, dtype=torch.float32)
    n = out.numel()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().yb_box_iou(a.data_ptr(), b.data_ptr(), n, 1 if box_format == "midpoint" else 0,
                                         1 if GIoU else 0, float(eps), out.data_ptr(), _lib.stream()))
    return out if src_dev.type == "cuda" else out.to(src_dev)
