"""cells_to_bboxes / non_max_suppression drop-ins (yolov5m_b200.boxes, csrc/nms.cu) against golden vectors of the real
reference (tests/golden/nms.npz) and against the oracle.

Bars: NMS output rows bit-exact (=> keep sets and their order bit-exact); decode: class index exact, floats 1e-5 rel
(fp32 expf differs from ATen's vectorised exp in the last ulp).
"""
import numpy as np
import pytest
import torch

import recipes
from oracle import model_ref, nms_ref

gpu = pytest.mark.gpu

NMS_CASES = {
    "realistic": (lambda: recipes.nms_boxes(1, 3, 4000, "realistic"), 0.45, 0.25),
    "allpass": (lambda: recipes.nms_boxes(2, 2, 2500, "allpass"), 0.45, 0.25),
    "ties": (lambda: recipes.nms_boxes(3, 2, 1500, "ties"), 0.45, 0.25),
    "clustered": (lambda: recipes.nms_boxes(4, 2, 3000, "clustered"), 0.6, 0.01),
    "none": (lambda: recipes.nms_boxes(5, 2, 100, "realistic") * torch.tensor([1, 0.0, 1, 1, 1, 1]), 0.45, 0.25),
}


@gpu
def test_decode_vs_reference(golden):
    from yolov5m_b200.boxes import cells_to_bboxes
    p = recipes.head_outputs(31, 2, 64, 96, scale=2.0)
    ref = golden["nms"]["decode"]
    out = cells_to_bboxes([t.cuda() for t in p], model_ref.head_anchors().cuda(), [8, 16, 32], is_pred=True, to_list=False)
    assert tuple(out.shape) == ref.shape and out.is_cuda
    o = out.cpu().numpy()
    assert np.array_equal(o[..., 0], ref[..., 0])          # class arg-max exact
    assert np.allclose(o[..., 1:], ref[..., 1:], rtol=1e-5, atol=1e-5)
    # CPU input -> CPU output, nested-list default, misspelt alias of plot_utils.py:77
    lst = cells_to_bboxes(p, model_ref.head_anchors(), [8, 16, 32], is_pred=True)
    assert isinstance(lst, list) and len(lst) == 2 and len(lst[0]) == ref.shape[1] and len(lst[0][0]) == 6
    t2 = cells_to_bboxes(p, model_ref.head_anchors(), [8, 16, 32], is_pred=True, list_output=False)
    assert torch.is_tensor(t2) and not t2.is_cuda


@gpu
def test_decode_targets_branch():
    """is_pred=False (plot_utils.py:29-34) on (B,3,H,W,6) target tensors."""
    from yolov5m_b200.boxes import cells_to_bboxes
    g = torch.Generator().manual_seed(1)
    t = [torch.rand(2, 3, 64 // s, 96 // s, 6, generator=g) for s in (8, 16, 32)]
    out = cells_to_bboxes([x.cuda() for x in t], model_ref.head_anchors().cuda(), [8, 16, 32], is_pred=False, to_list=False).cpu()
    exp = []
    for i, x in enumerate(t):
        ny, nx = x.shape[2:4]
        ys, xs = torch.meshgrid(torch.arange(ny), torch.arange(nx), indexing="ij")
        grid = torch.stack([xs, ys], -1).view(1, 1, ny, nx, 2)
        s = (8, 16, 32)[i]
        exp.append(torch.cat((x[..., 5:6], x[..., 4:5], (x[..., 0:2] + grid) * s, x[..., 2:4] * s), -1).reshape(2, -1, 6))
    assert torch.equal(out, torch.cat(exp, 1))


@gpu
@pytest.mark.parametrize("tag", list(NMS_CASES))
def test_nms_rows_bit_exact(golden, tag):
    from yolov5m_b200.boxes import non_max_suppression
    g = golden["nms"]
    mk, iou_t, thr = NMS_CASES[tag]
    bx = mk()
    before = bx.clone()
    res = non_max_suppression(bx.cuda(), iou_threshold=iou_t, threshold=thr, max_detections=300, tolist=True)
    assert torch.equal(bx, before)
    assert [len(r) for r in res] == list(g[f"{tag}_counts"])
    rows = np.array([row for r in res for row in r], np.float32).reshape(-1, 6)
    assert np.array_equal(rows, g[f"{tag}_rows"])
    cat = non_max_suppression(bx, iou_t, thr, 300, tolist=False)     # CPU tensor in -> one concatenated CPU tensor out
    assert torch.is_tensor(cat) and not cat.is_cuda and np.array_equal(cat.numpy(), g[f"{tag}_rows"])
    assert non_max_suppression(bx.tolist(), iou_t, thr, 300, to_list=True) == res  # nested-list input, detect.py:54 alias


@gpu
def test_nms_on_decoded_head_outputs(golden):
    from yolov5m_b200.boxes import non_max_suppression
    g = golden["nms"]
    dec = torch.from_numpy(g["decode"])
    res = non_max_suppression(dec.cuda(), 0.45, 0.25, 300, tolist=True)
    assert [len(r) for r in res] == list(g["decoded_counts"])
    assert np.array_equal(np.array([row for r in res for row in r], np.float32).reshape(-1, 6), g["decoded_rows"])


@gpu
@pytest.mark.parametrize("n,mode,maxdet", [(25200, "allpass", 300), (25200, "realistic", 300), (7000, "clustered", 50),
                                           (1, "allpass", 300), (513, "ties", 7)])
def test_nms_large_vs_oracle(n, mode, maxdet):
    """full-size single-level counts (25,200 candidates = one 640x640 image, every box above threshold) vs the oracle."""
    from yolov5m_b200.boxes import nms_device
    bx = recipes.nms_boxes(11, 2, n, mode)
    thr = 0.01 if mode == "clustered" else 0.25
    rows, counts, index = nms_device(bx.cuda(), 0.45, thr, maxdet, want_index=True)
    outs, idxs = nms_ref.non_max_suppression(bx, 0.45, thr, maxdet)
    counts = counts.tolist()
    assert counts == [len(o) for o in outs]
    for i, c in enumerate(counts):
        assert np.array_equal(rows[i, :c].cpu().numpy(), outs[i].astype(np.float32))
        assert np.array_equal(index[i, :c].cpu().numpy().astype(np.int64), idxs[i])


@gpu
def test_nms_full_size_properties():
    """BASELINE config-5 size (1280x1280: 100,800 boxes / image, all above threshold): size-independent properties --
    rows sorted by score (ties by index), no kept pair overlaps above the threshold, every kept row is an input row,
    idempotence (NMS of the kept rows keeps all of them)."""
    from yolov5m_b200.boxes import nms_device
    B, N = 4, 100800
    bx = recipes.nms_boxes(12, B, N, "allpass", size=1280.0).cuda()
    rows, counts, index = nms_device(bx, 0.45, 0.25, 300, want_index=True)
    for i, c in enumerate(counts.tolist()):
        assert c == 300
        r, ix = rows[i, :c], index[i, :c].long()
        sc = r[:, 1]
        assert bool(((sc[:-1] > sc[1:]) | ((sc[:-1] == sc[1:]) & (ix[:-1] < ix[1:]))).all())
        src = bx[i, ix]
        assert torch.equal(src[:, :2], r[:, :2])
        off = r[:, 2:] + r[:, 0:1]
        area = (off[:, 2] - off[:, 0]) * (off[:, 3] - off[:, 1])
        lt = torch.max(off[:, None, :2], off[None, :, :2]); rb = torch.min(off[:, None, 2:], off[None, :, 2:])
        wh = (rb - lt).clamp(min=0)
        inter = wh[..., 0] * wh[..., 1]
        iou = inter / (area[:, None] + area[None, :] - inter)
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= 0.45
    # the first kept row is the global arg-max score with the lowest index
    best = bx[:, :, 1].max(dim=1).values
    assert torch.equal(rows[:, 0, 1], best)


@gpu
def test_decode_class_argmax_ties_and_saturation():
    """class = argmax over sigmoid(logits), first maximum wins: duplicates, saturation to 1.0f (logits > ~17) and
    all-equal rows must resolve exactly like torch.argmax(torch.sigmoid(.))."""
    from yolov5m_b200.boxes import cells_to_bboxes
    g = torch.Generator().manual_seed(5)
    p = [torch.randn(2, 3, 8, 12, 85, generator=g) * 3 for _ in range(3)]
    q = p[0]
    q[0, 0, 0, 0, 5:] = 0.0                                   # all equal -> class 0
    q[0, 0, 0, 1, 5:] = -30.0; q[0, 0, 0, 1, 5 + 17] = -2.0   # single winner far below saturation
    q[0, 0, 0, 2, 5 + 40] = 25.0; q[0, 0, 0, 2, 5 + 7] = 19.0; q[0, 0, 0, 2, 5 + 63] = 30.0   # three saturated: first (7) wins
    q[0, 0, 0, 3, 5 + 11] = 4.5; q[0, 0, 0, 3, 5 + 3] = 4.5   # exact duplicate maximum -> lower index
    q[0, 0, 0, 4, 5:] = 40.0                                  # everything saturated
    out = cells_to_bboxes([t.cuda() for t in p], model_ref.head_anchors().cuda(), [8, 16, 32], is_pred=True, to_list=False).cpu()
    exp = torch.cat([torch.argmax(torch.sigmoid(t[..., 5:]), dim=-1).reshape(2, -1) for t in p], 1).float()
    assert torch.equal(out[..., 0], exp)
