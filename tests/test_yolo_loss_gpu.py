"""YOLO_LOSS drop-in (yolov5m_b200.yolo_loss; reference loss.py:20-246, the default loss of train.py:102-106) against golden
vectors produced by the REAL reference (tests/golden/yolo_loss.npz, make_golden_yolo_loss.py) and against the oracle.

Bit-exact: the dense target tensors of build_targets (cells, classes, objectness 1 / -1, fp32 box coordinates), including
the reference's in-place anchor decay across calls.  1e-5: loss and dL/dp (fp32 arithmetic, different summation order)."""
import os

import numpy as np
import pytest
import torch

import recipes
from oracle import model_ref
from oracle.yolo_loss_ref import YoloLossRef

gpu = pytest.mark.gpu
SEQ = [("c0", 2, 128, 128, 31), ("c1", 4, 160, 96, 32), ("c2", 3, 64, 64, 33), ("c3", 4, 128, 160, 34)]


class _Head:
    def __init__(self):
        self.nc, self.nl, self.naxs = 80, 3, 3
        self.anchors = model_ref.head_anchors()
        self.stride = [8, 16, 32]


class _FakeModel:
    def __init__(self):
        self.head = _Head()
        self._p = torch.nn.Parameter(torch.zeros(1, device="cuda"))

    def parameters(self):
        return iter([self._p])


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "yolo_loss.npz"))


@gpu
def test_yolo_loss_sequence_matches_reference(gold):
    from yolov5m_b200.yolo_loss import YOLO_LOSS
    loss_fn = YOLO_LOSS(_FakeModel(), rect_training=False)
    for tag, b, h, w, seed in SEQ:
        p = [t.cuda().requires_grad_(True) for t in recipes.head_outputs(40 + seed, b, h, w)]
        loss = loss_fn(p, recipes.yolo_labels(seed, b), pred_size=(h, w))
        assert loss.shape == (1,)
        ref = float(gold[tag + "_loss"][0])
        assert abs(loss.item() - ref) <= 1e-5 * abs(ref), (tag, loss.item(), ref)
        loss.backward()
        for i in range(3):
            g = p[i].grad.cpu()
            assert abs(float(g.norm()) - float(gold[f"{tag}_gnorm{i}"])) <= 1e-4 * float(gold[f"{tag}_gnorm{i}"]), (tag, i)
            nz = torch.from_numpy(gold[f"{tag}_gnz_idx{i}"])
            assert np.allclose(g.reshape(-1, 85)[nz].numpy(), gold[f"{tag}_gnz_val{i}"], rtol=1e-4, atol=1e-7), (tag, i)
            idx = torch.arange(0, g[..., 4].numel(), max(1, g[..., 4].numel() // 97))
            assert np.allclose(g[..., 4].flatten()[idx].numpy(), gold[f"{tag}_gobj{i}"], rtol=1e-4, atol=1e-8), (tag, i)
    # single-image target tensors through the public build_targets, continuing the same decay state
    p = [t.cuda() for t in recipes.head_outputs(77, 1, 96, 128)]
    for k, lab in enumerate(recipes.yolo_labels(35, 3)):
        tg = loss_fn.build_targets(p, lab, (96, 128))
        for i in range(3):
            assert np.array_equal(tg[i].numpy(), gold[f"bt{k}_l{i}"]), (k, i)


@gpu
def test_yolo_build_targets_first_boxes_bit_exact(gold):
    """a fresh loss object: box 1 sees correctly normalised anchors, the following ones anchors / 640**k"""
    from yolov5m_b200.yolo_loss import YOLO_LOSS
    fresh = YOLO_LOSS(_FakeModel(), rect_training=False)
    p = [t.cuda() for t in recipes.head_outputs(78, 1, 256, 256)]
    for k, lab in enumerate(recipes.yolo_labels(36, 4, max_boxes=5)):
        tg = fresh.build_targets(p, lab, (256, 256))
        for i in range(3):
            assert np.array_equal(tg[i].numpy(), gold[f"fresh{k}_l{i}"]), (k, i)


@gpu
def test_yolo_loss_larger_batch_vs_oracle_and_fixed_anchors():
    """bs=16 at 640x640 (~60 boxes), against the oracle: mirrored decay, and the `mirror_anchor_decay=False` mode"""
    from yolov5m_b200.yolo_loss import YOLO_LOSS
    for mirror in (True, False):
        loss_fn = YOLO_LOSS(_FakeModel(), rect_training=False, mirror_anchor_decay=mirror)
        ora = YoloLossRef(model_ref.head_anchors(), mirror_anchor_decay=mirror)
        for seed in (51, 52):
            labels = recipes.yolo_labels(seed, 16, max_boxes=8)
            pc = recipes.head_outputs(seed, 16, 640, 640)
            p = [t.cuda().requires_grad_(True) for t in pc]
            pr = [t.clone().requires_grad_(True) for t in pc]
            loss = loss_fn(p, labels, pred_size=(640, 640))
            ref, _ = ora(pr, labels)
            assert abs(loss.item() - float(ref.detach())) <= 1e-5 * abs(float(ref.detach())), (mirror, seed)
            loss.backward()
            ref.backward()
            for i in range(3):
                d = (p[i].grad.cpu().double() - pr[i].grad.double()).norm() / pr[i].grad.double().norm()
                assert float(d) < 1e-5, (mirror, seed, i, float(d))


@gpu
def test_yolo_loss_empty_batch_is_nan_like_the_reference():
    from yolov5m_b200.yolo_loss import YOLO_LOSS
    loss_fn = YOLO_LOSS(_FakeModel(), rect_training=False)
    p = [t.cuda() for t in recipes.head_outputs(5, 2, 64, 64)]
    loss = loss_fn(p, (np.zeros((0, 5)), []), pred_size=(64, 64))
    assert torch.isnan(loss).all()   # (1 - iou).mean() of an empty selection, loss.py:211
