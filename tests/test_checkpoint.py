"""Checkpoint wire format: the reference's own save_checkpoint / load_model_checkpoint / load_optim_checkpoint
(utils/utils.py:56-82, imported unmodified from baseline/_ref) round-trip the drop-in's state_dict and optimizer state
through a real .pth.tar file, in both directions."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import refshim  # noqa: E402
from oracle import model_ref  # noqa: E402

needs_ref = pytest.mark.skipif(refshim.ref_path() is None, reason="reference not vendored")
gpu = pytest.mark.gpu


@needs_ref
def test_state_dict_file_round_trip_both_ways_cpu(tmp_path, monkeypatch):
    import yolov5m_b200 as yb
    ref = refshim.import_reference("cpu")
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(3)
    rm = ref.model.YOLOV5m(first_out=48, nc=80, anchors=ref.config.ANCHORS, ch=(192, 384, 768))
    ours = yb.YOLOV5m(first_out=48, nc=80, anchors=ref.config.ANCHORS, ch=(192, 384, 768))
    # reference -> file -> drop-in
    ref.utils_utils.save_checkpoint({"state_dict": rm.state_dict(), "optimizer": {}}, "SAVED_CHECKPOINT", "model_1", 7)
    assert os.path.isfile(tmp_path / "SAVED_CHECKPOINT" / "model_1" / "checkpoint_epoch_7.pth.tar")
    ref.utils_utils.load_model_checkpoint("model_1", ours, 7)                 # strict load inside (utils.py:71)
    for (k, a), (k2, b) in zip(rm.state_dict().items(), ours.state_dict().items()):
        assert k == k2 and torch.equal(a, b), k
    # drop-in -> file -> reference
    with torch.no_grad():
        for p in ours.parameters():
            p.add_(0.01)
    ref.utils_utils.save_checkpoint({"state_dict": ours.state_dict(), "optimizer": {}}, "SAVED_CHECKPOINT", "model_1", 8)
    rm2 = ref.model.YOLOV5m(first_out=48, nc=80, anchors=ref.config.ANCHORS, ch=(192, 384, 768))
    ref.utils_utils.load_model_checkpoint("model_1", rm2, 8)
    for (k, a), (k2, b) in zip(ours.state_dict().items(), rm2.state_dict().items()):
        assert k == k2 and torch.equal(a, b), k


@needs_ref
@gpu
def test_optimizer_checkpoint_round_trip_gpu(tmp_path, monkeypatch):
    """train two steps with the fused optimiser, save {"state_dict", "optimizer"} like train.py:139-143, load it into the
    reference model + torch.optim.Adam with the reference's loaders, and back into a fresh drop-in: the next step matches"""
    import yolov5m_b200 as yb
    from yolov5m_b200.trainer import Adam, TrainStep
    ref = refshim.import_reference("cpu")
    monkeypatch.chdir(tmp_path)
    sd = model_ref.make_state_dict(0)
    m = yb.YOLOV5m(first_out=48, nc=80, anchors=yb.ANCHORS, ch=(192, 384, 768))
    m.load_state_dict({k: v.clone() for k, v in sd.items()})
    m = m.cuda().train()
    opt = Adam(m, lr=5e-4, weight_decay=5e-4)
    step = TrainStep(m, yb.ComputeLoss(m), opt, max_norm=10.0)
    g = torch.Generator().manual_seed(5)
    x = torch.randint(0, 256, (2, 3, 64, 64), dtype=torch.uint8, generator=g).cuda()
    tg = torch.tensor([[0, 3, 0.5, 0.5, 0.3, 0.3], [1, 7, 0.3, 0.6, 0.2, 0.4]])
    for _ in range(2):
        step(x, tg)
    ref.utils_utils.save_checkpoint({"state_dict": m.state_dict(), "optimizer": opt.state_dict()}, "SAVED_CHECKPOINT", "model_2", 1)
    # into the reference's classes through the reference's loaders
    rm = ref.model.YOLOV5m(first_out=48, nc=80, anchors=ref.config.ANCHORS, ch=(192, 384, 768))
    ropt = torch.optim.Adam(rm.parameters(), lr=1e-3)
    ref.utils_utils.load_model_checkpoint("model_2", rm, 1)
    ref.utils_utils.load_optim_checkpoint("model_2", ropt, 1)
    assert ropt.param_groups[0]["lr"] == 5e-4 and ropt.param_groups[0]["weight_decay"] == 5e-4
    st = ropt.state_dict()["state"]
    assert len(st) == len(list(rm.parameters())) and all(int(v["step"]) == 2 for v in st.values())
    # and back into a fresh drop-in: parameters, moments and step counter continue identically
    m2 = yb.YOLOV5m(first_out=48, nc=80, anchors=yb.ANCHORS, ch=(192, 384, 768))
    ref.utils_utils.load_model_checkpoint("model_2", m2, 1)
    m2 = m2.cuda().train()
    opt2 = Adam(m2)
    ref.config.DEVICE = "cuda"
    ref.utils_utils.load_optim_checkpoint("model_2", opt2, 1)
    ref.config.DEVICE = "cpu"
    assert torch.equal(m2.flat_params, m.flat_params) and torch.equal(opt2.m, opt.m) and torch.equal(opt2.v, opt.v)
    step2 = TrainStep(m2, yb.ComputeLoss(m2), opt2, max_norm=10.0)
    la, lb = step(x, tg), step2(x, tg)
    torch.cuda.synchronize()
    assert torch.equal(la, lb) and torch.equal(m2.flat_params, m.flat_params)
    assert opt.steps_taken() == 3 and opt2.steps_taken() == 3
