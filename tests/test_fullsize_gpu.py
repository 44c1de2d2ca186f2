"""Parity at BASELINE.json's full sizes (bs=64, 640x640, 512 targets) through size-independent properties: the oracle
cannot run a bs=64 fp32 step in seconds, so the full-size checks are
  * eval forward: batch-split invariance (64 images == two runs of 32, bit for bit: every op is per-image in eval mode),
    and decode + NMS of the full batch == the oracle's NMS on the same decoded tensor for a sample of images;
  * train step: build_targets / loss value vs the oracle ON THE GPU'S OWN LOGITS (bit-exact indices, 1e-4 loss),
    run-to-run determinism of the whole gradient bucket (fixed-order reductions everywhere, no atomics),
    head-bias gradient == column sums of the loss gradient, BN running statistics updated once;
  * multi-scale: model(x, size=s) == model(F.interpolate(x)) within bf16 staging round-off.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import recipes
import yolov5m_b200 as yb
from oracle import loss_ref, model_ref, nms_ref

pytestmark = pytest.mark.gpu
B, S, NT = 64, 640, 512


def rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def net():
    sd = model_ref.make_state_dict(0)
    m = yb.YOLOV5m(first_out=48, nc=80, anchors=yb.ANCHORS, ch=(192, 384, 768))
    m.load_state_dict({k: v.clone() for k, v in sd.items()})
    return m.cuda()


def test_eval_batch_split_invariance_and_nms(net):
    net.eval()
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).cuda()
    with torch.no_grad():
        full = net(x)
        halves = [net(x[:32]), net(x[32:])]
    for i in range(3):
        assert full[i].shape == (B, 3, S // (8 << i), S // (8 << i), 85)
        assert torch.equal(full[i], torch.cat([halves[0][i], halves[1][i]], 0)), f"level {i}: batch split changed bits"
    dec = yb.cells_to_bboxes(full, net.head.anchors, net.head.stride, is_pred=True, to_list=False)
    assert dec.shape == (B, 25200, 6)
    kept = yb.non_max_suppression(dec, iou_threshold=0.45, threshold=0.25, max_detections=300, tolist=True)
    dec_h = dec.cpu()
    for b in (0, 17, 63):  # the oracle NMS is O(N^2) python/numpy: a sample of images
        ref, _ = nms_ref.non_max_suppression(dec_h[b:b + 1], 0.45, 0.25, 300)
        assert np.array_equal(np.array(kept[b], np.float32).reshape(-1, 6), ref[0].astype(np.float32)), f"image {b}"
    net._engines.clear()


def test_train_step_full_size_properties(net):
    net.train()
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).cuda()
    tg = torch.cat([torch.randint(0, B, (NT, 1), generator=g).float(), torch.randint(0, 80, (NT, 1), generator=g).float(),
                    torch.rand(NT, 2, generator=g), torch.rand(NT, 2, generator=g) * 0.5 + 0.005], 1)
    loss_fn = yb.ComputeLoss(net)
    nbt0 = int(net.backbone[0].cbl[1].num_batches_tracked)
    buckets, losses = [], []
    for rep in range(2):
        for p in net.parameters():
            p.grad = None
        out = net(x)
        loss = loss_fn(out, tg, None)
        loss.backward()
        torch.cuda.synchronize()
        buckets.append(net.flat_grads.clone())
        losses.append(loss.detach().clone())
    assert int(net.backbone[0].cbl[1].num_batches_tracked) == nbt0 + 2
    assert torch.equal(losses[0], losses[1]) and torch.equal(buckets[0], buckets[1]), "train step is not deterministic"
    assert torch.isfinite(buckets[0]).all() and float(buckets[0].abs().max()) > 0
    # loss + target assignment vs the oracle on the GPU's own logits
    p_h = [o.detach().cpu() for o in out]
    anchors = model_ref.head_anchors()
    want = loss_ref.compute_loss(p_h, tg, anchors)
    assert rel(losses[1], want) < 1e-4, (losses[1].item(), want.item())
    tcls, tbox, idx, anch = loss_fn.build_targets(out, tg)
    ref = loss_ref.build_targets(tg.numpy(), anchors.numpy(), [tuple(o.shape) for o in p_h])
    for i in range(3):
        got_idx = torch.stack(idx[i], 0).cpu().numpy()
        assert got_idx.dtype == np.int64 and got_idx.shape[1] > 0
        assert np.array_equal(got_idx, np.stack([ref[i]["b"], ref[i]["a"], ref[i]["gj"], ref[i]["gi"]])), f"level {i}"
        assert np.array_equal(tcls[i].cpu().numpy(), ref[i]["tcls"])
        assert np.allclose(tbox[i].cpu().numpy(), ref[i]["tbox"], atol=1e-6)
        assert np.allclose(anch[i].cpu().numpy(), ref[i]["anch"], atol=1e-6)
    # head bias gradient == column sums of dL/dp (the oracle's autograd gives dL/dp for the same logits)
    p_req = [t.clone().requires_grad_(True) for t in p_h]
    loss_ref.compute_loss(p_req, tg, anchors).backward()
    for i in range(3):
        want_b = p_req[i].grad.sum(dim=(0, 2, 3)).reshape(-1)          # (3, 85) -> 255, channel a*85+o
        got_b = net.head.out_convs[i].bias.grad
        assert got_b is not None and rel(got_b.reshape(-1), want_b) < 2e-2  # bf16 head-gradient operand
    net._engines.clear()


def test_multi_scale_forward_matches_interpolate(net):
    net.eval()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(4, 3, 480, 640, generator=g)
    size = (352, 448)
    with torch.no_grad():
        got = net(x.cuda(), size=size)
        want = net(F.interpolate(x, size=size, mode="bilinear", align_corners=False).cuda())
    for i in range(3):
        assert got[i].shape == want[i].shape == (4, 3, size[0] // (8 << i), size[1] // (8 << i), 85)
        assert rel(got[i], want[i]) < 5e-3  # a few staged pixels differ by one bf16 ulp
    net._engines.clear()
