"""Pin the oracle (oracle/) against golden vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

import recipes
from oracle import loss_ref, model_ref, nms_ref


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64); b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_state_dict_layout():
    specs = model_ref.param_specs()
    assert len(specs) == 481
    n = sum(int(np.prod(s)) for _, s, k in specs if k in ("conv", "bn_w", "bn_b", "head_w", "head_b"))
    assert n == 21190557
    sd = model_ref.make_state_dict(0)
    assert torch.allclose(sd["head.anchors"][0, 0], torch.tensor([1.25, 1.625]))


@pytest.mark.parametrize("tag,shape", [("a", (2, 64, 96)), ("b", (1, 128, 128))])
def test_model_forward_eval_train(golden, tag, shape):
    g = golden["model"]
    x = recipes.model_input(11, *shape)
    sd = model_ref.make_state_dict(0)
    with torch.no_grad():
        pe = model_ref.forward(sd, x, train=False)
    for i in range(3):
        assert pe[i].shape == g[f"{tag}_eval_p{i}"].shape
        assert rel(pe[i], g[f"{tag}_eval_p{i}"]) < 1e-5
    with torch.no_grad():
        pt = model_ref.forward(sd, x, train=True)
    for i in range(3):
        assert rel(pt[i], g[f"{tag}_train_p{i}"]) < 1e-4
    for k, name in (("b0", "backbone.0"), ("n7", "neck.7.c_out")):
        assert rel(sd[name + ".cbl.1.running_mean"], g[f"{tag}_rm_{k}"]) < 1e-4
        assert rel(sd[name + ".cbl.1.running_var"], g[f"{tag}_rv_{k}"]) < 1e-4


def test_model_loss_and_grads(golden):
    g = golden["model"]
    tag, (b, h, w) = "a", (2, 64, 96)
    sd = model_ref.make_state_dict(0)
    names = [n for n in g[f"{tag}_grad_names"]]
    for n in names:
        sd[n] = sd[n].clone().requires_grad_(True)
    p = model_ref.forward(sd, recipes.model_input(11, b, h, w), train=True)
    loss = loss_ref.compute_loss(p, recipes.targets(5, b, 8 * b), sd["head.anchors"])
    assert rel(loss, g[f"{tag}_loss"]) < 1e-4
    loss.backward()
    norms = np.array([sd[n].grad.norm().item() for n in names])
    ref = g[f"{tag}_grad_norms"]
    big = ref > 1e-6 * ref.max()
    assert np.max(np.abs(norms[big] - ref[big]) / ref[big]) < 5e-3


@pytest.mark.parametrize("tag,dims", [("rand", (4, 160, 160)), ("many", (8, 128, 96)),
                                      ("zero", (2, 64, 64)), ("edge", (2, 640, 640))])
def test_build_targets_and_loss(golden, tag, dims):
    g = golden["loss"]
    b, h, w = dims
    tg = {"rand": lambda: recipes.targets(3, 4, 48), "many": lambda: recipes.targets(4, 8, 200),
          "zero": lambda: recipes.targets(0, 2, 0), "edge": lambda: recipes.edge_targets(2)}[tag]()
    p = [t.requires_grad_(True) for t in recipes.head_outputs(21, b, h, w)]
    anchors = model_ref.head_anchors()
    loss, parts, bt = loss_ref.compute_loss(p, tg, anchors, return_parts=True)
    for i in range(3):
        idx = np.stack([bt[i]["b"], bt[i]["a"], bt[i]["gj"], bt[i]["gi"]], 0)
        assert np.array_equal(idx, g[f"{tag}_idx{i}"]), f"level {i} indices differ"  # bit-exact, ordered
        assert np.array_equal(bt[i]["tcls"], g[f"{tag}_tcls{i}"])
        assert np.array_equal(bt[i]["tbox"], g[f"{tag}_tbox{i}"])  # fp32 bit-exact
        assert np.array_equal(bt[i]["anch"], g[f"{tag}_anch{i}"])
    assert rel(loss, g[f"{tag}_loss"]) < 1e-5
    loss.backward()
    for i in range(3):
        assert abs(p[i].grad.norm().item() - g[f"{tag}_gnorm{i}"]) <= 1e-4 * g[f"{tag}_gnorm{i}"] + 1e-12
        nz = g[f"{tag}_gnz_idx{i}"]
        if nz.size:
            assert rel(p[i].grad.reshape(-1, 85)[nz], g[f"{tag}_gnz_val{i}"]) < 1e-4


def test_giou(golden):
    gen = torch.Generator().manual_seed(9)
    a = torch.rand(256, 4, generator=gen); b = torch.rand(256, 4, generator=gen)
    a[:, 2:] += 0.05; b[:, 2:] += 0.05
    b[:8] = a[:8]; b[8:16, :2] += 5
    assert np.allclose(loss_ref.giou_midpoint(a, b).numpy(), golden["iou"]["giou"][:, 0], atol=1e-6)


def test_decode(golden):
    p = recipes.head_outputs(31, 2, 64, 96, scale=2.0)
    dec = nms_ref.cells_to_bboxes(p, model_ref.head_anchors(), [8, 16, 32])
    assert np.array_equal(dec.numpy(), golden["nms"]["decode"])


NMS_CASES = {
    "realistic": (lambda: recipes.nms_boxes(1, 3, 4000, "realistic"), 0.45, 0.25),
    "allpass": (lambda: recipes.nms_boxes(2, 2, 2500, "allpass"), 0.45, 0.25),
    "ties": (lambda: recipes.nms_boxes(3, 2, 1500, "ties"), 0.45, 0.25),
    "clustered": (lambda: recipes.nms_boxes(4, 2, 3000, "clustered"), 0.6, 0.01),
    "none": (lambda: recipes.nms_boxes(5, 2, 100, "realistic") * torch.tensor([1, 0.0, 1, 1, 1, 1]), 0.45, 0.25),
}


@pytest.mark.parametrize("tag", list(NMS_CASES))
def test_nms_keep_sets(golden, tag):
    g = golden["nms"]
    mk, iou_t, thr = NMS_CASES[tag]
    outs, _ = nms_ref.non_max_suppression(mk(), iou_t, thr, 300)
    assert [len(o) for o in outs] == list(g[f"{tag}_counts"])
    rows = np.concatenate(outs, 0) if outs else np.zeros((0, 6), np.float32)
    assert np.array_equal(rows.astype(np.float32), g[f"{tag}_rows"])  # bit-exact rows == bit-exact keep set


# ---- the reference's default loss YOLO_LOSS (loss.py), SURVEY.md 8(f) rank 3 -------------------------------------------
YOLO_SEQ = [("c0", 2, 128, 128, 31), ("c1", 4, 160, 96, 32), ("c2", 3, 64, 64, 33), ("c3", 4, 128, 160, 34)]


def test_yolo_loss_oracle_matches_reference_sequence():
    """the oracle reproduces a SEQUENCE of reference calls (the anchors decay by /640 per box, bboxes_utils.py:18):
    losses to 1e-6, gradients to 1e-5, single-image target tensors bit for bit"""
    import numpy as np
    import recipes
    from oracle import model_ref
    from oracle.yolo_loss_ref import YoloLossRef
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "yolo_loss.npz"))
    o = YoloLossRef(model_ref.head_anchors())
    for tag, b, h, w, seed in YOLO_SEQ:
        p = [t.requires_grad_(True) for t in recipes.head_outputs(40 + seed, b, h, w)]
        loss, _ = o(p, recipes.yolo_labels(seed, b))
        loss.backward()
        assert abs(float(loss.detach()) - float(g[tag + "_loss"][0])) <= 1e-6 * abs(float(g[tag + "_loss"][0])), tag
        for i in range(3):
            assert abs(float(p[i].grad.norm()) - float(g[f"{tag}_gnorm{i}"])) <= 1e-5 * float(g[f"{tag}_gnorm{i}"])
    for k, lab in enumerate(recipes.yolo_labels(35, 3)):
        tg = o.build_targets([(12, 16), (6, 8), (3, 4)], lab)
        for i in range(3):
            assert np.array_equal(tg[i], g[f"bt{k}_l{i}"]), (k, i)
    f = YoloLossRef(model_ref.head_anchors())
    for k, lab in enumerate(recipes.yolo_labels(36, 4, max_boxes=5)):
        tg = f.build_targets([(32, 32), (16, 16), (8, 8)], lab)
        for i in range(3):
            assert np.array_equal(tg[i], g[f"fresh{k}_l{i}"]), (k, i)
