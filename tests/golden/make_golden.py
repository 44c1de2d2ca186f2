"""Generate the golden fixtures by running the REAL reference (/root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Imports the reference under the stub shim of SURVEY.md Appendix C, feeds it the
seeded recipes of tests/golden/recipes.py and deterministic weights from
oracle.model_ref.make_state_dict, and writes small .npz fixtures next to this
file.  The tests then check the oracle (CPU) and the CUDA path against them.
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
from model import YOLOV5m  # noqa: E402  (reference)
from ultralytics_loss import ComputeLoss  # noqa: E402  (reference)
from utils.plot_utils import cells_to_bboxes  # noqa: E402  (reference)
from utils.bboxes_utils import non_max_suppression, intersection_over_union  # noqa: E402  (reference)

import recipes  # noqa: E402
from oracle import model_ref  # noqa: E402

torch.use_deterministic_algorithms(True)
torch.set_num_threads(8)


def sample_idx(numel, k=64, seed=7):
    g = torch.Generator().manual_seed(seed + numel % 9973)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def ref_model(sd):
    m = YOLOV5m(first_out=48, nc=80, anchors=config.ANCHORS, ch=(192, 384, 768))
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    return m


def golden_model():
    """G1: train-mode and eval-mode forward, loss, and parameter gradients."""
    out = {}
    for tag, (b, h, w) in {"a": (2, 64, 96), "b": (1, 128, 128)}.items():
        sd = model_ref.make_state_dict(seed=0)
        x = recipes.model_input(11, b, h, w)
        m = ref_model(sd)
        m.eval()
        with torch.no_grad():
            pe = m(x)
        m.train()
        p = m(x)
        tg = recipes.targets(5, b, 8 * b)
        loss_fn = ComputeLoss(m)
        config.IMAGE_SIZE = 640
        loss = loss_fn([t for t in p], tg, pred_size=None)
        loss.backward()
        for i in range(3):
            out[f"{tag}_eval_p{i}"] = pe[i].numpy()
            out[f"{tag}_train_p{i}"] = p[i].detach().numpy()
        out[f"{tag}_loss"] = loss.detach().numpy()
        names, norms, samples = [], [], []
        for n_, prm in m.named_parameters():
            names.append(n_)
            norms.append(prm.grad.norm().item())
            samples.append(prm.grad.flatten()[sample_idx(prm.numel())].numpy())
        out[f"{tag}_grad_names"] = np.array(names)
        out[f"{tag}_grad_norms"] = np.array(norms, np.float64)
        out[f"{tag}_grad_samples"] = np.concatenate(samples)
        sd2 = m.state_dict()
        out[f"{tag}_rm_b0"] = sd2["backbone.0.cbl.1.running_mean"].numpy()
        out[f"{tag}_rv_b0"] = sd2["backbone.0.cbl.1.running_var"].numpy()
        out[f"{tag}_rm_n7"] = sd2["neck.7.c_out.cbl.1.running_mean"].numpy()
        out[f"{tag}_rv_n7"] = sd2["neck.7.c_out.cbl.1.running_var"].numpy()
    np.savez_compressed(os.path.join(HERE, "model.npz"), **out)
    print("model.npz", {k: getattr(v, "shape", None) for k, v in list(out.items())[:6]})


class _Head:
    def __init__(self):
        self.nc, self.nl, self.naxs = 80, 3, 3
        self.anchors = model_ref.head_anchors()
        self.stride = [8, 16, 32]


class _FakeModel:
    def __init__(self):
        self.head = _Head()
        self._p = torch.nn.Parameter(torch.zeros(1))

    def parameters(self):
        return iter([self._p])


def golden_loss():
    """G2-G4: build_targets + loss parts + dL/dp samples on random head tensors."""
    out = {}
    cases = {
        "rand": (4, 160, 160, recipes.targets(3, 4, 48)),
        "many": (8, 128, 96, recipes.targets(4, 8, 200)),
        "zero": (2, 64, 64, recipes.targets(0, 2, 0)),
        "edge": (2, 640, 640, recipes.edge_targets(2)),
    }
    loss_fn = ComputeLoss(_FakeModel())
    for tag, (b, h, w, tg) in cases.items():
        p = [t.requires_grad_(True) for t in recipes.head_outputs(21, b, h, w)]
        tcls, tbox, indices, anch = loss_fn.build_targets(p, tg)
        for i in range(3):
            bb, aa, gj, gi = indices[i]
            out[f"{tag}_idx{i}"] = torch.stack([bb, aa, gj, gi], 0).numpy().astype(np.int64).reshape(4, -1)
            out[f"{tag}_tbox{i}"] = tbox[i].numpy().reshape(-1, 4)
            out[f"{tag}_anch{i}"] = anch[i].numpy().reshape(-1, 2)
            out[f"{tag}_tcls{i}"] = tcls[i].numpy().astype(np.int64).reshape(-1)
        # loss parts: re-run the reference body with logging hooks replaced by arithmetic
        loss = loss_fn(p, tg, pred_size=None)
        loss.backward()
        out[f"{tag}_loss"] = loss.detach().numpy()
        for i in range(3):
            gflat = p[i].grad.flatten()
            out[f"{tag}_gnorm{i}"] = np.array(gflat.norm().item())
            out[f"{tag}_gobj{i}"] = p[i].grad[..., 4].flatten()[sample_idx(p[i].grad[..., 4].numel())].numpy()
            nz = torch.nonzero(p[i].grad[..., 0].flatten()).flatten()[:32]
            out[f"{tag}_gnz_idx{i}"] = nz.numpy()
            out[f"{tag}_gnz_val{i}"] = p[i].grad.reshape(-1, 85)[nz].numpy()
    np.savez_compressed(os.path.join(HERE, "loss.npz"), **out)
    print("loss.npz", {t: out[f"{t}_loss"] for t in cases}, {t: out[f"{t}_idx0"].shape for t in cases})


def golden_iou():
    g = torch.Generator().manual_seed(9)
    a = torch.rand(256, 4, generator=g); b = torch.rand(256, 4, generator=g)
    a[:, 2:] += 0.05; b[:, 2:] += 0.05
    b[:8] = a[:8]  # identical boxes
    b[8:16, :2] += 5  # disjoint
    out = dict(giou=intersection_over_union(a, b, GIoU=True).numpy(),
               iou=intersection_over_union(a, b, GIoU=False).numpy())
    np.savez_compressed(os.path.join(HERE, "iou.npz"), **out)


def golden_decode_nms():
    """G5: decode + NMS keep sets (bit-exact) on seeded boxes."""
    out = {}
    anchors = model_ref.head_anchors()
    p = recipes.head_outputs(31, 2, 64, 96, scale=2.0)
    dec = cells_to_bboxes([t.clone() for t in p], anchors, [8, 16, 32], is_pred=True, to_list=False)
    out["decode"] = dec.numpy()
    cases = {
        "realistic": (recipes.nms_boxes(1, 3, 4000, "realistic"), 0.45, 0.25),
        "allpass": (recipes.nms_boxes(2, 2, 2500, "allpass"), 0.45, 0.25),
        "ties": (recipes.nms_boxes(3, 2, 1500, "ties"), 0.45, 0.25),
        "clustered": (recipes.nms_boxes(4, 2, 3000, "clustered"), 0.6, 0.01),
        "none": (recipes.nms_boxes(5, 2, 100, "realistic") * torch.tensor([1, 0.0, 1, 1, 1, 1]), 0.45, 0.25),
        "decoded": (dec, 0.45, 0.25),
    }
    for tag, (bx, iou_t, thr) in cases.items():
        before = bx.clone()
        res = non_max_suppression(bx, iou_threshold=iou_t, threshold=thr, max_detections=300, tolist=True)
        assert torch.equal(before, bx), "reference mutated its input"
        out[f"{tag}_counts"] = np.array([len(r) for r in res], np.int64)
        flat = [row for r in res for row in r]
        out[f"{tag}_rows"] = np.array(flat, np.float32).reshape(-1, 6)
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **out)
    print("nms.npz", {t: out[f"{t}_counts"] for t in cases})


if __name__ == "__main__":
    golden_iou()
    golden_loss()
    golden_decode_nms()
    golden_model()
