"""Seeded input recipes shared by the golden generator and the tests.

Everything is produced by CPU ``torch.Generator`` streams, which are
bit-identical on every machine with the same torch build, so only the
*outputs* of the reference need to be committed as fixtures.
"""
import torch


def model_input(seed, b, h, w):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, 3, h, w, generator=g)


def targets(seed, b, nt):
    """SURVEY 8(d) recipe: [img, cls, x, y, w, h] normalised."""
    g = torch.Generator().manual_seed(seed)
    if nt == 0:
        return torch.zeros(0, 6)
    return torch.cat([
        torch.randint(0, b, (nt, 1), generator=g).float(),
        torch.randint(0, 80, (nt, 1), generator=g).float(),
        torch.rand(nt, 2, generator=g),
        torch.rand(nt, 2, generator=g) * 0.5 + 0.005,
    ], 1)


def edge_targets(b=2):
    """G4: boxes on cell boundaries, within one cell of the border, and with
    w/h ratios exactly at the anchor_t=4 threshold (anchor (1.25,1.625) at P3/80)."""
    rows = [
        [0, 1, 0.5, 0.5, 0.1, 0.1],            # x*W integral at every level
        [0, 2, 1.0 / 80, 1.0 / 80, 0.05, 0.05],  # gxy == 1 at P3 (not > 1)
        [1, 3, 0.999, 0.999, 0.2, 0.3],        # last cell
        [1, 4, 0.0, 0.0, 0.3, 0.2],            # first cell, offsets would go negative
        [0, 5, 0.25, 0.75, 4 * 1.25 / 80, 4 * 1.625 / 80],   # ratio == 4 exactly (fails <)
        [1, 6, 0.25, 0.75, 1.25 / 80 / 4, 1.625 / 80 / 4],   # ratio == 1/4 exactly
        [0, 7, 0.50625, 0.49375, 0.15, 0.15],  # frac 0.5 boundary at P3: 40.5 / 39.5
        [1, 79, 0.3, 0.6, 0.9, 0.9],           # huge box, matches only P5 anchors
        [0, 0, 0.3, 0.6, 0.001, 0.001],        # tiny box, matches nothing
        [0, 9, 0.3, 0.6, 0.1, 0.1],
        [0, 10, 0.3, 0.6, 0.1, 0.1],           # duplicate cell (last write wins in tobj)
    ]
    t = torch.tensor(rows, dtype=torch.float32)
    t[:, 0] = t[:, 0].clamp(max=b - 1)
    return t


def head_outputs(seed, b, h, w, nc=80, scale=1.0):
    """Random stand-ins for the three head tensors (B,3,H/s,W/s,5+nc)."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(b, 3, h // s, w // s, 5 + nc, generator=g) * scale for s in (8, 16, 32)]


def nms_boxes(seed, b, n, mode="realistic", size=640.0):
    """(B,N,6) [cls, score, cx, cy, w, h] decoded boxes (SURVEY 8(d) config 5 recipe).
    modes: realistic (score = sigmoid(N(-5,2))), allpass (sigma ~ 0.5), ties (quantised
    scores => many exact ties), clustered (heavy overlap)."""
    g = torch.Generator().manual_seed(seed)
    cls = torch.randint(0, 80, (b, n, 1), generator=g).float()
    if mode == "realistic":
        score = torch.sigmoid(torch.randn(b, n, 1, generator=g) * 2 - 5)
    elif mode == "allpass":
        score = torch.sigmoid(torch.randn(b, n, 1, generator=g) * 0.02)
    elif mode == "ties":
        score = torch.round(torch.rand(b, n, 1, generator=g) * 20) / 20
    elif mode == "clustered":
        score = torch.rand(b, n, 1, generator=g)
    else:
        raise ValueError(mode)
    if mode == "clustered":
        cxy = 0.5 * size + torch.randn(b, n, 2, generator=g) * 20
        wh = 60 + torch.rand(b, n, 2, generator=g) * 20
        cls = torch.randint(0, 3, (b, n, 1), generator=g).float()
    else:
        cxy = torch.rand(b, n, 2, generator=g) * size
        wh = torch.exp(torch.randn(b, n, 2, generator=g) * 0.8 + 4)
    return torch.cat([cls, score, cxy, wh], -1)


def yolo_labels(seed, b, max_boxes=6):
    """per-image label arrays for the reference's default loss (loss.py:64): tuple of (n_i, 5) float64
    [class, x, y, w, h] like Training_Dataset.collate_fn yields (np.loadtxt rows); some images have no boxes,
    some boxes repeat a cell (exercises the `anchor_taken` / ignore rules)."""
    import numpy as np
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(b):
        n = int(torch.randint(0, max_boxes + 1, (1,), generator=g))
        cls = torch.randint(0, 80, (n, 1), generator=g).double()
        xy = torch.rand(n, 2, generator=g, dtype=torch.float64) * 0.98 + 0.01
        wh = torch.exp(torch.randn(n, 2, generator=g, dtype=torch.float64) * 0.9 - 2.2).clamp(0.01, 0.95)
        lab = torch.cat([cls, xy, wh], 1).numpy().astype(np.float64)
        if n >= 2 and i % 2 == 0:
            lab[1, 1:3] = lab[0, 1:3]        # same centre as box 0: same cells on every level
            lab[1, 3:5] = lab[0, 3:5] * 1.05  # near-identical shape: same anchor ranking
        out.append(lab)
    return tuple(out)


class StubModel:
    """stands in for the network in the eval-glue fixtures: returns fixed head tensors batch after batch"""

    class _Head:
        def __init__(self, anchors):
            self.nc, self.nl, self.naxs = 80, 3, 3
            self.anchors = anchors
            self.stride = [8, 16, 32]

    def __init__(self, outputs, device=None):
        from oracle import model_ref
        self.head = StubModel._Head(model_ref.head_anchors() if device is None else model_ref.head_anchors().to(device))
        self.outputs, self.k, self.device, self.training = outputs, 0, device, False

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        self.training = mode
        return self

    def __call__(self, images):
        out = self.outputs[self.k % len(self.outputs)]
        self.k += 1
        return [o.clone() if self.device is None else o.to(self.device) for o in out]


def eval_batches(nb=2, b=2, h=128, w=96):
    """[(uint8 images, [label tensors (B,3,H/s,W/s,6)], [head tensors])] for the eval-glue fixtures; the label tensors
    are Validation_Dataset-style dense targets [x_cell, y_cell, w_cell, h_cell, objectness, class]"""
    out = []
    for k in range(nb):
        g = torch.Generator().manual_seed(900 + k)
        img = torch.randint(0, 256, (b, 3, h, w), dtype=torch.uint8, generator=g)
        labels = []
        for s in (8, 16, 32):
            t = torch.zeros(b, 3, h // s, w // s, 6)
            n = 5
            bi = torch.randint(0, b, (n,), generator=g); ai = torch.randint(0, 3, (n,), generator=g)
            yi = torch.randint(0, h // s, (n,), generator=g); xi = torch.randint(0, w // s, (n,), generator=g)
            t[bi, ai, yi, xi, 0:2] = torch.rand(n, 2, generator=g)
            t[bi, ai, yi, xi, 2:4] = torch.rand(n, 2, generator=g) * 3 + 0.3
            t[bi, ai, yi, xi, 4] = 1.0
            t[bi, ai, yi, xi, 5] = torch.randint(0, 80, (n,), generator=g).float()
            labels.append(t)
        heads = head_outputs(950 + k, b, h, w, scale=1.5)
        for lv, t in zip(heads, labels):   # make the "network" right on most labelled cells so the accuracies are not trivial
            obj = t[..., 4] == 1
            cls = t[..., 5][obj].long()
            rows = lv[obj]
            rows[torch.arange(rows.shape[0]) % 4 != 0, 5 + cls[torch.arange(rows.shape[0]) % 4 != 0]] += 8.0
            lv[obj] = rows
        out.append((img, labels, heads))
    return out
