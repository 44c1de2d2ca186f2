"""Oracle: restatement of ComputeLoss (reference ultralytics_loss.py).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

build_targets  follows ultralytics_loss.py:122-311 -- written here as explicit
               numpy float32 loops so that the *row order* the reference gets
               from boolean-mask indexing is spelled out:
                 rows of level i = for off in 0..4: for a in 0..2: for t in 0..nt-1
                 (mask t[j] is row-major over (anchor, target) :213, then
                 t.repeat(5,1,1)[j] is offset-major :248).
compute_loss   follows ultralytics_loss.py:60-120 with
               intersection_over_union(GIoU=True) of utils/bboxes_utils.py:33-87.
"""
import numpy as np
import torch

F32 = np.float32
ANCHOR_T = F32(4.0)  # ultralytics_loss.py:35
BALANCE = [4.0, 1.0, 0.4]  # :37
OFF = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1]], dtype=F32) * F32(0.5)  # :151-160


def build_targets(targets, anchors, shapes):
    """targets (nt,6) float32 [img,cls,x,y,w,h]; anchors (nl,na,2) float32
    (stride-divided); shapes list of (B,na,H,W,no).
    Returns per level: b,a,gj,gi (int64), tbox (n,4) f32, anch (n,2) f32, tcls (n,) int64."""
    targets = np.asarray(targets, dtype=F32).reshape(-1, 6)
    anchors = np.asarray(anchors, dtype=F32)
    nt = targets.shape[0]
    na = anchors.shape[1]
    out = []
    for i, shape in enumerate(shapes):
        H, W = int(shape[2]), int(shape[3])
        gw, gh = F32(W), F32(H)
        rows = []  # (img, cls, gx, gy, gw, gh, a) after the anchor-ratio filter, (a, t) order
        if nt:
            for a in range(na):
                aw, ah = anchors[i, a, 0], anchors[i, a, 1]
                for t in range(nt):
                    x = targets[t, 2] * gw
                    y = targets[t, 3] * gh
                    w = targets[t, 4] * gw
                    h = targets[t, 5] * gh
                    rw = w / aw
                    rh = h / ah
                    m = max(max(rw, F32(1.0) / rw), max(rh, F32(1.0) / rh))
                    if m < ANCHOR_T:  # :195
                        rows.append((targets[t, 0], targets[t, 1], x, y, w, h, a))
            sel = []  # (row, offset index), offset-major
            for o in range(5):
                for r in rows:
                    gx, gy = r[2], r[3]
                    ix, iy = gw - gx, gh - gy  # gxi, :222-226
                    if o == 0:
                        take = True
                    elif o == 1:
                        take = (np.fmod(gx, F32(1.0)) < F32(0.5)) and gx > F32(1.0)
                    elif o == 2:
                        take = (np.fmod(gy, F32(1.0)) < F32(0.5)) and gy > F32(1.0)
                    elif o == 3:
                        take = (np.fmod(ix, F32(1.0)) < F32(0.5)) and ix > F32(1.0)
                    else:
                        take = (np.fmod(iy, F32(1.0)) < F32(0.5)) and iy > F32(1.0)
                    if take:
                        sel.append((r, o))
        else:
            sel = []  # ultralytics_loss.py:262-265 -> targets[0] is (0,7): no rows
        n = len(sel)
        b = np.zeros(n, np.int64); a_ = np.zeros(n, np.int64)
        gj = np.zeros(n, np.int64); gi = np.zeros(n, np.int64)
        tbox = np.zeros((n, 4), F32); anch = np.zeros((n, 2), F32); tcls = np.zeros(n, np.int64)
        for k, (r, o) in enumerate(sel):
            gx, gy = r[2], r[3]
            ci = np.int64(np.trunc(gx - OFF[o, 0]))  # .long() truncates toward zero, :278
            cj = np.int64(np.trunc(gy - OFF[o, 1]))
            ci = min(max(ci, 0), W - 1)  # clamp_ in place (:285) -> tbox sees the clamped cell
            cj = min(max(cj, 0), H - 1)
            b[k] = np.int64(np.trunc(r[0])); tcls[k] = np.int64(np.trunc(r[1])); a_[k] = r[6]
            gi[k] = ci; gj[k] = cj
            tbox[k] = (gx - F32(ci), gy - F32(cj), r[4], r[5])  # :296
            anch[k] = anchors[i, r[6]]
        out.append(dict(b=b, a=a_, gj=gj, gi=gi, tbox=tbox, anch=anch, tcls=tcls))
    return out


def giou_midpoint(p, t, eps=1e-7):
    """utils/bboxes_utils.py:33-87 with box_format='midpoint', GIoU=True. (n,4)x(n,4)->(n,)"""
    b1x1, b1x2 = p[:, 0] - p[:, 2] / 2, p[:, 0] + p[:, 2] / 2
    b1y1, b1y2 = p[:, 1] - p[:, 3] / 2, p[:, 1] + p[:, 3] / 2
    b2x1, b2x2 = t[:, 0] - t[:, 2] / 2, t[:, 0] + t[:, 2] / 2
    b2y1, b2y2 = t[:, 1] - t[:, 3] / 2, t[:, 1] + t[:, 3] / 2
    w1, h1, w2, h2 = b1x2 - b1x1, b1y2 - b1y1, b2x2 - b2x1, b2y2 - b2y1
    inter = (torch.min(b1x2, b2x2) - torch.max(b1x1, b2x1)).clamp(0) * \
            (torch.min(b1y2, b2y2) - torch.max(b1y1, b2y1)).clamp(0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = torch.max(b1x2, b2x2) - torch.min(b1x1, b2x1)
    ch = torch.max(b1y2, b2y2) - torch.min(b1y1, b2y1)
    c_area = cw * ch + eps
    return iou - (c_area - union) / c_area


def bce_logits(x, t):
    """BCEWithLogitsLoss(pos_weight=1), elementwise (ultralytics_loss.py:25-26)."""
    return x.clamp(min=0) - x * t + torch.log1p(torch.exp(-x.abs()))


def compute_loss(p, targets, anchors, nc=80, image_size=640, return_parts=False):
    """p: list of 3 float32 tensors (B,3,H,W,5+nc) (may require grad).
    Returns loss tensor shape (1,) = (lbox+lobj+lcls)*bs, optionally the three
    weighted parts and the per-level build_targets output."""
    nl = len(p)
    lam_cls = 0.5 * (nc / 80 * 3 / nl)  # :31
    lam_obj = 1.0 * ((image_size / 640) ** 2 * 3 / nl)  # :32
    lam_box = 0.05 * (3 / nl)  # :33
    tg = build_targets(targets.detach().cpu().numpy() if torch.is_tensor(targets) else targets,
                       anchors.detach().cpu().numpy() if torch.is_tensor(anchors) else anchors,
                       [tuple(pi.shape) for pi in p])
    lbox = torch.zeros(1); lobj = torch.zeros(1); lcls = torch.zeros(1)
    for i, pi in enumerate(p):
        t = tg[i]
        n = t["b"].shape[0]
        tobj = torch.zeros(pi.shape[:4], dtype=torch.float32)
        if n:
            b, a = torch.from_numpy(t["b"]), torch.from_numpy(t["a"])
            gj, gi = torch.from_numpy(t["gj"]), torch.from_numpy(t["gi"])
            ps = pi[b, a, gj, gi]  # (n, 5+nc)
            pxy = ps[:, 0:2].sigmoid() * 2 - 0.5  # :81
            pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * torch.from_numpy(t["anch"])  # :82
            giou = giou_midpoint(torch.cat((pxy, pwh), 1), torch.from_numpy(t["tbox"]))
            lbox = lbox + (1.0 - giou).mean()  # :85
            v = giou.detach().clamp(0).numpy()
            tobj_np = tobj.numpy()
            for k in range(n):  # duplicate cells: the LAST row wins (:89, CPU index_put_)
                tobj_np[t["b"][k], t["a"][k], t["gj"][k], t["gi"][k]] = v[k]
            if nc > 1:
                tc = torch.zeros(n, nc)
                tc[torch.arange(n), torch.from_numpy(t["tcls"])] = 1
                lcls = lcls + bce_logits(ps[:, 5:], tc).mean()  # :93-95
        lobj = lobj + bce_logits(pi[..., 4], tobj).mean() * BALANCE[i]  # :101-102
    lbox = lbox * lam_box; lobj = lobj * lam_obj; lcls = lcls * lam_cls
    bs = p[0].shape[0]
    loss = (lbox + lobj + lcls) * bs  # :120
    if return_parts:
        return loss, (lbox, lobj, lcls), tg
    return loss
