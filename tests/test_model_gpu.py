"""YOLOV5m drop-in (yolov5m_b200.model) against the golden vectors of the real reference and against the oracle.

Tolerances: the CUDA path stores activations / tensor-core operands in bf16 (fp32 accumulate, fp32 BN statistics,
fp32 master weights).  Integer / layout facts (state_dict keys, shapes) are exact.
  * eval mode (BN = running statistics) is well conditioned: end to end through all 82 convs the outputs agree with
    the reference golden vectors to 3e-2 of the output norm (1.5e-2 against the oracle with the same bf16 storage).
  * train mode with random-init weights is NOT well conditioned: batch-statistic BN re-normalises nearly constant
    feature maps at every layer, which amplifies ANY storage rounding multiplicatively (the oracle itself, evaluated
    with bf16 storage, is 29-58 % away from its own fp32 evaluation on the golden inputs -- measured in the header of
    test_train_forward_layerwise).  So train mode is pinned LAYER BY LAYER with teacher forcing: every layer of the
    engine is compared with the fp32 oracle applied to the engine's own input of that layer (5e-3 of the norm, i.e.
    bf16 output rounding), and the backward pass is compared with fp32 autograd through that same teacher-forced graph.
"""
import numpy as np
import pytest
import torch

import recipes
from oracle import loss_ref, model_ref

gpu = pytest.mark.gpu

# Per-layer forward tolerances (teacher-forced, relative to the layer's output norm) and backward tolerances (relative
# error of each parameter-gradient tensor against fp32 autograd through the teacher-forced graph).
# Measured on B200 (round 1, shapes (2,64,96) / (1,128,128)):
#   same-storage mirror (BN applied to the bf16-rounded conv output, like the engine): median 1.7e-3, max 6.0e-3.  The
#     max is always a residual layer (backbone.6.seq.5.c2): the engine rounds (h + a) and the test recovers h = out - a,
#     so the rounding of the larger sum is measured against the smaller h.
#   fp32-layer mirror (BN applied to the fp32 conv output): median 2.7e-3, max 9.5e-3 (backbone.9.c_out, 16 samples per
#     channel).  The excess over the same-storage figure is the bf16 storage of the RAW conv output amplified by the
#     batch-statistic normalisation (|mean|/std of a near-constant feature map) -- a property of the design (raw output
#     stored once in bf16, statistics taken from the fp32 accumulator), not of a kernel.  This bound is therefore
#     looser than plain bf16 output rounding (4e-3) and is stated as such.
#   heads (fp32 outputs): 2e-7 .. 7e-7;  backbone.0 running mean / var vs the real reference: 5.6e-4 / 4.6e-7.
#   backward: median 1.6e-2 / 1.9e-2, max 2.4e-2 / 2.6e-2 over the 185 / 195 parameter tensors with non-negligible
#     gradients (bf16 gradient storage through 80 layers).
FWD_TOL_SAME_STORAGE = 8e-3
FWD_TOL_FP32_LAYER = 1.5e-2
BWD_TOL_MEDIAN = 3e-2
BWD_TOL_MAX = 5e-2


def rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu(); b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def make_model(dev="cuda", seed=0):
    from yolov5m_b200.model import YOLOV5m
    m = YOLOV5m(first_out=48, nc=80, anchors=model_ref.ANCHORS, ch=(192, 384, 768))
    sd = model_ref.make_state_dict(seed)
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    return m.to(dev), sd


def test_state_dict_layout_cpu():
    """same 481 keys / shapes / order as the reference state_dict (checkpoint compatibility)."""
    from yolov5m_b200.model import YOLOV5m
    m = YOLOV5m(first_out=48, nc=80, anchors=model_ref.ANCHORS, ch=(192, 384, 768))
    sd = m.state_dict()
    specs = model_ref.param_specs()
    assert list(sd.keys()) == [n for n, _, _ in specs]
    for n, shape, _ in specs:
        assert tuple(sd[n].shape) == tuple(shape), n
    assert sum(p.numel() for p in m.parameters()) == 21190557
    assert torch.equal(m.head.anchors, model_ref.head_anchors())
    assert (m.head.nc, m.head.nl, m.head.naxs, m.head.stride) == (80, 3, 3, [8, 16, 32])
    with pytest.raises(Exception):
        m(torch.rand(1, 3, 64, 64))  # no CPU fallback
    with pytest.raises(AssertionError):
        m(torch.rand(1, 3, 65, 64))  # model.py:211


@gpu
@pytest.mark.parametrize("tag,shape", [("a", (2, 64, 96)), ("b", (1, 128, 128))])
def test_forward_eval(golden, tag, shape):
    g = golden["model"]
    m, sd = make_model()
    m.eval()
    x = recipes.model_input(11, *shape)
    with torch.no_grad():
        out = m(x.cuda())
        ref_q = model_ref.forward(sd, x, train=False, quant=True)
    assert isinstance(out, list) and len(out) == 3
    for i in range(3):
        assert tuple(out[i].shape) == g[f"{tag}_eval_p{i}"].shape and out[i].dtype == torch.float32
        assert out[i].is_contiguous()
        assert rel(out[i], ref_q[i]) < 1.5e-2, (i, rel(out[i], ref_q[i]))
        assert rel(out[i], g[f"{tag}_eval_p{i}"]) < 3e-2, (i, rel(out[i], g[f"{tag}_eval_p{i}"]))



class _Mirror(model_ref.Net):
    """Oracle network whose every CBL output is replaced (value only, gradient flows through the oracle op) by the
    engine's own stored output: `teacher forcing`.  Records the per-layer discrepancy before the substitution.
    quant=False: pure fp32 layer (BN applied to the fp32 conv output) -- differentiable, used for the backward check.
    quant=True : same bf16 storage point as the engine (BN applied to the bf16-rounded conv output) -- forward only."""

    def __init__(self, sd, eng, quant=False):
        super().__init__(sd, train=True, quant=quant, update_stats=False)
        self.errs, self.eng_out = {}, {}
        for rec in eng.tape:
            if rec[0] != "cbl":
                continue
            _, r, xin, out, y, res, up, ptrs = rec[:8]
            o = out.tensor().float().cpu().permute(0, 3, 1, 2)
            if res is not None:
                o = o - res.tensor().float().cpu().permute(0, 3, 1, 2)  # engine stores silu(bn(y)) + residual
            self.eng_out[r.name] = o

    def cbl(self, x, name, k, s, p):
        out = super().cbl(x, name, k, s, p)
        tgt = self.eng_out[name]
        self.errs[name] = ((out.detach().double() - tgt.double()).norm() / tgt.double().norm().clamp_min(1e-30)).item()
        return out + (tgt - out).detach()


def _mirror_state(m):
    sd = {k: v.detach().float().cpu().clone() for k, v in m.state_dict().items()}
    leaves = {}
    for n, p in m.named_parameters():
        v = sd[n]
        if v.dim() == 4:
            v = v.to(torch.bfloat16).float()  # tensor-core operands are the bf16 copies of the fp32 masters
        leaves[n] = v.contiguous().clone().requires_grad_(True)
        sd[n] = leaves[n]
    return sd, leaves


@gpu
@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 128, 128)])
def test_train_forward_layerwise_and_backward(golden, shape):
    """Measured on the golden inputs (CPU, oracle only): rel. distance between the oracle with bf16 storage and the
    oracle in fp32, train mode = 0.29 / 0.41 / 0.59 (P3/P4/P5) for (2,64,96) -- an end-to-end bf16-vs-fp32 comparison
    is meaningless there, hence the teacher-forced per-layer comparison below."""
    b, h, w = shape
    m, sd0 = make_model()
    m.train()
    x = recipes.model_input(11, b, h, w)
    out = m(x.cuda())
    eng = out[0]._yb_engine
    sd, leaves = _mirror_state(m)
    xq = x.to(torch.bfloat16).float()
    with torch.no_grad():                   # (1) same bf16 storage points as the engine
        mq = _Mirror({k: v.detach() for k, v in sd.items()}, eng, quant=True)
        mq.forward(xq)
    mir = _Mirror(sd, eng)                  # (2) pure fp32 layers, differentiable
    pt = mir.forward(xq)
    wq = max(mq.errs.items(), key=lambda kv: kv[1])
    wf = max(mir.errs.items(), key=lambda kv: kv[1])
    print("\nlayerwise fwd rel err %s: same-storage mirror median %.5f max %.5f (%s); fp32 mirror median %.5f max %.5f (%s)" % (
        shape, np.median(list(mq.errs.values())), wq[1], wq[0], np.median(list(mir.errs.values())), wf[1], wf[0]))
    head_errs = [rel(out[i], pt[i]) for i in range(3)]
    print("head rel err", head_errs)
    checks = [(len(mir.errs) == 79, "79 CBL layers"), (wq[1] < FWD_TOL_SAME_STORAGE, ("same-storage", wq)),
              (wf[1] < FWD_TOL_FP32_LAYER, ("fp32 layer", wf)), (max(head_errs) < 2e-3, ("heads", head_errs))]
    # running statistics of the first layer (well conditioned) against the real reference
    if shape == (2, 64, 96):
        g = golden["model"]
        s = m.state_dict()
        rs = (rel(s["backbone.0.cbl.1.running_mean"], g["a_rm_b0"]), rel(s["backbone.0.cbl.1.running_var"], g["a_rv_b0"]))
        print("backbone.0 running mean/var rel err vs reference", rs)
        checks.append((max(rs) < 5e-3, ("running stats", rs)))
        checks.append((int(s["backbone.0.cbl.1.num_batches_tracked"]) == 1, "num_batches_tracked"))
    # ---- backward: fp32 autograd through the teacher-forced graph vs the engine
    for t in pt:
        t.retain_grad()
    loss = loss_ref.compute_loss(pt, recipes.targets(5, b, 8 * b), sd["head.anchors"])
    loss.backward()
    torch.autograd.backward(out, [t.grad.cuda() for t in pt])
    prm = dict(m.named_parameters())
    gmax = max(v.grad.norm().item() for v in leaves.values())
    errs = {}
    for n, leaf in leaves.items():
        if leaf.grad.norm().item() < 1e-3 * gmax:
            continue
        errs[n] = rel(prm[n].grad, leaf.grad)
    vals = np.array(list(errs.values()))
    worst = max(errs.items(), key=lambda kv: kv[1])
    print("backward rel err: median %.4f max %.4f (%s) over %d tensors" % (np.median(vals), vals.max(), worst[0], len(vals)))
    checks.append((np.median(vals) < BWD_TOL_MEDIAN and vals.max() < BWD_TOL_MAX, ("backward", np.median(vals), worst)))
    failed = [what for ok, what in checks if not ok]
    assert not failed, failed


@gpu
def test_grad_accumulation_and_flat_views():
    """two backward calls accumulate into .grad like torch autograd; .grad tensors are views of the flat bucket."""
    m, _ = make_model()
    m.train()
    x = recipes.model_input(3, 1, 64, 64).cuda()
    out = m(x)
    sum(o.square().mean() for o in out).backward()
    g1 = {n: p.grad.clone() for n, p in m.named_parameters()}
    out = m(x)
    sum(o.square().mean() for o in out).backward()
    w = "neck.7.c_out.cbl.0.weight"
    p = dict(m.named_parameters())[w]
    assert p.grad.shape == p.shape
    assert rel(p.grad, 2 * g1[w]) < 1e-3   # same input, same batch statistics: exactly twice the first gradient
    assert m.flat_grads.numel() == m.flat_params.numel()


@gpu
def test_submodule_forward_matches_oracle_blocks():
    """Blocks called on their own (``model.backbone[0](img)``, reference model.py:27,49,89,106): inference semantics on the
    fused kernels, against the oracle's block functions with bf16 storage rounding (quant=True).  A block in training mode
    under grad refuses instead of returning a tensor that would not train."""
    from yolov5m_b200 import _lib
    m, sd = make_model()
    m.eval()
    ref = model_ref.Net({k: v.clone() for k, v in sd.items()}, train=False, quant=True)
    g = torch.Generator().manual_seed(5)
    img = torch.rand(2, 3, 64, 96, generator=g)
    with torch.no_grad():
        got = m.backbone[0](img.cuda())                       # stem CBL: the image through the space-to-depth staging
        want = ref.cbl(img, "backbone.0", 6, 2, 2)
        assert got.shape == want.shape and got.dtype == torch.float32
        assert rel(got.cpu(), want) < 1e-2
        x48 = torch.randn(2, 48, 32, 48, generator=g)
        got = m.backbone[1](x48.cuda()); want = ref.cbl(x48, "backbone.1", 3, 2, 1)
        assert got.shape == want.shape and rel(got.cpu(), want) < 1e-2
        x96 = torch.randn(2, 96, 16, 24, generator=g)
        got = m.backbone[2](x96.cuda()); want = ref.c3(x96, "backbone.2", 2, True)     # C3 with fused c1 || c_skipped
        assert got.shape == want.shape and rel(got.cpu(), want) < 2e-2
        x48b = torch.randn(2, 48, 16, 24, generator=g)
        blk = m.backbone[2].seq[0]                                                   # Bottleneck: c2(c1(x)) + x
        h = ref.cbl(x48b, "backbone.2.seq.0.c1", 1, 1, 0)
        want = ref.cbl(h, "backbone.2.seq.0.c2", 3, 1, 1) + x48b.to(torch.bfloat16).float()
        got = blk(x48b.cuda())
        assert got.shape == want.shape and rel(got.cpu(), want) < 2e-2
        x768 = torch.randn(2, 768, 4, 6, generator=g)
        got = m.backbone[9](x768.cuda()); want = ref.sppf(x768, "backbone.9")
        assert got.shape == want.shape and rel(got.cpu(), want) < 2e-2
        x384 = torch.randn(2, 384, 8, 12, generator=g)
        got = m.neck[3](x384.cuda()); want = ref.c3(x384, "neck.3", 2, False)       # neck C3 (no shortcut)
        assert got.shape == want.shape and rel(got.cpu(), want) < 2e-2
        # the full forward still works afterwards and a second call reuses the cached plan
        outs = m(img.cuda())
        assert [tuple(o.shape) for o in outs] == [(2, 3, 8, 12, 85), (2, 3, 4, 6, 85), (2, 3, 2, 3, 85)]
        again = m.backbone[1](x48.cuda())
        assert torch.equal(again, m.backbone[1](x48.cuda()))
    m.train()
    with pytest.raises(_lib.YBError):
        m.backbone[1](x48.cuda())
    with pytest.raises(_lib.YBError):
        m.backbone[1].cbl[0](x48.cuda())   # a bare conv / BN holder has no arithmetic of its own
    with pytest.raises(ValueError):
        with torch.no_grad():
            m.backbone[1](x96.cuda())      # wrong channel count
