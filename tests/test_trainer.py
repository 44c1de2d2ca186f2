"""Optimiser tail + data-parallel plumbing (yolov5m_b200.trainer).

GPU: fused unscale + clip + Adam against torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (training_utils.py:114-122,
train.py:61) on the same gradients: 1e-6 abs on the updated weights (fp32; identical formula, different op fusion).
CPU: the gradient exchange over a world_size-2 gloo group (the N>1 path of bench.py without GPUs).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import recipes
from oracle import model_ref

gpu = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from yolov5m_b200.model import YOLOV5m
    from yolov5m_b200.trainer import GradSync
    torch.manual_seed(rank)  # different init per rank: broadcast must make them equal
    m = YOLOV5m(first_out=48, nc=80, anchors=model_ref.ANCHORS, ch=(192, 384, 768))
    sync = GradSync(m, chunks=4)
    sync.broadcast_parameters(0)
    psum = float(m.flat_params.double().sum())
    g = m.flat_grads
    gen = torch.Generator().manual_seed(100 + rank)
    g.copy_(torch.randn(g.numel(), generator=gen))
    sync.all_reduce()
    exp = sum(torch.randn(g.numel(), generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
    ok = torch.allclose(g, exp, atol=1e-6)
    q.put((rank, sync.world, psum, bool(ok)))
    dist.destroy_process_group()


def test_grad_sync_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [2, 2]
    assert res[0][2] == res[1][2], "parameters differ after broadcast"
    assert all(r[3] for r in res), "all-reduced bucket != sum of the ranks' gradients"


@gpu
def test_fused_clip_adam_matches_torch():
    from yolov5m_b200.model import YOLOV5m
    from yolov5m_b200.trainer import Adam
    m = YOLOV5m(first_out=48, nc=80, anchors=model_ref.ANCHORS, ch=(192, 384, 768)).cuda()
    ref_p = [p.detach().clone().requires_grad_(True) for p in m.parameters()]
    ref_opt = torch.optim.Adam(ref_p, lr=5e-4, weight_decay=5e-4)
    opt = Adam(m, lr=5e-4, weight_decay=5e-4)
    gen = torch.Generator(device="cuda").manual_seed(0)
    scale = 1.0 / 128.0  # e.g. loss scale 64 x world 2
    for it in range(3):
        g = m.flat_grads
        g.copy_(torch.randn(g.numel(), device="cuda", generator=gen) * (50.0 if it == 0 else 0.01))
        for rp, gv in zip(ref_p, m._grad_views(g)):
            rp.grad = (gv * scale).contiguous().clone()
        norm = torch.nn.utils.clip_grad_norm_(ref_p, max_norm=10.0)
        ref_opt.step()
        opt.step(grad_scale=scale, max_norm=10.0)
        # padding slots of the flat bucket hold random junk in this test: compare norms over real parameters only
        mine = torch.cat([gv.reshape(-1) for gv in m._grad_views(g)]).double().norm() * scale
        assert abs(mine.item() - norm.item()) < 1e-4 * norm.item()
        for p, rp in zip(m.parameters(), ref_p):
            assert torch.allclose(p.detach(), rp.detach(), atol=2e-6, rtol=1e-5)
    # the bf16 forward operands follow the masters
    w = dict(m.named_parameters())["neck.7.c_out.cbl.0.weight"]
    r = m._rec_of[m.neck[7].c_out.cbl[0]]
    packed = m._wfwd[r.w_off:r.w_off + w.numel()].view(w.shape[0], 1, 1, w.shape[1]).permute(0, 3, 1, 2)
    assert torch.equal(packed, w.detach().to(torch.bfloat16))


@gpu
def test_dgrad_operand_repack_layout():
    """yb_repack_dgrad: every non-stem conv's bf16 dgrad operand is [Cin][tap][Cout_pad] of the fp32 master weight
    (zero in the padded output channels), bit for bit."""
    from yolov5m_b200.model import YOLOV5m
    m = YOLOV5m(first_out=48, nc=80, anchors=model_ref.ANCHORS, ch=(192, 384, 768)).cuda()
    m.refresh_packed(force=True)
    torch.cuda.synchronize()
    checked = 0
    for r in m._recs:
        if r.is_stem:
            continue
        w = r.conv.weight.detach()                                   # logical (Cout, Cin, k, k)
        co, ci, kh, kw = w.shape
        want = torch.zeros(ci, kh * kw, r.cout_pad, device="cuda", dtype=torch.bfloat16)
        want[:, :, :co] = w.permute(1, 2, 3, 0).reshape(ci, kh * kw, co).to(torch.bfloat16)
        got = m._wdg[r.wt_off:r.wt_off + want.numel()].view_as(want)
        assert torch.equal(got, want), r.name
        checked += 1
    assert checked == 81


@gpu
def test_adam_state_dict_interoperates_with_torch_adam():
    """checkpoint["optimizer"] wire format (reference train.py:140-143, utils/utils.py:74-82): the fused optimiser exports
    the torch.optim.Adam state_dict, a stock torch Adam loads it and continues identically, and the other way round."""
    from yolov5m_b200.model import YOLOV5m
    from yolov5m_b200.trainer import Adam
    m = YOLOV5m(first_out=48, nc=80, anchors=model_ref.ANCHORS, ch=(192, 384, 768)).cuda()
    opt = Adam(m, lr=5e-4, weight_decay=5e-4)
    assert opt.state_dict()["state"] == {} and len(opt.state_dict()["param_groups"][0]["params"]) == 243
    gen = torch.Generator(device="cuda").manual_seed(3)

    def fill():
        m.flat_grads.copy_(torch.randn(m.flat_grads.numel(), device="cuda", generator=gen) * 0.01)
    for _ in range(2):
        fill(); opt.step()
    sd = opt.state_dict()
    assert set(sd["state"]) == set(range(243)) and float(sd["state"][0]["step"]) == 2.0
    p0 = next(m.parameters())
    assert sd["state"][0]["exp_avg"].shape == p0.shape and sd["state"][0]["exp_avg"].is_contiguous()
    # fused -> torch
    ref_p = [p.detach().clone().requires_grad_(True) for p in m.parameters()]
    ref_opt = torch.optim.Adam(ref_p, lr=1e-3, weight_decay=0.0)   # hyper-parameters come from the loaded dict
    ref_opt.load_state_dict(sd)
    assert ref_opt.param_groups[0]["lr"] == 5e-4 and ref_opt.param_groups[0]["weight_decay"] == 5e-4
    fill()
    for rp, gv in zip(ref_p, m._grad_views(m.flat_grads)):
        rp.grad = gv.contiguous().clone()
    ref_opt.step(); opt.step()
    for p, rp in zip(m.parameters(), ref_p):
        assert torch.allclose(p.detach(), rp.detach(), atol=2e-6, rtol=1e-5)
    # torch -> fused (a fresh fused optimiser resumes from the torch optimiser's state)
    m2 = YOLOV5m(first_out=48, nc=80, anchors=model_ref.ANCHORS, ch=(192, 384, 768)).cuda()
    m2.load_state_dict(m.state_dict())
    opt2 = Adam(m2, lr=1.0, weight_decay=0.0)
    opt2.load_state_dict(ref_opt.state_dict())
    assert opt2.step_count == 3 and opt2.lr == 5e-4 and opt2.weight_decay == 5e-4
    fill()
    m2.flat_grads.copy_(m.flat_grads)
    opt.step(); opt2.step()
    for p2, p1 in zip(m2.parameters(), m.parameters()):  # (the alignment padding of the flat buffers is not state)
        assert torch.allclose(p2.detach(), p1.detach(), atol=2e-6, rtol=1e-5)
    # the flat format of earlier versions still loads
    opt2.load_state_dict({"step": 5, "exp_avg": opt.m, "exp_avg_sq": opt.v})
    assert opt2.step_count == 5 and torch.equal(opt2.m, opt.m)


@gpu
def test_train_step_runs_and_learns():
    """TrainStep = train_loop body: loss decreases over a few steps on a fixed batch; uint8 and float inputs agree."""
    import yolov5m_b200 as yb
    from yolov5m_b200.trainer import Adam, TrainStep
    torch.manual_seed(0)
    m = yb.YOLOV5m(first_out=48, nc=80, anchors=yb.ANCHORS, ch=(192, 384, 768)).cuda().train()
    loss_fn = yb.ComputeLoss(m)
    step = TrainStep(m, loss_fn, Adam(m, lr=1e-3), max_norm=10.0)
    x = (recipes.model_input(2, 4, 128, 128) * 255).to(torch.uint8)
    tg = recipes.targets(6, 4, 24)
    losses = [float(step(x.pin_memory(), tg)) for _ in range(8)]
    assert all(l == l for l in losses)
    assert losses[-1] < losses[0], losses
    m.eval()
    with torch.no_grad():
        a = m(x.cuda())
        b = m((x.float() / 255).cuda())
    for u, v in zip(a, b):
        assert torch.allclose(u, v, atol=1e-2, rtol=1e-2)


def test_multi_scale_size_follows_reference_rule():
    """utils/training_utils.py:11-28: longer side -> random multiple of 32 in [320, 640], shorter side scaled by the same
    factor and rounded UP to a multiple of 32 (the reference passes a float to random.randrange, which Python >= 3.12
    rejects; the draw is restated with the same integer bounds)."""
    import math
    import random
    from yolov5m_b200.trainer import multi_scale_size
    for (h, w) in [(640, 640), (480, 640), (640, 352), (1280, 1280)]:
        r1, r2 = random.Random(7), random.Random(7)
        seen = set()
        for _ in range(300):
            nh, nw = multi_scale_size(h, w, 640, 32, rng=r1)
            sz = r2.randrange(320, 672) // 32 * 32
            sf = sz / max(h, w)
            assert (nh, nw) == (math.ceil(h * sf / 32) * 32, math.ceil(w * sf / 32) * 32)
            assert nh % 32 == 0 and nw % 32 == 0 and max(nh, nw) == sz and 320 <= sz <= 640
            seen.add(sz)
        assert seen == set(range(320, 641, 32))


def test_chunk_ready_after_schedule():
    """Overlapped all-reduce: a bucket chunk is released right after the last backward op that writes into it."""
    from yolov5m_b200.trainer import chunk_ready_after
    writes = {0: [(90, 10)], 1: [(50, 40)], 2: [], 3: [(0, 50)], 4: [(60, 0)]}
    assert chunk_ready_after(writes, [0, 25, 50, 75, 100]) == {1: [2, 3], 3: [0, 1]}
    # a chunk nobody writes (alignment padding only) is released with the first op
    assert chunk_ready_after({5: [(0, 10)], 7: [(10, 10)]}, [0, 10, 20, 30]) == {5: [0, 2], 7: [1]}


@gpu
def test_engine_grad_chunk_schedule_covers_the_bucket():
    import yolov5m_b200 as yb
    from yolov5m_b200.trainer import GradSync
    m = yb.YOLOV5m(first_out=48, nc=80, anchors=yb.ANCHORS, ch=(192, 384, 768)).cuda().train()
    eng = m.engine(2, 64, 64, True)
    # every parameter slice is written by exactly one backward op
    got = sorted(sl for w in eng.bwd_writes.values() for sl in w)
    want = sorted((o, n) for o, n in m._poffs)
    assert got == want
    b = GradSync(m, chunks=4)._bounds(m.flat_params.numel())
    sched = eng.grad_chunk_schedule(b)
    order = [c for i in sorted(sched) for c in sched[i]]
    assert sorted(order) == [0, 1, 2, 3]
    first = {c: i for i in sched for c in sched[i]}
    assert first[3] < first[0], "the tail of the bucket (head / neck gradients) must be complete long before its head"
    assert max(first.values()) <= len(eng.bwd_ops) - 1


# ---- round-2 fixes of the advisor findings -----------------------------------------------------------------
def _small_model(seed=0):
    import yolov5m_b200 as yb
    from oracle import model_ref
    m = yb.YOLOV5m(first_out=48, nc=80, anchors=yb.ANCHORS, ch=(192, 384, 768))
    m.load_state_dict({k: v.clone() for k, v in model_ref.make_state_dict(seed).items()})
    return m.cuda().train()


@gpu
def test_train_step_accumulates_micro_batches_like_the_reference_loop():
    """TrainStep(accumulate=2): the bucket after two micro-batches = sum of the two single-batch gradients, and ONE
    optimiser step is taken (training_utils.py:88-90,:116); flush() steps on a partial window"""
    import yolov5m_b200 as yb
    from yolov5m_b200.trainer import Adam, TrainStep, nominal_accumulate
    assert nominal_accumulate(16) == 4 and nominal_accumulate(64) == 1 and nominal_accumulate(128) == 1
    xs = [(recipes.model_input(20 + i, 2, 64, 64) * 255).to(torch.uint8).cuda() for i in range(2)]
    tgs = [recipes.targets(30 + i, 2, 10) for i in range(2)]
    singles = []
    for x, tg in zip(xs, tgs):
        m = _small_model()
        m.expose_param_grads = False
        yb.ComputeLoss(m)(m(x), tg, None).backward()
        singles.append(m.flat_grads.clone())
    m = _small_model()
    opt = Adam(m)
    step = TrainStep(m, yb.ComputeLoss(m), opt, max_norm=0.0, accumulate=2)
    p0 = m.flat_params.clone()
    step(xs[0], tgs[0])
    assert opt.steps_taken() == 0 and torch.equal(m.flat_params, p0)        # no update inside the window
    want = singles[0] + singles[1]
    # intercept the bucket right before the optimiser consumes it
    seen = {}
    real = opt.step

    def spy(**kw):
        seen["g"] = m.flat_grads.clone()
        return real(**kw)
    opt.step = spy
    step(xs[1], tgs[1])
    torch.cuda.synchronize()
    # the second micro-batch sees the same weights but updated BN running statistics only: gradients are unaffected
    assert torch.allclose(seen["g"], want, rtol=1e-5, atol=1e-7)
    assert opt.steps_taken() == 1 and not torch.equal(m.flat_params, p0)
    step(xs[0], tgs[0])
    step.flush()
    assert opt.steps_taken() == 2


@gpu
def test_adam_skips_non_finite_gradients_like_grad_scaler():
    from yolov5m_b200.trainer import Adam
    m = _small_model()
    opt = Adam(m)
    m.flat_grads.normal_(0, 1e-3)
    opt.step(max_norm=10.0)
    p1, m1, v1 = m.flat_params.clone(), opt.m.clone(), opt.v.clone()
    m.flat_grads[12345] = float("inf")
    opt.step(max_norm=10.0)
    torch.cuda.synchronize()
    assert torch.equal(m.flat_params, p1) and torch.equal(opt.m, m1) and torch.equal(opt.v, v1)
    assert opt.steps_taken() == 1                                            # the skipped step does not count
    m.flat_grads[12345] = float("nan")
    opt.step(max_norm=0.0)
    assert torch.equal(m.flat_params, p1) and opt.steps_taken() == 1
    m.flat_grads.normal_(0, 1e-3)
    opt.step(max_norm=10.0)
    assert opt.steps_taken() == 2 and torch.isfinite(m.flat_params).all() and not torch.equal(m.flat_params, p1)


@gpu
def test_extra_gradients_on_the_head_outputs_are_added_not_dropped():
    """fast path (ComputeLoss writes the head gradient operand directly) + an auxiliary term on the same outputs, and two
    losses on the same outputs: the parameter gradients equal those of the generic dense path"""
    import yolov5m_b200 as yb
    x = (recipes.model_input(41, 2, 64, 64) * 255).to(torch.uint8).cuda()
    tg = recipes.targets(42, 2, 10)

    def grads(build_loss, detach_engine):
        m = _small_model()
        out = m(x)
        if detach_engine:            # break the fast-path link: ComputeLoss then returns dense fp32 gradients
            for o in out:
                o._yb_engine = None
        build_loss(m, out).backward()
        return torch.cat([p.grad.detach().flatten() for p in m.parameters()]).clone()

    def aux(m, out):
        return yb.ComputeLoss(m)(out, tg, None) + 1e-3 * sum(o.square().mean() for o in out)

    def twice(m, out):
        lf = yb.ComputeLoss(m)
        return lf(out, tg, None) + 0.5 * lf(out, tg, None)

    for build in (aux, twice):
        fast, dense = grads(build, False), grads(build, True)
        # a dropped contribution would show as ~0.33 (twice) / a missing term (aux); measured 2.2e-2 / < 2e-2: the bf16
        # rounding of the head-gradient operand (rounded once per contribution on the fast path) through the backward pass
        assert float((fast - dense).norm() / dense.norm()) < 5e-2, build.__name__


@gpu
def test_loss_workspace_is_not_leaked_by_forward_only_evaluations():
    import yolov5m_b200 as yb
    m = _small_model()
    lf = yb.ComputeLoss(m)
    x = (recipes.model_input(43, 2, 64, 64) * 255).to(torch.uint8).cuda()
    tg = recipes.targets(44, 2, 10)
    for _ in range(5):
        with torch.no_grad():
            lf(m(x), tg, None)           # validation-style evaluation
        loss = lf(m(x), tg, None)        # differentiable evaluation whose graph is dropped without backward
        del loss
    assert sum(len(v) for v in lf._ws.values()) <= 2
