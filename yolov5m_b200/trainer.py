"""Training-step plumbing around the hot path: the fused optimiser tail, the data-parallel gradient exchange and the
step function that mirrors the body of the reference's ``train_loop`` (reference utils/training_utils.py:97-122).

    reference (per batch)                                   here
    ---------------------------------------------------     -----------------------------------------------------------
    images.float()/255 ; .to(DEVICE)        (:98,:102)      uint8 H2D (4x fewer bytes) + /255 fused into the stem staging
    out = model(images)                     (:107)          YOLOV5m engine (csrc/conv_*.cu, elementwise.cu)
    loss = loss_fn(out, bboxes, ...)        (:108)          ComputeLoss kernels (csrc/loss.cu)
    scaler.scale(loss).backward()           (:114)          engine backward; gradients land in ONE flat fp32 bucket
    --- (no data parallelism in the reference) ---          all-reduce of that bucket over NCCL / NVLink
    scaler.unscale_ ; clip_grad_norm_(10)   (:117-118)      yb_grad_norm + yb_adam_step: unscale, 1/world, clip and
    scaler.step(optim) [Adam, L2 wd]        (:119, train.py:61)   Adam in one pass over the bucket, which also writes the
    optim.zero_grad(set_to_none=True)       (:121)          bf16 tensor-core operands of the next forward

One process per GPU; `torch.distributed` (NCCL) is only plumbing.  BatchNorm statistics stay local to each rank, like
the reference's plain nn.BatchNorm2d (model.py:17).
"""
import math
import os
import random

import torch
import torch.distributed as dist

from . import _lib


class Adam:
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) semantics (train.py:61: L2 weight decay added to the
    gradient of EVERY parameter) fused with gradient unscale + global-norm clipping, on the model's flat buffers."""

    def __init__(self, model, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=5e-4):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        flat = model.flat_params
        if not flat.is_cuda:
            raise _lib.YBError("yolov5m_b200.Adam: move the model to a CUDA device first (no CPU fallback)")
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self.step_count = 0
        self._partial = torch.zeros(4096, device=flat.device, dtype=torch.float32)
        self.grad_norm = torch.zeros(1, device=flat.device, dtype=torch.float32)
        self._step_dev = torch.zeros(1, device=flat.device, dtype=torch.int64)  # device-side counter (CUDA-graph replay)

    def zero_grad(self, set_to_none=True):
        if set_to_none:
            for p in self.model.parameters():
                p.grad = None
        else:
            self.model.flat_grads.zero_()

    def step(self, grad_scale=1.0, max_norm=0.0, grads=None, device_step=True):
        """grad_scale multiplies the stored gradient (1/loss_scale, 1/world_size); max_norm > 0 clips the global norm of
        the scaled gradient like torch.nn.utils.clip_grad_norm_ (training_utils.py:118).  A step whose gradient norm is
        inf / NaN is skipped on the device (parameters, moments, step counter untouched) like GradScaler.step
        (training_utils.py:119); ``device_step=False`` keeps the host-side counter only (no skipping bookkeeping)."""
        model = self.model
        with torch.cuda.device(model.flat_params.device):
            self._step(grad_scale, max_norm, grads, device_step)

    def _step(self, grad_scale, max_norm, grads, device_step):
        model = self.model
        L, st = _lib.lib(), _lib.stream()
        g = model.flat_grads if grads is None else grads
        p = model.flat_params
        n = p.numel()
        # the norm is always computed: it doubles as the finite check of the whole bucket
        _lib.check(L.yb_grad_norm(g.data_ptr(), n, grad_scale, self._partial.data_ptr(), self._partial.numel(),
                                  self.grad_norm.data_ptr(), st))
        self.step_count += 1
        if device_step:
            _lib.check(L.yb_counter_inc(self._step_dev.data_ptr(), self.grad_norm.data_ptr(), st))
        _lib.check(L.yb_adam_step(p.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), n, self.lr, self.betas[0],
                                  self.betas[1], self.eps, self.weight_decay, self.step_count,
                                  self._step_dev.data_ptr() if device_step else None, grad_scale, max_norm,
                                  self.grad_norm.data_ptr(), model._wfwd.data_ptr(), st))
        self._device_steps = device_step
        model._repack_derived(st)                     # dgrad-layout / stem operands from the updated masters
        if model.parity:
            model._pack_parity(st)
        model._packed_sig = model._param_signature()  # the bf16 operands are current

    def steps_taken(self):
        """optimiser steps actually applied (skipped non-finite steps excluded); synchronises when the device counter is live"""
        if getattr(self, "_device_steps", False):
            self.step_count = int(self._step_dev.item())
        return self.step_count

    def state_dict(self):
        """The ``torch.optim.Adam.state_dict()`` wire format (what the reference stores under checkpoint["optimizer"],
        train.py:140-143, and reloads with utils/utils.py:74-82): per-parameter ``step / exp_avg / exp_avg_sq`` in
        ``model.parameters()`` order plus one param group.  The moment tensors are copies in the parameters' logical
        (NCHW) layout, so the dict loads into a stock ``torch.optim.Adam`` of the reference model and vice versa."""
        model = self.model
        n = len(model._poffs)
        state = {}
        self.steps_taken()
        if self.step_count > 0:
            ms, vs = model._grad_views(self.m), model._grad_views(self.v)
            for i in range(n):
                state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": ms[i].contiguous().clone(),
                            "exp_avg_sq": vs[i].contiguous().clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "params": list(range(n))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts the torch.optim.Adam format above (e.g. the "optimizer" entry of a reference checkpoint) or the flat
        format of earlier versions of this class ({"step", "exp_avg", "exp_avg_sq"} over the flat buffer)."""
        if "param_groups" not in sd:  # flat format
            self.step_count = int(sd["step"])
            self.m.copy_(sd["exp_avg"]); self.v.copy_(sd["exp_avg_sq"])
            self._step_dev.fill_(self.step_count)
            return
        model = self.model
        n = len(model._poffs)
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != n:
            raise ValueError(f"yolov5m_b200.Adam: expected one param group with {n} parameters "
                             f"(got {len(groups)} groups / {sum(len(g['params']) for g in groups)} parameters)")
        g = groups[0]
        if g.get("amsgrad") or g.get("maximize"):
            raise ValueError("yolov5m_b200.Adam: amsgrad / maximize are not supported")
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(g["betas"]), float(g["eps"])
        self.weight_decay = float(g["weight_decay"])
        state = sd["state"]
        self.m.zero_(); self.v.zero_()
        steps = set()
        if state:
            ms, vs = model._grad_views(self.m), model._grad_views(self.v)
            for i, pid in enumerate(g["params"]):
                st = state.get(pid, state.get(str(pid)))
                if st is None:
                    continue
                ms[i].copy_(st["exp_avg"]); vs[i].copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"yolov5m_b200.Adam: parameters carry different step counts {sorted(steps)}; the fused step keeps one")
        self.step_count = steps.pop() if steps else 0
        self._step_dev.fill_(self.step_count)


def chunk_ready_after(writes, bounds):
    """writes: {op index: [(offset, length), ...]} -- the bucket slices each backward op writes, ops executed in index
    order; bounds: ascending chunk boundaries.  Returns {op index: [chunk, ...]}: chunk c = [bounds[c], bounds[c+1]) is
    complete right after the last op that writes into it (chunks nobody writes are ready after the first op)."""
    nchunks = len(bounds) - 1
    last = [min(writes) if writes else 0] * nchunks
    for i, sl in writes.items():
        for off, n in sl:
            if n <= 0:
                continue
            for c in range(nchunks):
                if off < bounds[c + 1] and off + n > bounds[c]:
                    last[c] = max(last[c], i)
    out = {}
    for c, i in enumerate(last):
        out.setdefault(i, []).append(c)
    return out


class GradSync:
    """Data-parallel gradient exchange: one all-reduce (sum) of the flat fp32 gradient bucket per step, split into a few
    large chunks so the collective is not latency-bound.  The 1/world_size average is folded into the
    optimiser's grad_scale, so the bucket is never rescaled in memory."""

    def __init__(self, model, chunks=4, group=None, mode=None):
        """mode: "p2p" (default on CUDA when world > 1, $YB_ALLREDUCE): ONE kernel of ours over NVLink peer memory
        (csrc/allreduce.cu: reduce-scatter + all-gather fused, the bucket lives in symmetric memory); "nccl": ncclAllReduce
        of the bucket in `chunks` pieces.  If the symmetric-memory rendezvous is not available the NCCL path is used and the
        reason is kept in ``self.p2p_error``."""
        self.model, self.group = model, group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.chunks = max(1, chunks)
        self._comm, self._b, self._ready, self._next = None, None, set(), -1
        self.mode = (mode or os.environ.get("YB_ALLREDUCE", "p2p")).lower()
        self._symm, self._peer_ptrs, self.p2p_error = None, None, None
        if self.world > 1 and self.mode == "p2p" and model.flat_params.is_cuda:
            self._setup_p2p()
        if self.world > 1 and self._symm is None:
            self.mode = "nccl"

    def _setup_p2p(self):
        """put the model's flat gradient bucket into symmetric memory and exchange the peer pointers (torch plumbing)"""
        import ctypes
        model = self.model
        try:
            import torch.distributed._symmetric_memory as symm_mem
            n = model.flat_params.numel()
            bucket = symm_mem.empty(n, dtype=torch.float32, device=model.flat_params.device)
            hdl = symm_mem.rendezvous(bucket, self.group if self.group is not None else dist.group.WORLD)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) != self.world or any(p == 0 for p in ptrs):
                raise RuntimeError(f"symmetric memory returned {ptrs}")
            bucket.zero_()
            model._gflat[0] = bucket          # every backward now writes its gradients straight into the peer-mapped bucket
            self._symm, self._bucket = hdl, bucket
            self._peer_ptrs = (ctypes.c_uint64 * self.world)(*ptrs)
            self.rank = dist.get_rank(self.group)
        except Exception as e:  # no symmetric memory on this box / build: NCCL carries the bucket instead
            self._symm, self.p2p_error = None, repr(e)

    def broadcast_parameters(self, src=0):
        if self.world > 1:
            dist.broadcast(self.model.flat_params, src, group=self.group)
            for b in self.model.buffers():
                dist.broadcast(b, src, group=self.group)
            if self.model.flat_params.is_cuda:
                self.model.refresh_packed(force=True)

    def _bounds(self, n):
        per = (n + self.chunks - 1) // self.chunks
        return [min(n, c * per) for c in range(self.chunks + 1)]

    prof = None  # set to a list to record (CUDA event, CUDA event) around every exchange (bench.py: grad_exchange.ms_per_step)

    def all_reduce(self, grads=None):
        if self.world == 1:
            return
        if self.prof is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self._all_reduce(grads)
            e1.record()
            self.prof.append((e0, e1))
            return
        self._all_reduce(grads)

    def _all_reduce(self, grads=None):
        g = self.model.flat_grads if grads is None else grads
        if self._symm is not None and g.data_ptr() == self._bucket.data_ptr():
            L, st = _lib.lib(), _lib.stream()
            self._symm.barrier(channel=0)   # every rank's backward has written its bucket
            _lib.check(L.yb_allreduce_p2p(self._peer_ptrs, self.rank, self.world, g.numel(), 0, st))
            self._symm.barrier(channel=1)   # every rank's stores have landed in every bucket
            return
        b = self._bounds(g.numel())
        # gradients are produced head-first (end of the bucket first): reduce from the tail
        for c in reversed(range(self.chunks)):
            if b[c + 1] > b[c]:
                dist.all_reduce(g[b[c]:b[c + 1]], op=dist.ReduceOp.SUM, group=self.group)

    # -- overlapped form (YB_OVERLAP_AR=1; off by default): a chunk is reduced on a communication stream as soon as the
    #    backward pass has issued the last kernel that writes into it, while the remaining layers' backward still runs.
    #    Measured on 2 x B200 (profiles/ab_overlap_allreduce_r1.json): identical loss, 28.49 vs 28.45 ms/step -- the
    #    collective of the 85 MB bucket costs ~0.5 ms after the backward pass and is not cheaper under it (the NCCL CTAs
    #    wait for SMs behind the persistent conv CTAs), so the simple post-backward form stays the default.
    def attach(self, engine):
        """install the chunk-ready hook on a training engine (idempotent); returns False when overlap is off"""
        if (self.world == 1 or self._symm is not None or os.environ.get("YB_OVERLAP_AR", "0") != "1"
                or not self.model.flat_params.is_cuda):
            engine.on_grad_chunks = None
            return False
        if engine.on_grad_chunks is None or engine.on_grad_chunks[1] != self._chunk_ready:
            self._b = self._bounds(self.model.flat_params.numel())
            engine.on_grad_chunks = (engine.grad_chunk_schedule(self._b), self._chunk_ready)
        if self._comm is None:
            self._comm = torch.cuda.Stream()
        self._ready, self._next = set(), self.chunks - 1
        return True

    def _chunk_ready(self, c, gflat):
        """Chunks are handed to NCCL strictly from the last one down (every rank must issue the same sequence of
        collectives, whatever order its backward pass completes them in): chunk c goes out once it and all later chunks
        are complete."""
        self._ready.add(c)
        main = torch.cuda.current_stream()
        while self._next >= 0 and self._next in self._ready:
            k = self._next
            self._next -= 1
            if self._b[k + 1] > self._b[k]:
                self._comm.wait_stream(main)  # everything issued so far, i.e. all writers of chunks >= k
                with torch.cuda.stream(self._comm):
                    dist.all_reduce(gflat[self._b[k]:self._b[k + 1]], op=dist.ReduceOp.SUM, group=self.group)

    def finish(self, gflat):
        """after backward: reduce whatever the hook did not see and make the main stream wait for the collectives"""
        for c in range(self.chunks - 1, -1, -1):
            if c not in self._ready:
                self._chunk_ready(c, gflat)
        torch.cuda.current_stream().wait_stream(self._comm)


def multi_scale_size(h, w, target_shape=640, max_stride=32, rng=random):
    """The output size the reference's multi_scale() draws (training_utils.py:11-28): a random side in
    [target_shape / 2, target_shape + max_stride) rounded down to a multiple of max_stride for the longer image side,
    the other side scaled by the same factor and rounded UP to a multiple of max_stride.  Returns (new_h, new_w)."""
    sz = rng.randrange(int(target_shape * 0.5), int(target_shape + max_stride)) // max_stride * max_stride
    sf = sz / max(h, w)
    return tuple(math.ceil(i * sf / max_stride) * max_stride for i in (h, w))


class TrainStep:
    """One optimisation step = the body of the reference train_loop (training_utils.py:97-122) for one batch.
    ``multi_scale=True`` mirrors ``multi_scale_training`` (:100): every batch is resampled to a random size; all ranks
    of a data-parallel job draw the same size (seeded ``random.Random``), so their step times stay aligned."""

    def __init__(self, model, loss_fn, optimizer, max_norm=10.0, sync=None, loss_scale=1.0, multi_scale=False,
                 target_shape=640, max_stride=32, seed=0, accumulate=1):
        """``accumulate`` = micro-batches per optimiser step: the reference accumulates gradients over
        ``max(round(64 / batch_size), 1)`` batches (training_utils.py:88-90,:116); pass that value (see
        :func:`nominal_accumulate`) for bs < 64.  ``flush()`` steps on a partial window (the reference's ``idx == nb-1``)."""
        self.model, self.loss_fn, self.opt = model, loss_fn, optimizer
        self.accumulate, self._micro = max(1, int(accumulate)), 0
        self.max_norm, self.loss_scale = max_norm, loss_scale
        self.sync = sync if sync is not None else GradSync(model)
        self.multi_scale, self.target_shape, self.max_stride = multi_scale, target_shape, max_stride
        self._rng = random.Random(seed)
        model.expose_param_grads = False  # the fused optimiser reads the flat bucket

    def __call__(self, images, targets):
        """images: (B,3,H,W) uint8 (0..255) or float32 (0..1), host or device; targets (nt,6) [img,cls,x,y,w,h]."""
        model = self.model
        dev = model.flat_params.device
        if not images.is_cuda:
            images = images.to(dev, non_blocking=True)
        size = None
        if self.multi_scale:
            size = multi_scale_size(images.shape[2], images.shape[3], self.target_shape, self.max_stride, self._rng)
        out = model(images, size=size)
        loss = self.loss_fn(out, targets, pred_size=size if size is not None else images.shape[2:4])
        last = self._micro + 1 >= self.accumulate
        # micro-batches after the first of a window are accumulated: their backward writes the second flat bucket, which is
        # then added into the first (model._NetFn.backward); the all-reduce overlap only applies to single-batch windows
        model._accumulate_grads = self._micro > 0
        overlapped = self.accumulate == 1 and self.sync.world > 1 and self.sync.attach(out[0]._yb_engine)
        if self.accumulate > 1:
            out[0]._yb_engine.on_grad_chunks = None
        try:
            if self.loss_scale != 1.0:
                (loss * self.loss_scale).backward()
            else:
                loss.backward()
        finally:
            model._accumulate_grads = False
        self._micro += 1
        if last:
            self._apply(overlapped)
        return loss

    def _apply(self, overlapped=False):
        model = self.model
        if overlapped:
            self.sync.finish(model.flat_grads)  # chunks were all-reduced while the backward pass ran
        else:
            self.sync.all_reduce()
        self.opt.step(grad_scale=1.0 / (self.loss_scale * self.sync.world), max_norm=self.max_norm)
        self.opt.zero_grad(set_to_none=True)
        self._micro = 0

    def flush(self):
        """optimiser step on a partially filled accumulation window (end of an epoch, training_utils.py:116 `idx == nb-1`)"""
        if self._micro > 0:
            self._apply(False)


def nominal_accumulate(batch_size, nbs=64):
    """micro-batches per optimiser step of the reference train_loop (training_utils.py:88-90)"""
    return max(round(nbs / batch_size), 1)
