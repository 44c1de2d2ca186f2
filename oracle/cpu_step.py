"""Oracle: the reference training step on host cores (TEST / BASELINE INFRASTRUCTURE ONLY, see oracle/__init__.py).

One step = the body of the reference train_loop (utils/training_utils.py:97-122) in fp32 on CPU PyTorch:
forward (oracle.model_ref, model.py:210-239) + ComputeLoss (oracle.loss_ref, ultralytics_loss.py:60-120) + backward +
clip_grad_norm_(10) + Adam(lr 5e-4, weight_decay 5e-4) (train.py:61).  Used by bench.py for `cpu_baseline` and for the
`--impl reference` arm (kind "port": the Python reference itself cannot travel to the GPU box).
"""
import os
import time

import torch

from . import loss_ref, model_ref


def synthetic_batch(seed, bs, size=640):
    """SURVEY.md 8(d) recipe: images U[0,1), nt = 8 boxes per image."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(bs, 3, size, size, generator=g)
    nt = 8 * bs
    t = torch.cat([torch.randint(0, bs, (nt, 1), generator=g).float(), torch.randint(0, 80, (nt, 1), generator=g).float(),
                   torch.rand(nt, 2, generator=g), torch.rand(nt, 2, generator=g) * 0.5 + 0.005], 1)
    return x, t


class CpuTrainer:
    def __init__(self, threads=None, seed=0):
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.sd = model_ref.make_state_dict(seed)
        self.params = []
        for name, _, kind in model_ref.param_specs():
            if kind in ("conv", "bn_w", "bn_b", "head_w", "head_b"):
                self.sd[name] = self.sd[name].clone().requires_grad_(True)
                self.params.append(self.sd[name])
        self.opt = torch.optim.Adam(self.params, lr=5e-4, weight_decay=5e-4)

    def step(self, x, targets):
        p = model_ref.forward(self.sd, x, train=True)
        loss = loss_ref.compute_loss(p, targets, self.sd["head.anchors"])
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.params, max_norm=10.0)
        self.opt.step()
        self.opt.zero_grad(set_to_none=True)
        return float(loss.detach())

    def time_steps(self, bs, steps, warmup, size=640, seed=1):
        x, t = synthetic_batch(seed, bs, size)
        for _ in range(warmup):
            self.step(x, t)
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step(x, t)
        dt = time.perf_counter() - t0
        return bs * steps / dt, dt / steps
