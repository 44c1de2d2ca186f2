"""BASELINE.json configs[1] ("single-GPU forward+ComputeLoss on synthetic bs=16 640x640 with random targets, fp32 vs
reference"), SURVEY.md 8(d) config 2 recipe: model.train(), generator seed 1, x = rand(16,3,640,640), nt = 128 targets.

The checker is the CPU oracle (oracle.model_ref / oracle.loss_ref, themselves pinned to the real reference by
tests/test_oracle_golden.py) evaluated in FLOAT64 on the host cores (~30 s): at this size fp32 is not a fixed point -- the
fp32 evaluation of the same oracle is itself 1.07e-3 (rel. L2 of all parameter gradients; 1.2e-3 median per tensor, outputs
1.8e-5 / 3.0e-5 / 5.2e-5) away from the float64 one (measured in the build container), i.e. AT north_star's 1e-3, so fp32
cannot referee a 1e-3 claim; float64 can.  Compared: the three head tensors, the loss and its three parts, build_targets
(bit-exact indices / classes), and the parameter gradients -- in fp32 parity mode against north_star's 1e-3, and in the
production bf16 mode with the MEASURED error reported beside it (written to gpurun_out/config2_parity.json for DESIGN.md).

Measured on B200 (round 2): parity mode vs the fp32 oracle: outputs 2.7e-5 / 4.7e-5 / 8.3e-5, loss 0, gradient norm 2.2e-5,
gradient rel. L2 1.10e-3 -- the last figure is the fp32 ORACLE's own distance to float64 (1.07e-3), hence this file's switch
to the float64 referee.  Production bf16 mode: loss 1.9e-4, loss parts <= 2.9e-3, gradient norm 4.0e-2; outputs 0.21 / 0.34
/ 0.48 and gradient direction 0.67 rel. L2: a random-initialised batch-norm network amplifies perturbations layer after
layer (every layer re-normalises; the known gradient / perturbation explosion of BN networks at initialisation), so ANY
reduced-precision storage -- bf16 here, fp16 autocast in the reference's own train_loop -- decorrelates the raw logits
while the loss, its parts and the gradient scale stay put.  Measured to make sure this is sensitivity and not a defect: the
CPU oracle with emulated bf16 storage (model_ref quant=True) is 0.21 / 0.34 / 0.49 away from its own fp32 evaluation at this
config, and the GPU is 0.15 / 0.25 / 0.38 away from that emulation -- two bf16 evaluations with different rounding details
are as far from each other as each is from fp32.  The per-layer (teacher-forced) comparison in test_model_gpu.py is what
pins the production kernels; this file pins the engine end to end through the parity mode.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import loss_ref, model_ref
from test_model_gpu import make_model, rel

gpu = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-3


def config2_inputs(bs=16, size=640, nt=128):
    g = torch.Generator().manual_seed(1)
    x = torch.rand(bs, 3, size, size, generator=g)
    t = torch.cat([torch.randint(0, bs, (nt, 1), generator=g).float(), torch.randint(0, 80, (nt, 1), generator=g).float(),
                   torch.rand(nt, 2, generator=g), torch.rand(nt, 2, generator=g) * 0.5 + 0.005], 1)
    return x, t


@pytest.fixture(scope="module")
def oracle_run():
    """float64 evaluation of the whole config: outputs, loss parts, targets, parameter gradients (CPU, all host threads)"""
    torch.set_num_threads(os.cpu_count() or 1)
    x, t = config2_inputs()
    sd = model_ref.make_state_dict(0)
    leaves = {}
    for name, _, kind in model_ref.param_specs():
        if sd[name].is_floating_point():
            sd[name] = sd[name].double()
        if kind in ("conv", "bn_w", "bn_b", "head_w", "head_b"):
            sd[name] = sd[name].clone().requires_grad_(True)
            leaves[name] = sd[name]
    p = model_ref.forward(sd, x.double(), train=True, update_stats=False)
    anchors = model_ref.head_anchors()
    loss, parts, tg = loss_ref.compute_loss(p, t, anchors, return_parts=True)
    loss.backward()
    return {"x": x, "t": t, "p": [q.detach() for q in p], "loss": float(loss.detach()), "parts": [float(v.detach()) for v in parts],
            "grads": {n: v.grad.detach() for n, v in leaves.items()}, "targets": tg}


def _run_gpu(parity, ref):
    import yolov5m_b200 as yb
    m, _ = make_model()
    m.parity = parity
    m.train()
    loss_fn = yb.ComputeLoss(m)
    out = m(ref["x"].cuda())
    loss = loss_fn(out, ref["t"], None)
    parts = [float(v) for v in loss_fn.last_parts.tolist()]
    tg = loss_fn.build_targets(out, ref["t"])
    loss.backward()
    prm = dict(m.named_parameters())
    res = {"out_rel": [rel(out[i], ref["p"][i]) for i in range(3)],
           "loss_rel": abs(loss.item() - ref["loss"]) / abs(ref["loss"]),
           "parts_rel": [abs(a - b) / max(abs(b), 1e-12) for a, b in zip(parts, ref["parts"])]}
    num = den = 0.0
    per = {}
    gmax = max(float(g.norm()) for g in ref["grads"].values())
    for n, g in ref["grads"].items():
        d = (prm[n].grad.detach().cpu().double() - g.double())
        num += float((d * d).sum()); den += float((g.double() ** 2).sum())
        if float(g.norm()) > 1e-3 * gmax:
            per[n] = float(d.norm() / g.double().norm())
    ours_norm = np.sqrt(sum(float(prm[n].grad.double().norm()) ** 2 for n in ref["grads"]))
    res["grad_norm_rel"] = abs(ours_norm - np.sqrt(den)) / np.sqrt(den)
    res["grad_rel_l2"] = float(np.sqrt(num / den))
    res["grad_per_tensor_median"] = float(np.median(list(per.values())))
    res["grad_per_tensor_max"] = max(per.items(), key=lambda kv: kv[1])
    return res, tg


def _check_targets(tg, ref_tg):
    tcls, tbox, indices, anch = tg
    for i in range(3):
        r = ref_tg[i]
        for a, key in zip(indices[i], ("b", "a", "gj", "gi")):
            assert np.array_equal(a.cpu().numpy(), r[key].astype(np.int64)), f"level {i}: target index {key} differs"
        assert np.array_equal(tcls[i].cpu().numpy(), r["tcls"].astype(np.int64))
        assert np.allclose(tbox[i].cpu().numpy(), r["tbox"], rtol=0, atol=1e-6)
        assert np.allclose(anch[i].cpu().numpy(), r["anch"], rtol=0, atol=1e-6)


def _record(tag, res):
    path = os.path.join(ROOT, "gpurun_out", "config2_parity.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        d = json.load(open(path)) if os.path.exists(path) else {}
        d[tag] = res
        json.dump(d, open(path, "w"), indent=1, default=str)
    except OSError:
        pass


@gpu
def test_config2_fp32_parity_mode(oracle_run):
    res, tg = _run_gpu(True, oracle_run)
    print("\nconfig 2, fp32 parity mode vs float64 oracle:", json.dumps(res, default=str))
    _record("parity_fp32", res)
    _check_targets(tg, oracle_run["targets"])
    assert max(res["out_rel"]) < TOL, res
    assert res["loss_rel"] < TOL and max(res["parts_rel"]) < TOL, res
    assert res["grad_norm_rel"] < TOL and res["grad_rel_l2"] < TOL, res
    assert res["grad_per_tensor_max"][1] < 5 * TOL, res


@gpu
def test_config2_production_bf16_measured(oracle_run):
    """production mode (bf16 activation storage, bf16 tensor-core operands): the measured distance to the fp32 reference at
    this well-conditioned size; the bounds below are that measurement with headroom, not north_star's fp32 figure"""
    res, tg = _run_gpu(False, oracle_run)
    print("\nconfig 2, production bf16 mode vs float64 oracle:", json.dumps(res, default=str))
    _record("production_bf16", res)
    _check_targets(tg, oracle_run["targets"])   # integer work is independent of the network's precision: still bit-exact
    # the quantities that are not chaotic at random initialisation (see the module docstring)
    assert res["loss_rel"] < 1e-3, res
    assert max(res["parts_rel"]) < 1e-2, res
    assert res["grad_norm_rel"] < 1e-1, res
