"""Worker of tests/test_dp_gpu.py: launched with torch.distributed.run on >= 2 GPUs of one box.

Checks, on every rank:
  1. the fused NVLink peer-memory all-reduce (csrc/allreduce.cu via trainer.GradSync mode "p2p") turns every rank's bucket
     into sum_r g_r, BIT-identical to the sum (in rank order) of the per-rank single-GPU gradients gathered over NCCL;
  2. the NCCL mode gives the same bucket;
  3. after 3 TrainStep steps (different batch per rank) the flat parameters are bit-identical on every rank and differ from
     the initial ones; the data-parallel gradient fed to the optimiser is the MEAN over ranks (grad_scale = 1/world).
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import yolov5m_b200 as yb
    from oracle import model_ref
    from yolov5m_b200.trainer import Adam, GradSync, TrainStep
    import recipes

    def make():
        m = yb.YOLOV5m(first_out=48, nc=80, anchors=yb.ANCHORS, ch=(192, 384, 768))
        m.load_state_dict({k: v.clone() for k, v in model_ref.make_state_dict(0).items()})
        return m.to(dev).train()

    res = {"rank": rank, "world": world}
    x = (recipes.model_input(100 + rank, 2, 128, 128) * 255).to(torch.uint8).to(dev)
    tg = recipes.targets(200 + rank, 2, 12)
    for mode in ("p2p", "nccl"):
        m = make()
        sync = GradSync(m, mode=mode)
        res[f"{mode}_mode_used"] = sync.mode
        res[f"{mode}_error"] = sync.p2p_error
        m.expose_param_grads = False
        loss_fn = yb.ComputeLoss(m)
        loss_fn(m(x), tg, None).backward()
        torch.cuda.synchronize()
        mine = m.flat_grads.clone()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        expect = parts[0].clone()
        for r in range(1, world):
            expect += parts[r]
        sync.all_reduce()
        torch.cuda.synchronize()
        res[f"{mode}_sum_bit_exact"] = bool(torch.equal(m.flat_grads, expect))
        res[f"{mode}_max_abs_diff"] = float((m.flat_grads - expect).abs().max())
        res[f"{mode}_grad_norm"] = float(expect.norm())
    # 3 optimisation steps, different data per rank, p2p exchange
    m = make()
    sync = GradSync(m)
    sync.broadcast_parameters(0)
    p0 = m.flat_params.clone()
    step = TrainStep(m, yb.ComputeLoss(m), Adam(m), max_norm=10.0, sync=sync)
    losses = [float(step(x, tg).detach()) for _ in range(3)]
    torch.cuda.synchronize()
    allp = [torch.empty_like(p0) for _ in range(world)]
    dist.all_gather(allp, m.flat_params.contiguous())
    res["params_identical_across_ranks"] = bool(all(torch.equal(allp[0], q) for q in allp[1:]))
    res["params_changed"] = bool((m.flat_params - p0).abs().max() > 0)
    res["losses"] = losses
    res["train_mode_used"] = sync.mode
    out = os.environ.get("YB_DP_OUT")
    if out:
        with open(f"{out}.rank{rank}.json", "w") as f:
            json.dump(res, f)
    print("DP_RESULT", json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
