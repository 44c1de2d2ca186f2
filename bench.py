#!/usr/bin/env python
"""bench.py -- training-step images/sec of the B200-native YOLOV5m hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2] / [3]): full train step, bf16 tensor-core compute with fp32 master weights, bs=64 per
GPU, 640x640, synthetic COCO-80-like targets (8 boxes / image), random-init weights: forward + ComputeLoss + backward +
(NCCL all-reduce of the flat gradient bucket when N > 1) + global-norm clip (10) + Adam(5e-4, wd 5e-4).

One JSON line on rank 0.  `value` = whole-job img/s with the batch resident in HBM; `e2e` = the same step through the
public API (yolov5m_b200.trainer.TrainStep) with the uint8 batch in pinned host memory copied to the device every step
and the loss read back every step; `roofline` = tensor-core roofline of the dominant kernel (conv_igemm_kernel: all
forward + dgrad convolutions) from CUDA events around each of its launches; `cpu_baseline` = the oracle port of the
reference step on the host cores (bounded sample).  `--impl reference` times that CPU path alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_step_images_per_sec_640x640_bs64_per_gpu"
UNIT = "img/s"
TRAIN_GFLOP_PER_IMG = 146.62  # 3 x 48.872 GFLOP (fwd + dgrad + wgrad of the 82 convs at 640x640)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "detect"],
                    help="train = BASELINE configs[2]/[3] (the headline metric); detect = configs[4]: bs=128 1280x1280 eval "
                         "forward + cells_to_bboxes + non_max_suppression(conf .25, iou .45), detect.py:50-54")
    ap.add_argument("--cand-frac", type=float, default=0.02,
                    help="detect: fraction of cells whose objectness passes conf 0.25 (D1 'realistic' of SURVEY.md 8d; 1.0 = D2)")
    ap.add_argument("--bs", type=int, default=None, help="images per GPU (train: 64, detect: 128 -- the metrics' configs)")
    ap.add_argument("--size", type=int, default=None, help="image side (train: 640, detect: 1280)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    if a.bs is None:
        a.bs = 64 if a.workload == "train" else 128
    if a.size is None:
        a.size = 640 if a.workload == "train" else 1280
    return a


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1400.0, 6650.0, "fallback"


def conv_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the conv kernels, from the committed ncu launch summary
    (profiles/launch_summary_*.json, written by tools/summarize_launches.py from an ncu capture of this same command)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "launch_summary_*.json")))
    if not files:
        return None
    try:
        with open(files[-1]) as f:
            d = json.load(f)
        ks = [k for k in d["kernels"] if "conv_igemm_kernel" in k["kernel"] or "conv_patch_kernel" in k["kernel"]]
        n = sum(k["launches_per_step"] for k in ks)
        b = sum(k["dram_read_GB_per_step"] + k["dram_write_GB_per_step"] for k in ks) * 1e9
        return b / n if n else None
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------- clocks sampler
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.th = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    def summary(self, windows):
        sm, mx, reasons, pw = [], 0.0, set(), 0.0
        for t, c in self.rows:
            if not any(a <= t <= b for a, b in windows) or len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx = max(mx, float(c[2])); pw = max(pw, float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "power_w_max": pw or None, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_trainer():
    """the reference's own train step on the host cores: the UNMODIFIED reference (baseline/_ref, vendored by
    __graft_entry__.build()) when it is there -- kind "reference" -- else the oracle port -- kind "port"."""
    try:
        from baseline.ref_step import RefTrainer
        return RefTrainer(), "reference", "the unmodified reference (baseline/_ref: model.py, ultralytics_loss.py, torch.optim.Adam)"
    except Exception as e:  # reference not vendored on this box
        sys.stderr.write(f"bench.py: reference not importable ({e!r}); timing the oracle port instead\n")
        from oracle.cpu_step import CpuTrainer
        return CpuTrainer(), "port", "oracle port of the reference step"


def cpu_reference(steps, warmup, budget_s, size):
    """the reference train step on the host cores; batch sized so the run fits the time budget."""
    tr, kind, what = cpu_trainer()
    _, t1 = tr.time_steps(1, 1, 1, size=size)                 # probe: seconds per image-step at bs=1
    bs = 8
    while bs > 1 and (steps + warmup) * t1 * bs * 0.8 > budget_s:
        bs //= 2
    ips, sec = tr.time_steps(bs, steps, warmup, size=size)
    return ips, sec, bs, tr.threads, kind, what


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ips, sec, bs, threads, kind, what = cpu_reference(a.steps, a.warmup, budget_s=150.0, size=a.size)
    sample = f"{a.steps} timed + {a.warmup} warm-up full train steps at bs={bs}, {a.size}x{a.size}, fp32, {what}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "full train step (fwd+ComputeLoss+bwd+clip+Adam), CPU host cores", "batch_per_step": bs,
                   "image": a.size},
        "cpu_baseline": {"value": ips, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))



# ---------------------------------------------------------------------------------------------------- detect workload
DETECT_METRIC = "detect_images_per_sec_1280x1280_bs128"
DETECT_GFLOP_PER_IMG_640 = 48.872  # forward convs at 640x640; scales with the pixel count


def calibrate_detector(model, cal_images, first_batch, cand_frac, conf=0.25):
    """D1 of SURVEY.md 8(d), made reproducible for a random-init network (both arms use exactly this recipe):
      1. ONE train-mode forward on `cal_images` with BatchNorm momentum 1.0: the running statistics become that batch's
         statistics, so the eval-mode activations are normalised (a freshly constructed network with running mean 0 / var 1
         lets the signal die out: every logit equals its bias and either all or none of the cells pass);
      2. the objectness bias of the three heads is shifted so that `cand_frac` of the cells of `first_batch` pass `conf`.
    Works on the reference nn.Module and on the drop-in alike (same attribute surface)."""
    import math
    import torch
    bns = [m for m in model.modules() if hasattr(m, "running_mean") and hasattr(m, "momentum")]
    old = [m.momentum for m in bns]
    for m in bns:
        m.momentum = 1.0
    model.train()
    with torch.no_grad():
        model(cal_images)
    for m, o in zip(bns, old):
        m.momentum = o
    model.eval()
    if cand_frac >= 1.0:
        return 1.0
    with torch.no_grad():
        out = model(first_batch)
        obj = torch.cat([o[..., 4].reshape(-1).float() for o in out])
        k = max(1, int(round(obj.numel() * (1.0 - cand_frac))))
        q = torch.kthvalue(obj.cpu(), k).values.item()
        shift = math.log(conf / (1.0 - conf)) - q
        for conv in model.head.out_convs:
            conv.bias.data.view(3, -1)[:, 4] += shift
        out = model(first_batch)
        frac = float(torch.cat([(o[..., 4].reshape(-1).float() > math.log(conf / (1.0 - conf))).float() for o in out]).mean())
    return frac


def detect_inputs(bs, size, seed=3):
    import torch
    g = torch.Generator().manual_seed(seed)
    cal = torch.randint(0, 256, (2, 3, 640, 640), dtype=torch.uint8, generator=g)
    x = torch.randint(0, 256, (bs, 3, size, size), dtype=torch.uint8, generator=g)
    return cal, x


def cpu_detect_sample(a, steps, warm):
    """bounded sample of the reference's own detect path (detect.py:50-54: model -> cells_to_bboxes -> non_max_suppression)
    on the host cores: (img/s, seconds per pass, cpu_baseline object, candidate fraction, kept per image)"""
    import torch
    from baseline.ref_step import RefDetector
    det = RefDetector()
    bs = 2
    cal, x = detect_inputs(bs, a.size)
    frac = calibrate_detector(det.model, cal.float() / 255, x.float() / 255, a.cand_frac)
    xf = x.float() / 255
    for _ in range(warm):
        det.detect(xf)
    t0 = time.perf_counter()
    for _ in range(steps):
        kept = det.detect(x.float() / 255)   # detect.py:47: img.float() / 255 is part of the path
    dt = (time.perf_counter() - t0) / steps
    ips = bs / dt
    sample = (f"{steps} timed + {warm} warm-up detect passes at bs={bs}, {a.size}x{a.size}, fp32, the unmodified reference "
              f"(model.py, utils/plot_utils.py cells_to_bboxes, utils/bboxes_utils.py non_max_suppression)")
    cpu = {"value": ips, "unit": UNIT, "cores": det.threads, "kind": "reference", "sample": sample}
    return ips, dt, cpu, frac, sum(len(k) for k in kept) / bs, bs


def run_reference_detect(a):
    """`--impl reference --workload detect`: the reference's own detect path on the host cores"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(a.steps, 3)), min(a.warmup, 1)
    try:
        ips, dt, cpu, frac, kept, bs = cpu_detect_sample(a, steps, warm)
    except Exception as e:
        print(json.dumps({"impl": "reference", "unavailable": f"reference not importable for the detect workload: {e!r}"}))
        return
    print(json.dumps({
        "impl": "reference", "metric": DETECT_METRIC, "value": ips, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "configs[4]: detect (eval fwd + cells_to_bboxes + NMS conf .25 iou .45), CPU host cores",
                   "batch_per_step": bs, "image": a.size, "candidate_fraction": frac, "kept_per_image": kept},
        "cpu_baseline": cpu,
        "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours_detect(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:  # replicas only: no collective on the detect path; the group is used for barrier / max
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    from yolov5m_b200.build import build
    build()
    import yolov5m_b200 as yb
    from yolov5m_b200 import _lib
    from yolov5m_b200.boxes import nms_device
    L = _lib.lib()
    torch.manual_seed(0)
    model = yb.YOLOV5m(first_out=yb.FIRST_OUT, nc=80, anchors=yb.ANCHORS, ch=(192, 384, 768)).to(dev)
    B, S = a.bs, a.size
    cal, x_host = detect_inputs(B, S, seed=3 + rank)
    x_host = x_host.pin_memory()
    x_dev = x_host.to(dev)
    frac = calibrate_detector(model, cal.to(dev), x_dev, a.cand_frac)
    model.eval()
    CONF, IOU, MAXDET = 0.25, 0.45, 300

    def detect(x):  # detect.py:50-54 on device tensors
        with torch.no_grad():
            out = model(x)
            dec = yb.cells_to_bboxes(out, model.head.anchors, model.head.stride, is_pred=True, to_list=False)
            return nms_device(dec, IOU, CONF, MAXDET), out, dec

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    windows = []
    for _ in range(a.warmup):
        detect(x_dev)
    barrier()
    l0 = L.yb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        (rows, counts), out, dec = detect(x_dev)
    e1.record()
    barrier()
    windows.append((w0, time.perf_counter()))
    launches = L.yb_launch_count() - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * a.steps / (ms * 1e-3)

    # per-stage device times (CUDA events on the launching stream) -> rooflines of the stages
    def stage(fn, n=3):
        for _ in range(3):  # warm-up: the caching allocator settles (a forward returns 4.4 GB of fresh logits)
            r = fn()
        del r
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n):
            r = fn()
        s1.record()
        torch.cuda.synchronize()
        return s0.elapsed_time(s1) / n * 1e-3, r

    with torch.no_grad():
        t_f, out = stage(lambda: model(x_dev))
        t_d, dec = stage(lambda: yb.cells_to_bboxes(out, model.head.anchors, model.head.stride, is_pred=True, to_list=False))
        t_n, (rows, counts) = stage(lambda: nms_device(dec, IOU, CONF, MAXDET))
    peak_tf, peak_hbm, peak_src = peaks()
    # the forward's floor: per conv launch max(flops / tensor peak, algorithmic bytes / HBM peak), one instrumented pass
    eng = model.engine(B, S, S, False)
    eng.prof = []
    with torch.no_grad():
        model(x_dev)
    torch.cuda.synchronize()
    fwd_floor = sum(max(f / (peak_tf * 1e12), nb / (peak_hbm * 1e9)) for _, f, _, _, nb in eng.prof)
    fwd_conv_t = sum(s0.elapsed_time(s1) for _, _, s0, s1, _ in eng.prof) * 1e-3
    eng.prof = None
    cells = sum(o.numel() // o.shape[-1] for o in out)
    dec_bytes = cells * (out[0].shape[-1] * 4 + 6 * 4)            # every logit read once, 6 floats written per cell
    fwd_flop = DETECT_GFLOP_PER_IMG_640 * 1e9 * (S * S) / (640.0 * 640.0) * B
    cand = int((dec[..., 1] > CONF).sum().item())

    # end to end through the public API: pinned uint8 batch -> H2D -> forward -> decode -> NMS -> rows + counts D2H
    e2e = None
    if not a.no_e2e:
        # double-buffered input: batch i+1 crosses PCIe on a copy stream while batch i is computed (every batch is copied
        # inside the timed region; the caller still consumes the detections of every batch before the next one starts)
        stage_x = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        copy_stream = torch.cuda.Stream()
        rows_h = torch.empty(B, MAXDET, 6, dtype=torch.float32).pin_memory()
        cnt_h = torch.empty(B, dtype=torch.int32).pin_memory()

        def h2d(k):
            with torch.cuda.stream(copy_stream):
                stage_x[k].copy_(x_host, non_blocking=True)
                ready[k].record(copy_stream)

        def e2e_loop(n):
            main = torch.cuda.current_stream()
            h2d(0)
            for i in range(n):
                k = i & 1
                main.wait_event(ready[k])
                if i + 1 < n:
                    h2d(1 - k)  # its last reader (batch i-1) has completed: the loop synchronises every batch
                (r, c), _, _ = detect(stage_x[k])
                rows_h.copy_(r, non_blocking=True)
                cnt_h.copy_(c, non_blocking=True)
                main.synchronize()  # the caller consumes the detections of every batch
            copy_stream.synchronize()
        e2e_loop(1)
        barrier()
        w0 = time.perf_counter()
        e0.record()
        e2e_loop(a.steps)
        e1.record()
        barrier()
        windows.append((w0, time.perf_counter()))
        ms_e = max_over_ranks(e0.elapsed_time(e1))
        e2e = {"value": world * B * a.steps / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel()) * world,
               "d2h_bytes_per_step": int(rows_h.numel() * 4 + cnt_h.numel() * 4) * world, "ms_per_step": ms_e / a.steps}

    cpu = None
    if rank == 0:
        clocks.stop()
        if world == 1 and not a.no_cpu_baseline:
            try:
                cpu = cpu_detect_sample(a, 2, 1)[2]
            except Exception as e:  # baseline/_ref not vendored on this box
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e!r}"}
        # keep sets of a few images against the oracle on the SAME decoded tensor (bit-exact contract)
        from oracle import nms_ref
        nchk = min(2, B)
        ref, _ = nms_ref.non_max_suppression(dec[:nchk].cpu(), IOU, CONF, MAXDET)
        import numpy as np
        exact = all(int(counts[i]) == len(ref[i]) and
                    np.array_equal(rows[i, : len(ref[i])].cpu().numpy(), ref[i].astype(np.float32)) for i in range(nchk))
        out_line = {
            "metric": DETECT_METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "configs[4]: detect.py path, bs=128 1280x1280 eval forward + cells_to_bboxes + "
                                   "non_max_suppression(conf .25, iou .45, max_det 300); replicas only for N > 1",
                       "batch_per_gpu": B, "image": S, "candidate_fraction": frac, "candidates_per_image": cand / B,
                       "kept_per_image": float(counts.float().mean().item()), "keep_sets_bit_exact_vs_oracle": bool(exact),
                       "l2_policy": "inputs larger than L2 (629 MB of images, 4.4 GB of logits per step)"},
            "roofline": {"bound": "hbm", "kernel": "decode_pred_kernel (cells_to_bboxes, 3 launches / step)",
                         "achieved": dec_bytes / t_d / 1e9, "peak": peak_hbm, "peak_source": peak_src + " hbm_gbs",
                         "unit": "GB/s", "frac": dec_bytes / t_d / 1e9 / peak_hbm, "traffic": None,
                         "algorithmic_bytes_per_step": dec_bytes, "ms_per_step": t_d * 1e3},
            "stages": {"forward": {"ms": t_f * 1e3, "tflops": fwd_flop / t_f / 1e12, "frac_of_bf16_peak": fwd_flop / t_f / 1e12 / peak_tf,
                                   "conv_launch_ms": fwd_conv_t * 1e3, "conv_floor_ms": fwd_floor * 1e3,
                                   "conv_frac_of_floor": fwd_floor / fwd_conv_t},
                       "decode": {"ms": t_d * 1e3, "GBps": dec_bytes / t_d / 1e9},
                       "nms": {"ms": t_n * 1e3, "candidates_per_image": cand / B}},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary(windows),
        }
        print(json.dumps(out_line))
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------- our arm (GPU)
def run_ours(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    from yolov5m_b200.build import build
    build()
    import yolov5m_b200 as yb
    from yolov5m_b200 import _lib
    from yolov5m_b200.trainer import Adam, GradSync, TrainStep
    L = _lib.lib()

    torch.manual_seed(0)
    model = yb.YOLOV5m(first_out=yb.FIRST_OUT, nc=80, anchors=yb.ANCHORS, ch=(192, 384, 768)).to(dev).train()
    loss_fn = yb.ComputeLoss(model)
    opt = Adam(model, lr=5e-4, weight_decay=5e-4)
    sync = GradSync(model)
    sync.broadcast_parameters(0)
    step = TrainStep(model, loss_fn, opt, max_norm=10.0, sync=sync)

    B, S = a.bs, a.size
    g = torch.Generator().manual_seed(1 + rank)
    nt = 8 * B
    nbuf = 2
    host_imgs = [torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).pin_memory() for _ in range(nbuf)]
    host_tgts = [torch.cat([torch.randint(0, B, (nt, 1), generator=g).float(), torch.randint(0, 80, (nt, 1), generator=g).float(),
                            torch.rand(nt, 2, generator=g), torch.rand(nt, 2, generator=g) * 0.5 + 0.005], 1).pin_memory()
                 for _ in range(nbuf)]
    dev_imgs = [h.to(dev) for h in host_imgs]
    dev_tgts = [h.to(dev) for h in host_tgts]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    windows = []

    # ---- (1) device-resident: inputs already in HBM
    for i in range(a.warmup):
        step(dev_imgs[i % nbuf], dev_tgts[i % nbuf])
    barrier()
    l0 = L.yb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for i in range(a.steps):
        loss = step(dev_imgs[i % nbuf], dev_tgts[i % nbuf])
    e1.record()
    barrier()
    windows.append((w0, time.perf_counter()))
    launches = L.yb_launch_count() - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * B * a.steps / (ms * 1e-3)
    last_loss = float(loss.item())

    # ---- (2) end to end through the public API: pinned host batch -> H2D -> step -> loss D2H, every step;
    #      the copy of batch i+1 overlaps the compute of batch i on a second stream (double buffering)
    e2e = None
    if not a.no_e2e:
        copy_stream = torch.cuda.Stream()
        stage_i = [torch.empty_like(d) for d in dev_imgs]
        stage_t = [torch.empty_like(d) for d in dev_tgts]
        ready = [torch.cuda.Event() for _ in range(nbuf)]
        consumed = [torch.cuda.Event() for _ in range(nbuf)]

        def prefetch(i):
            k = i % nbuf
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[k])
                stage_i[k].copy_(host_imgs[k], non_blocking=True)
                stage_t[k].copy_(host_tgts[k], non_blocking=True)
                ready[k].record(copy_stream)

        loss_h = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(nbuf)]
        loss_ev = [torch.cuda.Event() for _ in range(nbuf)]

        def e2e_loop(n):
            """every step: H2D copy of its batch (pinned host memory) and a D2H read of its loss.  The read is pipelined
            like a training loop's logging: step i's loss is copied to pinned memory asynchronously and CONSUMED on the host
            while step i+1 is already queued, so the host never drains the GPU between steps; every loss of the window is read
            inside the timed region (the last one before the closing barrier)."""
            for k in range(nbuf):
                consumed[k].record()
            prefetch(0)
            tot = 0.0
            for i in range(n):
                k = i % nbuf
                torch.cuda.current_stream().wait_event(ready[k])
                ls = step(stage_i[k], stage_t[k])
                consumed[k].record()
                loss_h[k].copy_(ls.detach(), non_blocking=True)  # device -> host read of the step's result, every step
                loss_ev[k].record()
                if i + 1 < n:
                    prefetch(i + 1)  # queued after this step's launches: the copy runs under step i, off its critical path
                if i > 0:            # consume the previous step's loss (its copy finished long ago)
                    loss_ev[(i - 1) % nbuf].synchronize()
                    tot += float(loss_h[(i - 1) % nbuf][0])
            loss_ev[(n - 1) % nbuf].synchronize()
            tot += float(loss_h[(n - 1) % nbuf][0])
            return tot

        e2e_loop(max(2, a.warmup // 2))
        barrier()
        w0 = time.perf_counter()
        e0.record()
        e2e_loop(a.steps)
        e1.record()
        barrier()
        windows.append((w0, time.perf_counter()))
        ms_e = max_over_ranks(e0.elapsed_time(e1))
        e2e = {"value": world * B * a.steps / (ms_e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(host_imgs[0].numel() + host_tgts[0].numel() * 4) * world,
               "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e / a.steps}

    # ---- (3) roofline of the dominant kernel: CUDA events around every conv launch (separate instrumented steps)
    #      every rank runs these steps (they contain the gradient all-reduce); only rank 0 records events
    roof, kern, roof_hbm = None, None, None
    eng = model.engine(B, S, S, True)
    nprof = min(3, a.steps)
    if rank == 0:
        eng.prof = []
        sync.prof = []
    for i in range(nprof):
        step(dev_imgs[i % nbuf], dev_tgts[i % nbuf])
    torch.cuda.synchronize()
    exch = {"mode": sync.mode if world > 1 else "none", "p2p_fallback_reason": sync.p2p_error}
    if rank == 0:
        if sync.prof:
            # includes the wait for the slowest rank (the exchange starts with a cross-rank barrier)
            exch["ms_per_step_incl_straggler_wait"] = sum(a.elapsed_time(b) for a, b in sync.prof) / len(sync.prof)
        sync.prof = None
        acc = {}
        peak_tf, peak_hbm, peak_src = peaks()
        for kind, flops, s0, s1, nbytes in eng.prof:
            d = acc.setdefault(kind, [0.0, 0.0, 0, 0.0])
            d[0] += flops; d[1] += s0.elapsed_time(s1) * 1e-3; d[2] += 1
            # the launch's own floor: whichever of the two rooflines binds it (bn_* records carry bytes in both fields)
            d[3] += nbytes / (peak_hbm * 1e9) if kind.startswith("bn_") else max(flops / (peak_tf * 1e12), nbytes / (peak_hbm * 1e9))
        eng.prof = None
        igemm_f = acc["fwd"][0] + acc["dgrad"][0]
        igemm_t = acc["fwd"][1] + acc["dgrad"][1]
        igemm_n = acc["fwd"][2] + acc["dgrad"][2]
        ach = igemm_f / igemm_t / 1e12
        roof = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM conv (conv_igemm_kernel + conv_patch_kernel: all fwd + dgrad launches)",
                "achieved": ach, "peak": peak_tf, "peak_source": peak_src + " bf16_tflops_sustained", "unit": "TFLOP/s",
                "frac": ach / peak_tf, "traffic": conv_traffic_per_launch(),
                "traffic_unit": "bytes/launch (dram read+write); PROFILE CONSTANT from the committed ncu launch list "
                                "profiles/launch_summary_*.json of this same command, not re-measured at bench time",
                "launches_per_step": igemm_n // nprof, "avg_launch_us": igemm_t / igemm_n * 1e6,
                "flop_per_launch": igemm_f / igemm_n, "ms_per_step": igemm_t / nprof * 1e3,
                # many of these launches are HBM-bound (1x1 convs on 160x160 / 80x80 maps): the per-launch floor
                # max(flops / tensor peak, algorithmic bytes / HBM peak), summed, is what the kernels can be held against
                "floor_ms_per_step": (acc["fwd"][3] + acc["dgrad"][3]) / nprof * 1e3,
                "frac_of_floor": (acc["fwd"][3] + acc["dgrad"][3]) / igemm_t}
        kern = {k: {"tflops": v[0] / v[1] / 1e12, "ms_per_step": v[1] / nprof * 1e3, "launches_per_step": v[2] // nprof,
                    "floor_ms_per_step": v[3] / nprof * 1e3, "frac_of_floor": v[3] / v[1]}
                for k, v in acc.items() if not k.startswith("bn_")}
        # second roofline entry: the HBM-bound share of the step (BN/SiLU passes), algorithmic bytes / CUDA-event time
        ew = [acc[k] for k in ("bn_fwd", "bn_bwd_reduce", "bn_bwd_apply") if k in acc]
        if ew:
            eb, et, en = sum(v[0] for v in ew), sum(v[1] for v in ew), sum(v[2] for v in ew)
            roof_hbm = {"bound": "hbm", "kernel": "bn_act_fwd + bn_act_bwd_reduce + bn_act_bwd_apply (elementwise.cu)",
                        "achieved": eb / et / 1e9, "peak": peak_hbm, "peak_source": peak_src + " hbm_gbs", "unit": "GB/s",
                        "frac": eb / et / 1e9 / peak_hbm, "traffic": None, "launches_per_step": en // nprof,
                        "bytes_per_launch": eb / en, "avg_launch_us": et / en * 1e6, "ms_per_step": et / nprof * 1e3,
                        "per_pass": {k: {"GBps": acc[k][0] / acc[k][1] / 1e9, "ms_per_step": acc[k][1] / nprof * 1e3}
                                     for k in ("bn_fwd", "bn_bwd_reduce", "bn_bwd_apply") if k in acc}}
        kern["whole_step_conv_tflops"] = value / world * TRAIN_GFLOP_PER_IMG * 1e9 / 1e12
    if world > 1:
        dist.barrier()

    cpu = None
    if rank == 0:
        clocks.stop()
        if world == 1 and not a.no_cpu_baseline:
            ips, sec, bs, threads, kind, what = cpu_reference(2, 1, budget_s=25.0, size=S)
            cpu = {"value": ips, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"2 timed + 1 warm-up full train steps at bs={bs}, {S}x{S}, fp32 ({what})"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "configs[2]: full train step bf16 bs=64/GPU 640x640 synthetic COCO-80 "
                                   "(fwd + ComputeLoss + bwd + grad all-reduce + clip(10) + Adam)",
                       "batch_per_gpu": B, "global_batch": B * world, "image": S, "targets_per_image": 8,
                       "parallelism": f"dp{world}", "grad_exchange": exch, "l2_policy": "inputs larger than L2 (activations of one step >> 126 MB)"},
            "roofline": roof, "roofline_hbm": roof_hbm, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks.summary(windows), "kernels": kern, "loss": last_loss,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    # the contract is ONE JSON line on stdout: anything a library prints there (e.g. the "NCCL version ..." banner of
    # ncclCommInit) is routed to stderr -- file descriptor 1 points at stderr while the bench runs, and the JSON line goes
    # to a private duplicate of the real stdout
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    try:
        if a.impl == "reference":
            run_reference_detect(a) if a.workload == "detect" else run_reference(a)
        else:
            run_ours_detect(a) if a.workload == "detect" else run_ours(a)
    finally:
        real_stdout.flush()


if __name__ == "__main__":
    main()
