"""Per-launch device times of ONE eval forward at the detect config (bs=128, 1280x1280) -- run under
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/detect_launches.csv \
        python tools/detect_layers.py
and summarised by layer (conv shape) with its tensor / HBM floor."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yolov5m_b200 as yb  # noqa: E402

torch.manual_seed(0)
bs = int(os.environ.get("BS", "128"))
m = yb.YOLOV5m(48, 80, yb.ANCHORS, (192, 384, 768)).cuda().eval()
x = torch.randint(0, 256, (bs, 3, 1280, 1280), dtype=torch.uint8, device="cuda")
with torch.no_grad():
    m(x)
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
eng = m.engine(bs, 1280, 1280, False)
print("fwd ops", len(eng.fwd_ops))
