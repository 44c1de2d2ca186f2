import sys, time, torch
sys.path.insert(0, '/root/repo')
import yolov5m_b200 as yb
from yolov5m_b200.trainer import Adam, TrainStep
dev = torch.device('cuda', 0)
torch.manual_seed(0)
m = yb.YOLOV5m(48, 80, yb.ANCHORS, (192, 384, 768)).to(dev).train()
step = TrainStep(m, yb.ComputeLoss(m), Adam(m), max_norm=10.0)
B = 64
g = torch.Generator().manual_seed(1)
x = torch.randint(0, 256, (B, 3, 640, 640), dtype=torch.uint8, generator=g).to(dev)
nt = 8 * B
t = torch.cat([torch.randint(0, B, (nt, 1), generator=g).float(), torch.randint(0, 80, (nt, 1), generator=g).float(), torch.rand(nt, 2, generator=g), torch.rand(nt, 2, generator=g) * 0.5 + 0.005], 1).to(dev)
for _ in range(3): step(x, t)
torch.cuda.synchronize()
for trial in range(3):
    t0 = time.perf_counter(); step(x, t); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"CPU enqueue of one step: {(t1 - t0) * 1e3:.2f} ms; until GPU done: {(t2 - t0) * 1e3:.2f} ms")
