"""Stage-1 training epochs only (train_one_epoch over a CelebA-shaped resident matrix): timing / profiling target.
env: N rows, BS batch size, EPOCHS timed epochs."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm
from dbmm import ops
from dbmm.modules import Adapter
N, D, H, G = int(os.environ.get("N", 162770)), int(os.environ.get("D", 1024)), 128, 4
bs, epochs = int(os.environ.get("BS", 1024)), int(os.environ.get("EPOCHS", 3))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
X = torch.randn(N, D, device=dev).half().float()
y = torch.randint(0, 2, (N,), device=dev, dtype=torch.int32)
g = torch.randint(0, 4, (N,), device=dev, dtype=torch.int32)
That = ops.normalize_text(torch.randn(D, 2, device=dev))
ad = Adapter(D, H).to(dev).tensors()
steps = (N + bs - 1) // bs
st = ops.BatchStatsBuffers(steps, G, device=dev)
buf = ops.TrainBuffers(D, H, device=dev)
order = torch.randperm(N, device=dev).to(torch.int32)
lrs = [0.01] * steps
for _ in range(2):
    ops.train_epoch(X, order, bs, y, g, ad, That, 100.0, buf, lrs, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(epochs):
    ops.train_epoch(X, order, bs, y, g, ad, That, 100.0, buf, lrs, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / epochs
print(f"train {N} rows bs {bs}: {ms:.3f} ms/epoch, {1e3 * ms / steps:.2f} us/step, {N / ms / 1e3:.2f} M emb/s")
