"""Batched-adapter stage-1 epochs (config 5): M members in lock step over one resident CelebA-shaped matrix: timing / profiling target.
env: N rows, BS batch size, M members, EPOCHS timed epochs."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm
from dbmm import ops
from dbmm.modules import Adapter
N, D, H, G = int(os.environ.get("N", 162770)), int(os.environ.get("D", 1024)), 128, 4
bs, epochs, M = int(os.environ.get("BS", 1024)), int(os.environ.get("EPOCHS", 2)), int(os.environ.get("M", 64))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
X = torch.randn(N, D, device=dev).half().float()
y = torch.randint(0, 2, (N,), device=dev, dtype=torch.int32)
g = torch.randint(0, 4, (N,), device=dev, dtype=torch.int32)
That = ops.normalize_text(torch.randn(D, 2, device=dev))
steps = (N + bs - 1) // bs
members = []
for m in range(M):
    torch.manual_seed(m)
    members.append(ops.SweepMember(order=torch.randperm(N, device=dev).to(torch.int32), ad=Adapter(D, H).to(dev).tensors(),
                                   buf=ops.TrainBuffers(D, H, device=dev), stats=ops.BatchStatsBuffers(steps, G, device=dev),
                                   lrs=np.full(steps, 0.01, np.float32)))
ops.train_epoch_batched(X, members, bs, y, g, That, 100.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(epochs):
    ops.train_epoch_batched(X, members, bs, y, g, That, 100.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / epochs
flop = 1.319e6 * N * M
print(f"batched M={M} {N} rows bs {bs}: {ms:.3f} ms/epoch, {1e3 * ms / steps / M:.2f} us/member-step, {N * M / ms / 1e3:.2f} M emb/s, "
      f"{flop / ms / 1e9:.1f} algorithmic TFLOP/s = {flop / ms / 1e9 / 1396:.3f} of sustained bf16")
