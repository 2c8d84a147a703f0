"""Eval forward only (validate()-style pass over a CelebA-shaped resident matrix): timing / profiling target."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm
from dbmm import ops
from dbmm.modules import Adapter
N, D, H, G = int(os.environ.get("N", 162770)), 1024, 128, 4
dev = torch.device("cuda", 0)
torch.manual_seed(0)
X = torch.randn(N, D, device=dev).half().float()
y = torch.randint(0, 2, (N,), device=dev, dtype=torch.int32)
g = torch.randint(0, 4, (N,), device=dev, dtype=torch.int32)
That = ops.normalize_text(torch.randn(D, 2, device=dev))
ad = Adapter(D, H).to(dev).tensors()
st = ops.BatchStatsBuffers((N + 511) // 512, G, device=dev)
reps = int(os.environ.get("REPS", 5))
for _ in range(2):
    ops.eval_fwd(X, y, g, ad, That, 100.0, st, 512, G=G)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.eval_fwd(X, y, g, ad, That, 100.0, st, 512, G=G)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"eval {N} rows: {ms:.3f} ms/pass, {N / ms / 1e3:.1f} M emb/s, {N * 4096 / ms / 1e6:.0f} GB/s")
