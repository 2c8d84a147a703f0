"""Eval forward over a CelebA-shaped resident matrix: fp32-resident (dbmm_eval_fwd) vs fp16-resident (dbmm_eval_fwd_f16)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm
from dbmm import ops
from dbmm.modules import Adapter
N, D, H = int(os.environ.get("N", 162770)), int(os.environ.get("D", 1024)), 128
dev = torch.device("cuda", 0)
torch.manual_seed(0)
X16 = torch.randn(N, D, device=dev).half(); X32 = X16.float()
y = torch.randint(0, 2, (N,), device=dev, dtype=torch.int32); g = torch.randint(0, 4, (N,), device=dev, dtype=torch.int32)
That = ops.normalize_text(torch.randn(D, 2, device=dev))
ad = Adapter(D, H).to(dev).tensors()
st = ops.BatchStatsBuffers((N + 1023) // 1024, 4, device=dev)
def bench(fn, reps=int(os.environ.get("REPS", 10))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
if os.environ.get("ONLY") != "f16":
    ms = bench(lambda: ops.eval_fwd(X32, y, g, ad, That, 100.0, st, 1024))
    print(f"fp32-resident eval {N} rows: {ms:.3f} ms, {N / ms / 1e3:.1f} M rows/s, {N * 4096 / ms / 1e6:.0f} GB/s algorithmic ({N * 4096 / ms / 1e6 / 6460:.3f} of HBM)")
ms = bench(lambda: ops.eval_fwd_f16(X16, y, g, ad, That, 100.0, st, 1024))
print(f"fp16-resident eval {N} rows: {ms:.3f} ms, {N / ms / 1e3:.1f} M rows/s, {N * 4096 / ms / 1e6:.0f} GB/s algorithmic ({N * 4096 / ms / 1e6 / 6460:.3f} of HBM); bytes moved {N * 2048 / ms / 1e6:.0f} GB/s")
