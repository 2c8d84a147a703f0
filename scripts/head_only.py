"""Zero-shot head (config 4 slice: N rows x 1024-d against C prompt columns): fp16-resident kind::f16 head vs fp32-resident tf32 head.
env: N rows, C prompts, REPS.  Timing / profiling target."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm
from dbmm import ops
N, D, C, reps = int(os.environ.get("N", 262144)), 1024, int(os.environ.get("C", 1000)), int(os.environ.get("REPS", 3))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
U16 = torch.randn(N, D, device=dev).half()
y = torch.randint(0, C, (N,), device=dev, dtype=torch.int32)
g = torch.randint(0, 4, (N,), device=dev, dtype=torch.int32)
Th = ops.normalize_text(torch.randn(D, C, device=dev))
st = ops.BatchStatsBuffers((N + 1023) // 1024, 4, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timed(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
fl = 2.0 * D * C * N
ms = timed(lambda: ops.logits_ce_f16(U16, y, g, Th, 100.0, st, 1024))
print(f"f16 head {N} x {D} x {C}: {ms:.3f} ms, {fl / ms / 1e9:.1f} algorithmic TFLOP/s = {fl / ms / 1e9 / 1396:.3f} of sustained bf16")
if os.environ.get("F32", "1") == "1":
    U = U16.float()
    ms = timed(lambda: ops.logits_ce(U, y, g, Th, 100.0, st, 1024))
    print(f"tf32 head: {ms:.3f} ms, {fl / ms / 1e9:.1f} algorithmic TFLOP/s = {fl / ms / 1e9 / 1396:.3f} of sustained bf16")
