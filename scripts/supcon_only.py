"""Contrastive regulariser (config 3 shape): forward + backward at B x d, timing / profiling target.  env: B, DD, REPS."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm
from dbmm import ops
B, d, reps = int(os.environ.get("B", 8192)), int(os.environ.get("DD", 768)), int(os.environ.get("REPS", 3))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
Z = torch.nn.functional.normalize(torch.randn(B, d, device=dev), dim=1).contiguous()
lab = torch.randint(0, 4, (B,), device=dev, dtype=torch.int32)
sc = ops.SupconState(device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def step():
    sc.zero_(); ops.supcon_fwd(Z, lab, sc); return ops.supcon_bwd(Z, sc)
for _ in range(2): step()
torch.cuda.synchronize(); e0.record()
for _ in range(reps): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
fl = 6.0 * B * d * B
print(f"supcon fwd+bwd B={B} d={d}: {ms:.3f} ms, {fl / ms / 1e9:.1f} algorithmic TFLOP/s = {fl / ms / 1e9 / 1396:.3f} of sustained bf16, loss {sc.loss():.6f}")
