"""BASELINE config 3 on N GPUs (torchrun): ViT-L/14-shaped 768-d rows, global batch 8192, contrastive regulariser with
all-gathered negatives (dbmm.parallel.supcon_distributed).  Checks loss and gradient against the same global batch
computed on one GPU, prints the time of a distributed forward + backward."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm
from dbmm import ops, parallel
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lr_ = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr_); dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)
B, d = int(os.environ.get("B", 8192)), 768
g = torch.Generator().manual_seed(0)
Z = torch.nn.functional.normalize(torch.randn(B, d, generator=g), dim=1).to(dev)
labels = torch.randint(0, 4, (B,), generator=g, dtype=torch.int32).to(dev)
Bl = B // world
Zl, ll = Z[rank * Bl:(rank + 1) * Bl].contiguous(), labels[rank * Bl:(rank + 1) * Bl].contiguous()
loss, dZl = parallel.supcon_distributed(Zl, ll, tau_cl=0.1)
# the same batch on this GPU alone
st = ops.SupconState(device=dev)
ops.supcon_fwd(Z, labels, st, tau_cl=0.1)
a, b = ops.supcon_bwd(Z, st, tau_cl=0.1)
ref = (a + b)[rank * Bl:(rank + 1) * Bl]
err = float((dZl - ref).abs().max() / ref.abs().max())
ok = abs(loss - st.loss()) <= 1e-5 * abs(st.loss()) and err < 1e-4
for _ in range(2):
    parallel.supcon_distributed(Zl, ll, tau_cl=0.1)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    parallel.supcon_distributed(Zl, ll, tau_cl=0.1)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world}: loss {loss:.6f} (one GPU {st.loss():.6f}), max rel gradient deviation {err:.2e}, "
          f"{float(ms):.3f} ms per step of {B} rows = {B / float(ms) / 1e3:.2f} M emb/s", "SUPCON DP OK" if int(flag) else "SUPCON DP FAILED")
dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
