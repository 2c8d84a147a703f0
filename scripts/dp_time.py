"""Data-parallel training epochs only (weak scaling: every rank its own rows): us per SGD step.  torchrun target."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm
from dbmm import ops, parallel
from dbmm.modules import Adapter
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
lr_ = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr_); dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)
N, D, H, G, bs, epochs = int(os.environ.get("N", 162770)), 1024, 128, 4, 1024, 3
torch.manual_seed(rank)
X = torch.randn(N, D, device=dev).half().float()
y = torch.randint(0, 2, (N,), device=dev, dtype=torch.int32); g = torch.randint(0, 4, (N,), device=dev, dtype=torch.int32)
torch.manual_seed(0)
That = ops.normalize_text(torch.randn(D, 2, device=dev))
ad = Adapter(D, H).to(dev).tensors()
steps = (N + bs - 1) // bs
st = ops.BatchStatsBuffers(steps, G, device=dev); buf = ops.TrainBuffers(D, H, device=dev)
order = torch.randperm(N, device=dev).to(torch.int32)
dp = parallel.DataParallelTrainer(local_batches=True)
lrs = [0.01] * steps
for _ in range(2):
    dp.train_epoch(X, order, bs, y, g, ad, That, 100.0, buf, lrs, st, G=G)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(epochs):
    dp.train_epoch(X, order, bs, y, g, ad, That, 100.0, buf, lrs, st, G=G)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / epochs], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world}: {float(ms):.3f} ms/epoch, {1e3 * float(ms) / steps:.2f} us/step, {world * N / float(ms) / 1e3:.2f} M emb/s")
dist.destroy_process_group()
