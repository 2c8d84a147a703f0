"""Data-parallel check on N GPUs (torchrun): N ranks training on shards of the same global batches must reproduce the
single-GPU weights (same global batch order; BatchNorm statistics / CE mean over the global batch).  Prints max relative
deviation per tensor; exits non-zero above 1e-3."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm  # noqa: E402
from dbmm import ops, parallel  # noqa: E402
from dbmm.modules import Adapter  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl", device_id=dev)
D, H, C, G, N, BS, STEPS = 1024, 128, 2, 4, 8192, 1000, 6
rng = np.random.default_rng(3)
mu = rng.standard_normal((4, D)).astype(np.float32)
g_np = rng.integers(0, 4, N).astype(np.int32)
x_np = (0.3 * mu[g_np] + rng.standard_normal((N, D))).astype(np.float16).astype(np.float32)
T_np = np.stack([mu[[0, 1]].mean(0), mu[[2, 3]].mean(0)], 1).astype(np.float32)
X, y, g = torch.from_numpy(x_np).to(dev), torch.from_numpy(g_np // 2).to(dev), torch.from_numpy(g_np).to(dev)
That = ops.normalize_text(torch.from_numpy(T_np).to(dev))
order = torch.from_numpy(rng.permutation(N).astype(np.int32)).to(dev)
lrs = np.full(STEPS, 0.5, np.float32)


def fresh():
    torch.manual_seed(1)
    a = Adapter(D, H).to(dev)
    return a, a.tensors()


def run(dp):
    mod, ad = fresh()
    buf, stats = ops.TrainBuffers(D, H, device=dev), ops.BatchStatsBuffers(STEPS, G, device=dev)
    if dp is None:
        ops.train_epoch(X, order[:STEPS * BS].contiguous(), BS, y, g, ad, That, 100.0, buf, lrs, stats, G=G)
    else:
        dp.train_epoch(X, order[:STEPS * BS].contiguous(), BS, y, g, ad, That, 100.0, buf, lrs, stats, G=G)
        dp.reduce_stats(stats)
    torch.cuda.synchronize()
    return ad.to_numpy(), stats.host()


p1, (l1, c1) = run(None)
trainer = parallel.DataParallelTrainer()
p2, (l2, c2) = run(trainer)
if rank == 0:
    import dbmm._lib as L
    print('fused peer-memory all-reduce active:', bool(L.load().dbmm_comm_has_p2p(trainer._comm)) if trainer._comm else None)
worst = 0.0
for k in ("W1", "b1", "gamma", "beta", "W2", "b2", "running_mean", "running_var"):
    err = float(np.abs(p1[k] - p2[k]).max() / (np.abs(p1[k]).max() + 1e-30))
    worst = max(worst, err)
    if rank == 0:
        print(f"{k:14s} max rel dev world={world} vs 1: {err:.2e}")
ok = worst < 1e-3 and np.allclose(l1, l2, rtol=1e-3) and np.abs(c1 - c2).max() <= 2
if rank == 0:
    print("loss per step single:", np.round(l1 / BS, 5), "dp:", np.round(l2 / BS, 5), "count delta", int(np.abs(c1 - c2).max()))
    print("DP CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
