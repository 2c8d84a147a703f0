"""Timeline of the training step inside the running epoch graph (needs a -DDBMM_TIMELINE build of libdbmm.so: see
profiles/r3_step_timeline.md for the command).  Prints, averaged over 40 steps in the middle of an epoch, when each kernel's first
CTA entered, when it passed its dependency wait and when its last CTA left, in us relative to the row kernel's wait of that step."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm
from dbmm import ops, _lib
from dbmm.modules import Adapter
N, D, H, G = int(os.environ.get("N", 162770)), 1024, 128, 4
bs = int(os.environ.get("BS", 1024))
world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
lrank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
dp = None
if world > 1:                  # torchrun: the data-parallel step (rank-local batches), rank 0 prints its own timeline
    import torch.distributed as dist
    from dbmm import parallel
    dist.init_process_group("nccl", device_id=dev)
    dp = parallel.DataParallelTrainer(local_batches=True)
torch.manual_seed(rank)
X = torch.randn(N, D, device=dev).half().float()
y = torch.randint(0, 2, (N,), device=dev, dtype=torch.int32)
g = torch.randint(0, 4, (N,), device=dev, dtype=torch.int32)
torch.manual_seed(0)
That = ops.normalize_text(torch.randn(D, 2, device=dev))
torch.manual_seed(1)
ad = Adapter(D, H).to(dev).tensors()
steps = (N + bs - 1) // bs
st = ops.BatchStatsBuffers(steps, G, device=dev); buf = ops.TrainBuffers(D, H, device=dev)
order = torch.randperm(N, device=dev).to(torch.int32)
for _ in range(3):
    if dp is not None:
        dp.train_epoch(X, order, bs, y, g, ad, That, 100.0, buf, [0.01] * steps, st, G=G)
    else:
        ops.train_epoch(X, order, bs, y, g, ad, That, 100.0, buf, [0.01] * steps, st)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    if rank != 0:
        dist.destroy_process_group(); sys.exit(0)
lib = _lib.load()
ring, kern = C.c_int(0), C.c_int(0)
out = np.zeros((1024, 8, 3), np.uint64); cnt = np.zeros(8, np.uint32)
rc = lib.dbmm_timeline_dump(out.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p), C.byref(ring), C.byref(kern))
assert rc == 0, lib.dbmm_last_error().decode()
names = ["gemm1", "reduce", "rows", "wgrad", "tail_w1", "sum_spart_g (S^T)", "hs_w2", "sum_gpart"]
print("launch counts:", dict(zip(names, cnt.tolist())))


def stamp(k, back):            # launch `back` from the end of kernel k's own ring (every kernel runs once per step)
    i = int(cnt[k]) - 1 - back
    return out[i % 1024, k].astype(np.int64)


acc = np.zeros((8, 3)); period = []; n = 0
for back in range(20, 60):     # 40 steps from the middle of the last epoch (exit stamps exist for all but the very last launch)
    t0 = stamp(2, back)[1]     # the row kernel of step s passes its wait
    rows = []
    for k in range(8):         # gemm1 / reduce of the NEXT step are shown (index back - 1): one whole period rows(s) .. rows(s + 1)
        rows.append((stamp(k, back - 1 if k in (0, 1) else back) - t0) / 1e3)
    acc += np.array(rows); n += 1
    period.append((stamp(2, back - 1)[1] - t0) / 1e3)
acc /= n
print(f"mean over {n} steps, us relative to the row kernel passing its dependency wait; period rows(s) -> rows(s+1): {np.mean(period):.1f} us")
print("| kernel | first CTA entered | passed its wait | last CTA left | busy after the wait |\n|---|---|---|---|---|")
for k in [2, 5, 6, 7, 3, 4, 0, 1]:
    if cnt[k] < 100:
        continue
    e, w, x = acc[k]
    tag = " (next step)" if k in (0, 1) else ""
    print(f"| {names[k]}{tag} | {e:.1f} | {w:.1f} | {x:.1f} | {x - w:.1f} |")
