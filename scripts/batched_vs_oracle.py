"""Debug: every member of a batched run vs the numpy oracle and vs the stand-alone run."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dbmm
from dbmm import ops
from oracle import adapter_math as am
D, H = 1024, 128
M, bs, n = 5, int(os.environ.get("BS", 1024)), int(os.environ.get("N", 3405))
epochs = int(os.environ.get("EPOCHS", 2))
rng = np.random.default_rng(55)
base = rng.standard_normal(D).astype(np.float32); mu = rng.standard_normal((4, D)).astype(np.float32)
g = rng.choice(4, n, p=[0.44, 0.41, 0.14, 0.01])
x = (base + 0.2 * mu[g] + rng.standard_normal((n, D)).astype(np.float32)).astype(np.float16).astype(np.float32)
T2 = (base[:, None] + np.stack([mu[[0, 1]].mean(0), mu[[2, 3]].mean(0)], 1)).astype(np.float32)
y = g // 2
dev = lambda a, dt=None: (torch.from_numpy(np.ascontiguousarray(a)).to(dt) if dt is not None else torch.from_numpy(np.ascontiguousarray(a))).cuda()
X, yd, gd = dev(x), dev(y, torch.int32), dev(g, torch.int32)
That = ops.normalize_text(dev(T2)); That_np = am.normalize_text(T2)
steps = (n + bs - 1) // bs
inits = [am.init_adapter_params(rng, D, H) for _ in range(M)]
orders = [rng.permutation(n).astype(np.int32) for _ in range(M)]
lrs = [np.linspace(0.02 * (m + 1), 0.01 * (m + 1), steps).astype(np.float32) for m in range(M)]
rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))
members = [ops.SweepMember(order=dev(orders[m]), ad=ops.AdapterTensors.from_numpy(inits[m]), buf=ops.TrainBuffers(D, H),
                           stats=ops.BatchStatsBuffers(steps, 4), lrs=lrs[m]) for m in range(M)]
for _ in range(epochs):
    ops.train_epoch_batched(X, members, bs, yd, gd, That, 100.0)
torch.cuda.synchronize()
for m in range(M):
    ad = ops.AdapterTensors.from_numpy(inits[m]); buf = ops.TrainBuffers(D, H); st = ops.BatchStatsBuffers(steps, 4)
    for _ in range(epochs):
        ops.train_epoch(X, dev(orders[m]), bs, yd, gd, ad, That, 100.0, buf, lrs[m], st)
    torch.cuda.synchronize()
    p, v = am.copy_params(inits[m]), None
    losses = []
    for ep in range(epochs):
        for s in range(steps):
            ii = orders[m][s * bs:(s + 1) * bs]
            r = am.train_step_single(x[ii], y[ii], p, v, That_np, 0.01, float(lrs[m][s])); v = r["v"]; losses.append(float(r["loss"]))
    gb, gs = members[m].ad.to_numpy(), ad.to_numpy()
    print(f"member {m} lr {lrs[m][0]:.2f}: W1 batched-oracle {rel(gb['W1'], p['W1']):.2e} single-oracle {rel(gs['W1'], p['W1']):.2e} batched-single {rel(gb['W1'], gs['W1']):.2e} | "
          f"W2 b-o {rel(gb['W2'], p['W2']):.2e} s-o {rel(gs['W2'], p['W2']):.2e} | loss first/last {losses[0]:.3e} {losses[-1]:.3e}", flush=True)
