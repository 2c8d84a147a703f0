"""ORACLE (test infrastructure / CPU baseline).  A torch-CPU restatement of the reference's modules and
step body, used (a) as the timed CPU baseline in bench.py (`cpu_baseline`, `--impl reference`; kind = "port":
the reference itself is Python that cannot travel to the GPU box) and (b) as a second, autograd-based
cross-check of oracle/adapter_math.py.  Follows final_main.py:53-174 (modules), 455-466 (step), 675-693
(eval) and demo/util.py:118-136 (optimizer); pinned to the reference by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn as nn


class PortAdapter(nn.Module):
    def __init__(self, D, H):
        super().__init__()
        self.layers = nn.Sequential(nn.Linear(D, H), nn.BatchNorm1d(H), nn.ReLU(), nn.Linear(H, D))

    def forward(self, x):
        return self.layers(x)


class PortCLIP(nn.Module):
    """CustomCLIP with in-memory prompt matrices (class / spurious / group), temperature tau."""

    def __init__(self, adapter, T_class, T_spurious=None, T_group=None, tau=0.01, old=None, w=0.5):
        super().__init__()
        self.adapter, self.old, self.w, self.tau = adapter, old, w, tau
        self.T = {"class": T_class, "spurious": T_spurious, "group": T_group}

    def forward(self, x, which="class"):
        z = self.adapter(x)
        u = z / z.norm(dim=-1, keepdim=True)
        if self.old is not None:
            zo = self.old(x)
            u = self.w * (zo / zo.norm(dim=-1, keepdim=True)).detach() + (1 - self.w) * u
        T = self.T[which]
        T = T / T.norm(dim=0, keepdim=True)
        return u @ T / self.tau


def load_params(adapter: PortAdapter, p: dict):
    sd = {"layers.0.weight": p["W1"], "layers.0.bias": p["b1"], "layers.1.weight": p["gamma"], "layers.1.bias": p["beta"],
          "layers.1.running_mean": p["running_mean"], "layers.1.running_var": p["running_var"],
          "layers.1.num_batches_tracked": np.int64(p["num_batches_tracked"]), "layers.3.weight": p["W2"],
          "layers.3.bias": p["b2"]}
    adapter.load_state_dict({k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()})


def update_group_meters(counts, logits, y, g, n_groups):
    """update_dict (final_main.py:383-391) with its per-group host syncs."""
    pred = torch.argmax(logits, dim=1)
    correct = pred == y
    for gv in np.unique(g.numpy()):
        m = g == gv
        counts[int(gv)][0] += correct[m].sum().item()
        counts[int(gv)][1] += m.sum().item()


def time_train_steps(X: np.ndarray, y: np.ndarray, g: np.ndarray, T_class: np.ndarray, H: int, batch_size: int,
                     n_steps: int, lr: float = 0.1, threads: int | None = None, seed: int = 0) -> dict:
    """Time `n_steps` SGD steps of the reference's step body on the host cores (tensors pre-loaded: the
    compute-only number of BASELINE.md, without the pandas/DataLoader overhead)."""
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(seed)
    D = X.shape[1]
    model = PortCLIP(PortAdapter(D, H), torch.from_numpy(T_class))
    opt = torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9, weight_decay=5e-5)
    crit = nn.CrossEntropyLoss()
    Xt, yt, gt = torch.from_numpy(X), torch.from_numpy(y), torch.from_numpy(g)
    counts = {k: [0, 0] for k in range(4)}
    model.train()
    n = X.shape[0]
    rows = 0
    t0 = time.perf_counter()
    for s in range(n_steps):
        lo = (s * batch_size) % max(n - batch_size + 1, 1)
        xb, yb, gb = Xt[lo:lo + batch_size], yt[lo:lo + batch_size], gt[lo:lo + batch_size]
        out = model(xb.detach())
        loss = crit(out, yb)
        loss.item()
        opt.zero_grad(); loss.backward(); opt.step()
        update_group_meters(counts, out, yb, gb, 4)
        rows += len(yb)
    dt = time.perf_counter() - t0
    return dict(seconds=dt, rows=rows, emb_per_s=rows / dt, threads=torch.get_num_threads())
