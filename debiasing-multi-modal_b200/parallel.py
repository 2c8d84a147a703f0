"""Data-parallel training step over one NVSwitch box: one process per GPU, torch.distributed (NCCL) for the
exchanges.  The reference is single-process; to keep its semantics the GLOBAL batch is
what BatchNorm and the CE mean see, so each step exchanges

    after GEMM-1 :  per-column sum / sum-of-squares of the pre-BN activations   fp64 [n_adapters][2][H]
    after rows   :  (dgamma, dbeta) partial sums (they also give BatchNorm's backward means) fp64 [2][H]
    after wgrad  :  the flat gradient  fp32 [2DH + 3H + D]  (1.05 MB at D=1024, H=128)

and, once per epoch, the per-batch loss sums and group counters.  Weights, momentum and BatchNorm buffers stay
replicated: every rank applies the same update to the same all-reduced gradient.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous split of a global batch of n rows; the first n % world ranks get one extra row."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def accum_views(ws: torch.Tensor, H: int, nad: int):
    lib = _lib.load(require_gpu=False)
    co, cc, do, dc = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    _lib.check(lib.dbmm_train_accum_layout(H, nad, C.byref(co), C.byref(cc), C.byref(do), C.byref(dc)))
    colsum = ws[co.value: co.value + 8 * cc.value].view(torch.float64)
    dgb = ws[do.value: do.value + 8 * dc.value].view(torch.float64)
    return colsum, dgb


class DataParallelTrainer:
    """Runs the phases of dbmm_train_step with the all-reduces in between.  `step_fn` and `all_reduce` are
    injectable so the sharding / reduction logic is testable on CPU with gloo (tests/test_parallel_cpu.py)."""

    def __init__(self, group=None, step_fn=None, all_reduce=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._step = step_fn or ops.train_step
        self._all_reduce = all_reduce or (lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group))

    def train_step(self, X, y, grp, ad, That, inv_tau, buf, lr, stats, slot, global_idx: torch.Tensor, *, old_ad=None,
                   ebd_weight=0.5, G=4, momentum=0.9, weight_decay=5e-5):
        """global_idx: the GLOBAL batch's row indices (identical on every rank); this rank takes its shard."""
        Bg = int(global_idx.numel())
        lo, hi = shard_bounds(Bg, self.world, self.rank)
        idx = global_idx[lo:hi].contiguous()
        kw = dict(idx=idx, B_global=Bg, old_ad=old_ad, ebd_weight=ebd_weight, G=G, momentum=momentum,
                  weight_decay=weight_decay)
        nad = 2 if old_ad is not None else 1
        self._step(X, y, grp, ad, That, inv_tau, buf, lr, stats, slot, phases=_lib.PHASE_GEMM1, **kw)
        ws = ops.workspace(0, X.device)
        colsum, dgb = accum_views(ws, ad.H, nad)
        if self.world > 1:
            self._all_reduce(colsum)
        self._step(X, y, grp, ad, That, inv_tau, buf, lr, stats, slot, phases=_lib.PHASE_ROWS, **kw)
        if self.world > 1:
            self._all_reduce(dgb)
        self._step(X, y, grp, ad, That, inv_tau, buf, lr, stats, slot, phases=_lib.PHASE_WGRAD, **kw)
        if self.world > 1:
            self._all_reduce(buf.grads)
        self._step(X, y, grp, ad, That, inv_tau, buf, lr, stats, slot, phases=_lib.PHASE_UPDATE, **kw)

    def train_epoch(self, X, order: torch.Tensor, batch_size_global: int, y, grp, ad, That, inv_tau, buf, lrs, stats, **kw):
        n = order.numel()
        steps = (n + batch_size_global - 1) // batch_size_global
        for s in range(steps):
            self.train_step(X, y, grp, ad, That, inv_tau, buf, float(np.float32(lrs[s])), stats, s,
                            order[s * batch_size_global:(s + 1) * batch_size_global], **kw)
        return steps

    def reduce_stats(self, stats: ops.BatchStatsBuffers):
        if self.world > 1:
            self._all_reduce(stats.loss_sum)
            self._all_reduce(stats.counts)
