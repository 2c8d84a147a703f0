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


def accum_views(ws: torch.Tensor, H: int, nad: int, dtype=torch.int64):
    """The two per-step accumulator blocks of a DBMM_OP_TRAIN workspace that a data-parallel caller all-reduces between
    the phases.  The kernels keep them as int64 fixed-point sums (include/dbmm.h: dbmm_train_accum_layout), so they are
    reduced as int64; an injected step function (CPU protocol test) may treat the same 8-byte slots as float64."""
    lib = _lib.load(require_gpu=False)
    co, cc, do, dc = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    _lib.check(lib.dbmm_train_accum_layout(H, nad, C.byref(co), C.byref(cc), C.byref(do), C.byref(dc)))
    colsum = ws[co.value: co.value + 8 * cc.value].view(dtype)
    dgb = ws[do.value: do.value + 8 * dc.value].view(dtype)
    return colsum, dgb


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process (and the pinned host buffers it allocates afterwards: first touch) to the CPUs of the NUMA node the
    GPU hangs off, read from sysfs.  With one process per GPU all started on node 0, every rank's host -> device staging
    crosses the socket interconnect (round 1: 13.6 GB/s per rank at 8 GPUs against 46.5 GB/s alone).  Returns what it did;
    never raises (containers may hide sysfs or forbid sched_setaffinity)."""
    import os
    info = {"device": device_index, "numa_node": None, "cpus": None, "bound": False}
    try:
        import ctypes
        buf = ctypes.create_string_buffer(32)
        if _lib.load().dbmm_device_pci_bus_id(device_index, buf, 32) != 0:
            return info
        bus = buf.value.decode()
        node_path = f"/sys/bus/pci/devices/{bus.lower()}/numa_node"
        if not os.path.exists(node_path):
            return info
        node = int(open(node_path).read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"], info["bound"] = len(allowed), True
    except Exception as exc:          # noqa: BLE001
        info["error"] = repr(exc)[:120]
    return info


class DataParallelTrainer:
    """Data-parallel training.  On CUDA the whole epoch runs inside libdbmm (`dbmm_train_epoch_dp`): the library owns an
    NCCL communicator (created here: rank 0's unique id is broadcast with torch.distributed), and the kernels of every
    step plus the three all-reduces between their phases are captured into one CUDA graph and replayed -- no host work
    per step.  `step_fn` / `all_reduce` are injectable so that the sharding / reduction protocol is testable on CPU with
    gloo, the oracle standing in for the kernels (tests/test_parallel_cpu.py); that path runs the phases of
    dbmm_train_step eagerly with the all-reduces in between, and is also what DBMM_DP=eager selects on CUDA."""

    def __init__(self, group=None, step_fn=None, all_reduce=None, local_batches=False):
        """local_batches=False: every rank passes the same GLOBAL index list and trains on its contiguous shard of it
        (same global batch as a single-GPU run).  local_batches=True: every rank passes the indices of its OWN rows (its
        own data shard, equal counts on all ranks) and the global batch is their union (weak scaling)."""
        self.local_batches = local_batches
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._custom_step = step_fn is not None
        self._step = step_fn or ops.train_step
        self._all_reduce = all_reduce or (lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group))
        self._comm = None

    # ---- native communicator --------------------------------------------------------------------------------------
    def _native_comm(self, device):
        if self._comm is None:
            lib = _lib.load()
            ident = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                buf = (C.c_ubyte * 128)()
                _lib.check(lib.dbmm_comm_unique_id(buf))
                ident = torch.tensor(list(buf), dtype=torch.uint8)
            ident = ident.to(device)
            dist.broadcast(ident, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            raw = bytes(ident.cpu().tolist())
            comm = C.c_void_p()
            _lib.check(lib.dbmm_comm_init(raw, self.world, self.rank, C.byref(comm)))
            self._comm = comm
        return self._comm

    def check(self):
        """Synchronise and raise if a peer-memory wait timed out (a rank never reached the same epoch call)."""
        if self._comm is not None:
            _lib.check(_lib.load().dbmm_comm_check(self._comm))

    def close(self):
        if self._comm is not None:
            _lib.check(_lib.load().dbmm_comm_destroy(self._comm))
            self._comm = None

    # ---- eager protocol (CPU tests, debugging) --------------------------------------------------------------------
    def train_step(self, X, y, grp, ad, That, inv_tau, buf, lr, stats, slot, global_idx: torch.Tensor, *, old_ad=None,
                   ebd_weight=0.5, G=4, momentum=0.9, weight_decay=5e-5):
        """global_idx: the GLOBAL batch's row indices (identical on every rank); this rank takes its shard."""
        if self.local_batches:
            idx, Bg = global_idx, int(global_idx.numel()) * self.world
        else:
            Bg = int(global_idx.numel())
            lo, hi = shard_bounds(Bg, self.world, self.rank)
            idx = global_idx[lo:hi].contiguous()
        kw = dict(idx=idx, B_global=Bg, old_ad=old_ad, ebd_weight=ebd_weight, G=G, momentum=momentum,
                  weight_decay=weight_decay)
        nad = 2 if old_ad is not None else 1
        self._step(X, y, grp, ad, That, inv_tau, buf, lr, stats, slot, phases=_lib.PHASE_GEMM1, **kw)
        ws = ops.workspace(0, X.device)
        colsum, dgb = accum_views(ws, ad.H, nad, torch.float64 if self._custom_step else torch.int64)
        if self.world > 1:
            self._all_reduce(colsum)
        self._step(X, y, grp, ad, That, inv_tau, buf, lr, stats, slot, phases=_lib.PHASE_ROWS, **kw)
        if self.world > 1:
            self._all_reduce(dgb)
        self._step(X, y, grp, ad, That, inv_tau, buf, lr, stats, slot, phases=_lib.PHASE_WGRAD, **kw)
        if self.world > 1:
            self._all_reduce(buf.grads)
        self._step(X, y, grp, ad, That, inv_tau, buf, lr, stats, slot, phases=_lib.PHASE_UPDATE, **kw)

    def _eager_epoch(self, X, order, batch_size_global, y, grp, ad, That, inv_tau, buf, lrs, stats, **kw):
        n = order.numel()
        steps = (n + batch_size_global - 1) // batch_size_global
        for s in range(steps):
            self.train_step(X, y, grp, ad, That, inv_tau, buf, float(np.float32(lrs[s])), stats, s,
                            order[s * batch_size_global:(s + 1) * batch_size_global], **kw)
        return steps

    # ---- epoch ----------------------------------------------------------------------------------------------------
    def train_epoch(self, X, order: torch.Tensor, batch_size_global: int, y, grp, ad, That, inv_tau, buf, lrs, stats, *,
                    old_ad=None, ebd_weight=0.5, G=4, momentum=0.9, weight_decay=5e-5, reduce_stats=False):
        """batch_size_global: rows of `order` consumed per step (the global batch when local_batches is False, this
        rank's share of it otherwise)."""
        import os
        native = (not self._custom_step and getattr(X, "is_cuda", False) and os.environ.get("DBMM_DP", "native") != "eager")
        if not native:
            steps = self._eager_epoch(X, order, batch_size_global, y, grp, ad, That, inv_tau, buf, lrs, stats, old_ad=old_ad,
                                      ebd_weight=ebd_weight, G=G, momentum=momentum, weight_decay=weight_decay)
            if reduce_stats:
                self.reduce_stats(stats)
            return steps
        lib = _lib.load()
        D, H, Cn = X.shape[1], ad.H, That.shape[1]
        n = order.numel()
        steps = (n + batch_size_global - 1) // batch_size_global
        lrs = np.ascontiguousarray(lrs, dtype=np.float32)
        if len(lrs) < steps or stats.n_slots < steps:
            raise _lib.DbmmError(f"need {steps} learning rates / stat slots")
        nad = 2 if old_ad is not None else 1
        ws = ops.workspace(lib.dbmm_workspace_bytes(_lib.OP_TRAIN, min(batch_size_global, n), D, H, Cn, nad), X.device)
        comm = self._native_comm(X.device) if self.world > 1 else None
        old_p = old_ad.ptrs() if old_ad is not None else None
        _lib.check(lib.dbmm_train_epoch_dp(comm, self.world, self.rank, 1 if self.local_batches else 0, X.data_ptr(), X.stride(0),
                                           order.data_ptr(), n, batch_size_global, y.data_ptr(),
                                           None if grp is None else grp.data_ptr(), D, H, Cn, G if grp is not None else 1,
                                           C.byref(old_p) if old_p is not None else None, C.byref(ad.ptrs()), ebd_weight,
                                           That.data_ptr(), inv_tau, buf.grads.data_ptr(), buf.momentum.data_ptr(),
                                           lrs.ctypes.data_as(C.POINTER(C.c_float)), momentum, weight_decay,
                                           1 if buf.first_step else 0, stats.c(), 1 if reduce_stats else 0,
                                           ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
        buf.first_step = False
        return steps

    def reduce_stats(self, stats: ops.BatchStatsBuffers):
        if self.world > 1:
            self._all_reduce(stats.loss_sum)
            self._all_reduce(stats.counts)


def supcon_distributed(Z_local: torch.Tensor, labels_local: torch.Tensor, *, tau_cl=0.1, group=None, compute=None):
    """Contrastive regulariser with GLOBAL negatives (BASELINE config 3): every rank holds B / world
    L2-normalised rows; the rows and labels are all-gathered, each rank scores ITS anchors against the whole batch
    (dbmm_supcon_fwd on the tcgen05 GEMM), loss sum and valid-anchor count are all-reduced, the backward produces the
    anchor-role gradient of the local rows and the contrast-role gradient of ALL rows, and the latter is reduce-scattered
    back to the owners.  Returns (mean loss over the global batch, dZ_local [B_local, d]).

    Equal shard sizes are required (all_gather_into_tensor); `compute` = (fwd, bwd) is the injection point for the gloo
    test, where the oracle stands in for the kernels (the product has no CPU path):
        fwd(Z_all, labels_all, row0, n_local) -> (loss_sum: float, n_valid: int, ctx)
        bwd(Z_all, ctx, row0, n_local, n_valid_global) -> (dZ_local [n_local, d], dZ_all [B, d])"""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    Bl, d = Z_local.shape
    if world > 1:
        Z_all = torch.empty((world * Bl, d), dtype=Z_local.dtype, device=Z_local.device)
        labels_all = torch.empty((world * Bl,), dtype=labels_local.dtype, device=labels_local.device)
        dist.all_gather_into_tensor(Z_all, Z_local.contiguous(), group=group)
        dist.all_gather_into_tensor(labels_all, labels_local.contiguous(), group=group)
    else:
        Z_all, labels_all = Z_local.contiguous(), labels_local.contiguous()
    row0 = rank * Bl
    if compute is None:
        state = ops.SupconState(device=Z_local.device)

        def fwd(Za, la, r0, nl):
            ops.supcon_fwd(Za, la, state, row0=r0, n_local=nl, tau_cl=tau_cl)
            return state.loss_sum, state.n_valid, None

        def bwd(Za, ctx, r0, nl, n_valid_global):
            state.n_valid.copy_(n_valid_global)
            return ops.supcon_bwd(Za, state, row0=r0, n_local=nl, tau_cl=tau_cl)
    else:
        fwd, bwd = compute
    loss_sum, n_valid, ctx = fwd(Z_all, labels_all, row0, Bl)
    red = torch.stack([torch.as_tensor(loss_sum, dtype=torch.float64, device=Z_local.device).reshape(()),
                       torch.as_tensor(n_valid, device=Z_local.device).to(torch.float64).reshape(())])
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.SUM, group=group)
    n_valid_global = red[1].to(torch.int32).reshape(1)
    dZ_local, dZ_all = bwd(Z_all, ctx, row0, Bl, n_valid_global)
    if world > 1:
        mine = torch.empty_like(dZ_local)
        dist.reduce_scatter_tensor(mine, dZ_all.contiguous(), op=dist.ReduceOp.SUM, group=group)
        dZ_local = dZ_local + mine
    else:
        dZ_local = dZ_local + dZ_all
    return float(red[0].item() / max(float(red[1].item()), 1.0)), dZ_local


def contrastive_step_distributed(X, labels, ad, buf, lr, idx_local, *, pre_norm=True, tau_cl=0.1, loss_weight=0.1, momentum=0.9,
                                 weight_decay=5e-5, group=None):
    """One data-parallel step of `--tl_method contrastive_adapter` with GLOBAL negatives (BASELINE config 3): every rank runs
    forward_ca on its own rows (BatchNorm statistics of its shard, as torch DDP without SyncBatchNorm), the normalised rows and
    labels are all-gathered, each rank scores its anchors against the whole global batch (tcgen05 similarity GEMMs), the
    contrast-role gradient is reduce-scattered back to the owners (supcon_distributed), every rank back-propagates its rows
    through its adapter copy, the flat gradients are all-reduced and every rank applies the same SGD step: the replicas stay
    identical.  Returns the global mean loss (unweighted)."""
    U, lab, ws = ops.contrastive_forward(X, labels, ad, idx=idx_local, pre_norm=pre_norm)
    loss, dU = supcon_distributed(U, lab, tau_cl=tau_cl, group=group)
    dU = dU.contiguous()
    ops.contrastive_backward(X, ad, dU, buf.grads, ws, idx=idx_local, pre_norm=pre_norm, loss_weight=loss_weight)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf.grads, op=dist.ReduceOp.SUM, group=group)
    ops.contrastive_apply(ad, buf, lr, U.shape[0], ws, momentum=momentum, weight_decay=weight_decay)
    return loss
