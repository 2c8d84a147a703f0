"""Operator layer: torch tensors in, libdbmm.so (C ABI, include/dbmm.h) calls out.

torch is used for device memory and streams only; every arithmetic step of the hot path runs in
the hand-written kernels.  Each wrapper names the reference code it stands in for.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import AdapterPtrs, BatchStats, DbmmError

_workspaces: dict = {}


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def _check(t: torch.Tensor, dtype, name: str, contiguous=True):
    if not t.is_cuda:
        raise DbmmError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise DbmmError(f"{name} must be {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise DbmmError(f"{name} must be contiguous")


def workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only per-device scratch buffer (the C ABI never allocates)."""
    key = torch.device(device).index or 0
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


@dataclass
class AdapterTensors:
    """Device tensors of one Adapter (final_main.py:160-174); shapes as in the state_dict."""
    W1: torch.Tensor
    b1: torch.Tensor
    gamma: torch.Tensor
    beta: torch.Tensor
    running_mean: torch.Tensor
    running_var: torch.Tensor
    num_batches_tracked: torch.Tensor
    W2: torch.Tensor
    b2: torch.Tensor

    @property
    def D(self):
        return self.W1.shape[1]

    @property
    def H(self):
        return self.W1.shape[0]

    def validate(self):
        H, D = self.W1.shape
        for n, shape, dt in (("W1", (H, D), torch.float32), ("b1", (H,), torch.float32), ("gamma", (H,), torch.float32),
                             ("beta", (H,), torch.float32), ("running_mean", (H,), torch.float32),
                             ("running_var", (H,), torch.float32), ("num_batches_tracked", (), torch.int64),
                             ("W2", (D, H), torch.float32), ("b2", (D,), torch.float32)):
            t = getattr(self, n)
            _check(t, dt, n)
            if tuple(t.shape) != shape:
                raise DbmmError(f"{n} has shape {tuple(t.shape)}, expected {shape}")

    def ptrs(self) -> AdapterPtrs:
        self.validate()
        return AdapterPtrs(*(getattr(self, n).data_ptr() for n in (
            "W1", "b1", "gamma", "beta", "running_mean", "running_var", "num_batches_tracked", "W2", "b2")))

    @staticmethod
    def from_numpy(p: dict, device="cuda") -> "AdapterTensors":
        def f(k):
            return torch.from_numpy(np.ascontiguousarray(p[k], dtype=np.float32)).to(device)
        return AdapterTensors(f("W1"), f("b1"), f("gamma"), f("beta"), f("running_mean"), f("running_var"),
                              torch.tensor(int(p["num_batches_tracked"]), dtype=torch.int64, device=device),
                              f("W2"), f("b2"))

    def to_numpy(self) -> dict:
        d = {k: getattr(self, k).detach().cpu().numpy().copy() for k in (
            "W1", "b1", "gamma", "beta", "running_mean", "running_var", "W2", "b2")}
        d["num_batches_tracked"] = np.int64(self.num_batches_tracked.item())
        return d


def param_count(D: int, H: int) -> int:
    return 2 * D * H + 3 * H + D


def flat_param_slices(D: int, H: int) -> dict:
    """Offsets of W1 | b1 | gamma | beta | W2 | b2 inside the flat gradient / momentum buffers."""
    o, out = 0, {}
    for k, n in (("W1", H * D), ("b1", H), ("gamma", H), ("beta", H), ("W2", D * H), ("b2", D)):
        out[k] = slice(o, o + n)
        o += n
    return out


class BatchStatsBuffers:
    """Per-batch loss sums and per-group correct/total counters (update_dict, final_main.py:383-391)."""

    def __init__(self, n_slots: int, G: int, device="cuda"):
        self.n_slots, self.G = n_slots, G
        self.loss_sum = torch.zeros(n_slots, dtype=torch.float64, device=device)
        self.counts = torch.zeros(n_slots, 2, G, dtype=torch.int64, device=device)

    def zero_(self):
        self.loss_sum.zero_()
        self.counts.zero_()

    def c(self) -> BatchStats:
        return BatchStats(self.loss_sum.data_ptr(), self.counts.data_ptr())

    def host(self):
        """One device->host read per epoch (the reference syncs 6-10 times per batch)."""
        return self.loss_sum.cpu().numpy(), self.counts.cpu().numpy()


def normalize_text(T: torch.Tensor) -> torch.Tensor:
    """final_main.py:77 / 136: text / text.norm(dim=0, keepdim=True), computed once instead of per forward."""
    lib = _lib.load()
    _check(T, torch.float32, "T")
    out = torch.empty_like(T)
    _lib.check(lib.dbmm_normalize_text(T.data_ptr(), out.data_ptr(), T.shape[0], T.shape[1], _stream_ptr()))
    return out


def _label_args(y, grp, G):
    if y is not None:
        _check(y, torch.int32, "y")
    if grp is not None:
        _check(grp, torch.int32, "grp")
    return (G if grp is not None else 1)


def eval_fwd(X: torch.Tensor, y: torch.Tensor, grp, ad: AdapterTensors, That: torch.Tensor, inv_tau: float,
             stats: BatchStatsBuffers | None, batch_size: int, *, idx=None, n_rows=None, old_ad: AdapterTensors | None = None,
             ebd_weight: float = 0.5, G: int = 4, want_logits=False, want_pred=False):
    """validate()/validate_zs() forward over many rows (final_main.py:655-803)."""
    lib = _lib.load()
    _check(X, torch.float32, "X", contiguous=False)
    if X.stride(1) != 1:
        raise DbmmError("X rows must be contiguous")
    _check(That, torch.float32, "That")
    D, H, Cn = X.shape[1], ad.H, That.shape[1]
    if That.shape[0] != D or ad.D != D:
        raise DbmmError("dimension mismatch between X, adapter and text prompts")
    G = _label_args(y, grp, G)
    if idx is not None:
        _check(idx, torch.int32, "idx")
    N = int(n_rows if n_rows is not None else (idx.numel() if idx is not None else X.shape[0]))
    nad = 2 if old_ad is not None else 1
    nbytes = lib.dbmm_workspace_bytes(_lib.OP_EVAL, max(N, 1), D, H, Cn, nad)
    ws = workspace(nbytes, X.device)
    logits = torch.empty((N, Cn), dtype=torch.float32, device=X.device) if want_logits else None
    pred = torch.empty((N,), dtype=torch.int32, device=X.device) if want_pred else None
    st = stats.c() if stats is not None else BatchStats(None, None)
    old_p = old_ad.ptrs() if old_ad is not None else None
    _lib.check(lib.dbmm_eval_fwd(X.data_ptr(), X.stride(0), _ptr(idx), _ptr(y), _ptr(grp), N, D, H, Cn, G,
                                 C.byref(old_p) if old_p is not None else None, C.byref(ad.ptrs()), ebd_weight,
                                 That.data_ptr(), inv_tau, batch_size, st, _ptr(logits), _ptr(pred),
                                 ws.data_ptr(), ws.numel(), _stream_ptr()))
    return logits, pred


def eval_f16_supported(D: int, H: int, Cn: int) -> bool:
    return bool(_lib.load().dbmm_eval_f16_supported(D, H, Cn))


def eval_fwd_f16(X16: torch.Tensor, y: torch.Tensor, grp, ad: AdapterTensors, That: torch.Tensor, inv_tau: float,
                 stats: BatchStatsBuffers | None, batch_size: int, *, old_ad: AdapterTensors | None = None,
                 ebd_weight: float = 0.5, G: int = 4, want_logits=False, want_pred=False):
    """validate()/validate_zs() forward over the fp16-resident copy of the embeddings (dbmm_eval_fwd_f16): same results as
    eval_fwd when the embeddings are fp16-valued, at half the bytes and twice the tensor-core rate."""
    lib = _lib.load()
    _check(X16, torch.float16, "X16", contiguous=False)
    if X16.stride(1) != 1:
        raise DbmmError("X16 rows must be contiguous")
    _check(That, torch.float32, "That")
    N, D, H, Cn = X16.shape[0], X16.shape[1], ad.H, That.shape[1]
    if That.shape[0] != D or ad.D != D:
        raise DbmmError("dimension mismatch between X, adapter and text prompts")
    G = _label_args(y, grp, G)
    nad = 2 if old_ad is not None else 1
    ws = workspace(lib.dbmm_eval_f16_workspace_bytes(max(N, 1), D, H, Cn, nad), X16.device)
    logits = torch.empty((N, Cn), dtype=torch.float32, device=X16.device) if want_logits else None
    pred = torch.empty((N,), dtype=torch.int32, device=X16.device) if want_pred else None
    st = stats.c() if stats is not None else BatchStats(None, None)
    old_p = old_ad.ptrs() if old_ad is not None else None
    _lib.check(lib.dbmm_eval_fwd_f16(X16.data_ptr(), X16.stride(0), _ptr(y), _ptr(grp), N, D, H, Cn, G,
                                     C.byref(old_p) if old_p is not None else None, C.byref(ad.ptrs()), ebd_weight,
                                     That.data_ptr(), inv_tau, batch_size, st, _ptr(logits), _ptr(pred),
                                     ws.data_ptr(), ws.numel(), _stream_ptr()))
    return logits, pred


class TrainBuffers:
    """Flat gradient + momentum of one optimizer (torch.optim.SGD state, demo/util.py:118-136)."""

    def __init__(self, D: int, H: int, device="cuda"):
        n = param_count(D, H)
        self.grads = torch.zeros(n, dtype=torch.float32, device=device)
        self.momentum = torch.zeros(n, dtype=torch.float32, device=device)
        self.first_step = True


def train_step(X, y, grp, ad: AdapterTensors, That, inv_tau, buf: TrainBuffers, lr: float, stats: BatchStatsBuffers,
               slot: int = 0, *, idx=None, n_rows=None, B_global=None, old_ad=None, ebd_weight=0.5, G=4,
               momentum=0.9, weight_decay=5e-5, phases=_lib.PHASE_ALL, fresh=True, lr_dev=None, first_step=None):
    """One iteration of the train loops' body (final_main.py:455-466 / 610-623).  `fresh` / `lr_dev`: see
    dbmm_train_step_ex (steps chained by the caller, learning rate read from device memory)."""
    lib = _lib.load()
    _check(X, torch.float32, "X", contiguous=False)
    _check(That, torch.float32, "That")
    D, H, Cn = X.shape[1], ad.H, That.shape[1]
    G = _label_args(y, grp, G)
    if idx is not None:
        _check(idx, torch.int32, "idx")
    B = int(n_rows if n_rows is not None else (idx.numel() if idx is not None else X.shape[0]))
    Bg = int(B_global if B_global is not None else B)
    nad = 2 if old_ad is not None else 1
    ws = workspace(lib.dbmm_workspace_bytes(_lib.OP_TRAIN, B, D, H, Cn, nad), X.device)
    old_p = old_ad.ptrs() if old_ad is not None else None
    first = buf.first_step if first_step is None else first_step
    _lib.check(lib.dbmm_train_step_ex(phases, 1 if fresh else 0, X.data_ptr(), X.stride(0), _ptr(idx), y.data_ptr(), _ptr(grp),
                                      B, Bg, D, H, Cn, G, C.byref(old_p) if old_p is not None else None, C.byref(ad.ptrs()),
                                      ebd_weight, That.data_ptr(), inv_tau, buf.grads.data_ptr(), buf.momentum.data_ptr(),
                                      lr, _ptr(lr_dev), momentum, weight_decay, 1 if first else 0, stats.c(), slot,
                                      ws.data_ptr(), ws.numel(), _stream_ptr()))
    if phases & _lib.PHASE_UPDATE and first_step is None:
        buf.first_step = False


def train_forward(X, ad: AdapterTensors, That, inv_tau, *, old_ad=None, ebd_weight=0.5, update_running_stats=True):
    """Train-mode logits [B, C] (batch-statistics BatchNorm; running statistics moved as torch does): the forward half of
    the nn.Module boundary, final_main.py:66-80 / 121-140 under classifier.train()."""
    lib = _lib.load()
    _check(X, torch.float32, "X", contiguous=False)
    _check(That, torch.float32, "That")
    if X.stride(1) != 1:
        raise DbmmError("X rows must be contiguous")
    B, D, H, Cn = X.shape[0], X.shape[1], ad.H, That.shape[1]
    nad = 2 if old_ad is not None else 1
    ws = workspace(lib.dbmm_workspace_bytes(_lib.OP_TRAIN, B, D, H, Cn, nad), X.device)
    logits = torch.empty((B, Cn), dtype=torch.float32, device=X.device)
    old_p = old_ad.ptrs() if old_ad is not None else None
    _lib.check(lib.dbmm_train_forward(X.data_ptr(), X.stride(0), None, B, D, H, Cn, C.byref(old_p) if old_p is not None else None,
                                      C.byref(ad.ptrs()), ebd_weight, That.data_ptr(), inv_tau, logits.data_ptr(),
                                      1 if update_running_stats else 0, ws.data_ptr(), ws.numel(), _stream_ptr()))
    return logits


def train_backward(X, ad: AdapterTensors, That, inv_tau, dlogits, *, old_ad=None, ebd_weight=0.5):
    """Flat gradient (W1 | b1 | gamma | beta | W2 | b2) of the trainable adapter from dL/dlogits [B, C]."""
    lib = _lib.load()
    _check(dlogits, torch.float32, "dlogits")
    B, D, H, Cn = X.shape[0], X.shape[1], ad.H, That.shape[1]
    nad = 2 if old_ad is not None else 1
    ws = workspace(lib.dbmm_workspace_bytes(_lib.OP_TRAIN, B, D, H, Cn, nad), X.device)
    grads = torch.empty(param_count(D, H), dtype=torch.float32, device=X.device)
    old_p = old_ad.ptrs() if old_ad is not None else None
    _lib.check(lib.dbmm_train_backward(X.data_ptr(), X.stride(0), None, B, D, H, Cn, C.byref(old_p) if old_p is not None else None,
                                       C.byref(ad.ptrs()), ebd_weight, That.data_ptr(), inv_tau, dlogits.data_ptr(), grads.data_ptr(),
                                       ws.data_ptr(), ws.numel(), _stream_ptr()))
    return grads


def train_workspace_bytes(batch_size: int, D: int, H: int, Cn: int, nad: int = 1) -> int:
    return int(_lib.load().dbmm_workspace_bytes(_lib.OP_TRAIN, batch_size, D, H, Cn, nad))


def train_epoch(X, order: torch.Tensor, batch_size: int, y, grp, ad: AdapterTensors, That, inv_tau, buf: TrainBuffers,
                lrs, stats: BatchStatsBuffers, *, old_ad=None, ebd_weight=0.5, G=4, momentum=0.9, weight_decay=5e-5, ws=None):
    """All steps of one epoch in a single C call (train_one_epoch / train_reg_seq_one_epoch loops).
    ws: caller-owned workspace (uint8 CUDA tensor of train_workspace_bytes) -- required when several members train
    concurrently on different streams; default: the shared per-device scratch buffer."""
    lib = _lib.load()
    _check(X, torch.float32, "X", contiguous=False)
    _check(order, torch.int32, "order")
    _check(That, torch.float32, "That")
    D, H, Cn = X.shape[1], ad.H, That.shape[1]
    G = _label_args(y, grp, G)
    n = order.numel()
    steps = (n + batch_size - 1) // batch_size
    lrs = np.ascontiguousarray(lrs, dtype=np.float32)
    if len(lrs) != steps or stats.n_slots < steps:
        raise DbmmError(f"need {steps} learning rates / stat slots, got {len(lrs)} / {stats.n_slots}")
    nad = 2 if old_ad is not None else 1
    if ws is None:
        ws = workspace(lib.dbmm_workspace_bytes(_lib.OP_TRAIN, min(batch_size, n), D, H, Cn, nad), X.device)
    old_p = old_ad.ptrs() if old_ad is not None else None
    _lib.check(lib.dbmm_train_epoch(X.data_ptr(), X.stride(0), order.data_ptr(), n, batch_size, y.data_ptr(), _ptr(grp),
                                    D, H, Cn, G, C.byref(old_p) if old_p is not None else None, C.byref(ad.ptrs()),
                                    ebd_weight, That.data_ptr(), inv_tau, buf.grads.data_ptr(), buf.momentum.data_ptr(),
                                    lrs.ctypes.data_as(C.POINTER(C.c_float)), momentum, weight_decay,
                                    1 if buf.first_step else 0, stats.c(), ws.data_ptr(), ws.numel(), _stream_ptr()))
    buf.first_step = False
    return steps


@dataclass
class SweepMember:
    """One member of a batched sweep epoch (dbmm_member): its own batch order, adapters, optimizer state, statistics slots
    and workspace; the data, batch size and prompts are shared with the other members."""
    order: torch.Tensor                     # int32 [n_rows] on the device
    ad: AdapterTensors
    buf: "TrainBuffers"
    stats: "BatchStatsBuffers"
    lrs: np.ndarray                         # [steps]
    old_ad: AdapterTensors | None = None
    ws: torch.Tensor | None = None          # uint8 workspace (allocated on first use, kept by the member)


_batched_ws: dict = {}


def train_epoch_batched(X, members, batch_size: int, y, grp, That, inv_tau, *, ebd_weight=0.5, G=4, momentum=0.9, weight_decay=5e-5):
    """One epoch of every member in lock step (dbmm_train_epoch_batched; the reference runs its sweep members one after
    another, run_multiple/final_main_iteration_wb.py:1129-1197).  All members must share n_rows, the stage (old_ad or not)
    and `first_step`."""
    lib = _lib.load()
    _check(X, torch.float32, "X", contiguous=False)
    _check(That, torch.float32, "That")
    M = len(members)
    if M < 1:
        raise DbmmError("train_epoch_batched: no members")
    D, H, Cn = X.shape[1], members[0].ad.H, That.shape[1]
    G = _label_args(y, grp, G)
    n = members[0].order.numel()
    steps = (n + batch_size - 1) // batch_size
    nad = 2 if members[0].old_ad is not None else 1
    first = members[0].buf.first_step
    need = int(lib.dbmm_workspace_bytes(_lib.OP_TRAIN, min(batch_size, n), D, H, Cn, nad))
    arr = (_lib.Member * M)()
    keep = []
    lr_all = np.empty((M, steps), dtype=np.float32)
    for i, mb in enumerate(members):
        _check(mb.order, torch.int32, "order")
        if mb.order.numel() != n or (mb.old_ad is not None) != (nad == 2) or mb.buf.first_step != first or mb.stats.n_slots < steps:
            raise DbmmError("train_epoch_batched: members differ in n_rows / stage / first_step, or too few stat slots")
        if mb.ws is None or mb.ws.numel() < need:
            mb.ws = torch.empty(need, dtype=torch.uint8, device=X.device)
        lr = np.asarray(mb.lrs, dtype=np.float32)
        if len(lr) != steps:
            raise DbmmError(f"member {i}: need {steps} learning rates, got {len(lr)}")
        lr_all[i] = lr
        ap = mb.ad.ptrs()
        op = mb.old_ad.ptrs() if mb.old_ad is not None else None
        keep += [ap, op]
        arr[i].order = mb.order.data_ptr()
        arr[i].old_ad = C.pointer(op) if op is not None else None
        arr[i].ad = C.pointer(ap)
        arr[i].grads = mb.buf.grads.data_ptr(); arr[i].momentum_buf = mb.buf.momentum.data_ptr()
        arr[i].stats = mb.stats.c()
        arr[i].ws = mb.ws.data_ptr(); arr[i].ws_bytes = mb.ws.numel()
    bbytes = int(lib.dbmm_batched_workspace_bytes(M, steps))
    key = (torch.device(X.device).index or 0, M, steps)
    bws = _batched_ws.get(key)
    if bws is None or bws.numel() < bbytes:
        bws = _batched_ws[key] = torch.empty(bbytes, dtype=torch.uint8, device=X.device)
    _lib.check(lib.dbmm_train_epoch_batched(M, arr, X.data_ptr(), X.stride(0), n, batch_size, y.data_ptr(), _ptr(grp), D, H, Cn, G,
                                            ebd_weight, That.data_ptr(), inv_tau, lr_all.ctypes.data_as(C.POINTER(C.c_float)),
                                            momentum, weight_decay, 1 if first else 0, bws.data_ptr(), bws.numel(), _stream_ptr()))
    for mb in members:
        mb.buf.first_step = False
    return steps


STEP_KERNELS = ("gemm1_tc", "reduce_stats", "rows_train", "wgrad_tc", "finalize_grads", "update")
STEP_KERNELS_FUSED = ("gemm1_tc", "reduce_stats", "rows_train", "wgrad_tc", "tail_w1", "tail_w2")     # fused step tail


def train_epoch_profile(X, order: torch.Tensor, batch_size: int, y, grp, ad: AdapterTensors, That, inv_tau, buf: TrainBuffers,
                        lrs, stats: BatchStatsBuffers, *, old_ad=None, ebd_weight=0.5, G=4, momentum=0.9, weight_decay=5e-5):
    """Measurement aid: one epoch with CUDA events between the kernels of every step -> {kernel: mean microseconds}."""
    lib = _lib.load()
    D, H, Cn = X.shape[1], ad.H, That.shape[1]
    G = _label_args(y, grp, G)
    n = order.numel()
    steps = (n + batch_size - 1) // batch_size
    lrs = np.ascontiguousarray(lrs, dtype=np.float32)
    nad = 2 if old_ad is not None else 1
    ws = workspace(lib.dbmm_workspace_bytes(_lib.OP_TRAIN, min(batch_size, n), D, H, Cn, nad), X.device)
    old_p = old_ad.ptrs() if old_ad is not None else None
    out = (C.c_float * 6)()
    _lib.check(lib.dbmm_train_epoch_profile(X.data_ptr(), X.stride(0), order.data_ptr(), n, batch_size, y.data_ptr(), _ptr(grp),
                                            D, H, Cn, G, C.byref(old_p) if old_p is not None else None, C.byref(ad.ptrs()),
                                            ebd_weight, That.data_ptr(), inv_tau, buf.grads.data_ptr(), buf.momentum.data_ptr(),
                                            lrs[:steps].ctypes.data_as(C.POINTER(C.c_float)), momentum, weight_decay, stats.c(),
                                            ws.data_ptr(), ws.numel(), _stream_ptr(), out))
    buf.first_step = False
    last = n - (steps - 1) * batch_size
    fused = lib.dbmm_train_tail_mode(min(batch_size, n), last, nad, D, H, Cn) > 0
    return dict(zip(STEP_KERNELS_FUSED if fused else STEP_KERNELS, [float(v) for v in out]))


def train_tail_mode(batch_size: int, last_batch: int, nad: int, D: int, H: int, Cn: int) -> int:
    """0: k_finalize_grads + k_update; 1: fused tail in line; 2: fused tail, W2 role on a second graph branch."""
    return int(_lib.load().dbmm_train_tail_mode(batch_size, last_batch, nad, D, H, Cn))


def sgd_step(p: torch.Tensor, g: torch.Tensor, v: torch.Tensor, lr, momentum=0.9, weight_decay=5e-5, first_step=False):
    """torch.optim.SGD on flat fp32 buffers (demo/util.py:118-136)."""
    lib = _lib.load()
    for t, n in ((p, "p"), (g, "g"), (v, "v")):
        _check(t, torch.float32, n)
    _lib.check(lib.dbmm_sgd_step(p.data_ptr(), g.data_ptr(), v.data_ptr(), p.numel(), lr, momentum, weight_decay,
                                 1 if first_step else 0, _stream_ptr()))


def export_embeddings(X: torch.Tensor, ad: AdapterTensors, *, old_ad=None, ebd_weight=0.5, normalize_single=False,
                      That_a=None, That_b=None, inv_tau=100.0, idx=None, n_rows=None):
    """Adapted embeddings [N, D] (+ logits against up to two prompt matrices) as validate_adapter_with_return forms them
    (demo/demo_visualization.ipynb:1117-1215): single adapter -> the un-normalised adapter output; with `old_ad` -> the
    MultipleAdapter mix of the two normalised outputs.  Returns (features, logits_a | None, logits_b | None)."""
    lib = _lib.load()
    _check(X, torch.float32, "X", contiguous=False)
    D, H = X.shape[1], ad.H
    N = int(n_rows if n_rows is not None else (idx.numel() if idx is not None else X.shape[0]))
    out = torch.empty((N, D), dtype=torch.float32, device=X.device)
    la = torch.empty((N, That_a.shape[1]), dtype=torch.float32, device=X.device) if That_a is not None else None
    lb = torch.empty((N, That_b.shape[1]), dtype=torch.float32, device=X.device) if That_b is not None else None
    nad = 2 if old_ad is not None else 1
    ws = workspace(lib.dbmm_workspace_bytes(_lib.OP_EVAL, max(N, 1), D, H, 1, nad), X.device)
    old_p = old_ad.ptrs() if old_ad is not None else None
    _lib.check(lib.dbmm_export_embeddings(X.data_ptr(), X.stride(0), _ptr(idx), N, D, H,
                                          C.byref(old_p) if old_p is not None else None, C.byref(ad.ptrs()), ebd_weight,
                                          1 if normalize_single else 0, _ptr(That_a), That_a.shape[1] if That_a is not None else 0,
                                          _ptr(That_b), That_b.shape[1] if That_b is not None else 0, inv_tau,
                                          out.data_ptr(), out.stride(0), _ptr(la), _ptr(lb), ws.data_ptr(), ws.numel(), _stream_ptr()))
    return out, la, lb


def widen_f16(src: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """fp16 rows [N, D] on the device -> fp32 rows (exact): the ingest step behind the fp16 packed store (pack.py)."""
    lib = _lib.load()
    if src.dtype != torch.float16 or src.dim() != 2 or not src.is_cuda or src.stride(1) != 1:
        raise DbmmError("widen_f16: a CUDA fp16 [N, D] tensor with unit column stride is required")
    if out is None:
        out = torch.empty(src.shape, dtype=torch.float32, device=src.device)
    if out.shape != src.shape or out.dtype != torch.float32 or out.stride(1) != 1 or out.device != src.device:
        raise DbmmError("widen_f16: output must be a CUDA fp32 tensor of the same shape")
    _lib.check(lib.dbmm_widen_f16(src.data_ptr(), src.stride(0), out.data_ptr(), out.stride(0), src.shape[0], src.shape[1],
                                  _stream_ptr()))
    return out


def group_counts(logits: torch.Tensor, y, grp, stats: BatchStatsBuffers, batch_size: int, G=4, want_pred=False):
    """update_dict (final_main.py:383-391) on given logits."""
    lib = _lib.load()
    _check(logits, torch.float32, "logits")
    G = _label_args(y, grp, G)
    N, Cn = logits.shape
    pred = torch.empty((N,), dtype=torch.int32, device=logits.device) if want_pred else None
    _lib.check(lib.dbmm_group_counts(logits.data_ptr(), y.data_ptr(), _ptr(grp), N, Cn, G, batch_size, stats.c(),
                                     _ptr(pred), _stream_ptr()))
    return pred


def logits_ce(U: torch.Tensor, y, grp, That: torch.Tensor, inv_tau: float, stats: BatchStatsBuffers | None, batch_size: int, *,
              idx=None, n_rows=None, G: int = 4, normalize_rows=True, want_pred=False, col_bias=None):
    """Zero-shot head on raw embeddings (validate_zs, final_main.py:757-768; BASELINE config 4): cosine logits against
    the prompt columns, CE, argmax and per-group counters without materialising the [N, C] logits."""
    lib = _lib.load()
    _check(U, torch.float32, "U", contiguous=False)
    if U.stride(1) != 1:
        raise DbmmError("U rows must be contiguous")
    _check(That, torch.float32, "That")
    D, Cn = U.shape[1], That.shape[1]
    if That.shape[0] != D:
        raise DbmmError("dimension mismatch between U and the text prompts")
    G = _label_args(y, grp, G)
    if idx is not None:
        _check(idx, torch.int32, "idx")
    N = int(n_rows if n_rows is not None else (idx.numel() if idx is not None else U.shape[0]))
    ws = workspace(lib.dbmm_head_workspace_bytes(max(N, 1), D, Cn, 1 if idx is not None else 0), U.device)
    pred = torch.empty((N,), dtype=torch.int32, device=U.device) if want_pred else None
    st = stats.c() if stats is not None else BatchStats(None, None)
    _lib.check(lib.dbmm_logits_ce(U.data_ptr(), U.stride(0), _ptr(idx), _ptr(y), _ptr(grp), N, D, Cn, G, That.data_ptr(),
                                  _ptr(col_bias), inv_tau, 1 if normalize_rows else 0, batch_size, st, _ptr(pred), ws.data_ptr(), ws.numel(), _stream_ptr()))
    return pred


def logits_ce_f16(U16: torch.Tensor, y, grp, That: torch.Tensor, inv_tau: float, stats: BatchStatsBuffers | None, batch_size: int, *,
                  G: int = 4, normalize_rows=True, want_pred=False, col_bias=None):
    """logits_ce over the fp16-resident copy of the embeddings (dbmm_logits_ce_f16: kind::f16 tensor-core head, rows read as
    stored).  Same loss / argmax / counters as logits_ce on the widened rows."""
    lib = _lib.load()
    _check(U16, torch.float16, "U16", contiguous=False)
    if U16.stride(1) != 1:
        raise DbmmError("U16 rows must be contiguous")
    _check(That, torch.float32, "That")
    N, D, Cn = U16.shape[0], U16.shape[1], That.shape[1]
    if That.shape[0] != D:
        raise DbmmError("dimension mismatch between U16 and the text prompts")
    G = _label_args(y, grp, G)
    ws = workspace(lib.dbmm_head_f16_workspace_bytes(max(N, 1), D, Cn), U16.device)
    pred = torch.empty((N,), dtype=torch.int32, device=U16.device) if want_pred else None
    st = stats.c() if stats is not None else BatchStats(None, None)
    _lib.check(lib.dbmm_logits_ce_f16(U16.data_ptr(), U16.stride(0), _ptr(y), _ptr(grp), N, D, Cn, G, That.data_ptr(), _ptr(col_bias),
                                      inv_tau, 1 if normalize_rows else 0, batch_size, st, _ptr(pred), ws.data_ptr(), ws.numel(),
                                      _stream_ptr()))
    return pred


def head_f16_supported(D: int) -> bool:
    return D >= 64 and D % 8 == 0


def contrastive_step(X, labels, ad: AdapterTensors, buf: "TrainBuffers", lr: float, *, idx=None, pre_norm=True, tau_cl=0.1,
                     loss_weight=0.1, momentum=0.9, weight_decay=5e-5, loss_out=None, n_valid_out=None):
    """One SGD step of `--tl_method contrastive_adapter` on the rows idx (or all rows) of X: u = L2(adapter(L2(x))), all-anchor
    supervised contrastive loss with the given labels, D-wide backward, SGD (dbmm_contrastive_step).  loss_out: float64 [1]
    device tensor the weighted mean loss is ADDED to."""
    lib = _lib.load()
    _check(X, torch.float32, "X", contiguous=False)
    _check(labels, torch.int32, "labels")
    if idx is not None:
        _check(idx, torch.int32, "idx")
    B = int(idx.numel() if idx is not None else X.shape[0])
    D, H = X.shape[1], ad.H
    ws = workspace(lib.dbmm_contrastive_workspace_bytes(B, D, H), X.device)
    _lib.check(lib.dbmm_contrastive_step(X.data_ptr(), X.stride(0), _ptr(idx), labels.data_ptr(), B, D, H, C.byref(ad.ptrs()),
                                         1 if pre_norm else 0, 1.0 / tau_cl, loss_weight, buf.grads.data_ptr(), buf.momentum.data_ptr(),
                                         lr, momentum, weight_decay, 1 if buf.first_step else 0, _ptr(loss_out), _ptr(n_valid_out),
                                         ws.data_ptr(), ws.numel(), _stream_ptr()))
    buf.first_step = False


def contrastive_forward(X, labels, ad: AdapterTensors, *, idx=None, pre_norm=True):
    """forward_ca of the rows (train-mode BatchNorm on THESE rows): returns (U [B, D] L2-normalised, labels of the rows [B],
    workspace handle to pass to contrastive_backward / contrastive_apply)."""
    lib = _lib.load()
    _check(X, torch.float32, "X", contiguous=False)
    _check(labels, torch.int32, "labels")
    B = int(idx.numel() if idx is not None else X.shape[0])
    D, H = X.shape[1], ad.H
    ws = torch.empty(int(lib.dbmm_contrastive_workspace_bytes(B, D, H)), dtype=torch.uint8, device=X.device)
    U = torch.empty((B, D), dtype=torch.float32, device=X.device)
    lab = torch.empty((B,), dtype=torch.int32, device=X.device)
    _lib.check(lib.dbmm_contrastive_forward(X.data_ptr(), X.stride(0), _ptr(idx), labels.data_ptr(), B, D, H, C.byref(ad.ptrs()),
                                            1 if pre_norm else 0, U.data_ptr(), lab.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()))
    return U, lab, ws


def contrastive_backward(X, ad: AdapterTensors, dU, grads, ws, *, idx=None, pre_norm=True, loss_weight=0.1):
    """Flat parameter gradient from dL/du (dU is overwritten); `ws` from contrastive_forward on the same rows."""
    lib = _lib.load()
    B, D, H = dU.shape[0], X.shape[1], ad.H
    _lib.check(lib.dbmm_contrastive_backward(X.data_ptr(), X.stride(0), _ptr(idx), B, D, H, C.byref(ad.ptrs()), 1 if pre_norm else 0,
                                             dU.data_ptr(), None, loss_weight, grads.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr()))


def contrastive_apply(ad: AdapterTensors, buf: "TrainBuffers", lr, B, ws, *, momentum=0.9, weight_decay=5e-5):
    lib = _lib.load()
    _lib.check(lib.dbmm_contrastive_apply(B, ad.D, ad.H, C.byref(ad.ptrs()), buf.grads.data_ptr(), buf.momentum.data_ptr(), lr, momentum,
                                          weight_decay, 1 if buf.first_step else 0, ws.data_ptr(), ws.numel(), _stream_ptr()))
    buf.first_step = False


class SupconState:
    """Device scalars of one contrastive step: sum of per-anchor losses and number of valid anchors."""

    def __init__(self, device="cuda"):
        self.loss_sum = torch.zeros(1, dtype=torch.float64, device=device)
        self.n_valid = torch.zeros(1, dtype=torch.int32, device=device)

    def zero_(self):
        self.loss_sum.zero_()
        self.n_valid.zero_()

    def loss(self) -> float:
        return float(self.loss_sum.item() / max(int(self.n_valid.item()), 1))


def supcon_fwd(Z_all: torch.Tensor, labels: torch.Tensor, state: SupconState, *, row0=0, n_local=None, tau_cl=0.1, want_row_loss=False):
    """All-anchor supervised contrastive loss (demo/visualizer_supcon.py:1532-1571 per anchor) of this rank's anchors
    [row0, row0 + n_local) against the global batch Z_all.  Adds into `state`; keeps the similarity gradient in the workspace."""
    lib = _lib.load()
    _check(Z_all, torch.float32, "Z_all")
    _check(labels, torch.int32, "labels")
    Bg, d = Z_all.shape
    Bl = int(n_local if n_local is not None else Bg - row0)
    ws = workspace(lib.dbmm_supcon_workspace_bytes(Bl, Bg, d), Z_all.device)
    row_loss = torch.empty(Bl, dtype=torch.float32, device=Z_all.device) if want_row_loss else None
    _lib.check(lib.dbmm_supcon_fwd(Z_all.data_ptr(), Bg, d, row0, Bl, labels.data_ptr(), 1.0 / tau_cl, state.loss_sum.data_ptr(),
                                   state.n_valid.data_ptr(), _ptr(row_loss), ws.data_ptr(), ws.numel(), _stream_ptr()))
    return row_loss


def supcon_bwd(Z_all: torch.Tensor, state: SupconState, *, row0=0, n_local=None, tau_cl=0.1, dZ_all=None, accumulate_all=False):
    """Gradient of the mean contrastive loss w.r.t. the normalised embeddings.  Returns (dZ_local [Bl, d], dZ_all [Bg, d]):
    anchor-role and contrast-role parts (on one GPU the gradient is their sum; under data parallelism dZ_all is
    reduce-scattered).  `state.n_valid` must already hold the GLOBAL number of valid anchors."""
    lib = _lib.load()
    Bg, d = Z_all.shape
    Bl = int(n_local if n_local is not None else Bg - row0)
    ws = workspace(lib.dbmm_supcon_workspace_bytes(Bl, Bg, d), Z_all.device)
    dZ_local = torch.empty((Bl, d), dtype=torch.float32, device=Z_all.device)
    if dZ_all is None:
        dZ_all = torch.empty((Bg, d), dtype=torch.float32, device=Z_all.device)
        accumulate_all = False
    _lib.check(lib.dbmm_supcon_bwd(Z_all.data_ptr(), Bg, d, row0, Bl, 1.0 / tau_cl, state.n_valid.data_ptr(), dZ_local.data_ptr(),
                                   dZ_all.data_ptr(), 1 if accumulate_all else 0, ws.data_ptr(), ws.numel(), _stream_ptr()))
    return dZ_local, dZ_all


def linear_train_epoch(X, order: torch.Tensor, batch_size: int, y, grp, W: torch.Tensor, b: torch.Tensor, grads: torch.Tensor,
                       momentum_buf: torch.Tensor, lrs, stats: BatchStatsBuffers, *, first_step=False, G=4, momentum=0.9,
                       weight_decay=5e-5):
    """One epoch of linear probing (LinearClassifier under train_one_epoch, final_main.py:43-49, 426-496)."""
    lib = _lib.load()
    _check(X, torch.float32, "X", contiguous=False)
    _check(order, torch.int32, "order")
    _check(W, torch.float32, "W"); _check(b, torch.float32, "b")
    Cn, D = W.shape
    G = _label_args(y, grp, G)
    n = order.numel()
    steps = (n + batch_size - 1) // batch_size
    lrs = np.ascontiguousarray(lrs, dtype=np.float32)
    if len(lrs) < steps or stats.n_slots < steps or grads.numel() < Cn * D + Cn or momentum_buf.numel() < Cn * D + Cn:
        raise DbmmError("linear_train_epoch: learning-rate table / stat slots / flat buffers too small")
    _lib.check(lib.dbmm_linear_train_epoch(X.data_ptr(), X.stride(0), order.data_ptr(), n, batch_size, y.data_ptr(), _ptr(grp),
                                           D, Cn, G, W.data_ptr(), b.data_ptr(), grads.data_ptr(), momentum_buf.data_ptr(),
                                           lrs.ctypes.data_as(C.POINTER(C.c_float)), momentum, weight_decay,
                                           1 if first_step else 0, stats.c(), _stream_ptr()))
    return steps


def linear_logits(X: torch.Tensor, Wt: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """LinearClassifier.forward (final_main.py:48-49): materialised logits through the tensor-core GEMM of the
    contrastive path would be overkill for C = 2; this is plumbing-sized, so it is the one place torch.addmm is used."""
    _check(X, torch.float32, "X")
    return torch.addmm(bias, X, Wt)
