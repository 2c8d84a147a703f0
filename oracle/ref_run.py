"""ORACLE (test infrastructure / CPU baseline).  Runs the UNMODIFIED reference modules of the hot path -- final_main.Adapter
/ CustomCLIP, demo.util.set_optimizer, nn.CrossEntropyLoss and the reference's own `train_one_epoch` loop with its meters
and host syncs (final_main.py:426-496) -- on synthetic embeddings, for bench.py's `--impl reference` arm, its
`cpu_baseline` leg and the "reference on the B200 under stock PyTorch eager" leg.

The reference is imported from /root/reference in the build container, or from the copy that oracle/stage_reference.py
placed under oracle/_ref/reference (git-ignored; it travels to the GPU box with the snapshot).  Nothing here is product code.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import tempfile
import time
import types
from functools import partial

import numpy as np
import torch

from .reference_shim import import_reference
from .stage_reference import staged_root

_ref = None


def available() -> bool:
    return staged_root() is not None


def reference():
    global _ref
    if _ref is None:
        root = staged_root()
        if root is None:
            raise RuntimeError("the reference is neither at /root/reference nor staged under oracle/_ref/reference")
        _ref = import_reference(root)
    return _ref


@contextlib.contextmanager
def on_host():
    """Keep the reference on the host cores of a GPU box: its unconditional `.cuda()` calls (final_main.py:62,64,447-448)
    become identities for the duration (shim 2)."""
    t_cuda, m_cuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_cuda, m_cuda


class _ListLoader(list):
    """Stands in for a DataLoader of pre-loaded batches: the reference's loops only need iteration, len() and
    .dataset.n_groups (final_main.py:436, 439)."""

    def __init__(self, batches, n_groups=4):
        super().__init__(batches)
        self.dataset = types.SimpleNamespace(n_groups=n_groups)


def _write_text(path, T):
    with open(path, "w") as f:
        json.dump({f"prompt {c}": [float(t) for t in T[:, c]] for c in range(T.shape[1])}, f)


def build_classifier(D, H, T_class, device="cpu", lr=0.1, seed=42):
    """final_main.set_model's adapter branch + demo.util.set_optimizer on in-memory prompts (written to temp JSON files,
    the only form CustomCLIP accepts)."""
    fm, ru = reference()
    tmp = tempfile.mkdtemp(prefix="dbmm_ref_")
    paths = [os.path.join(tmp, n) for n in ("class.json", "spurious.json", "group.json")]
    for p in paths:
        _write_text(p, T_class)
    torch.manual_seed(seed)
    ctx = on_host() if device == "cpu" else contextlib.nullcontext()
    with ctx:
        clf = fm.CustomCLIP(fm.Adapter(D, H), *paths, temperature=0.01)
        if device != "cpu":
            clf = clf.cuda()
    opt = types.SimpleNamespace(learning_rate=lr, learning_rate_reg=lr, momentum=0.9, weight_decay=5e-5, warm=False,
                                watch_batch_results=False, print_freq=10)
    optim = ru.set_optimizer(opt, clf)
    return fm, clf, opt, optim


def time_train_epochs(X, y, g, T_class, H, batch_size, n_batches, device="cpu", lr=0.1, epochs=1, threads=None):
    """`epochs` calls of the reference's train_one_epoch over `n_batches` pre-loaded batches (compute-only number of
    BASELINE.md section 2: no pandas / DataLoader).  device="cuda": the same code under stock PyTorch eager on the GPU,
    batches already resident -- the incumbent the kernels replace."""
    if threads:
        torch.set_num_threads(threads)
    D = X.shape[1]
    fm, clf, opt, optim = build_classifier(D, H, T_class, device=device, lr=lr)
    crit = torch.nn.CrossEntropyLoss()
    dev = torch.device(device)
    Xt, yt, gt = torch.from_numpy(X).to(dev), torch.from_numpy(y).long().to(dev), torch.from_numpy(g).long().to(dev)
    n = X.shape[0]
    batches = []
    for s in range(n_batches):
        lo = (s * batch_size) % max(n - batch_size + 1, 1)
        sl = slice(lo, lo + batch_size)
        batches.append((Xt[sl], {"class": yt[sl], "group": gt[sl]}, None))
    loader = _ListLoader(batches)
    get_yp = partial(fm.get_y_p, n_places=2)
    ctx = on_host() if device == "cpu" else contextlib.nullcontext()
    rows = 0
    with ctx, contextlib.redirect_stdout(io.StringIO()):
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        for ep in range(epochs):
            loss, acc, group_acc = fm.train_one_epoch(opt, loader, clf, crit, optim, ep + 1, get_yp, target="class")
            rows += sum(len(b[1]["class"]) for b in batches)
        if device != "cpu":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return dict(seconds=dt, rows=rows, emb_per_s=rows / dt, threads=torch.get_num_threads(), loss=float(loss),
                worst_acc=float(group_acc["worst_acc"]))


def time_dataloader_epoch(n_rows=4795, dim=1024, H=128, batch_size=1024, num_workers=None, seed=1234):
    """The user-visible number of BASELINE.md section 2 (i): the reference's own dataset class (pandas read_json of the
    embedding file, per-item DataFrame lookups) and DataLoader feeding its train_one_epoch, on a Waterbirds-shaped
    synthetic embedding file written in the reference's format.  Returns dataset-construction and epoch seconds."""
    import importlib
    synth = importlib.import_module("debiasing-multi-modal_b200.synth")
    fm, ru = reference()
    root = tempfile.mkdtemp(prefix="dbmm_ref_dl_")
    scale = n_rows / 4795.0
    ds = synth.make_dataset(name="waterbirds", dim=dim, seed=seed, scale=scale)
    paths = synth.write_reference_files(ds, root)
    from data.waterbirds_embeddings import load_waterbirds_embeddings
    if num_workers is None:
        num_workers = min(16, os.cpu_count() or 1)
    t0 = time.perf_counter()
    train_loader, _, _ = load_waterbirds_embeddings(paths["data_dir"], paths["image_embedding_dir"], bs_train=batch_size,
                                                     bs_val=batch_size, num_workers=num_workers)
    t_build = time.perf_counter() - t0
    with open(paths["text_embedding_dir"]) as f:
        tab = json.load(f)
    T = np.stack([np.asarray(v, np.float32) for v in tab.values()], 1)
    fm2, clf, opt, optim = build_classifier(dim, H, T, device="cpu", lr=1.0)
    crit = torch.nn.CrossEntropyLoss()
    get_yp = partial(fm.get_y_p, n_places=2)
    with on_host(), contextlib.redirect_stdout(io.StringIO()):
        t0 = time.perf_counter()
        fm.train_one_epoch(opt, train_loader, clf, crit, optim, 1, get_yp, target="class")
        t_epoch = time.perf_counter() - t0
    n = len(train_loader.dataset)
    return dict(rows=n, build_seconds=t_build, epoch_seconds=t_epoch, emb_per_s_epoch=n / t_epoch,
                emb_per_s_with_build=n / (t_epoch + t_build / 3.0), num_workers=num_workers, threads=torch.get_num_threads(),
                note="build_seconds = three Dataset instances (train / val / test), each parsing the whole JSON "
                     "(final_main.py:819-821); emb_per_s_with_build charges the train split's third of it to one epoch")
