"""Epoch loops of the adapter path with the reference's function names and return values
(final_main.py:346-379, 426-803), rebuilt around one C call per epoch.

Per epoch the reference issues ~40 tiny ATen kernels and 6-10 host syncs per batch; here an epoch is
(1) draw the batch order (same RNG protocol as DataLoader), (2) one `dbmm_train_epoch` / `dbmm_eval_fwd`
call that runs every batch on the device, (3) one device->host read of the per-batch loss sums and
per-group counters, replayed through the reference's meter arithmetic on the host (metrics.py).
"""
from __future__ import annotations

import os

import sys
import time

import numpy as np
import torch

from . import ops
from .data import EmbeddingLoader, Subset, resolve
from .metrics import eval_group_acc, get_y_p, replay_epoch, train_group_acc  # noqa: F401
from .optim import warmup_learning_rate, warmup_learning_rate_reg


def set_seed(seed):
    """demo/util.py:61-68."""
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True


def _n_groups(loader):
    base, _ = resolve(loader.dataset)
    return base.n_groups


def _check_criterion(criterion):
    if criterion is not None and not isinstance(criterion, torch.nn.CrossEntropyLoss):
        raise NotImplementedError("the fused step implements nn.CrossEntropyLoss() (mean), the reference's criterion")


_persistent: dict = {}
_member_tag = None          # set by the lock-step sweep driver (sweep.py): members keep separate order / statistics buffers


def _order_to_device(base, rows):
    """Batch order on the device, in a per-(device, length) buffer that keeps its address from epoch to epoch: the
    epoch's CUDA graph inside libdbmm is cached by argument addresses."""
    n = len(rows)
    key = ("order", _member_tag, str(base.device), n)
    buf = _persistent.get(key)
    if buf is None:
        buf = _persistent[key] = torch.empty(n, dtype=torch.int32, device=base.device)
    buf.copy_(torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int32)))
    return buf


def _stats_buffers(n_slots, n_groups, device, tag):
    key = ("stats", _member_tag, tag, str(device), n_slots, n_groups)
    st = _persistent.get(key)
    if st is None:
        st = _persistent[key] = ops.BatchStatsBuffers(n_slots, n_groups, device=device)
    else:
        st.zero_()
    return st


def _batch_sizes(n, bs):
    return [min(bs, n - s) for s in range(0, n, bs)]


def _lr_table(n_batches, optimizer, warm_fn):
    """Learning rate of every step of the epoch: call the per-batch warm-up hook exactly as the reference's
    loop does and record what it leaves in the optimizer."""
    lrs = []
    for idx in range(n_batches):
        warm_fn(idx)
        lrs.append(optimizer.param_groups[0]["lr"])
    return np.asarray(lrs, dtype=np.float64)


def _print_batches(opt, print_label, epoch, n_batches, losses_b, accs_b, elapsed, with_groups=None):
    if not getattr(opt, "watch_batch_results", False):
        return
    bt = elapsed / max(n_batches, 1)
    for idx in range(n_batches):
        if (idx + 1) % opt.print_freq == 0:
            line = (f"{print_label}: [{epoch}][{idx + 1}/{n_batches}]\tBT {bt:.3f} ({bt:.3f})\tDT 0.000 (0.000)\t"
                    f"loss {losses_b[idx][0]:.3f} ({losses_b[idx][1]:.3f})\tAcc@1 {accs_b[idx][0]:.3f} ({accs_b[idx][1]:.3f})")
            if with_groups is not None:
                line += f"\tGroup Acc {with_groups[idx]}"
            print(line)
    sys.stdout.flush()


class TrainJob:
    """One training epoch of one model, prepared (batch order drawn, learning-rate table built, buffers chosen) but not yet
    run.  The epoch functions below are generators that yield their job: `drive` runs it on the spot (the stand-alone
    path, identical to calling dbmm_train_epoch directly); the lock-step sweep driver (sweep.py) collects the jobs of all
    members and runs those with the same signature as ONE dbmm_train_epoch_batched call."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def signature(self):
        """Jobs with equal signatures share data, shapes, prompts and hyper-parameters and can run in lock step."""
        if self.linear:
            return None
        return (self.base.x.data_ptr(), self.labels.data_ptr(), self.n, self.bs, self.old is None, self.prompt_key,
                float(self.inv_tau), float(self.w), float(self.optimizer.momentum), float(self.optimizer.weight_decay),
                bool(self.optimizer.buffers.first_step), self.ad.D, self.ad.H)

    def member(self):
        return ops.SweepMember(order=self.order, ad=self.ad, buf=self.optimizer.buffers, stats=self.stats, lrs=self.lrs,
                               old_ad=self.old, ws=getattr(self.optimizer, "_sweep_ws", None))

    def run(self):
        base = self.base
        if self.linear:
            c, o = self.classifier, self.optimizer
            ops.linear_train_epoch(base.x, self.order, self.bs, self.labels, base.labels["group"], c.fc.weight.data, c.fc.bias.data,
                                   o.grads, o.momentum_buf, self.lrs, self.stats, first_step=o.first_step, G=base.n_groups,
                                   momentum=o.momentum, weight_decay=o.weight_decay)
            o.first_step = False
            return
        ops.train_epoch(base.x, self.order, self.bs, self.labels, base.labels["group"], self.ad, self.That, self.inv_tau,
                        self.optimizer.buffers, self.lrs, self.stats, old_ad=self.old, ebd_weight=self.w, G=base.n_groups,
                        momentum=self.optimizer.momentum, weight_decay=self.optimizer.weight_decay)


def run_jobs_batched(jobs):
    """Jobs of equal signature() in one lock-step call; their statistics land in each job's own buffers."""
    j0 = jobs[0]
    members = [j.member() for j in jobs]
    ops.train_epoch_batched(j0.base.x, members, j0.bs, j0.labels, j0.base.labels["group"], j0.That, j0.inv_tau, ebd_weight=j0.w,
                            G=j0.base.n_groups, momentum=j0.optimizer.momentum, weight_decay=j0.optimizer.weight_decay)
    for j, m in zip(jobs, members):
        j.optimizer._sweep_ws = m.ws              # the member keeps its workspace (and so its epoch-graph key) across epochs


def drive(gen):
    """Run an epoch / training generator to completion, executing every TrainJob it yields immediately."""
    try:
        job = next(gen)
        while True:
            job.run()
            job = gen.send(None)
    except StopIteration as e:
        return e.value


def _run_train_epoch(opt, loader, classifier, optimizer, target, use_group, warm_fn, get_yp_func, print_label, epoch,
                     count_metrics=True):
    """Generator: prepares the epoch, yields its TrainJob, then reads the statistics back."""
    classifier.train()
    base, rows = loader.base_rows(loader.draw_order())
    n = len(rows)
    bs = loader.batch_size
    sizes = _batch_sizes(n, bs)
    lrs = _lr_table(len(sizes), optimizer, warm_fn)
    stats = _stats_buffers(len(sizes), base.n_groups, base.device, "train")
    t0 = time.time()
    if hasattr(classifier, "fc"):                                 # linear probing (final_main.py:43-49)
        job = TrainJob(linear=True, base=base, order=_order_to_device(base, rows), n=n, bs=bs, lrs=lrs, stats=stats,
                       labels=base.labels[target], classifier=classifier, optimizer=optimizer)
    else:
        old, ad, w = classifier.kernel_adapters()
        job = TrainJob(linear=False, base=base, order=_order_to_device(base, rows), n=n, bs=bs, lrs=lrs, stats=stats,
                       labels=base.labels["group"] if use_group else base.labels[target], classifier=classifier,
                       optimizer=optimizer, old=old, ad=ad, w=w, That=classifier.prompt_matrix(use_group=use_group),
                       prompt_key=(classifier.text_group_embedding_dir, "group") if use_group else (classifier.text_embedding_dir, "class"),
                       inv_tau=1.0 / classifier.temperature)
    yield job
    loss_sum, counts = stats.host()
    return loss_sum, counts, sizes, time.time() - t0


def train_one_epoch_gen(opt, train_loader, classifier, criterion, optimizer, epoch, get_yp_func, target,
                        print_label='Train', predict_group=True):
    """Stage-1 / plain adapter epoch (final_main.py:426-496), as a generator that yields its TrainJob."""
    _check_criterion(criterion)
    loss_sum, counts, sizes, elapsed = yield from _run_train_epoch(
        opt, train_loader, classifier, optimizer, target, False,
        lambda idx: warmup_learning_rate(opt, epoch, idx, len(train_loader), optimizer), get_yp_func, print_label, epoch)
    losses, acc, acc_groups = replay_epoch(loss_sum, counts, sizes, _n_groups(train_loader))
    group_acc = train_group_acc(acc_groups, get_yp_func)
    print(f"{print_label}:", str(group_acc))
    return losses.avg, acc.avg, group_acc


def train_one_epoch(*a, **k):
    """final_main.train_one_epoch (426-496): same arguments and return value."""
    return drive(train_one_epoch_gen(*a, **k))


def train_reg_seq_one_epoch_gen(opt, train_loader, classifier, criterion, optimizer, epoch, get_yp_func, target,
                                print_label='Train', predict_group=True, use_group=False):
    """Stage-2 epoch on the held-out regularisation split (final_main.py:571-653): labels are the group ids when
    `use_group`, and the warm-up hook is the `_reg` one with the stage-relative epoch."""
    _check_criterion(criterion)
    rel_epoch = epoch - opt.epochs_feature_learning
    loss_sum, counts, sizes, elapsed = yield from _run_train_epoch(
        opt, train_loader, classifier, optimizer, target, use_group,
        lambda idx: warmup_learning_rate_reg(opt, rel_epoch, idx, len(train_loader), optimizer),
        get_yp_func, print_label, epoch)
    losses, acc, acc_groups = replay_epoch(loss_sum, counts, sizes, _n_groups(train_loader))
    group_acc = train_group_acc(acc_groups, get_yp_func)
    print(f"{print_label}:", str(group_acc))
    return losses.avg, acc.avg, group_acc


def train_reg_seq_one_epoch(*a, **k):
    """final_main.train_reg_seq_one_epoch (571-653): same arguments and return value."""
    return drive(train_reg_seq_one_epoch_gen(*a, **k))


def train_reg_one_epoch_gen(opt, train_loader1, train_loader2, classifier, criterion, optimizer, epoch, get_yp_func, target,
                            group_prompt=True, print_label='Train'):
    """`adapter_reg` epoch (final_main.py:498-569): a pass over the train loader with class prompts, then a pass
    over the reg loader with group (or class) prompts, one optimizer; passes with class prompts feed the meters."""
    _check_criterion(criterion)
    n_groups = _n_groups(train_loader1)
    merged = None
    for loader, use_group in ((train_loader1, False), (train_loader2, group_prompt)):
        loss_sum, counts, sizes, _ = yield from _run_train_epoch(
            opt, loader, classifier, optimizer, target, use_group is True,
            lambda idx, L=loader: warmup_learning_rate(opt, epoch, idx, len(L), optimizer), get_yp_func, print_label, epoch)
        if use_group is False:
            part = (loss_sum, counts, sizes)
            merged = part if merged is None else (np.concatenate([merged[0], part[0]]),
                                                  np.concatenate([merged[1], part[1]]), merged[2] + part[2])
    losses, acc, acc_groups = replay_epoch(merged[0], merged[1], merged[2], n_groups)
    group_acc = train_group_acc(acc_groups, get_yp_func)
    print(f"{print_label}:", str(group_acc))
    return losses.avg, acc.avg, group_acc


def train_reg_one_epoch(*a, **k):
    """final_main.train_reg_one_epoch (498-569): same arguments and return value."""
    return drive(train_reg_one_epoch_gen(*a, **k))


def train_one_epoch_cl(opt, train_loader, batches, classifier, optimizer, epoch, print_label='Train'):
    """Contrastive epoch (demo/visualizer_supcon.py:412-508): one optimizer step per contrastive loader batch (batch_factor
    anchor groups), the same per-batch warm-up hook; the batch is scored by ONE all-anchor B x B contraction with the rows'
    class labels (dbmm_contrastive_step) instead of the reference's Python loop over single anchors.  Returns the mean loss."""
    classifier.train()
    base, _ = train_loader.base_rows(np.arange(1))
    ad = classifier.adapter.tensors()
    loss = torch.zeros(1, dtype=torch.float64, device=base.device)
    n_used = 0
    t0 = time.time()
    for idx, rows in enumerate(batches):
        if idx >= opt.ca_update:
            continue
        warmup_learning_rate(opt, epoch, idx, len(batches), optimizer)
        ops.contrastive_step(base.x, base.labels["class"], ad, optimizer.buffers, optimizer.lr, idx=_order_to_device(base, rows),
                             pre_norm=not opt.no_ca_pre_norm, tau_cl=opt.cl_temperature, loss_weight=opt.contrastive_weight,
                             momentum=optimizer.momentum, weight_decay=optimizer.weight_decay, loss_out=loss)
        n_used += 1
    avg = float(loss.item()) / max(n_used, 1)
    print(f"Loss in {print_label}: {avg:.3f} ({n_used} batches, {time.time() - t0:.2f} s)")
    return avg


def _run_eval(loader, classifier, target, spurious_prompts=False):
    classifier.eval()
    base, rows = loader.base_rows(loader.draw_order())
    n, bs = len(rows), loader.batch_size
    sizes = _batch_sizes(n, bs)
    stats = _stats_buffers(len(sizes), base.n_groups, base.device, "eval")
    contiguous = len(rows) == len(base) and np.array_equal(rows, np.arange(len(base)))
    idx = None if contiguous else _order_to_device(base, rows)
    if hasattr(classifier, "fc"):                                 # linear probe: logits = x W^T + b on the raw embeddings
        ops.logits_ce(base.x, base.labels[target], base.labels["group"], classifier.fc.weight.data.t().contiguous(), 1.0, stats, bs,
                      idx=idx, n_rows=n, G=base.n_groups, normalize_rows=False, col_bias=classifier.fc.bias.data)
        loss_sum, counts = stats.host()
        return loss_sum, counts, sizes, base.n_groups
    old, ad, w = classifier.kernel_adapters()
    That = classifier.prompt_matrix(spurious=spurious_prompts)
    if getattr(base, "x16", None) is not None and n > 0 and ops.eval_f16_supported(base.x16.shape[1], ad.H, That.shape[1]) \
            and os.environ.get("DBMM_EVAL_F16", "1") != "0":
        # fp16-resident rows through the kind::f16 kernels; a Subset (the val half of the reference's stratified split) is
        # gathered once into its own contiguous copy and kept with the loader
        if contiguous:
            x16, yy, gg = base.x16, base.labels[target], base.labels["group"]
        else:
            cache = loader.__dict__.setdefault("_f16_rows", {})
            key = (hash(rows.tobytes()), target)
            if key not in cache:
                il = idx.long()
                cache[key] = (base.x16[il].contiguous(), base.labels[target][il].contiguous(), base.labels["group"][il].contiguous())
            x16, yy, gg = cache[key]
        ops.eval_fwd_f16(x16, yy, gg, ad, That, 1.0 / classifier.temperature, stats, bs, old_ad=old, ebd_weight=w, G=base.n_groups)
    else:
        ops.eval_fwd(base.x, base.labels[target], base.labels["group"], ad, That, 1.0 / classifier.temperature, stats, bs,
                     idx=idx, n_rows=n, old_ad=old, ebd_weight=w, G=base.n_groups)
    loss_sum, counts = stats.host()
    return loss_sum, counts, sizes, base.n_groups


def validate(opt, val_loader, classifier, criterion, get_yp_func, train_group_ratio, target, print_label='Test'):
    """Eval-mode pass with group meters and the train-ratio-weighted mean (final_main.py:655-719)."""
    _check_criterion(criterion)
    loss_sum, counts, sizes, n_groups = _run_eval(val_loader, classifier, target)
    losses, acc, acc_groups = replay_epoch(loss_sum, counts, sizes, n_groups)
    group_acc = eval_group_acc(acc_groups, get_yp_func, train_group_ratio)
    print(f"{print_label}:", str(group_acc))
    return losses.avg, acc.avg, group_acc


def validate_adapter_with_return(opt, val_loader, classifier, criterion, get_yp_func, train_group_ratio, target, print_label='Test'):
    """Validation that also returns the adapted embeddings for the visualisation notebooks
    (demo/demo_visualization.ipynb:1117-1215): same return value,
        (None, acc.avg, group_acc), (total_embeddings [N, D] numpy, {"targets", "spuriouss", "groups", "predictions",
                                                                     "predictions_spurious"})
    with the notebook's scoring rule -- logits = features @ That / temperature on the EXPORTED features, which for
    tl_method == "adapter" are the un-normalised adapter outputs (unlike validate())."""
    if "adapter" not in opt.tl_method:
        raise AssertionError("validate_adapter_with_return is defined for the adapter methods")
    classifier.eval()
    base, rows = val_loader.base_rows(val_loader.draw_order())
    n, bs = len(rows), val_loader.batch_size
    sizes = _batch_sizes(n, bs)
    contiguous = len(rows) == len(base) and np.array_equal(rows, np.arange(len(base)))
    idx = None if contiguous else _order_to_device(base, rows)
    old, ad, w = classifier.kernel_adapters()
    if opt.tl_method == "adapter":
        old = None
    feats, logits, logits_sp = ops.export_embeddings(base.x, ad, old_ad=old, ebd_weight=w, That_a=classifier.prompt_matrix(),
                                                     That_b=classifier.prompt_matrix(spurious=True),
                                                     inv_tau=1.0 / classifier.temperature, idx=idx, n_rows=n)

    def visit(t):
        return t if idx is None else t.index_select(0, idx.long())
    y_t, g_t, p_t = visit(base.labels[target]), visit(base.labels["group"]), visit(base.labels["spurious"])
    stats = _stats_buffers(len(sizes), base.n_groups, base.device, "export")
    stats.zero_()
    pred = ops.group_counts(logits, y_t, g_t, stats, bs, G=base.n_groups, want_pred=True)
    scratch = _stats_buffers(len(sizes), base.n_groups, base.device, "export_sp")
    pred_sp = ops.group_counts(logits_sp, p_t, g_t, scratch, bs, G=base.n_groups, want_pred=True)
    loss_sum, counts = stats.host()
    _, acc, acc_groups = replay_epoch(loss_sum, counts, sizes, base.n_groups)
    group_acc = eval_group_acc(acc_groups, get_yp_func, train_group_ratio)
    print(f"{print_label}:", str(group_acc))
    meta = {"targets": list(y_t.cpu().numpy().astype(np.int64)), "spuriouss": list(p_t.cpu().numpy().astype(np.int64)),
            "groups": list(g_t.cpu().numpy().astype(np.int64)), "predictions": list(pred.cpu().numpy().astype(np.int64)),
            "predictions_spurious": list(pred_sp.cpu().numpy().astype(np.int64))}
    return (None, acc.avg, group_acc), (feats.cpu().numpy(), meta)


def validate_zs(opt, val_loader, classifier, criterion, get_yp_func, train_group_ratio, target,
                print_label='Zero-shot Prediction (Test) (Class)'):
    """Feature-quality check with class or spurious prompts (final_main.py:725-803)."""
    _check_criterion(criterion)
    if target not in ("class", "spurious"):
        raise ValueError(target)
    if opt.tl_method == "linear_probing":
        # same as the CLIP zero-shot head on the raw embeddings (final_main.py:730-738, 757-759)
        from .modules import get_text_embedding
        raw = get_text_embedding(opt.text_embedding_dir if target == "class" else opt.text_spurious_embedding_dir)
        classifier.eval()
        base, rows = val_loader.base_rows(val_loader.draw_order())
        n, bs = len(rows), val_loader.batch_size
        sizes = _batch_sizes(n, bs)
        That = ops.normalize_text(raw.to(base.device, torch.float32).contiguous())
        stats = _stats_buffers(len(sizes), base.n_groups, base.device, "eval")
        contiguous = len(rows) == len(base) and np.array_equal(rows, np.arange(len(base)))
        if contiguous and getattr(base, "x16", None) is not None and ops.head_f16_supported(base.x16.shape[1]):
            # fp16-resident rows (CLIP emits fp16): kind::f16 head on the rows as stored, same results
            ops.logits_ce_f16(base.x16, base.labels[target], base.labels["group"], That, 1.0 / opt.zs_temperature, stats, bs,
                              G=base.n_groups, normalize_rows=True)
        else:
            ops.logits_ce(base.x, base.labels[target], base.labels["group"], That, 1.0 / opt.zs_temperature, stats, bs,
                          idx=None if contiguous else _order_to_device(base, rows), n_rows=n, G=base.n_groups, normalize_rows=True)
        loss_sum, counts = stats.host()
        n_groups = base.n_groups
    else:
        loss_sum, counts, sizes, n_groups = _run_eval(val_loader, classifier, target, spurious_prompts=(target == "spurious"))
    losses, acc, acc_groups = replay_epoch(loss_sum, counts, sizes, n_groups)
    group_acc = eval_group_acc(acc_groups, get_yp_func, train_group_ratio)
    print(f"{print_label}:", str(group_acc))
    return losses.avg, acc.avg, group_acc


def balance_val(val_loader, opt, print_procedure=False):
    """Per-epoch group-balanced resampling of the regularisation split (final_main.py:346-379): each group's
    positions are shuffled with the GLOBAL numpy RNG, cut to the smallest group, interleaved g0,g1,g2,g3,...;
    the loader is sequential with batch size min(batch_size_reg, #balanced)."""
    sub_dataset = val_loader.dataset
    n_groups = sub_dataset.dataset.n_groups
    groups_here = sub_dataset.dataset.group_array[sub_dataset.indices]
    per_group = [np.where(groups_here == g)[0] for g in range(n_groups)]
    smallest = min(len(ix) for ix in per_group)
    for i, ix in enumerate(per_group):
        np.random.shuffle(ix)
        if print_procedure:
            print(f"(Group {i}): ", len(ix), ix[:10])
        per_group[i] = ix[:smallest]
    balanced_indices = np.stack(per_group, axis=1).reshape(-1)
    if print_procedure:
        print(f"Balanced sample indices : {len(balanced_indices)} per epoch ({balanced_indices[:16]})")
    bs = opt.batch_size_reg if opt.batch_size_reg <= len(balanced_indices) else len(balanced_indices)
    return EmbeddingLoader(Subset(sub_dataset, balanced_indices), shuffle=False, batch_size=bs)
