"""Group metrics from the kernels' integer counters.

Mirrors the observable behaviour of `AverageMeter`/`accuracy` (reference demo/util.py:18-46) and
`update_dict`/`get_results`/`get_y_p` (reference final_main.py:383-412): the kernels return, per batch,
the loss sum and per-group (#correct, #rows); this module replays the reference's meter protocol on
those integers on the host, once per epoch, so the rounded 4-decimal dictionaries are identical.
"""
from __future__ import annotations

import numpy as np

new_order_for_print = ["weighted_mean_acc", "worst_acc", "acc_0_0", "acc_0_1", "acc_1_0", "acc_1_1", "mean_acc"]


class AverageMeter:
    """Running value / sum / count / avg with the reference's float accumulation order."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def get_y_p(g, n_places):
    return g // n_places, g % n_places


def get_results(acc_groups, get_yp_func):
    results = {}
    for g in acc_groups.keys():
        y, p = get_yp_func(g)
        results[f"acc_{y}_{p}"] = acc_groups[g].avg
    total_correct = sum(acc_groups[g].sum for g in acc_groups.keys())
    total_rows = sum(acc_groups[g].count for g in acc_groups.keys())
    results["mean_acc"] = total_correct / total_rows
    results["worst_acc"] = min(results.values())      # includes mean_acc, like the reference
    return results


def replay_epoch(loss_sum: np.ndarray, counts: np.ndarray, batch_sizes, n_groups: int):
    """Feed per-batch kernel outputs through the meters exactly as the reference's loops do.

    loss_sum[b]: sum of the rows' NLL; counts[b, 0, g] correct, counts[b, 1, g] rows.
    Returns (losses, acc, acc_groups) meters.
    """
    losses, acc = AverageMeter(), AverageMeter()
    acc_groups = {g: AverageMeter() for g in range(n_groups)}
    for b, bsz in enumerate(batch_sizes):
        bsz = int(bsz)
        # criterion(output, labels).item(): fp32 mean over the batch
        losses.update(float(np.float32(loss_sum[b] / bsz)), bsz)
        acc.update(int(counts[b, 0].sum()) / bsz, bsz)
        for g in range(n_groups):              # np.unique(g): only the groups present in the batch
            n = int(counts[b, 1, g])
            if n > 0:
                acc_groups[g].update(int(counts[b, 0, g]) / n, n)
    return losses, acc, acc_groups


def train_group_acc(acc_groups, get_yp_func):
    r = get_results(acc_groups, get_yp_func)
    r = {k: r[k] for k in new_order_for_print[1:]}
    return {k: np.round(v, 4) for k, v in r.items()}


def eval_group_acc(acc_groups, get_yp_func, train_group_ratio):
    r = get_results(acc_groups, get_yp_func)
    indiv = [r["acc_{}_{}".format(*get_yp_func(g))] for g in range(len(acc_groups))]
    r["weighted_mean_acc"] = (np.array(indiv) * np.array(train_group_ratio)).sum()
    r = {k: r[k] for k in new_order_for_print}
    return {k: np.round(v, 4) for k, v in r.items()}
