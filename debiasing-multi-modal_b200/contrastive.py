"""Contrastive-adapter batch construction and the epoch loop of `--tl_method contrastive_adapter`.

Host side, numpy, same sampling protocol as the reference (demo/visualizer_supcon.py): the train split is sliced by the
zero-shot prediction stored with the embeddings (`compute_slice_indices`, :1100-1146); inside a slice the mispredicted rows
are the ANCHORS, the correctly predicted rows the (hard) negatives, the mispredicted rows of the other slice are added as easy
negatives, and the correctly predicted rows of a class are that class's POSITIVES (`prepare_contrastive_points`, :1148-1340);
for every anchor one group [anchor (+ num_anchor - 1 more of its class) | num_positive positives | num_negative negatives] is
drawn with the global numpy RNG in the reference's call order (`construct_contrastive_data`, :1342-1435), the groups are
balanced / shuffled (`load_contrastive_loader`, :1437-1484) and `batch_factor` groups form one loader batch.

What differs, by design (BASELINE.json north star): a loader batch is scored by ONE all-anchor B x B tensor-core contraction
(dbmm_contrastive_step) with the rows' class labels instead of a Python loop over single anchors with two forward passes
each (`train_one_epoch_cl`, :458-485) -- every row of the batch is an anchor, positives are the rows of its class.
"""
from __future__ import annotations

import numpy as np


def compute_slice_indices(y_pred: np.ndarray, labels: np.ndarray):
    """visualizer_supcon.py:1100-1146: one slice per zero-shot prediction; the rows' correctness inside each slice."""
    y_pred, labels = np.asarray(y_pred), np.asarray(labels)
    correct = y_pred == labels
    sliced_data_indices, all_correct = [], []
    for label in np.unique(y_pred):
        group = np.where(y_pred == label)[0]
        sliced_data_indices.append(group)
        all_correct.append(correct[group])
    return sliced_data_indices, all_correct


def prepare_contrastive_points(train_targets, train_spurious, sliced_data_indices, sliced_data_correct):
    """visualizer_supcon.py:1148-1340 (binary classes, as the reference: `another_slice_ix = |slice_ix - 1|`)."""
    train_targets, train_spurious = np.asarray(train_targets), np.asarray(train_spurious)
    incorrect = [np.logical_not(np.asarray(c, bool)) for c in sliced_data_correct]
    slice_anchors = [None] * len(sliced_data_indices)
    slice_negatives = [None] * len(sliced_data_indices)
    positives_by_class = {}
    for s, data_indices in enumerate(sliced_data_indices):
        ix = np.where(incorrect[s])[0]
        anchors = {"ix": data_indices[ix], "target": train_targets[data_indices][ix], "incorrect": incorrect[s][ix],
                   "source": np.ones(len(ix)).astype(int) * s, "spurious": train_spurious[data_indices][ix], "ix_by_class": {}}
        for t in np.unique(train_targets[data_indices][ix]):
            tix = np.where(train_targets[data_indices][ix] == t)[0]
            anchors["ix_by_class"][t] = data_indices[ix][tix]
        nix = np.setdiff1d(np.arange(len(data_indices)), ix)
        negatives = {"ix": list(data_indices[nix]), "target": list(train_targets[data_indices][nix]),
                     "incorrect": list(incorrect[s][nix]), "source": list(np.ones(len(nix)).astype(int) * s),
                     "spurious": list(train_spurious[data_indices][nix])}
        correct_data_indices = data_indices[nix]
        for c in np.unique(train_targets[data_indices][nix]):
            pix = np.where(train_targets[correct_data_indices] == c)[0]
            new = {"ix": list(correct_data_indices[pix]), "target": list(train_targets[correct_data_indices][pix]),
                   "correct": list(np.asarray(sliced_data_correct[s])[nix][pix]), "source": list(np.ones(len(pix)).astype(int) * s),
                   "spurious": list(train_spurious[correct_data_indices][pix])}
            if c in positives_by_class:
                for k, v in new.items():
                    positives_by_class[c][k].extend(v)
            else:
                positives_by_class[c] = new
        slice_anchors[s], slice_negatives[s] = anchors, negatives
    for s, data_indices in enumerate(sliced_data_indices):           # easy negatives: the other slice's mispredicted rows
        other = abs(s - 1)
        ix = np.where(incorrect[s])[0]
        slice_negatives[other]["ix"].extend(data_indices[ix])
        slice_negatives[other]["target"].extend(train_targets[data_indices][ix])
        slice_negatives[other]["incorrect"].extend(incorrect[s][ix])
        slice_negatives[other]["source"].extend(list(np.ones(len(ix)).astype(int) * s))
        slice_negatives[other]["spurious"].extend(list(train_spurious[data_indices][ix]))
    for c in positives_by_class:
        positives_by_class[c] = {k: np.array(v) for k, v in positives_by_class[c].items()}
    slice_negatives = [{k: np.array(v) for k, v in d.items()} for d in slice_negatives]
    return slice_anchors, slice_negatives, positives_by_class


def adjust_num_pos_neg(positives_by_class, slice_negatives, n_cls, num_anchor, num_positive, num_negative):
    """visualizer_supcon.py:1051-1080: cap the requested group composition by what the slices hold."""
    num_pos = min(int(num_positive), int(np.min([len(positives_by_class[c]["target"]) for c in range(n_cls)])))
    num_neg = min(int(num_negative), int(np.min([len(d["target"]) for d in slice_negatives])))
    return min(int(num_anchor), num_pos, num_neg), num_pos, num_neg


def construct_contrastive_data(slice_anchors, slice_negatives, positives_by_class, num_anchor, num_positive, num_negative):
    """visualizer_supcon.py:1342-1435: one group per anchor, drawn with the GLOBAL numpy RNG in the reference's call order
    (anchors of the class, positives, negatives; np.random.shuffle of the slice's groups at the end)."""
    batch_samples = []
    for s, anchor_dict in enumerate(slice_anchors):
        negative_dict = slice_negatives[s]
        per_slice = []
        for aix, anchor_ix in enumerate(anchor_dict["ix"]):
            anchor_class = anchor_dict["target"][aix]
            pool = anchor_dict["ix_by_class"][anchor_class]
            extra = np.random.choice(pool, size=num_anchor - 1, replace=(num_anchor - 1) > len(pool), p=None)
            anchors = np.concatenate([[anchor_ix], extra])
            pd = positives_by_class[anchor_class]
            pi = np.random.choice(np.arange(len(pd["ix"])), size=num_positive, replace=num_positive > len(pd["ix"]), p=None)
            positives = pd["ix"][pi]
            negatives = np.random.choice(negative_dict["ix"], size=num_negative, replace=num_negative > len(negative_dict["ix"]), p=None)
            per_slice.append(np.concatenate([anchors, positives, negatives]))
        np.random.shuffle(per_slice)
        batch_samples.append(per_slice)
    return batch_samples


def assemble_groups(batch_samples, balance_by_zs_pred=False, re_shuffle_ca_loader=True, maintain_alternative_ordering=False):
    """visualizer_supcon.py:1437-1467: [n_groups, 1 + P + N] index matrix in loader order."""
    if balance_by_zs_pred:
        if re_shuffle_ca_loader:
            for s in range(len(batch_samples)):
                np.random.shuffle(batch_samples[s])
        groups = np.array(list(zip(*batch_samples)))
        groups = groups.reshape(-1, groups.shape[-1])
        if not maintain_alternative_ordering and re_shuffle_ca_loader:
            np.random.shuffle(groups)
    else:
        groups = np.concatenate(batch_samples)
        if re_shuffle_ca_loader:
            np.random.shuffle(groups)
    return groups


def contrastive_batches(y, spurious, y_pred, *, n_cls=2, num_anchor=1, num_positive=64, num_negative=64, batch_factor=8,
                        balance_by_zs_pred=False, re_shuffle_ca_loader=True, maintain_alternative_ordering=False):
    """Everything above in the reference's order (train_all_epochs of visualizer_supcon.py); returns (groups [G, 1+P+N],
    list of loader batches: index arrays of batch_factor groups each, the adjusted (num_anchor, num_positive, num_negative))."""
    sl_ix, sl_ok = compute_slice_indices(y_pred, y)
    anchors, negatives, positives = prepare_contrastive_points(y, spurious, sl_ix, sl_ok)
    na, npos, nneg = adjust_num_pos_neg(positives, negatives, n_cls, num_anchor, num_positive, num_negative)
    samples = construct_contrastive_data(anchors, negatives, positives, na, npos, nneg)
    groups = assemble_groups(samples, balance_by_zs_pred, re_shuffle_ca_loader, maintain_alternative_ordering)
    flat = np.concatenate(groups)
    per = groups.shape[1] * int(batch_factor)
    batches = [flat[i:i + per] for i in range(0, len(flat), per)]
    return groups, batches, (na, npos, nneg)
