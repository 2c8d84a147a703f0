"""ORACLE (test infrastructure, not product code).

A CPU/numpy restatement of the reference's regularized-adapter hot path, written from the
reference's behaviour (file:line citations are into /root/reference).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg may import this package; the product
package never does.  Parity pinning: every function here is checked in `tests/test_oracle_golden.py`
against fixtures produced by running the reference's own PyTorch code (`oracle/make_golden.py`).

All functions take a `dtype` (np.float32 mirrors the reference's arithmetic type, np.float64 is
the high-precision cross-check).
"""
from __future__ import annotations

import math

import numpy as np

BN_EPS = 1e-5          # torch.nn.BatchNorm1d default, final_main.py:169
BN_MOMENTUM = 0.1      # torch.nn.BatchNorm1d default


# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
PARAM_KEYS = ("W1", "b1", "gamma", "beta", "W2", "b2")


def init_adapter_params(rng: np.random.Generator, D: int, H: int, dtype=np.float32) -> dict:
    """Same *distribution* as nn.Linear/nn.BatchNorm1d defaults (final_main.py:167-172); the
    stream is numpy's, and is injected into both sides of every parity test."""
    k1 = 1.0 / math.sqrt(D)
    k2 = 1.0 / math.sqrt(H)
    return {
        "W1": rng.uniform(-k1, k1, (H, D)).astype(dtype),
        "b1": rng.uniform(-k1, k1, H).astype(dtype),
        "gamma": np.ones(H, dtype),
        "beta": np.zeros(H, dtype),
        "W2": rng.uniform(-k2, k2, (D, H)).astype(dtype),
        "b2": rng.uniform(-k2, k2, D).astype(dtype),
        "running_mean": np.zeros(H, dtype),
        "running_var": np.ones(H, dtype),
        "num_batches_tracked": np.int64(0),
    }


def copy_params(p: dict) -> dict:
    return {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in p.items()}


# ----------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------
def normalize_text(T: np.ndarray) -> np.ndarray:
    """final_main.py:77 -- text matrix [D, C] is normalised per column on every call."""
    return T / np.sqrt((T * T).sum(axis=0, keepdims=True))


def adapter_forward(X, p, train: bool, dtype=np.float32) -> dict:
    """Adapter.forward (final_main.py:160-174) + the row L2-normalise of CustomCLIP.forward
    (final_main.py:67-68).  Train mode uses batch statistics (biased variance), eval mode the
    running statistics."""
    X = X.astype(dtype)
    W1, b1, W2, b2 = (p[k].astype(dtype) for k in ("W1", "b1", "W2", "b2"))
    gamma, beta = p["gamma"].astype(dtype), p["beta"].astype(dtype)
    a = X @ W1.T + b1
    if train:
        mu = a.mean(axis=0)
        var = ((a - mu) ** 2).mean(axis=0)
    else:
        mu = p["running_mean"].astype(dtype)
        var = p["running_var"].astype(dtype)
    rstd = 1.0 / np.sqrt(var + dtype(BN_EPS))
    ahat = (a - mu) * rstd
    pre = gamma * ahat + beta
    h = np.maximum(pre, 0)
    z = h @ W2.T + b2
    n = np.sqrt((z * z).sum(axis=1, keepdims=True))
    u = z / n
    return dict(a=a, mu=mu, var=var, rstd=rstd, ahat=ahat, pre=pre, h=h, z=z, n=n, u=u)


def bn_running_update(p: dict, mu, var, B: int) -> None:
    """nn.BatchNorm1d train-mode buffer update: momentum 0.1, unbiased variance."""
    dt = p["running_mean"].dtype
    p["running_mean"] = ((1 - BN_MOMENTUM) * p["running_mean"] + BN_MOMENTUM * mu).astype(dt)
    p["running_var"] = ((1 - BN_MOMENTUM) * p["running_var"] + BN_MOMENTUM * var * (B / (B - 1))).astype(dt)
    p["num_batches_tracked"] = np.int64(p["num_batches_tracked"] + 1)


def clip_logits(u, That, tau, dtype=np.float32):
    """final_main.py:78 -- cosine logits divided by the temperature."""
    return (u.astype(dtype) @ That.astype(dtype)) / dtype(tau)


def log_softmax(l):
    m = l.max(axis=1, keepdims=True)
    s = l - m
    return s - np.log(np.exp(s).sum(axis=1, keepdims=True))


def cross_entropy(logits, y):
    """nn.CrossEntropyLoss() default: mean over the batch (final_main.py:302)."""
    ls = log_softmax(logits)
    return -ls[np.arange(len(y)), y].mean()


def softmax(l):
    return np.exp(log_softmax(l))


# ----------------------------------------------------------------------------------------------
# backward (checked against the reference's autograd by the golden tests)
# ----------------------------------------------------------------------------------------------
def adapter_backward(X, p, fw: dict, du, dtype=np.float32) -> dict:
    """Gradients of the six adapter tensors given dL/du (u = row-normalised adapter output)."""
    X = X.astype(dtype)
    W2 = p["W2"].astype(dtype)
    gamma = p["gamma"].astype(dtype)
    u, n, h, pre, ahat, rstd = fw["u"], fw["n"], fw["h"], fw["pre"], fw["ahat"], fw["rstd"]
    dz = (du - u * (u * du).sum(axis=1, keepdims=True)) / n
    g = {}
    g["b2"] = dz.sum(axis=0)
    g["W2"] = dz.T @ h
    dh = dz @ W2
    dpre = dh * (pre > 0)
    g["gamma"] = (dpre * ahat).sum(axis=0)
    g["beta"] = dpre.sum(axis=0)
    dahat = dpre * gamma
    da = (dahat - dahat.mean(axis=0) - ahat * (dahat * ahat).mean(axis=0)) * rstd
    g["W1"] = da.T @ X
    g["b1"] = da.sum(axis=0)
    return g


def sgd_step(p: dict, g: dict, v: dict | None, lr, momentum=0.9, wd=5e-5):
    """torch.optim.SGD (demo/util.py:118-136): g += wd*p; v = g on the first step else m*v + g;
    p -= lr*v.  Returns the momentum dict."""
    first = v is None
    if first:
        v = {}
    for k in PARAM_KEYS:
        dt = p[k].dtype
        gk = (g[k].astype(dt) + dt.type(wd) * p[k]).astype(dt)
        if first:
            v[k] = gk.copy()
        else:
            v[k] = (dt.type(momentum) * v[k] + gk).astype(dt)
        p[k] = (p[k] - dt.type(lr) * v[k]).astype(dt)
    return v


def train_step_single(X, y, p, v, That, tau, lr, momentum=0.9, wd=5e-5, dtype=np.float32):
    """One iteration of train_one_epoch's body (final_main.py:455-466) for CustomCLIP(Adapter)."""
    B = X.shape[0]
    fw = adapter_forward(X, p, train=True, dtype=dtype)
    logits = clip_logits(fw["u"], That, tau, dtype)
    loss = cross_entropy(logits, y)
    dl = softmax(logits)
    dl[np.arange(B), y] -= 1
    dl /= B
    du = (dl @ That.astype(dtype).T) / dtype(tau)
    g = adapter_backward(X, p, fw, du, dtype)
    bn_running_update(p, fw["mu"], fw["var"], B)
    v = sgd_step(p, g, v, lr, momentum, wd)
    return dict(loss=float(loss), logits=logits, grads=g, v=v)


def multiple_adapter_logits(X, p_old, p_new, That, tau, train, w=0.5, dtype=np.float32):
    """MultipleAdapter.forward (final_main.py:121-140): 0.5/0.5 mix of the two normalised outputs,
    mix not re-normalised."""
    fo = adapter_forward(X, p_old, train, dtype)
    fn = adapter_forward(X, p_new, train, dtype)
    u = dtype(w) * fo["u"] + dtype(1 - w) * fn["u"]
    return clip_logits(u, That, tau, dtype), fo, fn


def train_step_multiple(X, y, p_old, p_new, v, That, tau, lr, w=0.5, momentum=0.9, wd=5e-5, dtype=np.float32):
    """Stage-2 step (final_main.py:610-623): both adapters run batch-stat BN and update their running
    stats; only new_adapter gets gradients (demo/util.py:128)."""
    B = X.shape[0]
    logits, fo, fn = multiple_adapter_logits(X, p_old, p_new, That, tau, True, w, dtype)
    loss = cross_entropy(logits, y)
    dl = softmax(logits)
    dl[np.arange(B), y] -= 1
    dl /= B
    du = dtype(1 - w) * (dl @ That.astype(dtype).T) / dtype(tau)
    g = adapter_backward(X, p_new, fn, du, dtype)
    bn_running_update(p_old, fo["mu"], fo["var"], B)
    bn_running_update(p_new, fn["mu"], fn["var"], B)
    v = sgd_step(p_new, g, v, lr, momentum, wd)
    return dict(loss=float(loss), logits=logits, grads=g, v=v)


def eval_logits(X, p, That, tau, p_new=None, w=0.5, dtype=np.float32):
    """validate()'s forward (final_main.py:680): eval-mode BN; MultipleAdapter if p_new is given."""
    if p_new is None:
        fw = adapter_forward(X, p, False, dtype)
        return clip_logits(fw["u"], That, tau, dtype)
    return multiple_adapter_logits(X, p, p_new, That, tau, False, w, dtype)[0]


def export_features(X, p, p_new=None, w=0.5, dtype=np.float32):
    """validate_adapter_with_return (demo/demo_visualization.ipynb:1117-1215): the adapted embeddings handed to the
    visualisation notebooks.  tl_method == "adapter": the UN-normalised adapter output `classifier.adapter(x)`; otherwise
    the MultipleAdapter mix `w * u_old + (1 - w) * u_new` of the two L2-normalised outputs (not re-normalised).  The
    notebook scores these features directly: logits = features @ That / tau (class and spurious prompts)."""
    if p_new is None:
        return adapter_forward(X, p, False, dtype)["z"]
    fo, fn = adapter_forward(X, p, False, dtype), adapter_forward(X, p_new, False, dtype)
    return dtype(w) * fo["u"] + dtype(1 - w) * fn["u"]


# ----------------------------------------------------------------------------------------------
# group metrics (final_main.py:383-412, demo/util.py:18-46)
# ----------------------------------------------------------------------------------------------
class AverageMeter:
    """demo/util.py:18-33, including its float accumulation `sum += val * n`."""

    def __init__(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def group_counts(logits, y, g, n_groups):
    """Integer content of update_dict (final_main.py:383-391): per group, #rows and #correct."""
    pred = np.argmax(logits, axis=1)
    corr = pred == y
    total = np.bincount(g, minlength=n_groups).astype(np.int64)
    correct = np.bincount(g, weights=corr, minlength=n_groups).astype(np.int64)
    return correct, total, pred


def update_dict(acc_groups, y, g, logits):
    """final_main.py:383-391 -- groups absent from the batch are skipped (np.unique)."""
    pred = np.argmax(logits, axis=1)
    corr = pred == y
    for gv in np.unique(g):
        mask = g == gv
        n = int(mask.sum())
        c = int(corr[mask].sum())
        acc_groups[int(gv)].update(c / n, n)


def get_results(acc_groups, n_places=2):
    """final_main.py:395-412; worst_acc is the min over every entry including mean_acc."""
    res = {f"acc_{g // n_places}_{g % n_places}": acc_groups[g].avg for g in acc_groups}
    all_correct = sum(acc_groups[g].sum for g in acc_groups)
    all_total = sum(acc_groups[g].count for g in acc_groups)
    res["mean_acc"] = all_correct / all_total
    res["worst_acc"] = min(res.values())
    return res


PRINT_ORDER = ["weighted_mean_acc", "worst_acc", "acc_0_0", "acc_0_1", "acc_1_0", "acc_1_1", "mean_acc"]


def finalize_train_results(acc_groups):
    """final_main.py:491-493 (no weighted mean in the train dict)."""
    r = get_results(acc_groups)
    return {k: np.round(r[k], 4) for k in PRINT_ORDER[1:]}


def finalize_eval_results(acc_groups, train_group_ratio):
    """final_main.py:704-716; train_group_ratio is the float32 tensor of data/*_embeddings.py:54-55."""
    r = get_results(acc_groups)
    indiv = [r[f"acc_{g // 2}_{g % 2}"] for g in range(len(acc_groups))]
    r["weighted_mean_acc"] = (np.array(indiv) * np.array(train_group_ratio)).sum()
    return {k: np.round(r[k], 4) for k in PRINT_ORDER}


def evaluate_batches(batches_logits_y_g, n_groups, train_group_ratio=None):
    """Run the meter protocol of validate()/train_*_epoch over a list of (logits, y, g, loss)."""
    meters = {g: AverageMeter() for g in range(n_groups)}
    losses, acc = AverageMeter(), AverageMeter()
    for logits, y, g, loss in batches_logits_y_g:
        bsz = len(y)
        losses.update(float(loss), bsz)
        acc.update(float((np.argmax(logits, 1) == y).sum()) / bsz, bsz)
        update_dict(meters, y, g, logits)
    if train_group_ratio is None:
        return losses.avg, acc.avg, finalize_train_results(meters)
    return losses.avg, acc.avg, finalize_eval_results(meters, train_group_ratio)


# ----------------------------------------------------------------------------------------------
# schedules (demo/util.py:70-115, final_main.py:262-284)
# ----------------------------------------------------------------------------------------------
def epoch_lr(lr0, epoch, decay_epochs, decay_rate, cosine=False, total_epochs=None):
    """adjust_learning_rate / adjust_learning_rate_reg, step branch (and stage-1 cosine branch)."""
    lr = lr0
    if cosine:
        eta_min = lr * (decay_rate ** 3)
        return eta_min + (lr - eta_min) * (1 + math.cos(math.pi * epoch / total_epochs)) / 2
    steps = int(np.sum(epoch > np.asarray(decay_epochs)))
    if steps > 0:
        lr = lr * (decay_rate ** steps)
    return lr


def warmup_lr(epoch_in_stage, batch_id, total_batches, warm_epochs, warmup_from, warmup_to):
    """warmup_learning_rate[_reg]; returns None when the warm-up no longer applies."""
    if epoch_in_stage > warm_epochs:
        return None
    p = (batch_id + (epoch_in_stage - 1) * total_batches) / (warm_epochs * total_batches)
    return warmup_from + p * (warmup_to - warmup_from)


# ----------------------------------------------------------------------------------------------
# sampling (final_main.py:346-379, data/waterbirds_embeddings_reg.py:97-109)
# ----------------------------------------------------------------------------------------------
def balance_val_indices(sub_groups: np.ndarray, n_groups: int, batch_size_reg: int, rng=np.random):
    """balance_val: shuffle each group's positions with the *global* numpy RNG, truncate to the
    smallest group, interleave.  Returns (indices into the reg subset, adjusted batch size)."""
    g_idx = [np.where(sub_groups == g)[0] for g in range(n_groups)]
    min_g = min(len(g) for g in g_idx)
    for i, g in enumerate(g_idx):
        rng.shuffle(g)
        g_idx[i] = g[:min_g]
    bal = np.array(list(zip(*g_idx))).reshape(-1)
    bs = batch_size_reg if batch_size_reg <= len(bal) else len(bal)
    return bal, bs


def stratified_halves(group_array: np.ndarray):
    """stratified_split_dataset: sklearn's train_test_split(test_size=0.5, random_state=42, stratify)."""
    from sklearn.model_selection import train_test_split
    reg_idx, val_idx = train_test_split(np.arange(len(group_array)), test_size=0.5, random_state=42,
                                        stratify=group_array)
    return reg_idx, val_idx


# ----------------------------------------------------------------------------------------------
# contrastive formula source (demo/visualizer_supcon.py:1532-1571); PARITY UNPINNED by the
# reference's own tests -- the reference has no runnable caller for it.
# ----------------------------------------------------------------------------------------------
def supcon_single_anchor(feats, n_pos, n_neg, tau_cl=0.1, dtype=np.float64):
    """feats: [1+P+N, d] already L2-normalised rows (forward_ca output); anchor first."""
    f = feats.astype(dtype)
    a = f[0]
    pos = f[1:1 + n_pos]
    neg = f[len(f) - n_neg:]

    def cos(rows):
        return (rows @ a) / (np.linalg.norm(rows, axis=1) * np.linalg.norm(a)) / tau_cl

    sp = cos(pos)
    m = sp.max()
    ep = np.exp(sp - m)
    en = np.exp(cos(neg) - m)
    log_probs = np.log(ep) - np.log(en.sum() + ep.sum())
    return float((-log_probs).mean())


def supcon_all_anchors(Z, labels, tau_cl=0.1, dtype=np.float64):
    """B x B generalisation: every row is an anchor; positives = other rows with the same label,
    negatives = rows with a different label.  Mean over anchors that have at least one positive and
    one negative of the single-anchor loss."""
    Z = Z.astype(dtype)
    B = len(Z)
    losses = []
    for i in range(B):
        pos = [j for j in range(B) if j != i and labels[j] == labels[i]]
        neg = [j for j in range(B) if labels[j] != labels[i]]
        if not pos or not neg:
            continue
        feats = np.concatenate([Z[i:i + 1], Z[pos], Z[neg]], 0)
        losses.append(supcon_single_anchor(feats, len(pos), len(neg), tau_cl, dtype))
    return float(np.mean(losses)) if losses else 0.0


def supcon_all_anchors_grad(Z, labels, tau_cl=0.1, dtype=np.float64):
    """Vectorised all-anchor loss and its gradient w.r.t. the (already normalised) rows, s_ij = z_i . z_j / tau:
    loss_i = logsumexp_{j != i} s_ij - mean_{p in P_i} s_ip; G_ij = softmax_{j != i}(s_i)_j - [j in P_i] / |P_i|;
    dZ = (G + G^T) Z / (tau * n_valid).  Checked against autograd of the loop form in tests/test_oracle_golden.py."""
    Z = Z.astype(dtype)
    labels = np.asarray(labels)
    B = len(Z)
    S = (Z @ Z.T) / tau_cl
    eye = np.eye(B, dtype=bool)
    same = (labels[:, None] == labels[None, :]) & ~eye
    npos = same.sum(1)
    nneg = (labels[:, None] != labels[None, :]).sum(1)
    valid = (npos > 0) & (nneg > 0)
    Sm = np.where(eye, -np.inf, S)
    mx = Sm.max(1, keepdims=True)
    E = np.exp(Sm - mx)
    se = E.sum(1, keepdims=True)
    lse = (np.log(se) + mx)[:, 0]
    pos_mean = np.where(npos > 0, (np.where(same, S, 0.0).sum(1)) / np.maximum(npos, 1), 0.0)
    row_loss = np.where(valid, lse - pos_mean, 0.0)
    n_valid = int(valid.sum())
    G = E / se - same / np.maximum(npos, 1)[:, None]
    G[~valid] = 0.0
    dZ = (G + G.T) @ Z / (tau_cl * max(n_valid, 1))
    loss = float(row_loss.sum() / max(n_valid, 1))
    return dict(loss=loss, row_loss=row_loss, n_valid=n_valid, G=G, dZ=dZ)


def head_logits(U, That, tau, normalize_rows=True, dtype=np.float64):
    """validate_zs on raw embeddings (final_main.py:757-768): row-normalise, cosine logits / temperature."""
    U = U.astype(dtype)
    if normalize_rows:
        U = U / np.sqrt((U * U).sum(1, keepdims=True))
    return (U @ That.astype(dtype)) / tau


def linear_probe_epoch(X, y, order, batch_size, W, b, lrs, momentum=0.9, wd=5e-5, dtype=np.float64):
    """LinearClassifier (final_main.py:43-49) trained by train_one_epoch (final_main.py:455-466) with torch SGD semantics."""
    W, b = W.astype(dtype).copy(), b.astype(dtype).copy()
    vW = vb = None
    losses, logits_all = [], []
    for s, lo in enumerate(range(0, len(order), batch_size)):
        rows = order[lo:lo + batch_size]
        x, yy = X[rows].astype(dtype), y[rows]
        logits = x @ W.T + b
        losses.append(-log_softmax(logits)[np.arange(len(rows)), yy].sum())
        logits_all.append(logits)
        dl = softmax(logits)
        dl[np.arange(len(rows)), yy] -= 1
        dl /= len(rows)
        gW, gb = dl.T @ x + wd * W, dl.sum(0) + wd * b
        vW = gW if vW is None else momentum * vW + gW
        vb = gb if vb is None else momentum * vb + gb
        W -= lrs[s] * vW
        b -= lrs[s] * vb
    return W, b, np.array(losses), logits_all
