"""Drop-in boundary: the REFERENCE's own training loop --
final_main.train_one_epoch (final_main.py:426-496: `output = classifier(embeddings.detach())`, criterion, backward, a stock
torch.optim.SGD built by demo.util.set_optimizer, update_dict / get_results meters) -- runs unchanged on the dbmm modules, whose
train-mode forward / backward are the CUDA kernels behind a torch.autograd.Function, and lands where the reference's own
modules land from the same initial weights on the same batches.  The reference is the unmodified copy staged under
oracle/_ref/reference by oracle/stage_reference.py (git-ignored, travels with the snapshot)."""
import contextlib
import io
import json
import os
import types
from functools import partial

import numpy as np
import pytest
import torch

from oracle import ref_run

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_run.available(), reason="the reference is not staged (oracle/_ref/reference)")]

D, H, B = 1024, 128, 256


def _data(n, seed=3):
    rng = np.random.default_rng(seed)
    base = rng.standard_normal(D).astype(np.float32); mu = rng.standard_normal((4, D)).astype(np.float32)
    g = rng.choice(4, n, p=[0.44, 0.41, 0.12, 0.03])
    x = (base + 0.25 * mu[g] + rng.standard_normal((n, D)).astype(np.float32)).astype(np.float16).astype(np.float32)
    T2 = (base[:, None] + np.stack([mu[[0, 1]].mean(0), mu[[2, 3]].mean(0)], 1)).astype(np.float32)
    T4 = (base[:, None] + mu.T).astype(np.float32)
    return x, g.astype(np.int64), T2, T4


def _prompt_files(tmp_path, T2, T4):
    paths = []
    for name, T in (("class", T2), ("spurious", T2[:, ::-1]), ("group", T4)):
        p = os.path.join(str(tmp_path), name + ".json")
        with open(p, "w") as f:
            json.dump({f"prompt {c}": [float(t) for t in T[:, c]] for c in range(T.shape[1])}, f)
        paths.append(p)
    return paths


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_reference_train_one_epoch_runs_on_dbmm_modules(tmp_path):
    import dbmm
    from dbmm import modules as M
    fm, ru = ref_run.reference()
    x, g, T2, T4 = _data(5 * B + 77)
    y = g // 2
    paths = _prompt_files(tmp_path, T2, T4)
    torch.manual_seed(0)
    ref_clf = fm.CustomCLIP(fm.Adapter(D, H), *paths, temperature=0.01).cuda()
    our_clf = M.CustomCLIP(M.Adapter(D, H), *paths, temperature=0.01).cuda()
    our_clf.load_state_dict(ref_clf.state_dict(), strict=True)             # same key set, same initial weights
    opt = types.SimpleNamespace(learning_rate=0.1, learning_rate_reg=0.1, momentum=0.9, weight_decay=5e-5, warm=False,
                                watch_batch_results=False, print_freq=10)
    Xt, yt, gt = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(g).cuda()
    batches = [(Xt[s:s + B], {"class": yt[s:s + B], "group": gt[s:s + B]}, None) for s in range(0, len(x), B)]
    loader = ref_run._ListLoader(batches)
    crit = torch.nn.CrossEntropyLoss()
    get_yp = partial(fm.get_y_p, n_places=2)
    results = []
    for clf in (ref_clf, our_clf):
        optim = ru.set_optimizer(opt, clf)                                  # the reference's optimizer factory on either module
        with contextlib.redirect_stdout(io.StringIO()):
            for ep in range(2):
                loss, acc, group_acc = fm.train_one_epoch(opt, loader, clf, crit, optim, ep + 1, get_yp, target="class")
        results.append((float(loss), float(acc), dict(group_acc)))
    (l_ref, a_ref, ga_ref), (l_our, a_our, ga_our) = results
    assert l_our == pytest.approx(l_ref, rel=1e-3, abs=1e-5)
    for k, v in ga_ref.items():
        assert abs(float(ga_our[k]) - float(v)) <= 2.0 / 30 + 1e-4, k       # at most a couple of borderline samples of the smallest group
    sd_ref, sd_our = ref_clf.state_dict(), our_clf.state_dict()
    assert list(sd_ref.keys()) == list(sd_our.keys())
    for k in sd_ref:
        if sd_ref[k].dtype.is_floating_point:
            assert _rel(sd_our[k], sd_ref[k]) < 2e-3, k
        else:
            assert int(sd_our[k]) == int(sd_ref[k]), k                      # num_batches_tracked


def test_train_mode_forward_backward_match_reference_autograd(tmp_path):
    """One batch through classifier(x) in .train(): logits, parameter gradients and BatchNorm running statistics against the
    reference modules under torch autograd; single adapter and MultipleAdapter (use_group=True, forward_spurious)."""
    import dbmm
    from dbmm import modules as M
    fm, _ = ref_run.reference()
    x, g, T2, T4 = _data(699, seed=5)
    paths = _prompt_files(tmp_path, T2, T4)
    Xt, gt = torch.from_numpy(x).cuda(), torch.from_numpy(g).cuda()
    torch.manual_seed(1)
    ref_clf = fm.CustomCLIP(fm.Adapter(D, H), *paths, temperature=0.01).cuda()
    our_clf = M.CustomCLIP(M.Adapter(D, H), *paths, temperature=0.01).cuda()
    our_clf.load_state_dict(ref_clf.state_dict())
    ref_ma = fm.MultipleAdapter(ref_clf, fm.Adapter(D, H), init_near_identity=False).cuda()
    our_ma = M.MultipleAdapter(our_clf, M.Adapter(D, H), init_near_identity=False).cuda()
    our_ma.load_state_dict(ref_ma.state_dict())
    crit = torch.nn.CrossEntropyLoss()
    for ref, our, call, labels in ((ref_clf, our_clf, lambda m: m(Xt), gt // 2),
                                   (ref_ma, our_ma, lambda m: m(Xt, use_group=True), gt),
                                   (ref_ma, our_ma, lambda m: m.forward_spurious(Xt), gt % 2)):
        ref.train(); our.train()
        for m in (ref, our):
            m.zero_grad(set_to_none=True)
        lo_ref, lo_our = call(ref), call(our)
        assert _rel(lo_our, lo_ref) < 1e-4
        crit(lo_ref, labels).backward(); crit(lo_our, labels).backward()
        p_ref, p_our = dict(ref.named_parameters()), dict(our.named_parameters())
        for k, pr in p_ref.items():
            if pr.grad is None:
                assert p_our[k].grad is None or float(p_our[k].grad.abs().max()) == 0.0, k
                continue
            if k.endswith("layers.0.bias"):
                continue                                                    # db1 is analytically 0 (BatchNorm): rounding noise on both sides
            assert _rel(p_our[k].grad, pr.grad) < 1e-3, k
        for k, br in dict(ref.named_buffers()).items():
            bo = dict(our.named_buffers())[k]
            if br.dtype.is_floating_point:
                assert _rel(bo, br) < 1e-5, k
            else:
                assert int(bo) == int(br), k
