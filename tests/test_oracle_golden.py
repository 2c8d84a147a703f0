"""The numpy oracle port (oracle/adapter_math.py) against fixtures produced by the reference's own
PyTorch code (oracle/make_golden.py).  CPU-only."""
import json
import os

import numpy as np
import pytest

from oracle import adapter_math as am
from oracle import cases

RTOL = 2e-4   # fp32 port vs fp32 torch: different summation order only


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "kernel_cases.npz"))


def _digest_close(p, gold, prefix, rtol=RTOL):
    d = cases.param_digest(p)
    ref = gold[f"{prefix}/sample"]
    scale = np.abs(ref).max()
    assert np.abs(d["sample"] - ref).max() <= rtol * scale
    assert abs(d["abs_sum"] - gold[f"{prefix}/abs_sum"]) <= rtol * gold[f"{prefix}/abs_sum"]
    for k in ("running_mean", "running_var"):     # scale-relative: lr=1.0 amplifies fp32 rounding noise
        ref_k = gold[f"{prefix}/{k}"]
        assert np.abs(d[k] - ref_k).max() <= rtol * max(np.abs(ref_k).max(), 1e-3), k
    assert d["nbt"] == gold[f"{prefix}/nbt"]


@pytest.mark.parametrize("name", list(cases.TRAIN_CASES))
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_stage1_steps_match_reference(gold, name, dtype):
    c = cases.make_case(name)
    p = {k: (v.astype(dtype) if isinstance(v, np.ndarray) else v) for k, v in c["p_old"].items()}
    That = am.normalize_text(c["T_class"].astype(dtype))
    v = None
    losses = []
    for s in range(c["steps"]):
        r = am.train_step_single(c["X"][s], c["Y"][s], p, v, That, 0.01, c["lr"], dtype=dtype)
        v = r["v"]
        losses.append(r["loss"])
        if s == 0:
            np.testing.assert_allclose(r["logits"], gold[f"{name}/s1_logits0"], rtol=RTOL, atol=2e-3)
            key = {"W1": "layers.0.weight", "b1": "layers.0.bias", "gamma": "layers.1.weight",
                   "beta": "layers.1.bias", "W2": "layers.3.weight", "b2": "layers.3.bias"}
            for k, tk in key.items():
                ref = gold[f"{name}/s1_grad0/{tk}"]
                got = r["grads"][k] if ref.shape == r["grads"][k].shape else r["grads"][k].reshape(-1)[::41]
                if k == "b1":      # analytically zero (BatchNorm removes it); both sides are rounding noise
                    assert np.abs(got).max() < 1e-5
                    continue
                assert np.abs(got - ref).max() <= RTOL * np.abs(ref).max(), k
    np.testing.assert_allclose(losses, gold[f"{name}/s1_losses"], rtol=1e-3 if dtype == np.float32 else 5e-4)
    _digest_close(p, gold, f"{name}/s1_final", rtol=2e-3)

    # eval with the trained single adapter (class / group / spurious prompts)
    for tag, T in (("", c["T_class"]), ("_group", c["T_group"]), ("_spurious", c["T_spurious"])):
        le = am.eval_logits(c["Xe"], p, am.normalize_text(T.astype(dtype)), 0.01, dtype=dtype)
        ref = gold[f"{name}/s1_eval_logits{tag}"]
        np.testing.assert_allclose(le, ref, rtol=2e-3, atol=2e-2)


@pytest.mark.parametrize("name", list(cases.TRAIN_CASES))
def test_stage2_multiple_adapter_matches_reference(gold, name):
    dtype = np.float32
    c = cases.make_case(name)
    p_old = am.copy_params(c["p_old"])
    That_c = am.normalize_text(c["T_class"])
    That_g = am.normalize_text(c["T_group"])
    v = None
    for s in range(c["steps"]):          # reproduce stage 1 to obtain the frozen adapter
        v = am.train_step_single(c["X"][s], c["Y"][s], p_old, v, That_c, 0.01, c["lr"])["v"]
    p_new = am.copy_params(c["p_new"])
    v2, losses = None, []
    for s in range(c["steps"]):
        ug = s % 2 == 1
        r = am.train_step_multiple(c["X"][s], c["G"][s] if ug else c["Y"][s], p_old, p_new, v2,
                                   That_g if ug else That_c, 0.01, c["lr"], dtype=dtype)
        v2 = r["v"]
        losses.append(r["loss"])
        if s <= 1:
            np.testing.assert_allclose(r["logits"], gold[f"{name}/s2_logits{s}"], rtol=2e-3, atol=2e-2)
    np.testing.assert_allclose(losses, gold[f"{name}/s2_losses"], rtol=2e-3)
    _digest_close(p_new, gold, f"{name}/s2_final_new", rtol=3e-3)
    _digest_close(p_old, gold, f"{name}/s2_final_old", rtol=3e-3)   # frozen weights, drifting BN buffers
    le = am.eval_logits(c["Xe"], p_old, That_c, 0.01, p_new=p_new)
    np.testing.assert_allclose(le, gold[f"{name}/s2_eval_logits"], rtol=3e-3, atol=3e-2)


def test_metrics_protocol_matches_reference(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "metrics_cases.json")))
    for name, c in gold.items():
        batches = []
        for b in c["batches"]:
            logits = np.array(b["logits"], np.float32)
            y = np.array(b["y"]); g = np.array(b["g"])
            batches.append((logits, y, g, am.cross_entropy(logits.astype(np.float64), y)))
        ratio = np.array(c["train_group_ratio"], np.float32)
        loss_avg, acc_avg, ga = am.evaluate_batches(batches, 4, ratio)
        assert acc_avg == pytest.approx(c["acc_avg"], abs=1e-12), name
        assert loss_avg == pytest.approx(c["loss_avg"], rel=1e-5), name
        for k, v in c["group_acc"].items():
            assert float(ga[k]) == pytest.approx(v, abs=1e-9), (name, k)


def test_schedules_match_reference(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "schedule_cases.json")))
    for name, c in gold.items():
        cfg = c["config"]
        FL = cfg["epochs_feature_learning"]
        W = 2 if cfg["dataset"] == "celeba" else 10
        for e, per_batch in enumerate(c["lrs"], start=1):
            if e <= FL:
                base = am.epoch_lr(cfg["learning_rate"], e, cfg["lr_decay_epochs"], cfg["lr_decay_rate"])
                exp = [base] * cfg["n_train"]
            else:
                base = am.epoch_lr(cfg["learning_rate_reg"], e, cfg["lr_decay_epochs"], cfg["lr_decay_rate"])
                exp = []
                for b in range(cfg["n_reg"]):
                    w = am.warmup_lr(e - FL, b, cfg["n_reg"], W, cfg["learning_rate_reg"] / 1e2, cfg["learning_rate_reg"])
                    exp.append(base if w is None else w)
            assert exp == per_batch, (name, e)


def test_sampling_matches_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, "sampling_cases.npz"))
    for name in ("waterbirds", "celeba"):
        g = gold[f"{name}/group_array"].astype(np.int64)
        reg_idx, val_idx = am.stratified_halves(g)
        assert np.array_equal(reg_idx, gold[f"{name}/reg_idx"])
        assert np.array_equal(val_idx, gold[f"{name}/val_idx"])
        np.random.seed(42)
        for bsr in (4, 256, 100000):
            for ep in range(3):
                idx, bs = am.balance_val_indices(g[reg_idx], 4, bsr)
                assert np.array_equal(idx, gold[f"{name}/balanced_bsr{bsr}_ep{ep}"])
                assert bs == gold[f"{name}/balanced_bsr{bsr}_ep{ep}_bs"]


def test_supcon_single_anchor_matches_reference(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "supcon_cases.json")))
    for name, c in gold.items():
        feats = np.array(c["feats"], np.float32)
        got = am.supcon_single_anchor(feats, c["P"], c["N"], 0.1)
        assert got == pytest.approx(c["loss"], rel=1e-5), name


def test_supcon_all_anchors_reduces_to_single_anchor():
    rng = np.random.default_rng(3)
    Z = rng.standard_normal((12, 16)); Z /= np.linalg.norm(Z, axis=1, keepdims=True)
    labels = np.array([0] * 5 + [1] * 7)
    i = 0
    pos = [j for j in range(12) if j != i and labels[j] == labels[i]]
    neg = [j for j in range(12) if labels[j] != labels[i]]
    single = am.supcon_single_anchor(np.concatenate([Z[i:i + 1], Z[pos], Z[neg]]), len(pos), len(neg))
    assert np.isfinite(single)
    assert np.isfinite(am.supcon_all_anchors(Z, labels))


def test_supcon_vectorised_matches_loop_and_autograd():
    """The B x B form used as the GPU kernels' oracle == the loop over the reference-pinned single-anchor formula, and its
    analytic gradient == torch autograd of that loop (unit rows, one label with a single member -> invalid anchor)."""
    import torch
    rng = np.random.default_rng(5)
    Z = rng.standard_normal((14, 16)); Z /= np.linalg.norm(Z, axis=1, keepdims=True)
    labels = np.array([0] * 5 + [1] * 6 + [2] * 2 + [3])
    r = am.supcon_all_anchors_grad(Z, labels, 0.1)
    assert r["n_valid"] == 13
    assert r["loss"] == pytest.approx(am.supcon_all_anchors(Z, labels, 0.1), rel=1e-10)
    Zt = torch.tensor(Z, dtype=torch.float64, requires_grad=True)
    losses = []
    for i in range(14):
        pos = [j for j in range(14) if j != i and labels[j] == labels[i]]
        neg = [j for j in range(14) if labels[j] != labels[i]]
        if not pos or not neg:
            continue
        sp = (Zt[pos] @ Zt[i]) / 0.1
        sn = (Zt[neg] @ Zt[i]) / 0.1
        m = sp.max().detach()                                   # demo/visualizer_supcon.py:1546 (detached max of the positives)
        lp = torch.log(torch.exp(sp - m)) - torch.log(torch.exp(sn - m).sum() + torch.exp(sp - m).sum())
        losses.append((-lp).mean())
    torch.stack(losses).mean().backward()
    np.testing.assert_allclose(r["dZ"], Zt.grad.numpy(), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", ["tiny_b33", "vitl_b256"])
def test_export_features_match_reference_modules(golden_dir, name):
    """validate_adapter_with_return's feature / logit lines (demo/demo_visualization.ipynb:1117-1215) executed on the
    reference's modules (oracle/make_golden.py::gen_export) vs the oracle restatement."""
    gold = np.load(os.path.join(golden_dir, "export_cases.npz"))
    c = cases.make_case(name)
    run = gold[f"{name}/running"]
    p_old, p_new = am.copy_params(c["p_old"]), am.copy_params(c["p_new"])
    p_old["running_mean"], p_old["running_var"], p_new["running_mean"], p_new["running_var"] = run[0], run[1], run[2], run[3]
    X = c["Xe"][:96]
    Tc, Ts = am.normalize_text(c["T_class"]), am.normalize_text(c["T_spurious"])
    for tag, feats in (("adapter", am.export_features(X, p_old)), ("multi", am.export_features(X, p_old, p_new, 0.5))):
        ref = gold[f"{name}/{tag}/features"]
        assert np.abs(feats - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-6
        for key, T in (("logits", Tc), ("logits_spurious", Ts)):
            rl = gold[f"{name}/{tag}/{key}"]
            ol = feats @ T / np.float32(0.01)
            assert np.abs(ol - rl).max() <= 1e-3 * np.abs(rl).max() + 1e-3
