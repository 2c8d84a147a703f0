"""`--tl_method contrastive_adapter` step (dbmm_contrastive_step): loss, all six parameter gradients, the SGD update and the
BatchNorm running statistics against torch fp64 autograd of
    L = w * SupCon_all_anchors(L2(adapter_train(L2(x))), labels; tau)
(forward_ca of workspace/jinsu/SupCon.ipynb:109-113, per-anchor formula of demo/visualizer_supcon.py:1532-1571 applied to
every anchor of the batch; parity unpinned by the reference's own tests: it has no runnable contrastive path)."""
import numpy as np
import pytest
import torch

from oracle import adapter_math as am

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import dbmm
    return dbmm.ops


def _torch_reference(x, labels, p, tau, w, pre_norm):
    t = {k: torch.tensor(np.asarray(v, np.float64), requires_grad=k in ("W1", "b1", "gamma", "beta", "W2", "b2")) for k, v in p.items()
         if k in ("W1", "b1", "gamma", "beta", "W2", "b2")}
    X = torch.tensor(x.astype(np.float64))
    if pre_norm:
        X = X / X.norm(dim=1, keepdim=True)
    a = X @ t["W1"].T + t["b1"]
    mu, var = a.mean(0), a.var(0, unbiased=False)
    ah = (a - mu) / torch.sqrt(var + 1e-5)
    h = torch.relu(ah * t["gamma"] + t["beta"])
    z = h @ t["W2"].T + t["b2"]
    u = z / z.norm(dim=1, keepdim=True)
    S = (u @ u.T) / tau
    y = torch.tensor(labels)
    B = len(labels)
    eye = torch.eye(B, dtype=torch.bool)
    pos = (y[:, None] == y[None, :]) & ~eye
    neg = y[:, None] != y[None, :]
    valid = (pos.sum(1) > 0) & (neg.sum(1) > 0)
    lse = torch.logsumexp(S.masked_fill(eye, -float("inf")), dim=1)
    per = -((S - lse[:, None]) * pos).sum(1) / pos.sum(1).clamp_min(1)
    loss = w * per[valid].mean()
    loss.backward()
    return float(loss), {k: v.grad.numpy() for k, v in t.items()}, (mu.detach().numpy(), a.var(0, unbiased=True).detach().numpy()), int(valid.sum())


@pytest.mark.parametrize("B,D,pre_norm", [(300, 1024, True), (256, 768, False), (1024, 1024, True)])
def test_contrastive_step_matches_torch_autograd(ops, B, D, pre_norm):
    H = 128
    rng = np.random.default_rng(B + D)
    base = rng.standard_normal(D).astype(np.float32); mu4 = rng.standard_normal((4, D)).astype(np.float32)
    g = rng.choice(4, B, p=[0.4, 0.4, 0.15, 0.05])
    x = (base + 0.3 * mu4[g] + rng.standard_normal((B, D)).astype(np.float32)).astype(np.float16).astype(np.float32)
    labels = (g // 2).astype(np.int32)
    p = am.init_adapter_params(rng, D, H)
    p["gamma"] = (1.0 + 0.1 * rng.standard_normal(H)).astype(np.float32); p["beta"] = (0.1 * rng.standard_normal(H)).astype(np.float32)
    p["b2"] = (0.05 * rng.standard_normal(D)).astype(np.float32)
    tau, w, lr, wd = 0.1, 0.1, 0.05, 5e-5
    loss_ref, g_ref, (mu_ref, var_ref), n_valid_ref = _torch_reference(x, labels, p, tau, w, pre_norm)

    ad = ops.AdapterTensors.from_numpy(p)
    buf = ops.TrainBuffers(D, H)
    loss = torch.zeros(1, dtype=torch.float64, device="cuda"); nv = torch.zeros(1, dtype=torch.int32, device="cuda")
    X = torch.from_numpy(x).cuda()
    order = torch.from_numpy(rng.permutation(B).astype(np.int32)).cuda()       # through the index list, like an epoch does
    xs = torch.from_numpy(x[order.cpu().numpy()]).cuda()
    del xs
    ops.contrastive_step(X, torch.from_numpy(labels).cuda(), ad, buf, lr, idx=order, pre_norm=pre_norm, tau_cl=tau, loss_weight=w,
                         loss_out=loss, n_valid_out=nv)
    torch.cuda.synchronize()
    assert int(nv.item()) == n_valid_ref
    assert float(loss.item()) == pytest.approx(loss_ref, rel=1e-4)
    sl = ops.flat_param_slices(D, H)
    grads = buf.grads.cpu().numpy()

    def rel(a, b):
        return np.abs(a.astype(np.float64) - b).max() / max(np.abs(b).max(), 1e-30)
    for k in ("W1", "gamma", "beta", "W2", "b2"):
        assert rel(grads[sl[k]], g_ref[k].reshape(-1)) < 1e-3, k
    got = ad.to_numpy()
    for k in ("W1", "b1", "gamma", "beta", "W2", "b2"):             # first step: v = g + wd p, p -= lr v
        gk = np.zeros_like(p[k], dtype=np.float64) if k == "b1" else g_ref[k]
        want = p[k].astype(np.float64) - lr * (gk + wd * p[k].astype(np.float64))
        assert rel(got[k], want) < 1e-4, k
    assert rel(got["running_mean"], 0.9 * p["running_mean"] + 0.1 * mu_ref) < 1e-4
    assert rel(got["running_var"], 0.9 * p["running_var"] + 0.1 * var_ref) < 1e-4
    assert int(got["num_batches_tracked"]) == int(p["num_batches_tracked"]) + 1


def test_contrastive_adapter_cli_runs(tmp_path):
    """`--tl_method contrastive_adapter` end to end through the drop-in CLI: groups from the stored zero-shot predictions,
    contrastive epochs, the usual val / test evaluation and model selection; the loss falls."""
    import contextlib
    import io
    import dbmm
    from dbmm import cli, synth
    ds = synth.make_dataset(name="waterbirds", dim=1024, seed=1234, scale=0.3, k=0.085, k_text=0.5, text_noise=0.02)
    paths = synth.write_reference_files(ds, str(tmp_path))
    argv = ["--dataset", "waterbirds", "--tl_method", "contrastive_adapter", "--batch_size", "256", "--learning_rate", "0.05",
            "--epochs", "3", "--lr_decay_rate", "0.1", "--lr_decay_epochs", "3", "--train_target", "class", "--random_seed", "7",
            "--num_positive", "32", "--num_negative", "32", "--batch_factor", "4"]
    for k, v in paths.items():
        argv += [f"--{k}", v]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        (tr, va, te), _ = cli.train_all_epochs(cli.parse_option(argv))
    out = buf.getvalue()
    losses = [float(l.split(":")[1].split("(")[0]) for l in out.splitlines() if l.startswith("Loss in Train (Contrastive)")]
    assert len(losses) == 3 and losses[-1] < losses[0], losses
    assert 0.0 <= float(te["worst_acc"]) <= 1.0 and "Contrastive groups:" in out
