"""Data-parallel host logic on CPU: two gloo ranks run dbmm.parallel.DataParallelTrainer with the oracle's arithmetic
standing in for the CUDA phases (the product has no CPU path; `step_fn` is the injection point for exactly this test).
Checked: contiguous sharding of the global batch, the three all-reduces of a step (BatchNorm column sums, dgamma /
dbeta sums, flat gradient), the B_local / B_global scaling of the dgamma / dbeta entries of the flat gradient, the
per-epoch counter reduction -- the two-rank result must equal the oracle's single-process step on the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import adapter_math as am

D, H, C, G = 32, 8, 2, 4
TAU = 0.01


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    rng = np.random.default_rng(21)
    N = 64
    X = rng.standard_normal((N, D))
    g = rng.integers(0, G, N)
    y = g // 2
    T = rng.standard_normal((D, C))
    p = am.init_adapter_params(rng, D, H, np.float64)
    order = rng.permutation(N)
    return X, y, g, am.normalize_text(T), p, order


class OracleStep:
    """The four phases of dbmm_train_step (include/dbmm.h), numpy fp64, exchanging data through the same workspace views
    (dbmm_train_accum_layout) and flat gradient buffer the CUDA kernels use."""

    def __init__(self, parallel, ops):
        self.parallel, self.ops, self.s = parallel, ops, {}

    def __call__(self, X, y, grp, ad, That, inv_tau, buf, lr, stats, slot, *, phases, idx, B_global, old_ad=None,
                 ebd_weight=0.5, G=4, momentum=0.9, weight_decay=5e-5):
        from dbmm import _lib
        s = self.s
        ws = self.ops.workspace(0, "cpu")
        colsum, dgb = self.parallel.accum_views(ws, H, 1, torch.float64)      # the oracle's sums are plain doubles
        p = ad                                                   # dict of float64 numpy arrays (this rank's replica)
        rows = idx.numpy()
        Xb, yb, gb = X.numpy()[rows], y[rows], grp[rows]
        Bl, Bg = len(rows), B_global
        if phases == _lib.PHASE_GEMM1:
            s["a"] = Xb @ p["W1"].T + p["b1"]
            colsum.zero_(); dgb.zero_()
            colsum[:H] = torch.from_numpy(s["a"].sum(0)); colsum[H:2 * H] = torch.from_numpy((s["a"] ** 2).sum(0))
        elif phases == _lib.PHASE_ROWS:
            cs = colsum.numpy()
            mu = cs[:H] / Bg
            var = np.maximum(cs[H:2 * H] / Bg - mu * mu, 0.0)
            rstd = 1.0 / np.sqrt(var + am.BN_EPS)
            ahat = (s["a"] - mu) * rstd
            pre = p["gamma"] * ahat + p["beta"]
            h = np.maximum(pre, 0)
            z = h @ p["W2"].T + p["b2"]
            n = np.sqrt((z * z).sum(1, keepdims=True))
            u = z / n
            logits = u @ That / TAU
            dl = am.softmax(logits)
            dl[np.arange(Bl), yb] -= 1
            dl /= Bg                                             # CE mean over the GLOBAL batch
            du = dl @ That.T / TAU
            dz = (du - u * (u * du).sum(1, keepdims=True)) / n
            dpre = (dz @ p["W2"]) * (pre > 0)
            s.update(mu=mu, var=var, rstd=rstd, ahat=ahat, h=h, dz=dz, dahat=dpre * p["gamma"])
            dgb[:H] = torch.from_numpy((dpre * ahat).sum(0)); dgb[H:] = torch.from_numpy(dpre.sum(0))
            correct, total, _ = am.group_counts(logits, yb, gb, G)
            stats.loss_sum[slot] += float(-am.log_softmax(logits)[np.arange(Bl), yb].sum())
            stats.counts[slot, 0] += torch.from_numpy(correct); stats.counts[slot, 1] += torch.from_numpy(total)
        elif phases == _lib.PHASE_WGRAD:
            d = dgb.numpy()                                      # GLOBAL sums after the all-reduce
            m1, m2 = p["gamma"] * d[H:] / Bg, p["gamma"] * d[:H] / Bg
            da = (s["dahat"] - m1 - s["ahat"] * m2) * s["rstd"]
            scale = Bl / Bg                                      # every rank holds the global sums: counted once after the reduce
            flat = np.concatenate([(da.T @ Xb).ravel(), np.zeros(H), d[:H] * scale, d[H:] * scale,
                                   (s["dz"].T @ s["h"]).ravel(), s["dz"].sum(0)])
            buf.grads.copy_(torch.from_numpy(flat))
        elif phases == _lib.PHASE_UPDATE:
            gflat = buf.grads.numpy()
            sl = self.ops.flat_param_slices(D, H)
            grads = {k: gflat[sl[k]].reshape(p[k].shape) for k in am.PARAM_KEYS}
            am.bn_running_update(p, s["mu"], s["var"], Bg)
            buf.v = am.sgd_step(p, grads, getattr(buf, "v", None), lr, momentum, weight_decay)
        else:
            raise AssertionError(phases)


class _Params(dict):
    H = H


class _Buf:
    def __init__(self, n):
        self.grads = torch.zeros(n, dtype=torch.float64)


class _Stats:
    def __init__(self, n_slots):
        self.loss_sum = torch.zeros(n_slots, dtype=torch.float64)
        self.counts = torch.zeros(n_slots, 2, G, dtype=torch.int64)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import dbmm
        from dbmm import ops, parallel
        X, y, g, That, p, order = _problem()
        p = _Params(am.copy_params(p))
        X = torch.from_numpy(X)
        trainer = parallel.DataParallelTrainer(step_fn=OracleStep(parallel, ops))
        assert (trainer.world, trainer.rank) == (world, rank)
        buf, stats = _Buf(ops.param_count(D, H)), _Stats(3)
        steps = trainer.train_epoch(X, torch.from_numpy(order), 24, y, g, p, That, 1.0 / TAU, buf, [0.5, 0.25, 0.125], stats, G=G)
        assert steps == 3                                        # 24 + 24 + 16 rows; rank shards of 12 / 12 / 8
        trainer.reduce_stats(stats)
        if rank == 0:
            out.put(dict(p={k: v for k, v in p.items()}, loss=stats.loss_sum.numpy(), counts=stats.counts.numpy()))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    import dbmm
    from dbmm import parallel
    for n in (1, 7, 24, 1024, 699):
        for world in (1, 2, 3, 8):
            cuts = [parallel.shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_epoch_equals_single_process_oracle():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = out.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    # single-process oracle on the same global batches
    X, y, g, That, p, order = _problem()
    v, loss, counts = None, [], []
    for s, lr in enumerate([0.5, 0.25, 0.125]):
        rows = order[s * 24:(s + 1) * 24]
        r = am.train_step_single(X[rows], y[rows], p, v, That, TAU, lr, dtype=np.float64)
        v = r["v"]
        loss.append(r["loss"] * len(rows))
        c, t, _ = am.group_counts(r["logits"], y[rows], g[rows], G)
        counts.append(np.stack([c, t]))
    for k in ("W1", "b1", "gamma", "beta", "W2", "b2", "running_mean", "running_var"):
        np.testing.assert_allclose(res["p"][k], p[k], rtol=1e-9, atol=1e-12, err_msg=k)
    assert int(res["p"]["num_batches_tracked"]) == 3
    np.testing.assert_allclose(res["loss"], loss, rtol=1e-10)
    assert np.array_equal(res["counts"], np.stack(counts))


# ---------------------------------------------------------------------------------------------------------------------
# contrastive regulariser with all-gathered negatives (BASELINE config 3): gather / all-reduce / reduce-scatter protocol
# ---------------------------------------------------------------------------------------------------------------------
def _supcon_problem():
    rng = np.random.default_rng(4)
    B, d = 48, 16
    Z = rng.standard_normal((B, d))
    Z /= np.linalg.norm(Z, axis=1, keepdims=True)
    return Z, rng.integers(0, 3, B)


def _oracle_supcon_compute():
    """(fwd, bwd) with the oracle's arithmetic: this rank's anchors against the gathered batch."""
    def parts(Z, labels, r0, nl):
        Z = Z.numpy().astype(np.float64); labels = labels.numpy()
        B = len(Z)
        S = Z[r0:r0 + nl] @ Z.T / 0.1
        rows = np.arange(r0, r0 + nl)
        self_mask = rows[:, None] == np.arange(B)[None, :]
        same = (labels[rows][:, None] == labels[None, :]) & ~self_mask
        valid = (same.sum(1) > 0) & ((labels[rows][:, None] != labels[None, :]).sum(1) > 0)
        Sm = np.where(self_mask, -np.inf, S)
        mx = Sm.max(1, keepdims=True)
        E = np.exp(Sm - mx)
        lse = np.log(E.sum(1)) + mx[:, 0]
        loss = lse - (np.where(same, S, 0).sum(1) / np.maximum(same.sum(1), 1))
        Gm = E / E.sum(1, keepdims=True) - same / np.maximum(same.sum(1), 1)[:, None]
        Gm[~valid] = 0
        return float(loss[valid].sum()), int(valid.sum()), (Z, Gm)

    def fwd(Z_all, labels_all, r0, nl):
        return parts(Z_all, labels_all, r0, nl)

    def bwd(Z_all, ctx, r0, nl, n_valid_global):
        Z, Gm = ctx
        scale = 1.0 / (0.1 * float(n_valid_global.item()))
        return torch.from_numpy(Gm @ Z * scale), torch.from_numpy(Gm.T @ Z[r0:r0 + nl] * scale)
    return fwd, bwd


def _supcon_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dbmm
    from dbmm import parallel
    Z, labels = _supcon_problem()
    Bl = len(Z) // world
    loss, dZ = parallel.supcon_distributed(torch.from_numpy(Z[rank * Bl:(rank + 1) * Bl]), torch.from_numpy(labels[rank * Bl:(rank + 1) * Bl]),
                                           compute=_oracle_supcon_compute())
    out.put((rank, loss, dZ.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_supcon_equals_single_process_oracle():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_supcon_worker, args=(r, 2, port, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = dict()
    for _ in range(2):
        r, loss, dZ = out.get(timeout=240)
        got[r] = (loss, dZ)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    Z, labels = _supcon_problem()
    ref = am.supcon_all_anchors_grad(Z, labels, 0.1)
    assert got[0][0] == pytest.approx(ref["loss"], rel=1e-10) and got[1][0] == pytest.approx(ref["loss"], rel=1e-10)
    np.testing.assert_allclose(np.concatenate([got[0][1], got[1][1]]), ref["dZ"], rtol=1e-9, atol=1e-12)
