"""fp16-resident eval forward (dbmm_eval_fwd_f16, csrc/eval_f16.cuh) against the numpy oracle and the fp32-resident path:
logits within 1e-4 relative (+1e-3 absolute on logits of magnitude ~100), argmax / group counters bit-exact, per-slot losses;
single adapter and MultipleAdapter (stage 2), ragged row counts, class / group / spurious prompts."""
import numpy as np
import pytest
import torch

from oracle import adapter_math as am
from oracle import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import dbmm
    return dbmm.ops


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype is not None else t).cuda()


def _trained(c, steps=2):
    """A few oracle steps so that BatchNorm running statistics, W2 and b2 are not at their initial values."""
    p = am.copy_params(c["p_old"]); v = None
    That = am.normalize_text(c["T_class"])
    for s in range(steps):
        v = am.train_step_single(c["X"][s], c["Y"][s], p, v, That, 0.01, 0.1)["v"]
    return p


@pytest.mark.parametrize("name", ["rn50_b699", "vitl_b256"])
def test_eval_f16_matches_oracle_and_fp32_path(ops, name):
    c = cases.make_case(name)
    p = _trained(c)
    ad = ops.AdapterTensors.from_numpy(p)
    xe = c["Xe"]
    assert np.array_equal(xe.astype(np.float16).astype(np.float32), xe), "the case's embeddings are fp16-valued by construction"
    X16, X32 = dev(xe.astype(np.float16)), dev(xe)
    ge = dev(c["Ge"], torch.int32)
    assert ops.eval_f16_supported(c["D"], c["H"], 4)
    for tag, T, C in (("class", c["T_class"], 2), ("group", c["T_group"], 4), ("spurious", c["T_spurious"], 2)):
        labels = c["Ge"] if C == 4 else (c["Pe"] if tag == "spurious" else c["Ye"])
        That = ops.normalize_text(dev(T))
        st = ops.BatchStatsBuffers(5, 4); st32 = ops.BatchStatsBuffers(5, 4)
        logits, pred = ops.eval_fwd_f16(X16, dev(labels, torch.int32), ge, ad, That, 100.0, st, 128, want_logits=True, want_pred=True)
        l32, p32 = ops.eval_fwd(X32, dev(labels, torch.int32), ge, ad, That, 100.0, st32, 128, want_logits=True, want_pred=True)
        orl = am.eval_logits(xe, p, am.normalize_text(T), 0.01)
        lo = logits.cpu().numpy()
        assert np.abs(lo - orl).max() <= 1e-4 * np.abs(orl).max() + 1e-3, tag
        assert np.abs(lo - l32.cpu().numpy()).max() <= 1e-4 * np.abs(orl).max() + 1e-3, tag
        correct, total, opred = am.group_counts(orl, labels, c["Ge"], 4)
        assert np.array_equal(pred.cpu().numpy(), opred) and torch.equal(pred, p32)
        ls, cn = st.host(); ls32, cn32 = st32.host()
        assert np.array_equal(cn, cn32)
        assert np.array_equal(cn[:, 0].sum(0), correct) and np.array_equal(cn[:, 1].sum(0), total)
        np.testing.assert_allclose(ls, ls32, rtol=1e-4, atol=1e-3)


def test_eval_f16_multiple_adapter(ops):
    c = cases.make_case("rn50_b699")
    p_old, p_new = _trained(c), am.copy_params(c["p_new"])
    That_g = am.normalize_text(c["T_group"]); v = None
    for s in range(2):       # stage-2 oracle steps: the frozen adapter's running statistics drift, the new adapter trains
        v = am.train_step_multiple(c["X"][s], c["G"][s], p_old, p_new, v, That_g, 0.01, 0.1)["v"]
    old, new = ops.AdapterTensors.from_numpy(p_old), ops.AdapterTensors.from_numpy(p_new)
    xe = c["Xe"]
    X16, X32, ge = dev(xe.astype(np.float16)), dev(xe), dev(c["Ge"], torch.int32)
    That = ops.normalize_text(dev(c["T_group"]))
    st = ops.BatchStatsBuffers(3, 4); st32 = ops.BatchStatsBuffers(3, 4)
    logits, pred = ops.eval_fwd_f16(X16, ge, ge, new, That, 100.0, st, 200, old_ad=old, ebd_weight=0.5, want_logits=True, want_pred=True)
    l32, p32 = ops.eval_fwd(X32, ge, ge, new, That, 100.0, st32, 200, old_ad=old, ebd_weight=0.5, want_logits=True, want_pred=True)
    assert np.abs(logits.cpu().numpy() - l32.cpu().numpy()).max() <= 1e-4 * float(l32.abs().max()) + 1e-3
    assert torch.equal(pred, p32)
    assert np.array_equal(st.host()[1], st32.host()[1])
    np.testing.assert_allclose(st.host()[0], st32.host()[0], rtol=1e-4, atol=1e-3)


def test_eval_f16_celeba_shape_counts_equal_fp32_path(ops):
    """162,770 x 1024 (BASELINE config 2 shape): the two resident formats give the same predictions and counters."""
    n, D, H = 162770, 1024, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    X16 = (torch.randn(n, D, device="cuda", generator=g) * 0.5).half()
    X32 = X16.float()
    y = torch.randint(0, 2, (n,), device="cuda", generator=g, dtype=torch.int32)
    grp = torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.int32)
    rng = np.random.default_rng(3)
    ad = ops.AdapterTensors.from_numpy(am.init_adapter_params(rng, D, H))
    That = ops.normalize_text(torch.randn(D, 2, device="cuda", generator=g))
    st, st32 = ops.BatchStatsBuffers(159, 4), ops.BatchStatsBuffers(159, 4)
    _, pred = ops.eval_fwd_f16(X16, y, grp, ad, That, 100.0, st, 1024, want_pred=True)
    _, p32 = ops.eval_fwd(X32, y, grp, ad, That, 100.0, st32, 1024, want_pred=True)
    differ = int((pred != p32).sum())
    assert differ <= 2, f"{differ} of {n} predictions differ between the fp16- and fp32-resident paths (near-ties only)"
    assert np.abs(st.host()[1] - st32.host()[1]).max() <= 2
    np.testing.assert_allclose(st.host()[0], st32.host()[0], rtol=1e-3, atol=1e-2)
