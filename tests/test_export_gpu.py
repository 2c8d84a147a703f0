"""Adapted-embedding export: dbmm_export_embeddings / engine.validate_adapter_with_return against
the fixture produced by running the notebook's feature lines on the reference's modules (tests/golden/export_cases.npz,
oracle/make_golden.py::gen_export) and against the oracle restatement.
Bars: features and logits within 1e-3 relative; argmax predictions equal to the oracle's on rows whose top-2 logit gap
exceeds the fp32 noise floor (all rows of these cases)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import adapter_math as am
from oracle import cases

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype is not None else t).cuda()


@pytest.mark.parametrize("name", ["tiny_b33", "vitl_b256"])
def test_export_matches_reference_fixture(golden_dir, name):
    import dbmm
    from dbmm import ops
    gold = np.load(os.path.join(golden_dir, "export_cases.npz"))
    c = cases.make_case(name)
    run = gold[f"{name}/running"]
    p_old, p_new = am.copy_params(c["p_old"]), am.copy_params(c["p_new"])
    p_old["running_mean"], p_old["running_var"], p_new["running_mean"], p_new["running_var"] = run[0], run[1], run[2], run[3]
    old, new = ops.AdapterTensors.from_numpy(p_old), ops.AdapterTensors.from_numpy(p_new)
    X = dev(c["Xe"][:96])
    Tc, Ts = ops.normalize_text(dev(c["T_class"])), ops.normalize_text(dev(c["T_spurious"]))
    for tag, kw in (("adapter", dict(ad=old)), ("multi", dict(ad=new, old_ad=old, ebd_weight=0.5))):
        ad = kw.pop("ad")
        feats, la, lb = ops.export_embeddings(X, ad, That_a=Tc, That_b=Ts, inv_tau=100.0, **kw)
        ref = gold[f"{name}/{tag}/features"]
        assert np.abs(feats.cpu().numpy() - ref).max() <= 1e-3 * np.abs(ref).max()
        for got, key in ((la, "logits"), (lb, "logits_spurious")):
            rl = gold[f"{name}/{tag}/{key}"]
            assert np.abs(got.cpu().numpy() - rl).max() <= 1e-3 * np.abs(rl).max() + 1e-3
            assert np.array_equal(got.cpu().numpy().argmax(1), rl.argmax(1))


def test_export_ragged_indexed_rows_and_normalised_single():
    """Index lists, a row count that is no multiple of the 8-row CTA tile, and the normalised single-adapter variant."""
    import dbmm
    from dbmm import ops
    c = cases.make_case("rn50_b699")
    p = c["p_old"]
    ad = ops.AdapterTensors.from_numpy(p)
    X = dev(c["Xe"])
    order = torch.randperm(X.shape[0], generator=torch.Generator().manual_seed(3))[:203].to(torch.int32).cuda()
    feats, la, lb = ops.export_embeddings(X, ad, normalize_single=True, idx=order, That_a=ops.normalize_text(dev(c["T_group"])))
    assert lb is None and feats.shape == (203, 1024) and la.shape == (203, 4)
    fw = am.adapter_forward(c["Xe"][order.cpu().numpy()], p, False)
    assert np.abs(feats.cpu().numpy() - fw["u"]).max() <= 1e-3 * np.abs(fw["u"]).max()
    ol = am.clip_logits(fw["u"], am.normalize_text(c["T_group"]), 0.01)
    assert np.abs(la.cpu().numpy() - ol).max() <= 1e-3 * np.abs(ol).max() + 1e-3
    empty, _, _ = ops.export_embeddings(X[:0], ad)
    assert empty.shape == (0, 1024)


def test_validate_adapter_with_return_contract(tmp_path):
    """Return structure of the notebook function; for MultipleAdapter its scoring rule coincides with validate()'s, so the
    group dictionaries must be identical; the exported features reproduce the classifier's own logits."""
    import dbmm
    from dbmm import data, engine, metrics, synth
    from dbmm.modules import Adapter, CustomCLIP, MultipleAdapter
    ds = synth.make_dataset(name="waterbirds", dim=64, group_sizes=((300, 40, 30, 130), (52, 51, 26, 27), (53, 52, 25, 29)), seed=5)
    paths = synth.write_reference_files(ds, str(tmp_path))
    _, _, val_loader, test_loader = data.loaders_from_synthetic(ds, 128, 100)
    torch.manual_seed(0)
    clf = CustomCLIP(Adapter(64, 16), paths["text_embedding_dir"], paths["text_spurious_embedding_dir"], paths["text_group_embedding_dir"]).cuda()
    ma = MultipleAdapter(clf, Adapter(64, 16), init_near_identity=False).cuda()
    ratio = test_loader.dataset.group_ratio
    yp = lambda g: metrics.get_y_p(g, 2)
    opt = types.SimpleNamespace(tl_method="adapter_reg_seq_alter")
    (none, acc, group_acc), (emb, meta) = engine.validate_adapter_with_return(opt, test_loader, ma, torch.nn.CrossEntropyLoss(), yp, ratio, "class")
    _, acc_v, group_acc_v = engine.validate(opt, test_loader, ma, torch.nn.CrossEntropyLoss(), yp, ratio, "class")
    n = len(test_loader.dataset)
    assert none is None and emb.shape == (n, 64) and emb.dtype == np.float32
    assert sorted(meta) == ["groups", "predictions", "predictions_spurious", "spuriouss", "targets"] and all(len(v) == n for v in meta.values())
    assert group_acc == group_acc_v and acc == pytest.approx(acc_v)
    ma.eval()
    own = ma(test_loader.dataset.x).argmax(1).cpu().numpy()
    assert np.array_equal(np.array(meta["predictions"]), own)
    assert np.array_equal(np.array(meta["predictions_spurious"]), ma.forward_spurious(test_loader.dataset.x).argmax(1).cpu().numpy())
    assert np.array_equal(np.array(meta["groups"]), test_loader.dataset.group_array)
    # single adapter: un-normalised features, scored as the notebook scores them
    opt1 = types.SimpleNamespace(tl_method="adapter")
    (_, acc1, ga1), (emb1, meta1) = engine.validate_adapter_with_return(opt1, val_loader, clf, torch.nn.CrossEntropyLoss(), yp, ratio, "class")
    base, rows = data.resolve(val_loader.dataset)
    z = am.export_features(base.x.cpu().numpy()[rows], clf.adapter.tensors().to_numpy())
    assert np.abs(emb1 - z).max() <= 1e-3 * np.abs(z).max()
    Tn = am.normalize_text(ds.text_class)
    assert np.array_equal(np.array(meta1["predictions"]), (z @ Tn).argmax(1))
    assert set(ga1) == set(metrics.new_order_for_print)
