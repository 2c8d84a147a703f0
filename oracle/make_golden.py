"""ORACLE (test infrastructure).  Generates tests/golden/* by running the UNMODIFIED reference
(/root/reference, imported through oracle/reference_shim.py) on the seeded cases of oracle/cases.py.

Run in the build container only:   python -m oracle.make_golden [--only kernels|metrics|schedule|sampling|e2e|supcon|export]

The fixtures pin (a) the numpy oracle port and (b) the CUDA path to the reference's own PyTorch
arithmetic: Adapter/CustomCLIP/MultipleAdapter + nn.CrossEntropyLoss + optim.SGD (final_main.py:53-174,
demo/util.py:118-136), the meter protocol (final_main.py:383-412, 655-719), the LR schedules
(demo/util.py:70-115), balance_val (final_main.py:346-379), the stratified split
(data/waterbirds_embeddings_reg.py:97-109) and whole-run behaviour of train_all_epochs.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import cases  # noqa: E402
from oracle.reference_shim import import_reference  # noqa: E402

synth = importlib.import_module("debiasing-multi-modal_b200.synth")


def _write_text_json(path, T):
    with open(path, "w") as f:
        json.dump({f"prompt {c}": [float(t) for t in T[:, c]] for c in range(T.shape[1])}, f)


def _load_into(adapter, p):
    sd = {
        "layers.0.weight": torch.from_numpy(p["W1"].copy()), "layers.0.bias": torch.from_numpy(p["b1"].copy()),
        "layers.1.weight": torch.from_numpy(p["gamma"].copy()), "layers.1.bias": torch.from_numpy(p["beta"].copy()),
        "layers.1.running_mean": torch.from_numpy(p["running_mean"].copy()),
        "layers.1.running_var": torch.from_numpy(p["running_var"].copy()),
        "layers.1.num_batches_tracked": torch.tensor(int(p["num_batches_tracked"])),
        "layers.3.weight": torch.from_numpy(p["W2"].copy()), "layers.3.bias": torch.from_numpy(p["b2"].copy()),
    }
    adapter.load_state_dict(sd, strict=True)


def _dump_from(adapter):
    sd = adapter.state_dict()
    return {
        "W1": sd["layers.0.weight"].numpy().copy(), "b1": sd["layers.0.bias"].numpy().copy(),
        "gamma": sd["layers.1.weight"].numpy().copy(), "beta": sd["layers.1.bias"].numpy().copy(),
        "running_mean": sd["layers.1.running_mean"].numpy().copy(),
        "running_var": sd["layers.1.running_var"].numpy().copy(),
        "num_batches_tracked": np.int64(sd["layers.1.num_batches_tracked"].item()),
        "W2": sd["layers.3.weight"].numpy().copy(), "b2": sd["layers.3.bias"].numpy().copy(),
    }


def gen_kernels(fm, ru):
    out = {}
    tmp = tempfile.mkdtemp(prefix="dbmm_gold_")
    for name in cases.TRAIN_CASES:
        c = cases.make_case(name)
        D, H = c["D"], c["H"]
        tc, ts, tg = (os.path.join(tmp, f"{name}_{k}.json") for k in ("class", "spurious", "group"))
        _write_text_json(tc, c["T_class"]); _write_text_json(ts, c["T_spurious"]); _write_text_json(tg, c["T_group"])
        opt = types.SimpleNamespace(learning_rate=c["lr"], learning_rate_reg=c["lr"], momentum=0.9, weight_decay=5e-5)
        crit = torch.nn.CrossEntropyLoss()

        # ---- stage 1: CustomCLIP(Adapter), class prompts, SGD (train_one_epoch body, final_main.py:455-466)
        clf = fm.CustomCLIP(fm.Adapter(D, H), tc, ts, tg, temperature=0.01)
        _load_into(clf.adapter, c["p_old"])
        optim = ru.set_optimizer(opt, clf)
        clf.train()
        losses, grads0 = [], None
        for s in range(c["steps"]):
            x = torch.from_numpy(c["X"][s]); y = torch.from_numpy(c["Y"][s])
            o = clf(x.detach())
            loss = crit(o, y)
            if s == 0:
                out[f"{name}/s1_logits0"] = o.detach().numpy().copy()
            optim.zero_grad(); loss.backward()
            if s == 0:
                grads0 = {k: v.grad.numpy().copy() for k, v in clf.adapter.named_parameters()}
            optim.step()
            losses.append(loss.item())
        out[f"{name}/s1_losses"] = np.array(losses, np.float64)
        for k, v in grads0.items():
            out[f"{name}/s1_grad0/{k}"] = v if v.size <= 4096 else v.reshape(-1)[::41].copy()
        p1 = _dump_from(clf.adapter)
        for k, v in cases.param_digest(p1).items():
            out[f"{name}/s1_final/{k}"] = v
        if name == "tiny_b33":
            for k in ("W1", "b1", "gamma", "beta", "W2", "b2"):
                out[f"{name}/s1_final_full/{k}"] = p1[k]

        # ---- eval with the stage-1 result (validate forward, final_main.py:675-681) + forward_spurious
        clf.eval()
        with torch.no_grad():
            xe = torch.from_numpy(c["Xe"])
            le = clf(xe); lg = clf(xe, use_group=True); lsp = clf.forward_spurious(xe)
            out[f"{name}/s1_eval_logits"] = le.numpy().copy()
            out[f"{name}/s1_eval_logits_group"] = lg.numpy().copy()
            out[f"{name}/s1_eval_logits_spurious"] = lsp.numpy().copy()
            out[f"{name}/s1_eval_loss"] = np.float64(crit(le, torch.from_numpy(c["Ye"])).item())

        # ---- stage 2: MultipleAdapter (old frozen, new trainable), alternating class/group prompts
        new_ad = fm.Adapter(D, H)
        _load_into(new_ad, c["p_new"])
        ma = fm.MultipleAdapter(clf, new_ad, init_near_identity=False, ebd_weight=0.5)
        optim2 = ru.set_optimizer_reg(opt, ma)
        ma.train()
        losses2 = []
        for s in range(c["steps"]):
            use_group = (s % 2 == 1)
            x = torch.from_numpy(c["X"][s])
            y = torch.from_numpy(c["G"][s] if use_group else c["Y"][s])
            o = ma(x.detach(), use_group)
            loss = crit(o, y)
            if s <= 1:
                out[f"{name}/s2_logits{s}"] = o.detach().numpy().copy()
            optim2.zero_grad(); loss.backward(); optim2.step()
            losses2.append(loss.item())
        out[f"{name}/s2_losses"] = np.array(losses2, np.float64)
        for tag, ad in (("old", ma.old_cls.adapter), ("new", ma.new_adapter)):
            for k, v in cases.param_digest(_dump_from(ad)).items():
                out[f"{name}/s2_final_{tag}/{k}"] = v
        ma.eval()
        with torch.no_grad():
            le = ma(xe); lsp = ma.forward_spurious(xe)
            out[f"{name}/s2_eval_logits"] = le.numpy().copy()
            out[f"{name}/s2_eval_logits_spurious"] = lsp.numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "kernel_cases.npz"), **out)
    print("kernel_cases.npz:", len(out), "arrays")


class _ListLoader(list):
    """Stands in for a DataLoader: the reference loops only need iteration, len() and .dataset.n_groups."""

    def __init__(self, batches, n_groups=4):
        super().__init__(batches)
        self.dataset = types.SimpleNamespace(n_groups=n_groups)


class _PassThrough(torch.nn.Module):
    """classifier(embeddings) -> the first C columns: lets a test dictate the logits."""

    def __init__(self, C):
        super().__init__()
        self.C = C

    def forward(self, x, use_group=False):
        return x[:, :self.C]


def gen_metrics(fm, ru):
    from functools import partial
    rng = np.random.default_rng(77)
    get_yp = partial(fm.get_y_p, n_places=2)
    opt = types.SimpleNamespace()
    crit = torch.nn.CrossEntropyLoss()
    out = {}
    specs = {
        "even": dict(sizes=[64, 64, 64], C=2, groups=4),
        "ragged_missing_group": dict(sizes=[50, 7, 1, 33], C=2, groups=4, drop_group=2),
        "ties": dict(sizes=[40, 40], C=2, groups=4, ties=True),
        "four_way": dict(sizes=[128, 100], C=4, groups=4),
        "celeba_like": dict(sizes=[512] * 5 + [123], C=2, groups=4, skew=True),
    }
    for name, sp in specs.items():
        batches, raw = [], []
        for n in sp["sizes"]:
            C = sp["C"]
            logits = (rng.standard_normal((n, C)) * 3).astype(np.float32)
            if sp.get("ties"):
                logits[::3, 1] = logits[::3, 0]
            pr = [0.44, 0.41, 0.14, 0.01] if sp.get("skew") else [0.25] * 4
            g = rng.choice(4, n, p=pr).astype(np.int64)
            if "drop_group" in sp:
                g[g == sp["drop_group"]] = 0
            y = (g // 2) if C == 2 else g
            batches.append((torch.from_numpy(logits), {"class": torch.from_numpy(y), "group": torch.from_numpy(g)},
                            ["f"] * n))
            raw.append(dict(logits=logits.tolist(), y=y.tolist(), g=g.tolist()))
        ratio = torch.tensor([0.7295, 0.0384, 0.0117, 0.2204])
        loss_avg, acc_avg, group_acc = fm.validate(opt, _ListLoader(batches), _PassThrough(sp["C"]), crit, get_yp,
                                                   ratio, target="class", print_label=name)
        out[name] = dict(batches=raw, train_group_ratio=[float(t) for t in ratio.numpy()],
                         loss_avg=float(loss_avg), acc_avg=float(acc_avg),
                         group_acc={k: float(v) for k, v in group_acc.items()})
    with open(os.path.join(GOLD, "metrics_cases.json"), "w") as f:
        json.dump(out, f)
    print("metrics_cases.json:", list(out))


def gen_schedule(fm, ru):
    out = {}
    configs = {
        "waterbirds": dict(dataset="waterbirds", learning_rate=1.0, learning_rate_reg=1.0, epochs=100,
                           epochs_feature_learning=40, lr_decay_epochs=[90, 95], lr_decay_rate=0.1, n_train=5, n_reg=3),
        "celeba": dict(dataset="celeba", learning_rate=0.1, learning_rate_reg=1.0, epochs=65,
                       epochs_feature_learning=40, lr_decay_epochs=[62, 64], lr_decay_rate=0.1, n_train=159, n_reg=91),
    }
    for name, cfg in configs.items():
        opt = types.SimpleNamespace(cosine=False, warm=False, warm_reg=True, **cfg)
        opt.warmup_from_reg = opt.learning_rate_reg / 1e2          # final_main.py:272-284
        opt.warm_epochs_reg = 2 if opt.dataset == "celeba" else 10
        opt.warmup_to_reg = opt.learning_rate_reg
        o1 = types.SimpleNamespace(param_groups=[{"lr": None}])
        o2 = types.SimpleNamespace(param_groups=[{"lr": None}])
        lrs = []
        for epoch in range(1, opt.epochs + 1):
            ru.adjust_learning_rate(opt, o1, epoch)
            if epoch <= opt.epochs_feature_learning:
                per_batch = []
                for b in range(cfg["n_train"]):
                    ru.warmup_learning_rate(opt, epoch, b, cfg["n_train"], o1)
                    per_batch.append(o1.param_groups[0]["lr"])
            else:
                ru.adjust_learning_rate_reg(opt, o2, epoch)
                per_batch = []
                for b in range(cfg["n_reg"]):
                    ru.warmup_learning_rate_reg(opt, epoch - opt.epochs_feature_learning, b, cfg["n_reg"], o2)
                    per_batch.append(float(o2.param_groups[0]["lr"]))
            lrs.append([float(t) for t in per_batch])
        out[name] = dict(config={k: v for k, v in cfg.items()}, lrs=lrs)
    with open(os.path.join(GOLD, "schedule_cases.json"), "w") as f:
        json.dump(out, f)
    print("schedule_cases.json:", list(out))


def gen_sampling(fm, ru):
    sys.path.insert(0, "/root/reference")
    from data.waterbirds_embeddings_reg import stratified_split_dataset
    out = {}
    for name, sizes in (("waterbirds", synth.WATERBIRDS_GROUPS[1]), ("celeba", synth.CELEBA_GROUPS[1])):
        rng = np.random.default_rng(5)
        g = np.concatenate([np.full(n, i, dtype=np.int64) for i, n in enumerate(sizes)])
        g = g[rng.permutation(len(g))]
        ds = types.SimpleNamespace(group_array=g, n_groups=4)
        reg_subset, val_subset = stratified_split_dataset(ds)
        out[f"{name}/group_array"] = g.astype(np.int8)
        out[f"{name}/reg_idx"] = np.asarray(reg_subset.indices, np.int32)
        out[f"{name}/val_idx"] = np.asarray(val_subset.indices, np.int32)
        # balance_val draws from the global numpy RNG, seeded once by set_seed (demo/util.py:66)
        np.random.seed(42)
        loader = types.SimpleNamespace(dataset=reg_subset)
        reg_subset.dataset = ds
        for bsr in (4, 256, 100000):
            opt = types.SimpleNamespace(batch_size_reg=bsr)
            for ep in range(3):
                bl = fm.balance_val(loader, opt)
                out[f"{name}/balanced_bsr{bsr}_ep{ep}"] = np.asarray(bl.dataset.indices, np.int32)
                out[f"{name}/balanced_bsr{bsr}_ep{ep}_bs"] = np.int64(bl.batch_size)
    np.savez_compressed(os.path.join(GOLD, "sampling_cases.npz"), **out)
    print("sampling_cases.npz:", len(out), "arrays")


def gen_supcon(fm, ru):
    sys.path.insert(0, "/root/reference/demo")
    # demo/visualizer_supcon.py imports matplotlib/umap/easydict at module level (absent here): load only the class.
    src = open("/root/reference/demo/visualizer_supcon.py").read()
    start = src.index("class SupervisedContrastiveLoss")
    end = src.index("def skim_dataloader_by_group")
    ns = {"nn": torch.nn, "torch": torch}
    exec(compile(src[start:end], "visualizer_supcon_excerpt", "exec"), ns)   # executes reference code in place
    Loss = ns["SupervisedContrastiveLoss"]
    out = {}
    rng = np.random.default_rng(99)
    for name, (P, N, d) in {"p4n4": (4, 4, 32), "p7n12": (7, 12, 128), "p1n30": (1, 30, 64)}.items():
        feats = rng.standard_normal((1 + P + N, d)).astype(np.float32)
        feats /= np.linalg.norm(feats, axis=1, keepdims=True)
        args = types.SimpleNamespace(cl_temperature=0.1, num_positive=P, num_negative=N, tl_method="contrastive_adapter")
        model = types.SimpleNamespace(forward_ca=lambda f: f)
        loss = Loss(args)(model, torch.from_numpy(feats))[0]
        out[name] = dict(P=P, N=N, d=d, seed=99, loss=float(loss.item()), feats=feats.tolist())
    with open(os.path.join(GOLD, "supcon_cases.json"), "w") as f:
        json.dump(out, f)
    print("supcon_cases.json:", {k: v["loss"] for k, v in out.items()})


def gen_export(fm, ru):
    """Adapted-embedding export (demo/demo_visualization.ipynb:1117-1215, validate_adapter_with_return): its feature lines
    run on the reference's own modules in eval mode.  The notebook cannot be imported, so the three statements that form
    `image_features` and the two logit lines are executed here verbatim on final_main's classes."""
    out = {}
    tmp = tempfile.mkdtemp(prefix="dbmm_gold_")
    for name in ("tiny_b33", "vitl_b256"):
        c = cases.make_case(name)
        D, H = c["D"], c["H"]
        tc, ts, tg = (os.path.join(tmp, f"{name}_{k}.json") for k in ("class", "spurious", "group"))
        _write_text_json(tc, c["T_class"]); _write_text_json(ts, c["T_spurious"]); _write_text_json(tg, c["T_group"])
        clf = fm.CustomCLIP(fm.Adapter(D, H), tc, ts, tg, temperature=0.01)
        _load_into(clf.adapter, c["p_old"])
        new_ad = fm.Adapter(D, H)
        _load_into(new_ad, c["p_new"])
        ma = fm.MultipleAdapter(clf, new_ad, init_near_identity=False, ebd_weight=0.5)
        rng = np.random.default_rng(5)
        for ad in (clf.adapter, new_ad):                      # non-trivial running statistics
            bn = ad.layers[1]
            bn.running_mean.copy_(torch.from_numpy(rng.standard_normal(H).astype(np.float32) * 0.3))
            bn.running_var.copy_(torch.from_numpy((0.5 + rng.random(H)).astype(np.float32)))
        out[f"{name}/running"] = np.stack([clf.adapter.layers[1].running_mean.numpy(), clf.adapter.layers[1].running_var.numpy(),
                                           new_ad.layers[1].running_mean.numpy(), new_ad.layers[1].running_var.numpy()])
        clf.eval(); ma.eval()
        with torch.no_grad():
            embeddings = torch.from_numpy(c["Xe"][:96])
            for tag, classifier, single in (("adapter", clf, True), ("multi", ma, False)):
                if single:
                    image_features = classifier.adapter(embeddings)
                else:
                    old_image_features = classifier.old_cls.adapter(embeddings)
                    old_image_features = old_image_features / old_image_features.norm(dim=-1, keepdim=True)
                    new_image_features = classifier.new_adapter(embeddings)
                    new_image_features = new_image_features / new_image_features.norm(dim=-1, keepdim=True)
                    image_features = classifier.ebd_weight * old_image_features + (1 - classifier.ebd_weight) * new_image_features
                text_features_normalized = classifier.text_features / classifier.text_features.norm(dim=0, keepdim=True)
                logits = image_features @ text_features_normalized / classifier.temperature
                tsn = classifier.text_spurious_features / classifier.text_spurious_features.norm(dim=0, keepdim=True)
                logits_spurious = image_features @ tsn / classifier.temperature
                out[f"{name}/{tag}/features"] = image_features.numpy().copy()
                out[f"{name}/{tag}/logits"] = logits.numpy().copy()
                out[f"{name}/{tag}/logits_spurious"] = logits_spurious.numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "export_cases.npz"), **out)
    print("export_cases.npz:", len(out), "arrays")


def gen_checkpoint_layout(fm, ru):
    path = [p for p in os.listdir("/root/reference/trained_model") if p.endswith(".pth")][0]
    sd = torch.load(os.path.join("/root/reference/trained_model", path), map_location="cpu")
    layout = {k: dict(shape=list(v.shape), dtype=str(v.dtype)) for k, v in sd.items()}
    jpath = [p for p in os.listdir("/root/reference/trained_model") if p.endswith(".json")][0]
    res = json.load(open(os.path.join("/root/reference/trained_model", jpath)))

    def schema(d):
        return {k: schema(v) for k, v in d.items()} if isinstance(d, dict) else type(d).__name__
    res_schema = schema(res)
    # the per-epoch block repeats 100x; keep one epoch of it
    first = sorted(res_schema["All Results (all epoch)"])[0]
    res_schema["All Results (all epoch)"] = {first: res_schema["All Results (all epoch)"][first],
                                             "_n_epochs": len(res["All Results (all epoch)"])}
    with open(os.path.join(GOLD, "checkpoint_layout.json"), "w") as f:
        json.dump(dict(file=path, state_dict=layout, results_file=jpath, results_schema=res_schema), f, indent=1)
    print("checkpoint_layout.json:", len(layout), "keys")


# ------------------------------------------------------------------------------------------------------
# whole-run fixtures: the reference's train_all_epochs on synthetic files, same seed as the test will use
# ------------------------------------------------------------------------------------------------------
E2E_CASES = {
    "waterbirds_small": dict(
        synth=dict(name="waterbirds", dim=1024, seed=1234, scale=0.5, k=0.085, k_text=0.5, text_noise=0.02),
        argv=["--dataset", "waterbirds", "--tl_method", "adapter_reg_seq_alter", "--add_adapter", "--warm_reg",
              "--batch_size", "256", "--batch_size_reg", "64", "--learning_rate", "1.0", "--learning_rate_reg", "1.0",
              "--epochs", "12", "--epochs_feature_learning", "6", "--lr_decay_rate", "0.1", "--lr_decay_epochs", "10,11",
              "--train_target", "class", "--save_results", "--random_seed", "42"]),
    "celeba_small_balval": dict(
        synth=dict(name="celeba", dim=1024, seed=4321, k=0.085, k_text=0.5, text_noise=0.02,
                   group_sizes=[[1400, 1300, 450, 60], [340, 330, 115, 24], [390, 300, 100, 24]]),
        argv=["--dataset", "celeba", "--tl_method", "adapter_reg_seq_alter", "--add_adapter", "--warm_reg", "--balance_val",
              "--batch_size", "512", "--batch_size_reg", "4", "--learning_rate", "0.1", "--learning_rate_reg", "1.0",
              "--epochs", "10", "--epochs_feature_learning", "5", "--lr_decay_rate", "0.1", "--lr_decay_epochs", "8,9",
              "--train_target", "class", "--save_results", "--random_seed", "32"]),
    # --tl_method adapter_reg (train_reg_one_epoch, final_main.py:498-569): one optimizer, a pass over the train loader with
    # class prompts then a pass over the reg loader with group prompts (GP) or class prompts (CP) in every epoch
    "waterbirds_small_reg_gp": dict(
        synth=dict(name="waterbirds", dim=1024, seed=1234, scale=0.5, k=0.085, k_text=0.5, text_noise=0.02),
        argv=["--dataset", "waterbirds", "--tl_method", "adapter_reg", "--batch_size", "256", "--batch_size_reg", "64",
              "--learning_rate", "0.5", "--epochs", "6", "--lr_decay_rate", "0.1", "--lr_decay_epochs", "5,6",
              "--train_target", "class", "--save_results", "--random_seed", "42"]),
    "waterbirds_small_reg_cp": dict(
        synth=dict(name="waterbirds", dim=1024, seed=1234, scale=0.5, k=0.085, k_text=0.5, text_noise=0.02),
        argv=["--dataset", "waterbirds", "--tl_method", "adapter_reg", "--use_cls_prompt_in_reg", "--balance_val",
              "--batch_size", "256", "--batch_size_reg", "64", "--learning_rate", "0.5", "--epochs", "6",
              "--lr_decay_rate", "0.1", "--lr_decay_epochs", "5,6", "--train_target", "class", "--save_results",
              "--random_seed", "22"]),
    # BASELINE.json configs[0] at full size, the reference's own launch script (run_final_main.sh:1-31): 4,795 train rows,
    # bs 1024 / bsr 256, lr 1.0 / 1.0, 100 epochs of which 40 feature learning, decay 0.1 at 90 and 95
    "waterbirds_full": dict(
        synth=dict(name="waterbirds", dim=1024, seed=1234, scale=1.0, k=0.085, k_text=0.5, text_noise=0.02),
        argv=["--dataset", "waterbirds", "--tl_method", "adapter_reg_seq_alter", "--add_adapter", "--warm_reg",
              "--batch_size", "1024", "--batch_size_reg", "256", "--learning_rate", "1.0", "--learning_rate_reg", "1.0",
              "--epochs", "100", "--epochs_feature_learning", "40", "--lr_decay_rate", "0.1", "--lr_decay_epochs", "90,95",
              "--train_target", "class", "--save_results", "--random_seed", "42"]),
}


def e2e_paths_argv(case, root):
    ds = synth.make_dataset(**case["synth"])
    paths = synth.write_reference_files(ds, root)
    argv = list(case["argv"])
    for k, v in paths.items():
        argv += [f"--{k}", v]
    return ds, argv


def gen_e2e(fm, ru, only_cases=None):
    import io
    import contextlib
    path = os.path.join(GOLD, "e2e_cases.json")
    out = json.load(open(path)) if (only_cases and os.path.exists(path)) else {}
    for name, case in E2E_CASES.items():
        if only_cases and name not in only_cases:
            continue
        root = tempfile.mkdtemp(prefix=f"dbmm_e2e_{name}_")
        ds, argv = e2e_paths_argv(case, root)
        old_argv = sys.argv
        sys.argv = ["final_main.py"] + argv
        try:
            opt = fm.parse_option()
        finally:
            sys.argv = old_argv
        # record what train_all_epochs computes per epoch by wrapping the reference's own epoch functions
        log = dict(train=[], val_test=[], zs=[])
        t1, t2, t3, tv, tz = fm.train_one_epoch, fm.train_reg_seq_one_epoch, fm.train_reg_one_epoch, fm.validate, fm.validate_zs

        def wrap(fn, key):
            def inner(*a, **k):
                r = fn(*a, **k)
                log[key].append(dict(label=k.get("print_label", ""), loss=float(r[0]), acc=float(r[1]),
                                     group_acc={kk: float(vv) for kk, vv in r[2].items()}))
                return r
            return inner
        fm.train_one_epoch, fm.train_reg_seq_one_epoch, fm.train_reg_one_epoch = wrap(t1, "train"), wrap(t2, "train"), wrap(t3, "train")
        fm.validate, fm.validate_zs = wrap(tv, "val_test"), wrap(tz, "zs")
        buf = io.StringIO()
        try:
            with contextlib.redirect_stdout(buf):
                res = fm.train_all_epochs(opt)
        finally:
            fm.train_one_epoch, fm.train_reg_seq_one_epoch, fm.train_reg_one_epoch, fm.validate, fm.validate_zs = t1, t2, t3, tv, tz
        stdout = buf.getvalue()
        best_epoch = int([ln for ln in stdout.splitlines() if ln.startswith("best epoch")][0].split(":")[1])
        folder = os.path.dirname(opt.image_embedding_dir).replace("data", "results")
        files = sorted(os.listdir(folder))
        sd = torch.load(os.path.join(folder, [f for f in files if f.endswith(".pth")][0]), map_location="cpu")
        results_json = json.load(open(os.path.join(folder, [f for f in files if f.endswith(".json")][0])))
        out[name] = dict(
            synth=case["synth"], argv=case["argv"], best_epoch=best_epoch, files=files,
            train=log["train"], val_test=log["val_test"], zs=log["zs"],
            final={k: float(v) for part in res[0] for k, v in part.items()} and
                  [{k: float(v) for k, v in part.items()} for part in res[0]],
            state_dict_keys=list(sd.keys()),
            state_dict_abs_sums={k: float(v.double().abs().sum()) for k, v in sd.items()},
            results_json=results_json)
        print(name, "best epoch", best_epoch, "val/test dicts", len(log["val_test"]), files)
    with open(path, "w") as f:
        json.dump(out, f)
    print("e2e_cases.json written")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="all")
    ap.add_argument("--e2e-cases", default="", help="comma-separated subset of E2E_CASES to (re)generate; the others are kept")
    a = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    fm, ru = import_reference()
    torch.set_num_threads(8)
    todo = dict(kernels=gen_kernels, metrics=gen_metrics, schedule=gen_schedule, sampling=gen_sampling,
                supcon=gen_supcon, layout=gen_checkpoint_layout, e2e=gen_e2e, export=gen_export)
    for k, fn in todo.items():
        if a.only in ("all", k):
            if k == "e2e" and a.e2e_cases:
                fn(fm, ru, only_cases=a.e2e_cases.split(","))
            else:
                fn(fm, ru)


if __name__ == "__main__":
    main()




def contrastive_construction_golden(out_path=None):
    """Golden for debiasing-multi-modal_b200/contrastive.py: the reference's own sampling functions
    (demo/visualizer_supcon.py:1051-1080, 1100-1146 semantics, 1148-1340, 1342-1435, 1437-1467) executed IN PLACE from an excerpt
    of the source file (the module itself cannot be imported here: it drags in matplotlib / umap), on seeded synthetic labels."""
    import types
    src = open(os.path.join("/root/reference", "demo", "visualizer_supcon.py")).read().split("\n")

    def excerpt(lo, hi):
        return "\n".join(src[lo - 1:hi])
    class _NP:                                   # numpy >= 1.24 refuses the ragged np.array(list of per-slice arrays) at :1162
        def __getattr__(self, k):                # (the reference's pinned numpy made an object array of it)
            return getattr(np, k)

        @staticmethod
        def array(x, *a, **k):
            try:
                return np.array(x, *a, **k)
            except ValueError:
                o = np.empty(len(x), dtype=object)
                for i, v in enumerate(x):
                    o[i] = v
                return o
    ns = {"np": _NP(), "tqdm": (lambda it, **k: it), "print": (lambda *a, **k: None)}
    exec(excerpt(1051, 1081), ns)            # adjust_num_pos_neg_
    exec(excerpt(1148, 1341), ns)            # prepare_contrastive_points
    exec(excerpt(1342, 1435), ns)            # construct_contrastive_data
    rng = np.random.default_rng(77)
    n = 1500
    g = rng.choice(4, n, p=[0.55, 0.1, 0.05, 0.3])
    y, sp = g // 2, g % 2
    y_pred = np.where(rng.random(n) < 0.8, y, 1 - y)          # zero-shot prediction: 80 % correct
    # compute_slice_indices (1100-1146) reads a pandas frame of the dataset; its arithmetic on (pseudo_labels, labels):
    correct = y_pred == y
    sl_ix = [np.where(y_pred == lab)[0] for lab in np.unique(y_pred)]
    sl_ok = [correct[ix] for ix in sl_ix]
    ds = types.SimpleNamespace(y_array=y, confounder_array=sp)
    anchors, negatives, positives, _ = ns["prepare_contrastive_points"](ds, sl_ix, sl_ok)
    args = types.SimpleNamespace(num_anchor=2, num_positive=12, num_negative=10, n_cls=2)
    ns["adjust_num_pos_neg_"](positives, negatives, args)
    np.random.seed(123)
    samples = ns["construct_contrastive_data"](anchors, negatives, positives, args)
    groups = np.concatenate(samples)
    np.random.shuffle(groups)                                  # load_contrastive_loader, no balancing, re_shuffle_ca_loader
    out = dict(y=y, spurious=sp, y_pred=y_pred, groups=groups, adjusted=np.array([args.num_anchor, args.num_positive, args.num_negative]),
               neg_counts=np.array([len(d["ix"]) for d in negatives]), pos_counts=np.array([len(positives[c]["ix"]) for c in range(2)]),
               anchor_counts=np.array([len(a["ix"]) for a in anchors]))
    out_path = out_path or os.path.join(GOLD, "contrastive_construction.npz")
    np.savez_compressed(out_path, **out)
    return out
