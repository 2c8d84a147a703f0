"""Command line and training driver: same flags, defaults, derived fields, epoch schedule, model selection
and .json / .pth artefacts as the reference's final_main.py (parse_option 176-297, set_model 300-343,
train_all_epochs 805-1128), re-written around the fused kernels.
"""
from __future__ import annotations

import argparse
import json
import math
import os
from copy import deepcopy
from functools import partial

import numpy as np
import torch

from . import data as D
from . import engine as E
from .modules import Adapter, CustomCLIP, LinearClassifier, MultipleAdapter
from .optim import adjust_learning_rate, adjust_learning_rate_reg, set_optimizer, set_optimizer_reg

# embedding width per encoder; the reference only knows 'resnet50' (final_main.py:31), ViT-L/14 is an extension
model_dict = {"resnet50": [None, 1024], "vit_l14": [None, 768]}
REG_METHODS = ("adapter_reg", "adapter_reg_seq", "adapter_reg_seq_alter")


def build_parser():
    p = argparse.ArgumentParser("argument for training")
    p.add_argument("--print_freq", type=int, default=10, help="print frequency")
    p.add_argument("--save_freq", type=int, default=50, help="save frequency (accepted, unused as in the reference)")
    p.add_argument("--batch_size", type=int, default=128, help="batch_size")
    p.add_argument("--batch_size_reg", type=int, default=128, help="batch_size for adapter_reg")
    p.add_argument("--num_workers", type=int, default=16, help="accepted for compatibility; data is GPU-resident")
    p.add_argument("--epochs", type=int, default=10, help="number of training epochs")
    p.add_argument("--learning_rate", type=float, default=1e-1, help="learning rate")
    p.add_argument("--learning_rate_reg", type=float, default=1e-3, help="learning rate in stage 2")
    p.add_argument("--lr_decay_epochs", type=str, default="60,75,90", help="where to decay lr, can be a list")
    p.add_argument("--lr_decay_rate", type=float, default=1, help="decay rate for learning rate")
    p.add_argument("--weight_decay", type=float, default=5e-5, help="weight decay")
    p.add_argument("--momentum", type=float, default=0.9, help="momentum")
    p.add_argument("--model", type=str, default="resnet50")
    p.add_argument("--dataset", type=str, default="waterbirds", choices=["celeba", "waterbirds"], help="dataset")
    p.add_argument("--cosine", action="store_true", help="using cosine annealing")
    p.add_argument("--warm", action="store_true", help="warm-up for large batch training")
    p.add_argument("--warm_reg", action="store_true", help="warm-up for stable stage-2 training")
    p.add_argument("--image_embedding_dir", type=str, help="extracted image embedding")
    p.add_argument("--text_embedding_dir", type=str, help="extracted text embedding")
    p.add_argument("--text_group_embedding_dir", type=str, help="extracted group embedding")
    p.add_argument("--text_spurious_embedding_dir", type=str, help="extracted text embedding (spurious attributes)")
    p.add_argument("--train_target", type=str, default="class", choices=["class", "spurious", "group"])
    p.add_argument("--data_dir", type=str, help="folder in which metadata.csv exists")
    p.add_argument("--tl_method", type=str, default="linear_probing",
                   choices=["linear_probing", "adapter", "adapter_reg", "adapter_reg_seq", "adapter_reg_seq_alter",
                            "contrastive_adapter"], help="transfer learning method")
    # contrastive_adapter (demo/visualizer_supcon.py; the reference's final_main accepts the method name only)
    p.add_argument("--ca_head", type=str, default=None, help="projection head of forward_ca (only None = identity is built)")
    p.add_argument("--contrastive_weight", type=float, default=0.1, help="weight of the contrastive loss (visualizer_supcon.py:255)")
    p.add_argument("--cl_temperature", type=float, default=0.1, help="contrastive temperature (visualizer_supcon.py:250)")
    p.add_argument("--num_anchor", type=int, default=1)
    p.add_argument("--num_positive", type=int, default=64)
    p.add_argument("--num_negative", type=int, default=64)
    p.add_argument("--batch_factor", type=int, default=8, help="anchor groups per contrastive loader batch")
    p.add_argument("--ca_update", type=int, default=1 << 30, help="contrastive loader batches used per epoch")
    p.add_argument("--balance_by_zs_pred", action="store_true")
    p.add_argument("--no_ca_pre_norm", action="store_true", help="skip the input normalisation of forward_ca")
    p.add_argument("--balance_val", action="store_true", help="Balancing Val-reg loader.")
    p.add_argument("--resample_ce", action="store_true", help="accepted; only tags the result file name (reference no-op)")
    p.add_argument("--use_cls_prompt_in_reg", action="store_true", help="use class prompts in regularization")
    p.add_argument("--add_adapter", action="store_true", default=False, help="additional adapter in stage 2")
    p.add_argument("--init_near_identity", action="store_true", help="initialise the new adapter from the old one")
    p.add_argument("--epochs_feature_learning", type=int, help="epochs of stage 1 in 'adapter_reg_seq*'")
    p.add_argument("--continue_from_best", action="store_true", help="stage 2 starts from the best worst-acc model")
    p.add_argument("--adapter_feat_dim", type=int, default=128, help="reduced dimension in adapter")
    p.add_argument("--zs_temperature", type=float, default=0.01, help="temperature in zero-shot prediction")
    p.add_argument("--watch_batch_results", action="store_true", help="print per-batch results every print_freq")
    p.add_argument("--save_results", action="store_true", help="save results (.json) and best model (.pth)")
    p.add_argument("--random_seed", type=int, default=42, help="random seed")
    return p


def finalize_options(opt):
    """Derived fields and consistency checks of parse_option (final_main.py:253-297)."""
    E.set_seed(opt.random_seed)
    if isinstance(opt.lr_decay_epochs, str):
        opt.lr_decay_epochs = [int(t) for t in opt.lr_decay_epochs.split(",")]
    if opt.warm:
        opt.warmup_from = 0.01
        opt.warm_epochs = 10
        if opt.cosine:
            eta_min = opt.learning_rate * (opt.lr_decay_rate ** 3)
            opt.warmup_to = eta_min + (opt.learning_rate - eta_min) * (
                1 + math.cos(math.pi * opt.warm_epochs / opt.epochs)) / 2
        else:
            opt.warmup_to = opt.learning_rate
    if opt.warm_reg:
        opt.warmup_from_reg = opt.learning_rate_reg / 1e2
        opt.warm_epochs_reg = 2 if opt.dataset == "celeba" else 10
        if opt.cosine:
            eta_min = opt.learning_rate_reg * (opt.lr_decay_rate ** 3)
            opt.warmup_to_reg = eta_min + (opt.learning_rate_reg - eta_min) * (
                1 + math.cos(math.pi * opt.warm_epochs_reg / (opt.epochs - opt.epochs_feature_learning))) / 2
        else:
            opt.warmup_to_reg = opt.learning_rate_reg
    if opt.dataset not in ("celeba", "waterbirds"):
        raise ValueError("dataset not supported: {}".format(opt.dataset))
    opt.n_cls = 2
    if opt.tl_method != "linear_probing" and not (4 <= opt.adapter_feat_dim <= 128 and opt.adapter_feat_dim % 4 == 0):
        # the reference accepts any width (final_main.py:241); the fused kernels hold one hidden vector per warp slot set
        raise ValueError(f"--adapter_feat_dim {opt.adapter_feat_dim}: the B200 kernels support multiples of 4 up to 128 "
                         "(DBMM_MAX_H, include/dbmm.h); the reference's default is 128")
    if opt.tl_method == "adapter":
        assert not opt.add_adapter
        assert not opt.balance_val
    return opt


def parse_option(argv=None):
    return finalize_options(build_parser().parse_args(argv))


def set_model(opt):
    criterion = torch.nn.CrossEntropyLoss()
    _, input_dim = model_dict[opt.model]
    if opt.tl_method == "linear_probing":
        print("Off-the-shelf classifier : [Linear Classifier]")
        classifier = LinearClassifier(input_dim=input_dim, num_classes=opt.n_cls)
    elif opt.tl_method in ("adapter", "contrastive_adapter") or opt.tl_method in REG_METHODS:
        if opt.tl_method == "contrastive_adapter" and opt.ca_head not in (None, "none"):
            raise NotImplementedError("--ca_head linear / mlp is not built: forward_ca uses the adapter output itself (head = identity)")
        tail = " with group regularized training" if opt.tl_method in REG_METHODS else ""
        print("Off-the-shelf classifier : [Adapter + (temperatured) image-text jointly normalized prediction]" + tail)
        adapter = Adapter(input_dim=input_dim, hidden_dim=opt.adapter_feat_dim)
        classifier = CustomCLIP(adapter, opt.text_embedding_dir, opt.text_spurious_embedding_dir,
                                opt.text_group_embedding_dir, temperature=opt.zs_temperature)
    else:
        raise NotImplementedError(f"--tl_method {opt.tl_method} has no model in the reference either (final_main.py:306-316)")
    return classifier.cuda(), criterion.cuda()


def set_model_multiple_adapter(opt, erm_classifier):
    criterion = torch.nn.CrossEntropyLoss()
    _, input_dim = model_dict[opt.model]
    assert opt.tl_method in ["adapter_reg_seq", "adapter_reg_seq_alter"]
    print("================== Stage 2) New adapter for Balanced-Text-Prompt ==================")
    new_adapter = Adapter(input_dim=input_dim, hidden_dim=opt.adapter_feat_dim)
    new_classifier = MultipleAdapter(erm_classifier, new_adapter, init_near_identity=opt.init_near_identity, ebd_weight=0.5)
    return new_classifier.cuda(), criterion.cuda()


def build_loaders(opt):
    """(trainset, train_loader, reg_loader | None, val_loader, test_loader); the JSON is parsed once."""
    reg = opt.tl_method in REG_METHODS
    load = D.load_waterbirds_embeddings if opt.dataset == "waterbirds" else D.load_celeba_embeddings
    print(f"Load image embedding of {opt.dataset}: {opt.image_embedding_dir}")
    print("Load Data Loader (train, validation, test)")
    loaders = load(opt.data_dir, opt.image_embedding_dir, opt.batch_size, opt.batch_size_reg if reg else opt.batch_size,
                   reg=reg)
    if reg:
        train_loader, reg_loader, val_loader, test_loader = loaders
    else:
        (train_loader, val_loader, test_loader), reg_loader = loaders, None
    return train_loader.dataset, train_loader, reg_loader, val_loader, test_loader


def result_file_stem(opt):
    """Result file naming rule of final_main.py:1062-1096."""
    folder = os.path.dirname(opt.image_embedding_dir).replace("data", "results")
    img = os.path.basename(opt.image_embedding_dir).split(".")[0]
    txt = os.path.basename(opt.text_embedding_dir).split(".")[0]
    name = f"im_{img}_t_{txt}_tl_{opt.tl_method}_t_{opt.train_target}_lr_{opt.learning_rate}_bs_{opt.batch_size}"
    if "reg" in opt.tl_method:
        name += f"_lrr{opt.learning_rate_reg}_bsr_{opt.batch_size_reg}"
        if opt.balance_val:
            name += "_balval"
        if opt.tl_method != "adapter_reg_seq_alter":
            name += "_CP" if opt.use_cls_prompt_in_reg else "_GP"
        if opt.add_adapter:
            name += "_MA" + ("+ni" if opt.init_near_identity else "+rn")
        if opt.continue_from_best and ("seq" in opt.tl_method):
            name += "_cont"
    if opt.resample_ce:
        name += "_rs"
    return folder, name


def _to_jsonable(d):
    return {k: float(v) for k, v in d.items()}


def train_all_epochs(opt, loaders=None):
    """final_main.train_all_epochs (805-1128): same arguments, prints, artefacts and return value."""
    return E.drive(train_all_epochs_gen(opt, loaders))


def train_all_epochs_gen(opt, loaders=None):
    """train_all_epochs as a generator over its training epochs: every epoch's prepared TrainJob is yielded instead of run,
    so that the sweep driver can run the jobs of many members in lock step (sweep.py); `train_all_epochs` drives it alone."""
    best_acc, best_epoch, best_model = 0, 0, None
    print(f"> Start Transfer Learning using [{opt.tl_method}]")
    print("========================================================================")
    trainset, train_loader, reg_loader, val_loader, test_loader = loaders if loaders is not None else build_loaders(opt)
    print(f"Training target : {opt.train_target}")
    reg_mode = opt.tl_method in REG_METHODS
    balancing = opt.balance_val and reg_mode
    if balancing:
        print("Using [Balanced] Validation loader for regularized training")
        origin_reg_loader = reg_loader
    if opt.resample_ce:
        print("Using [Resampled] Train loader for erm/feature laerning (reference builds it and never uses it)")

    get_yp_func = partial(E.get_y_p, n_places=trainset.n_places)
    train_group_ratio = trainset.group_ratio

    classifier, criterion = set_model(opt)
    print("Set Optimizer: SGD (default)")
    print("========================================================================")
    optimizer = set_optimizer(opt, classifier)
    multiple_adapter, optimizer_reg = None, None
    train_group_accs, val_group_accs, test_group_accs = [], [], []
    FL = opt.epochs_feature_learning
    ca_batches = None
    if opt.tl_method == "contrastive_adapter":
        # anchor / positive / negative groups from the zero-shot predictions stored with the embeddings, built once with the
        # global numpy RNG (visualizer_supcon.py:1100-1484; contrastive.py)
        from . import contrastive as CA
        base, rows = train_loader.base_rows(np.arange(len(trainset)))
        groups, ca_batches, adj = CA.contrastive_batches(
            base.y_array[rows], base.confounder_array[rows], base.y_pred_array[rows], n_cls=opt.n_cls, num_anchor=opt.num_anchor,
            num_positive=opt.num_positive, num_negative=opt.num_negative, batch_factor=opt.batch_factor,
            balance_by_zs_pred=opt.balance_by_zs_pred)
        ca_batches = [rows[b] for b in ca_batches]                 # positions in the train split -> rows of the resident matrix
        print(f"Contrastive groups: {groups.shape[0]} x (anchors {adj[0]} + positives {adj[1]} + negatives {adj[2]}), "
              f"{len(ca_batches)} loader batches of {opt.batch_factor} groups")

    for epoch in range(1, opt.epochs + 1):
        adjust_learning_rate(opt, optimizer, epoch)
        print(f"--- Epoch {epoch} ---")
        if balancing:
            reg_loader = E.balance_val(origin_reg_loader, opt, print_procedure=False)

        if opt.tl_method == "adapter_reg":
            gp = not opt.use_cls_prompt_in_reg
            _, _, group_acc = yield from E.train_reg_one_epoch_gen(
                opt, train_loader, reg_loader, classifier, criterion, optimizer, epoch, get_yp_func,
                target=opt.train_target, group_prompt=gp,
                print_label="Train (Alternative Learning)(" + ("Group" if gp else "Class") + " prompt)")
        elif opt.tl_method in ("adapter_reg_seq", "adapter_reg_seq_alter"):
            if epoch <= FL:
                _, _, group_acc = yield from E.train_one_epoch_gen(opt, train_loader, classifier, criterion, optimizer, epoch,
                                                                   get_yp_func, target=opt.train_target,
                                                                   print_label="Train-1 (Feature Learning)")
            else:
                if epoch == FL + 1:
                    if opt.continue_from_best:
                        print("Load Best (Worst-acc) Model.")
                        classifier = deepcopy(best_model)
                    if opt.add_adapter:
                        multiple_adapter, criterion = set_model_multiple_adapter(opt, classifier)
                        optimizer_reg = set_optimizer_reg(opt, multiple_adapter)
                    else:
                        optimizer_reg = set_optimizer_reg(opt, classifier)
                adjust_learning_rate_reg(opt, optimizer_reg, epoch)
                if opt.tl_method == "adapter_reg_seq_alter":
                    use_group = (epoch % 2) == 0          # odd epochs: class prompts, even epochs: group prompts
                else:
                    use_group = not opt.use_cls_prompt_in_reg
                model = multiple_adapter if opt.add_adapter else classifier
                label = "Train-2 (Balanced Learning)" + ("(new adapter)" if opt.add_adapter else "") + \
                        ("(Group prompt)" if use_group else "(Class prompt)")
                _, _, group_acc = yield from E.train_reg_seq_one_epoch_gen(opt, reg_loader, model, criterion, optimizer_reg, epoch,
                                                                           get_yp_func, target=opt.train_target, print_label=label,
                                                                           use_group=use_group)
        elif opt.tl_method == "contrastive_adapter":
            E.train_one_epoch_cl(opt, train_loader, ca_batches, classifier, optimizer, epoch, print_label="Train (Contrastive)")
            _, _, group_acc = E.validate(opt, train_loader, classifier, criterion, get_yp_func, train_group_ratio,
                                         target=opt.train_target, print_label="Train (eval pass)")
        else:
            _, _, group_acc = yield from E.train_one_epoch_gen(opt, train_loader, classifier, criterion, optimizer, epoch, get_yp_func,
                                                               target=opt.train_target, print_label=f"Train({opt.train_target})")
        train_group_accs.append(group_acc)

        stage2_ma = bool(opt.add_adapter and FL is not None and epoch > FL)
        eval_model = multiple_adapter if stage2_ma else classifier
        tag = "(new adapter)" if stage2_ma else ""
        _, _, val_group_acc = E.validate(opt, val_loader, eval_model, criterion, get_yp_func, train_group_ratio,
                                         target=opt.train_target, print_label=f"Val({opt.train_target}){tag}")
        val_group_accs.append(val_group_acc)
        if val_group_acc["worst_acc"] > best_acc:        # strict improvement from 0 (final_main.py:1001)
            best_acc = val_group_acc["worst_acc"]
            best_epoch = epoch
            best_model = deepcopy(eval_model)
        _, _, test_group_acc = E.validate(opt, test_loader, eval_model, criterion, get_yp_func, train_group_ratio,
                                          target="class", print_label=f"Test({opt.train_target}){tag}")
        test_group_accs.append(test_group_acc)

    print("========================================================================")
    print("> end of training. \n")
    print("best epoch : {}".format(best_epoch))
    if best_model is None:
        raise RuntimeError("validation worst-group accuracy never rose above 0, so no best model was selected "
                           "(the reference crashes in validate_zs at this point, final_main.py:728)")
    best_train_group_acc = train_group_accs[best_epoch - 1]
    best_val_group_acc = val_group_accs[best_epoch - 1]
    best_test_group_acc = test_group_accs[best_epoch - 1]
    print(f"best training accuracy on [{opt.train_target}]: {best_train_group_acc}")
    print(f"best validation accuracy on [{opt.train_target}]: {best_val_group_acc}")
    print(f"best test accuracy on [{opt.train_target}]: {best_test_group_acc}")

    print("========================================================================")
    print("> start evaluating feature quality of best model. (using zero-shot prediction)\n")
    _, _, zs_group_acc = E.validate_zs(opt, test_loader, best_model, criterion, get_yp_func, train_group_ratio,
                                       target="class", print_label="zero-shot prediction (test) (class)")
    _, _, zs_group_acc_spurious = E.validate_zs(opt, test_loader, best_model, criterion, get_yp_func, train_group_ratio,
                                                target="spurious", print_label="zero-shot prediction (test) (spurious)")
    print("========================================================================")
    if opt.save_results:
        print("> Save results\n")
        all_results = {}
        for epoch in range(1, opt.epochs + 1):
            # "Val" holds the TEST accuracies in the reference's file (final_main.py:1055); kept for schema parity
            all_results[f"Epoch {epoch}"] = {"Train": _to_jsonable(train_group_accs[epoch - 1]),
                                             "Val": _to_jsonable(test_group_accs[epoch - 1]),
                                             "Test": _to_jsonable(test_group_accs[epoch - 1])}
        final_results = {
            "Final Results (best epoch)": {f"Epoch {best_epoch}": {"Train": _to_jsonable(best_train_group_acc),
                                                                   "Val": _to_jsonable(best_val_group_acc),
                                                                   "Test": _to_jsonable(best_test_group_acc)}},
            "Feature Quality (using zs)": {"class": _to_jsonable(zs_group_acc),
                                           "spurious": _to_jsonable(zs_group_acc_spurious)},
            "All Results (all epoch)": all_results}
        folder, stem = result_file_stem(opt)
        os.makedirs(folder, exist_ok=True)
        json_path = os.path.join(folder, stem + ".json")
        model_path = os.path.join(folder, stem + ".pth")
        print("final result path: ", json_path)
        print("final model path: ", model_path)
        with open(json_path, "w") as f:
            json.dump(final_results, f, indent=4)
        torch.save(best_model.state_dict(), model_path)
    print("========================================================================")
    print("> end")
    train_all_epochs.last_run = dict(best_epoch=best_epoch, train=train_group_accs, val=val_group_accs,
                                     test=test_group_accs, best_model=best_model)
    return (best_train_group_acc, best_val_group_acc, best_test_group_acc), (zs_group_acc, zs_group_acc_spurious)
