"""nn.Module surface of the adapter path, same class names, constructor arguments, attribute names
and state_dict keys as the reference (final_main.py:43-174):

    Adapter(input_dim, hidden_dim)                  .layers = Sequential(Linear, BatchNorm1d, ReLU, Linear)
    CustomCLIP(adapter, text_embedding_dir, text_spurious_embedding_dir, text_group_embedding_dir, temperature)
    MultipleAdapter(old_cls, new_adapter, init_near_identity=True, ebd_weight=0.5)
    LinearClassifier(input_dim, num_classes)

The torch modules only OWN the tensors (so state_dict / load_state_dict / deepcopy / .cuda() behave as in
the reference and a checkpoint written here loads into the reference's classes with strict=True).  All
arithmetic goes through libdbmm.so: `forward` in eval mode runs the fused eval kernel; `forward` in train mode is a
torch.autograd.Function over dbmm_train_forward / dbmm_train_backward, so the REFERENCE's own training loop
(output = classifier(x); loss = criterion(output, y); loss.backward(); optimizer.step(), final_main.py:455-466) runs
unchanged on these modules with a stock torch optimizer; the fast path is `engine.train_one_epoch` & co., which run the
whole epoch (forward, CE, backward, SGD) in the fused kernels on the same storage.
"""
from __future__ import annotations

import json

import torch
import torch.nn as nn

from . import ops


def get_text_embedding(text_embedding_dir):
    """{prompt: [D floats]} JSON -> [D, C] tensor, insertion order = class index (final_main.py:414-424)."""
    with open(text_embedding_dir, "r") as f:
        table = json.load(f)
    cols = [torch.tensor(vec) for vec in table.values()]
    feats = torch.stack(cols, dim=1)
    return feats.cuda() if torch.cuda.is_available() else feats


class LinearClassifier(nn.Module):
    """Linear probing head (final_main.py:43-49)."""

    def __init__(self, input_dim, num_classes=2):
        super().__init__()
        self.fc = nn.Linear(input_dim, num_classes)

    def forward(self, features):
        """logits = x W^T + b (eval / no-grad; training runs through engine.train_one_epoch's fused linear-probe step)."""
        x = features.contiguous()
        That = self.fc.weight.data.t().contiguous()               # [D, C]
        return ops.linear_logits(x, That, self.fc.bias.data)


class Adapter(nn.Module):
    """Linear(D,H) -> BatchNorm1d(H) -> ReLU -> Linear(H,D); no residual (final_main.py:160-174)."""

    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Linear(input_dim, hidden_dim),
            nn.BatchNorm1d(hidden_dim),
            nn.ReLU(),
            nn.Linear(hidden_dim, input_dim),
        )

    def tensors(self) -> ops.AdapterTensors:
        """Zero-copy view of the module's storage for the kernels (which update it in place)."""
        lin1, bn, lin2 = self.layers[0], self.layers[1], self.layers[3]
        return ops.AdapterTensors(lin1.weight.data, lin1.bias.data, bn.weight.data, bn.bias.data,
                                  bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                  lin2.weight.data, lin2.bias.data)

    def forward(self, features):
        """The un-normalised adapter output [N, D] (final_main.py:173-174).  Eval mode: the export kernel (running-stat
        BatchNorm), which is what the visualisation notebooks call (`classifier.adapter(embeddings)`,
        demo/demo_visualization.ipynb:1146).  The training loops never need it: the D-wide activation is not
        materialised on the fused path (CustomCLIP / MultipleAdapter .forward return logits straight from the kernels).
        Train mode is kept for interface completeness as the plain torch composition of the four layers."""
        if self.training:
            return self.layers(features)
        out, _, _ = ops.export_embeddings(features.contiguous(), self.tensors())
        return out


class _PromptMixin:
    """Prompt matrices are plain attributes (not in the state_dict), as in the reference; the column
    normalisation of final_main.py:77 is done once per matrix by a kernel and cached."""

    def _init_prompts(self):
        self._that_cache = {}

    def _that(self, kind: str) -> torch.Tensor:
        if kind == "group":
            # the reference re-reads this JSON on every forward (final_main.py:72); read once, same numbers
            if "group_raw" not in self._that_cache:
                self._that_cache["group_raw"] = get_text_embedding(self.text_group_embedding_dir)
            raw = self._that_cache["group_raw"]
        elif kind == "spurious":
            raw = self.text_spurious_features
        else:
            raw = self.text_features
        dev = self._device()
        key = (kind, dev, raw.data_ptr())
        if key not in self._that_cache:
            self._that_cache[key] = ops.normalize_text(raw.to(dev, torch.float32).contiguous())
        return self._that_cache[key]

    def prompt_matrix(self, use_group=False, spurious=False) -> torch.Tensor:
        return self._that("spurious" if spurious else ("group" if use_group else "class"))

    def __deepcopy__(self, memo):
        # best_model = deepcopy(classifier) (final_main.py:1005-1008): caches hold device tensors keyed by pointers
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == "_that_cache" else copy.deepcopy(v, memo)
        return new


class _TrainLogits(torch.autograd.Function):
    """Train-mode forward of CustomCLIP / MultipleAdapter under torch autograd: logits from dbmm_train_forward (batch-stat
    BatchNorm, running statistics moved), parameter gradients from dbmm_train_backward.  Only the trainable adapter's six
    tensors get gradients (the frozen adapter of a MultipleAdapter and the embeddings do not, as in the reference:
    final_main.py:128 `.detach()` on the old features, 455 `embeddings.detach()`)."""

    @staticmethod
    def forward(ctx, x, That, inv_tau, old_ad, w, ad, W1, b1, gamma, beta, W2, b2):
        x = x.detach().contiguous()
        logits = ops.train_forward(x, ad, That, inv_tau, old_ad=old_ad, ebd_weight=w)
        ctx.x, ctx.That, ctx.inv_tau, ctx.old_ad, ctx.w, ctx.ad = x, That, inv_tau, old_ad, w, ad
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        ad = ctx.ad
        g = ops.train_backward(ctx.x, ad, ctx.That, ctx.inv_tau, dlogits.contiguous().float(), old_ad=ctx.old_ad, ebd_weight=ctx.w)
        sl = ops.flat_param_slices(ad.D, ad.H)
        return (None, None, None, None, None, None, g[sl["W1"]].view(ad.H, ad.D), g[sl["b1"]], g[sl["gamma"]], g[sl["beta"]],
                g[sl["W2"]].view(ad.D, ad.H), g[sl["b2"]])


def _train_logits(x, adapter_module, That, inv_tau, old_ad=None, w=0.5):
    lin1, bn, lin2 = adapter_module.layers[0], adapter_module.layers[1], adapter_module.layers[3]
    return _TrainLogits.apply(x, That, inv_tau, old_ad, w, adapter_module.tensors(), lin1.weight, lin1.bias, bn.weight, bn.bias,
                              lin2.weight, lin2.bias)


def _no_label_eval(x, ad, That, inv_tau, old_ad=None, w=0.5):
    x = x.contiguous()
    logits, _ = ops.eval_fwd(x, None, None, ad, That, inv_tau, None, max(1, x.shape[0]), old_ad=old_ad, ebd_weight=w,
                             want_logits=True)
    return logits


class CustomCLIP(nn.Module, _PromptMixin):
    """adapter -> row L2-normalise -> cosine logits / temperature against text prompts (final_main.py:53-92)."""

    def __init__(self, adapter, text_embedding_dir, text_spurious_embedding_dir, text_group_embedding_dir,
                 temperature=0.01):
        super().__init__()
        self.text_embedding_dir = text_embedding_dir
        self.text_spurious_embedding_dir = text_spurious_embedding_dir
        self.text_group_embedding_dir = text_group_embedding_dir
        self.adapter = adapter
        self.temperature = temperature
        self.text_features = get_text_embedding(self.text_embedding_dir)
        self.n_cls = self.text_features.shape[0]
        self.text_spurious_features = get_text_embedding(self.text_spurious_embedding_dir)
        self._init_prompts()

    def _device(self):
        return self.adapter.layers[0].weight.device

    def kernel_adapters(self):
        """(old adapter or None, trainable adapter, ebd_weight) for the kernels."""
        return None, self.adapter.tensors(), 0.5

    def forward(self, features, use_group=False):
        if self.training:      # the reference's own loop (final_main.py:455-466) runs on this: autograd over the kernels
            return _train_logits(features, self.adapter, self.prompt_matrix(use_group), 1.0 / self.temperature)
        return _no_label_eval(features, self.adapter.tensors(), self.prompt_matrix(use_group), 1.0 / self.temperature)

    def forward_spurious(self, features):
        if self.training:
            return _train_logits(features, self.adapter, self.prompt_matrix(spurious=True), 1.0 / self.temperature)
        return _no_label_eval(features, self.adapter.tensors(), self.prompt_matrix(spurious=True), 1.0 / self.temperature)


class MultipleAdapter(nn.Module, _PromptMixin):
    """Frozen stage-1 classifier + trainable second adapter; logits from the 0.5/0.5 mix of the two
    normalised outputs, mix not re-normalised (final_main.py:97-158)."""

    def __init__(self, old_cls, new_adapter, init_near_identity=True, ebd_weight=0.5):
        super().__init__()
        self.old_cls = old_cls
        self.text_embedding_dir = self.old_cls.text_embedding_dir
        self.text_spurious_embedding_dir = self.old_cls.text_spurious_embedding_dir
        self.text_group_embedding_dir = self.old_cls.text_group_embedding_dir
        self.text_features = get_text_embedding(self.text_embedding_dir)
        self.n_cls = self.text_features.shape[0]
        self.text_spurious_features = get_text_embedding(self.text_spurious_embedding_dir)
        self.new_adapter = new_adapter
        self.ebd_weight = ebd_weight
        if init_near_identity:
            print("Initialize paramters of [New adapter] from [Old adapter]")
            self.new_adapter.load_state_dict(self.old_cls.adapter.state_dict())
        self.temperature = self.old_cls.temperature
        self._init_prompts()

    def _device(self):
        return self.new_adapter.layers[0].weight.device

    def kernel_adapters(self):
        return self.old_cls.adapter.tensors(), self.new_adapter.tensors(), float(self.ebd_weight)

    def forward(self, features, use_group=False):
        old, new, w = self.kernel_adapters()
        if self.training:      # both adapters run batch-stat BatchNorm, only the new one gets gradients (final_main.py:121-140)
            return _train_logits(features, self.new_adapter, self.prompt_matrix(use_group), 1.0 / self.temperature, old_ad=old, w=w)
        return _no_label_eval(features, new, self.prompt_matrix(use_group), 1.0 / self.temperature, old_ad=old, w=w)

    def forward_spurious(self, features):
        old, new, w = self.kernel_adapters()
        if self.training:
            return _train_logits(features, self.new_adapter, self.prompt_matrix(spurious=True), 1.0 / self.temperature, old_ad=old, w=w)
        return _no_label_eval(features, new, self.prompt_matrix(spurious=True), 1.0 / self.temperature, old_ad=old, w=w)
