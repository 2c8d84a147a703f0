"""nn.Module surface of the adapter path, same class names, constructor arguments, attribute names
and state_dict keys as the reference (final_main.py:43-174):

    Adapter(input_dim, hidden_dim)                  .layers = Sequential(Linear, BatchNorm1d, ReLU, Linear)
    CustomCLIP(adapter, text_embedding_dir, text_spurious_embedding_dir, text_group_embedding_dir, temperature)
    MultipleAdapter(old_cls, new_adapter, init_near_identity=True, ebd_weight=0.5)
    LinearClassifier(input_dim, num_classes)

The torch modules only OWN the tensors (so state_dict / load_state_dict / deepcopy / .cuda() behave as in
the reference and a checkpoint written here loads into the reference's classes with strict=True).  All
arithmetic goes through libdbmm.so: `forward` in eval mode runs the fused eval kernel, training runs through
`engine.train_one_epoch` & co. which call the fused train-step kernels on the same storage.
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
        raise NotImplementedError("the un-normalised D-wide adapter output is never materialised on the B200 path; "
                                  "call CustomCLIP / MultipleAdapter forward (logits) instead")


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
        if self.training:
            raise NotImplementedError("train-mode forward goes through engine.train_one_epoch (fused step)")
        return _no_label_eval(features, self.adapter.tensors(), self.prompt_matrix(use_group), 1.0 / self.temperature)

    def forward_spurious(self, features):
        if self.training:
            raise NotImplementedError("train-mode forward goes through engine.train_one_epoch (fused step)")
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
        if self.training:
            raise NotImplementedError("train-mode forward goes through engine.train_reg_seq_one_epoch (fused step)")
        old, new, w = self.kernel_adapters()
        return _no_label_eval(features, new, self.prompt_matrix(use_group), 1.0 / self.temperature, old_ad=old, w=w)

    def forward_spurious(self, features):
        if self.training:
            raise NotImplementedError("train-mode forward goes through engine.train_reg_seq_one_epoch (fused step)")
        old, new, w = self.kernel_adapters()
        return _no_label_eval(features, new, self.prompt_matrix(spurious=True), 1.0 / self.temperature, old_ad=old, w=w)
