"""ORACLE (test infrastructure).  Imports the unmodified reference from /root/reference so that
`oracle/make_golden.py` can run it to produce the fixtures under `tests/golden/`.

This only works in the build container: /root/reference does not exist on the GPU box, and nothing
in `tests/ -m gpu`, `smoke()` or `bench.py` goes through this module.

Shims: a stub `visualizer_supcon` module (final_main.py:26 imports it and it
drags in matplotlib/umap), identity `.cuda()` on a CPU-only host (final_main.py:62,64,447-448,675),
and `torch.cuda.is_available() -> True` so set_model_multiple_adapter binds its return value
(final_main.py:338-343).
"""
from __future__ import annotations

import sys
import types

REFERENCE_ROOT = "/root/reference"


def import_reference(root=None):
    """root: where the reference lives -- /root/reference in the build container, or the copy staged by
    oracle/stage_reference.py under oracle/_ref/reference (the only form that reaches the GPU box)."""
    import torch

    root = root or REFERENCE_ROOT
    if root not in sys.path:
        sys.path.insert(0, root)
    if "visualizer_supcon" not in sys.modules:
        stub = types.ModuleType("visualizer_supcon")
        stub.skim_dataloader_by_group = lambda *a, **k: None
        sys.modules["visualizer_supcon"] = stub
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
        torch.cuda.is_available = lambda: True
        torch.cuda.manual_seed = lambda *a, **k: None
    # pandas >= 3 hands out read-only `.values` (copy-on-write); data/celeba_embeddings.py:35-36 assigns
    # into them in place (written for pandas 1.x).  Give the reference writable copies.
    import numpy as np
    import pandas as pd
    if not getattr(pd.Series, "_dbmm_writable_values", False):
        _orig_values = pd.Series.values.fget
        pd.Series.values = property(lambda self: np.array(_orig_values(self)))
        pd.Series._dbmm_writable_values = True
    import final_main  # noqa: E402  (the reference's module)
    import demo.util as ref_util  # noqa: E402
    return final_main, ref_util
