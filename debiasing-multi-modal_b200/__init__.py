"""dbmm-b200: B200-native regularized-adapter training/evaluation over cached CLIP embeddings.

The directory name follows the project's naming rule (`debiasing-multi-modal_b200`); because of the
hyphens it is imported through the top-level alias module `dbmm` (``import dbmm``).
"""
from . import _lib, ops, synth, metrics, data, optim, modules, engine, cli, parallel  # noqa: F401

__all__ = ["_lib", "ops", "synth", "metrics", "data", "optim", "modules", "engine", "cli", "parallel"]
