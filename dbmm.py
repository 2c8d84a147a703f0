"""Import alias: ``import dbmm`` -> the package in ``debiasing-multi-modal_b200/`` (hyphens are not importable)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("debiasing-multi-modal_b200")
for _k, _m in list(sys.modules.items()):
    if _k.startswith("debiasing-multi-modal_b200."):
        sys.modules["dbmm." + _k.split(".", 1)[1]] = _m
sys.modules[__name__] = _pkg
