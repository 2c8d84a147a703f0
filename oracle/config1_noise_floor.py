"""ORACLE (test infrastructure).  Noise floor of the FULL config-1 run (run_final_main.sh:1-31) of the UNMODIFIED reference
against ITSELF: the same code, seed and data, run with a different number of CPU threads (a different fp32 summation order
inside its matmuls / reductions -- the kind of difference any other machine or BLAS introduces).  Counts, per epoch, whether
the 4-decimal test dictionary equals the one in tests/golden/e2e_cases.json (generated at 8 threads), and records the selected
epoch and the final test dictionary.  Written to tests/golden/config1_noise_floor.json; read by
tests/test_e2e_gpu.py::test_full_config1_worst_group_accuracy_delta to put the CUDA path's own deviation in context.

Run in the build container only:   python -m oracle.config1_noise_floor
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_golden as mg  # noqa: E402


def run_reference(fm, threads: int):
    torch.set_num_threads(threads)
    case = mg.E2E_CASES["waterbirds_full"]
    root = tempfile.mkdtemp(prefix=f"dbmm_noise_{threads}_")
    _, argv = mg.e2e_paths_argv(case, root)
    old = sys.argv
    sys.argv = ["final_main.py"] + argv
    try:
        opt = fm.parse_option()
    finally:
        sys.argv = old
    tests = []
    tv = fm.validate

    def inner(*a, **k):
        r = tv(*a, **k)
        tests.append({kk: float(vv) for kk, vv in r[2].items()})
        return r
    fm.validate = inner
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            res = fm.train_all_epochs(opt)
    finally:
        fm.validate = tv
    best = int([ln for ln in buf.getvalue().splitlines() if ln.startswith("best epoch")][0].split(":")[1])
    return tests[1::2], best, {k: float(v) for k, v in res[0][2].items()}


def main():
    fm, _ = mg.import_reference()
    gold = json.load(open(os.path.join(mg.GOLD, "e2e_cases.json")))["waterbirds_full"]
    out = {"golden_threads": 8, "runs": []}
    for threads in (8, 1, 3):
        test_dicts, best, final_test = run_reference(fm, threads)
        exact = sum(int(all(d[k] == v for k, v in gold["val_test"][2 * e + 1]["group_acc"].items())) for e, d in enumerate(test_dicts))
        first_diff = next((e + 1 for e, d in enumerate(test_dicts)
                           if any(d[k] != v for k, v in gold["val_test"][2 * e + 1]["group_acc"].items())), None)
        out["runs"].append(dict(threads=threads, epochs=len(test_dicts), epochs_exact=exact, first_differing_epoch=first_diff,
                                best_epoch=best, test_worst_acc=final_test["worst_acc"], test_mean_acc=final_test["mean_acc"],
                                worst_acc_delta_vs_golden=final_test["worst_acc"] - gold["final"][2]["worst_acc"]))
        print(out["runs"][-1], flush=True)
    with open(os.path.join(mg.GOLD, "config1_noise_floor.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
