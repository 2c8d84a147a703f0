"""ORACLE (test infrastructure).  Seeded input cases shared by `oracle/make_golden.py` (which feeds
them to the real reference) and by the tests (which feed the same arrays to the oracle port and to
the CUDA path).  Inputs are regenerated from the seed on both sides; only the reference's OUTPUTS
are stored under tests/golden/.
"""
from __future__ import annotations

import numpy as np

from .adapter_math import init_adapter_params

# name -> (D, H, B, steps, lr, seed).  B values include ragged sizes (4795 % 1024 = 699).
TRAIN_CASES = {
    "rn50_b699": dict(D=1024, H=128, B=699, steps=3, lr=1.0, seed=11),
    "vitl_b256": dict(D=768, H=128, B=256, steps=3, lr=0.1, seed=12),
    "tiny_b33": dict(D=64, H=16, B=33, steps=4, lr=0.5, seed=13),
    "rn50_b4": dict(D=1024, H=128, B=4, steps=3, lr=1.0, seed=14),       # CelebA bsr=4 batches
}
EVAL_ROWS = 515   # deliberately not a multiple of any tile size


def make_case(name: str) -> dict:
    c = dict(TRAIN_CASES[name])
    rng = np.random.default_rng(c["seed"])
    D, H, B, steps = c["D"], c["H"], c["B"], c["steps"]
    mu = rng.standard_normal((4, D)).astype(np.float32)
    base = rng.standard_normal(D).astype(np.float32)

    def rows(n):
        g = rng.integers(0, 4, n)
        x = base + 0.3 * mu[g] + rng.standard_normal((n, D)).astype(np.float32)
        return x.astype(np.float16).astype(np.float32), g.astype(np.int64)

    xs, gs = zip(*[rows(B) for _ in range(steps)])
    xe, ge = rows(EVAL_ROWS)
    T2 =(base[:, None] + np.stack([mu[[0, 1]].mean(0), mu[[2, 3]].mean(0)], 1)).astype(np.float32)
    Tsp = (base[:, None] + np.stack([mu[[0, 2]].mean(0), mu[[1, 3]].mean(0)], 1)).astype(np.float32)
    T4 = (base[:, None] + mu.T).astype(np.float32)
    c.update(
        X=list(xs), G=list(gs), Y=[g // 2 for g in gs], Xe=xe, Ge=ge, Ye=ge // 2, Pe=ge % 2,
        T_class=T2, T_spurious=Tsp, T_group=T4,
        p_old=init_adapter_params(rng, D, H), p_new=init_adapter_params(rng, D, H),
    )
    return c


def flat_params(p: dict) -> np.ndarray:
    return np.concatenate([np.asarray(p[k], np.float32).reshape(-1)
                           for k in ("W1", "b1", "gamma", "beta", "W2", "b2")])


def param_digest(p: dict, stride: int = 41) -> dict:
    """A compact, order-stable summary of a parameter set (full tensors are ~1 MB each)."""
    f = flat_params(p).astype(np.float64)
    return dict(sample=f[::stride].astype(np.float32), sum=np.float64(f.sum()), abs_sum=np.float64(np.abs(f).sum()),
                running_mean=np.asarray(p["running_mean"], np.float32),
                running_var=np.asarray(p["running_var"], np.float32),
                nbt=np.int64(p["num_batches_tracked"]))
