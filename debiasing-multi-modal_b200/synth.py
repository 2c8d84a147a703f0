"""Deterministic synthetic stand-ins for the cached CLIP embeddings.

The real Waterbirds / CelebA embedding files are not shipped with the reference
(`/root/reference/.MISSING_LARGE_BLOBS`), so tests, the oracle fixtures and the bench all use
this generator.  Shapes and group sizes follow SURVEY.md section 8d; the on-disk writer
reproduces the reference formats (`clip_inference.py:237-269` image JSON, `clip_inference.py:68-106`
text JSON, `metadata.csv` columns read at `data/waterbirds_embeddings.py:26-41`, CelebA CSVs read
at `data/celeba_embeddings.py:23-41`).

    x   = base + k * mu_group + eps          (rounded through fp16: CLIP emits fp16)
    t_c = base + k' * mean_{g in c} mu_g + small noise
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np

# (train, val, test) group sizes, group id = 2*y + spurious
WATERBIRDS_GROUPS = ((3498, 184, 56, 1057), (467, 466, 133, 133), (2255, 2255, 642, 642))
CELEBA_GROUPS = ((71629, 66874, 22880, 1387), (8535, 8276, 2874, 182), (9767, 7535, 2480, 180))

CLASS_PROMPTS = {
    "waterbirds": ("a photo of a landbird.", "a photo of a waterbird."),
    "celeba": ("a photo of a celebrity with non-blond hair.", "a photo of a celebrity with blond hair."),
}
SPURIOUS_PROMPTS = {
    "waterbirds": ("a photo of a land background.", "a photo of a water background."),
    "celeba": ("a photo of a female.", "a photo of a male."),
}
GROUP_PROMPTS = {
    "waterbirds": (
        "a photo of a landbird on land background.", "a photo of a landbird on water background.",
        "a photo of a waterbird on land background.", "a photo of a waterbird on water background."),
    "celeba": (
        "a photo of a female celebrity with non-blond hair.", "a photo of a male celebrity with non-blond hair.",
        "a photo of a female celebrity with blond hair.", "a photo of a male celebrity with blond hair."),
}


@dataclass
class SyntheticSplit:
    x: np.ndarray          # [N, D] float32 (fp16-valued)
    y: np.ndarray          # [N] int64 class
    p: np.ndarray          # [N] int64 spurious attribute
    g: np.ndarray          # [N] int64 group = 2*y + p
    y_pred: np.ndarray     # [N] int64 zero-shot prediction
    filenames: list


@dataclass
class SyntheticDataset:
    name: str
    dim: int
    splits: dict                       # 'train' | 'val' | 'test' -> SyntheticSplit
    text_class: np.ndarray             # [D, 2] float32, unnormalised
    text_spurious: np.ndarray          # [D, 2]
    text_group: np.ndarray             # [D, 4]
    prompts: dict = field(default_factory=dict)


def scaled_group_sizes(groups, scale: float):
    """Shrink the per-split group sizes (at least 2 rows per group so a stratified 50/50 split works)."""
    return tuple(tuple(max(2, int(round(n * scale))) for n in split) for split in groups)


def make_dataset(name: str = "waterbirds", dim: int = 1024, seed: int = 1234, scale: float = 1.0,
                 k: float = 0.25, k_text: float = 1.0, text_noise: float = 0.05,
                 group_sizes=None, shuffle_rows: bool = True) -> SyntheticDataset:
    rng = np.random.default_rng(seed)
    base_sizes = WATERBIRDS_GROUPS if name == "waterbirds" else CELEBA_GROUPS
    sizes = group_sizes if group_sizes is not None else (
        base_sizes if scale == 1.0 else scaled_group_sizes(base_sizes, scale))
    base = rng.standard_normal(dim).astype(np.float32)
    mu = rng.standard_normal((4, dim)).astype(np.float32)

    def text(rows):
        cols = []
        for grp in rows:
            t = base + k_text * mu[list(grp)].mean(0) + text_noise * rng.standard_normal(dim).astype(np.float32)
            cols.append(t)
        return np.stack(cols, axis=1).astype(np.float16).astype(np.float32)

    text_class = text(((0, 1), (2, 3)))
    text_spurious = text(((0, 2), (1, 3)))
    text_group = text(((0,), (1,), (2,), (3,)))
    tc_hat = text_class / np.linalg.norm(text_class, axis=0, keepdims=True)

    splits = {}
    counter = 0
    for split_name, split_sizes in zip(("train", "val", "test"), sizes):
        g = np.concatenate([np.full(n, gi, dtype=np.int64) for gi, n in enumerate(split_sizes)])
        if shuffle_rows:
            g = g[rng.permutation(len(g))]
        n = len(g)
        x = np.empty((n, dim), dtype=np.float32)
        chunk = 16384
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            eps = rng.standard_normal((e - s, dim), dtype=np.float32)
            x[s:e] = base[None, :] + k * mu[g[s:e]] + eps
        x = x.astype(np.float16).astype(np.float32)
        y = g // 2
        p = g % 2
        y_pred = np.empty(n, dtype=np.int64)
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            xs = x[s:e]
            xs = xs / np.linalg.norm(xs, axis=1, keepdims=True)
            y_pred[s:e] = np.argmax(xs @ tc_hat, axis=1)
        if name == "waterbirds":
            fnames = [f"{(counter + i) % 200:03d}.Species/img_{counter + i:07d}.jpg" for i in range(n)]
        else:
            fnames = [f"{counter + i + 1:06d}.jpg" for i in range(n)]
        counter += n
        splits[split_name] = SyntheticSplit(x=x, y=y, p=p, g=g, y_pred=y_pred, filenames=fnames)

    prompts = {"class": CLASS_PROMPTS[name], "spurious": SPURIOUS_PROMPTS[name], "group": GROUP_PROMPTS[name]}
    return SyntheticDataset(name=name, dim=dim, splits=splits, text_class=text_class,
                            text_spurious=text_spurious, text_group=text_group, prompts=prompts)


def _fmt_vec(v: np.ndarray) -> str:
    # repr() of a Python float round-trips; values are fp16-exact so the strings stay short.
    return "[" + ", ".join(repr(float(t)) for t in v) + "]"


def write_reference_files(ds: SyntheticDataset, root: str) -> dict:
    """Write `ds` in the reference's own on-disk formats; returns the CLI path arguments."""
    emb_dir = os.path.join(root, "data", "embeddings_unnormalized", ds.name)
    os.makedirs(os.path.join(emb_dir, "RN50"), exist_ok=True)
    data_dir = os.path.join(root, "data", ds.name)
    os.makedirs(data_dir, exist_ok=True)
    split_id = {"train": 0, "val": 1, "test": 2}
    ykey, pkey = ("y", "place") if ds.name == "waterbirds" else ("blond", "male")

    image_path = os.path.join(emb_dir, "RN50", "clip.json")
    with open(image_path, "w") as f:
        f.write("{")
        first = True
        for sname, sp in ds.splits.items():
            for i, fn in enumerate(sp.filenames):
                if not first:
                    f.write(", ")
                first = False
                f.write(json.dumps(fn) + ": {")
                f.write(f'"{ykey}": "{int(sp.y[i])}", "{pkey}": "{int(sp.p[i])}", "group": "{int(sp.g[i])}", '
                        f'"split": "{split_id[sname]}", "y_pred": "{int(sp.y_pred[i])}", '
                        f'"image_embedding": {_fmt_vec(sp.x[i])}' + "}")
        f.write("}")

    def dump_text(mat, prompts, fname):
        path = os.path.join(emb_dir, fname)
        with open(path, "w") as f:
            json.dump({pr: [float(t) for t in mat[:, c]] for c, pr in enumerate(prompts)}, f)
        return path

    paths = {
        "image_embedding_dir": image_path,
        "text_embedding_dir": dump_text(ds.text_class, ds.prompts["class"], "clip_class.json"),
        "text_spurious_embedding_dir": dump_text(ds.text_spurious, ds.prompts["spurious"], "clip_spurious.json"),
        "text_group_embedding_dir": dump_text(ds.text_group, ds.prompts["group"], "clip_group.json"),
        "data_dir": data_dir,
    }

    if ds.name == "waterbirds":
        with open(os.path.join(data_dir, "metadata.csv"), "w") as f:
            f.write("img_id,img_filename,y,split,place,place_filename\n")
            k = 0
            for sname, sp in ds.splits.items():
                for i, fn in enumerate(sp.filenames):
                    k += 1
                    f.write(f"{k},{fn},{int(sp.y[i])},{split_id[sname]},{int(sp.p[i])},/x/{k}.jpg\n")
    else:
        with open(os.path.join(data_dir, "list_attr_celeba.csv"), "w") as fa, \
                open(os.path.join(data_dir, "list_eval_partition.csv"), "w") as fp:
            fa.write("image_id,Blond_Hair,Male\n")
            fp.write("image_id,partition\n")
            for sname, sp in ds.splits.items():
                for i, fn in enumerate(sp.filenames):
                    fa.write(f"{fn},{1 if sp.y[i] else -1},{1 if sp.p[i] else -1}\n")
                    fp.write(f"{fn},{split_id[sname]}\n")
    return paths
