"""Cached-embedding datasets, GPU-resident.

Replaces the reference's pandas/DataLoader path (data/waterbirds_embeddings[_reg].py,
data/celeba_embeddings[_reg].py): the embedding JSON is parsed ONCE (the reference parses it 4-8 times,
final_main.py:819-849), joined with the metadata CSV, and kept as one device-resident fp32 matrix per
split plus int32 label vectors.  Batches are index lists into that matrix, so a "loader" here is an
order generator; iterating it still yields the reference's batch tuple
`(embeddings [B,D] fp32, {"class","group","spurious","ebd_y_pred"}: int64 [B], filenames)`
(data/waterbirds_embeddings.py:87) for code that wants it.

Shuffling consumes the torch global RNG in the same sequence as `torch.utils.data.DataLoader` +
`RandomSampler` (one base-seed draw per iterator, one sampler-seed draw when shuffling, then
`torch.randperm` on a private generator), so a run with the same `--random_seed` visits the same
batches as the reference.
"""
from __future__ import annotations

import csv
import json
import os

import numpy as np
import torch

SPLIT_ID = {"train": 0, "val": 1, "test": 2}
_json_cache: dict = {}


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def read_embedding_json(path: str) -> dict:
    """{filename: {label fields as strings, 'image_embedding': [D floats]}} (clip_inference.py:237-269)."""
    key = (os.path.abspath(path), os.path.getmtime(path))
    if key not in _json_cache:
        _json_cache.clear()
        with open(path, "r") as f:
            _json_cache[key] = json.load(f)
    return _json_cache[key]


class EmbeddingDataset:
    """One split of cached embeddings; attribute names follow WaterbirdsEmbeddings / CelebaEmbeddings."""

    n_classes = 2
    n_groups = 4
    n_places = 2

    def __init__(self, x: np.ndarray, y, place, y_pred, filenames, split="train", device=None, data_dir=None,
                 embedding_dir=None):
        self.split = split
        self.data_dir = data_dir
        self.embedding_dir = embedding_dir
        self.split_dict = dict(SPLIT_ID)
        self.y_array = np.asarray(y).astype(np.int64)
        self.confounder_array = np.asarray(place).astype(np.int64)
        self.group_array = (self.y_array * 2 + self.confounder_array).astype("int")
        self.y_pred_array = np.asarray(y_pred).astype(np.int64)
        self.filename_array = np.asarray(filenames)
        self.targets = torch.tensor(self.y_array)
        self.targets_group = torch.tensor(self.group_array)
        self.targets_spurious = torch.tensor(self.confounder_array)
        self.group_counts = (torch.arange(self.n_groups).unsqueeze(1) == torch.from_numpy(self.group_array)).sum(1).float()
        self.group_ratio = self.group_counts / len(self)
        # device-resident store
        self.device = device if device is not None else _device()
        # x16: fp16-resident copy for the eval forward (dbmm_eval_fwd_f16) whenever the values are fp16-representable -- CLIP
        # emits fp16, so the reference's JSON files and the packed store both are; None otherwise (fp32 path only)
        self.x16 = None
        on_gpu = torch.device(self.device).type == "cuda"
        if x.dtype == np.float16 and on_gpu:
            # fp16 packed store (pack.py): half the host -> device bytes, widened exactly on the device (dbmm_widen_f16)
            from . import ops
            with torch.cuda.device(self.device):
                self.x16 = torch.from_numpy(np.ascontiguousarray(x)).to(self.device)
                self.x = ops.widen_f16(self.x16)
        else:
            x32 = np.ascontiguousarray(x, dtype=np.float32)
            self.x = torch.from_numpy(x32).to(self.device)
            if on_gpu and x32.shape[1] % 8 == 0:
                h = x32.astype(np.float16)
                if np.array_equal(h.astype(np.float32), x32):
                    self.x16 = torch.from_numpy(h).to(self.device)
        self.labels = {
            "class": torch.from_numpy(self.y_array.astype(np.int32)).to(self.device),
            "spurious": torch.from_numpy(self.confounder_array.astype(np.int32)).to(self.device),
            "group": torch.from_numpy(self.group_array.astype(np.int32)).to(self.device),
        }

    @property
    def dim(self):
        return self.x.shape[1]

    def __len__(self):
        return len(self.filename_array)

    def __getitem__(self, idx):
        idx = int(idx)
        return (self.x[idx], {"class": self.targets[idx], "group": self.targets_group[idx],
                              "spurious": self.targets_spurious[idx], "ebd_y_pred": int(self.y_pred_array[idx])},
                self.filename_array[idx])


class Subset:
    """torch.utils.data.Subset look-alike (`.dataset`, `.indices`) that can nest (balance_val builds a
    Subset of a Subset, final_main.py:376)."""

    def __init__(self, dataset, indices):
        self.dataset = dataset
        self.indices = np.asarray(indices)

    def __len__(self):
        return len(self.indices)

    def __getitem__(self, i):
        return self.dataset[self.indices[i]]


def resolve(ds):
    """(base EmbeddingDataset, base-row index array) for a dataset or any nesting of Subsets."""
    idx = None
    while isinstance(ds, Subset):
        idx = ds.indices if idx is None else ds.indices[idx]
        ds = ds.dataset
    if idx is None:
        idx = np.arange(len(ds))
    return ds, np.asarray(idx)


class EmbeddingLoader:
    """DataLoader look-alike over a device-resident dataset: `.dataset`, `.batch_size`, `len()`, iteration."""

    def __init__(self, dataset, batch_size=1, shuffle=False, num_workers=0):
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.shuffle = bool(shuffle)
        self.num_workers = num_workers
        self._forced_orders = []          # test hook: injected batch orders, consumed first

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def inject_order(self, order):
        self._forced_orders.append(np.asarray(order, dtype=np.int64))

    def draw_order(self) -> np.ndarray:
        """Positions (into `self.dataset`) in visiting order for one pass; consumes the global torch RNG
        exactly like creating a DataLoader iterator would."""
        n = len(self.dataset)
        torch.empty((), dtype=torch.int64).random_()                   # the iterator's base seed
        if self.shuffle:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
            gen = torch.Generator()
            gen.manual_seed(seed)
            order = torch.randperm(n, generator=gen).numpy()
        else:
            order = np.arange(n)
        if self._forced_orders:
            order = self._forced_orders.pop(0)
        return order

    def base_rows(self, order: np.ndarray):
        base, idx = resolve(self.dataset)
        return base, idx[order]

    def __iter__(self):
        base, rows = self.base_rows(self.draw_order())
        for s in range(0, len(rows), self.batch_size):
            r = rows[s:s + self.batch_size]
            rt = torch.from_numpy(r).to(base.device)
            labels = {"class": base.targets[r], "group": base.targets_group[r], "spurious": base.targets_spurious[r],
                      "ebd_y_pred": torch.from_numpy(base.y_pred_array[r])}
            yield base.x.index_select(0, rt), labels, list(base.filename_array[r])


def stratified_split_dataset(dataset, test_size=0.5):
    """50/50 stratified split of the validation split into (reg, val) with sklearn's
    train_test_split(random_state=42) -- the same call the reference makes
    (data/waterbirds_embeddings_reg.py:97-109), so the index sets are identical."""
    from sklearn.model_selection import train_test_split
    reg_idx, val_idx = train_test_split(np.arange(len(dataset.group_array)), test_size=test_size, random_state=42,
                                        stratify=dataset.group_array)
    return Subset(dataset, reg_idx), Subset(dataset, val_idx)


# ------------------------------------------------------------------------------------------------------
# file ingest (reference formats)
# ------------------------------------------------------------------------------------------------------
def _read_csv(path):
    with open(path, newline="") as f:
        return list(csv.DictReader(f))


def read_split_arrays(name, data_dir, embedding_dir, split):
    """(x fp32 [n, D], y, place, y_pred, filenames) of one split from the reference's own files."""
    emb = read_embedding_json(embedding_dir)
    sid = SPLIT_ID[split]
    if name == "waterbirds":
        rows = [r for r in _read_csv(os.path.join(data_dir, "metadata.csv")) if int(r["split"]) == sid]
        files = [r["img_filename"] for r in rows]
        y = np.array([int(r["y"]) for r in rows]); place = np.array([int(r["place"]) for r in rows])
        ykey, pkey = "y", "place"
    else:
        attr = _read_csv(os.path.join(data_dir, "list_attr_celeba.csv"))
        part = _read_csv(os.path.join(data_dir, "list_eval_partition.csv"))
        keep = [i for i, r in enumerate(part) if int(r["partition"]) == sid]
        files = [attr[i]["image_id"] for i in keep]
        y = np.array([max(int(attr[i]["Blond_Hair"]), 0) for i in keep])
        place = np.array([max(int(attr[i]["Male"]), 0) for i in keep])
        ykey, pkey = "blond", "male"
    D = len(next(iter(emb.values()))["image_embedding"])
    x = np.empty((len(files), D), dtype=np.float32)
    y_pred = np.empty(len(files), dtype=np.int64)
    for i, fn in enumerate(files):
        e = emb[fn]
        # same consistency check the reference asserts per item (data/waterbirds_embeddings.py:84-85)
        if int(e[ykey]) != y[i] or int(e[pkey]) != place[i] or int(e["group"]) != 2 * y[i] + place[i]:
            raise AssertionError(f"inconsistency between metadata in {data_dir} and {embedding_dir} for {fn}")
        x[i] = e["image_embedding"]
        y_pred[i] = int(e["y_pred"])
    return x, y, place, y_pred, files


def _build_split(name, data_dir, embedding_dir, split, device=None):
    """One split as a device-resident dataset: from the packed store next to the JSON when there is a fresh one
    (pack.py: no JSON parse at all), else from the reference's files."""
    from . import pack
    pk = pack.usable_pack(name, data_dir, embedding_dir)
    x, y, place, y_pred, files = pk.split_arrays(split, raw=True) if pk is not None else \
        read_split_arrays(name, data_dir, embedding_dir, split)
    return EmbeddingDataset(x, y, place, y_pred, files, split=split, device=device, data_dir=data_dir,
                            embedding_dir=embedding_dir)


class WaterbirdsEmbeddings(EmbeddingDataset):
    def __new__(cls, data_dir, split, embedding_dir, transform=None, device=None):
        return _build_split("waterbirds", data_dir, embedding_dir, split, device)


class CelebaEmbeddings(EmbeddingDataset):
    def __new__(cls, data_dir, split, embedding_dir, transform=None, device=None):
        return _build_split("celeba", data_dir, embedding_dir, split, device)


def _loaders(name, data_dir, embedding_dir, bs_train, bs_val, reg: bool, num_workers=16):
    train = _build_split(name, data_dir, embedding_dir, "train")
    val = _build_split(name, data_dir, embedding_dir, "val")
    test = _build_split(name, data_dir, embedding_dir, "test")
    train_loader = EmbeddingLoader(train, bs_train, shuffle=True, num_workers=num_workers)
    test_loader = EmbeddingLoader(test, bs_val, shuffle=False, num_workers=num_workers)
    if not reg:
        return train_loader, EmbeddingLoader(val, bs_val, shuffle=False, num_workers=num_workers), test_loader
    reg_set, val_set = stratified_split_dataset(val, test_size=0.5)
    reg_loader = EmbeddingLoader(reg_set, bs_val, shuffle=True, num_workers=num_workers)
    val_loader = EmbeddingLoader(val_set, bs_val, shuffle=False, num_workers=num_workers)
    return train_loader, reg_loader, val_loader, test_loader


def load_waterbirds_embeddings(data_dir, embedding_dir, bs_train=512, bs_val=256, num_workers=16, transform=None,
                               reg=False):
    return _loaders("waterbirds", data_dir, embedding_dir, bs_train, bs_val, reg, num_workers)


def load_celeba_embeddings(data_dir, embedding_dir, bs_train=512, bs_val=512, num_workers=16, transform=None,
                           reg=False):
    return _loaders("celeba", data_dir, embedding_dir, bs_train, bs_val, reg, num_workers)


def loaders_from_synthetic(ds, bs_train, bs_val, reg=True, device=None):
    """Same loader set from an in-memory synthetic dataset (synth.make_dataset) -- bench and tests."""
    def mk(split):
        sp = ds.splits[split]
        return EmbeddingDataset(sp.x, sp.y, sp.p, sp.y_pred, sp.filenames, split=split, device=device)
    train, val, test = mk("train"), mk("val"), mk("test")
    train_loader = EmbeddingLoader(train, bs_train, shuffle=True)
    test_loader = EmbeddingLoader(test, bs_val, shuffle=False)
    if not reg:
        return train_loader, EmbeddingLoader(val, bs_val, shuffle=False), test_loader
    reg_set, val_set = stratified_split_dataset(val)
    return (train_loader, EmbeddingLoader(reg_set, bs_val, shuffle=True),
            EmbeddingLoader(val_set, bs_val, shuffle=False), test_loader)
