"""Packed embedding store: a one-time conversion of the reference's on-disk inputs -- the embedding
JSON written by clip_inference.py:237-269 joined with metadata.csv / list_attr_celeba.csv + list_eval_partition.csv
(data/waterbirds_embeddings.py:30-67, data/celeba_embeddings.py) -- into ONE binary file that loads in milliseconds:

    header   64 bytes  magic "DBMMPACK", version, N, D, dtype code, section sizes, CRC-32 of everything after the header
    x        [N, D]    fp16 when every value is fp16-representable (CLIP emits fp16, clip/model.py:375-396), else fp32
    y, place, y_pred, split   int8 [N] each (split: 0 train / 1 val / 2 test, rows sorted by split, metadata order inside)
    names    uint32 [N + 1] offsets + UTF-8 blob (the filename column of the batch tuple, data/waterbirds_embeddings.py:87)
    source   JSON blob: dataset name and (path, size, mtime) of every input file -> staleness check

The reference re-parses the JSON 4x (Waterbirds) or 8x (CelebA, ~3 GB) per run (final_main.py:819-849); here it is parsed
once, at conversion.  `load_split` maps the file, checks the checksum and hands the rows of one split to
EmbeddingDataset (GPU-resident fp32), so datasets built from a pack are identical -- arrays, label vectors, filenames,
group ratios -- to those built from the JSON (tests/test_pack.py).

CLI:  python -m dbmm.pack --dataset waterbirds --data_dir DIR --image_embedding_dir clip.json [--out clip.json.dbmm]
`data._loaders` picks up `<image_embedding_dir>.dbmm` automatically when it exists and is not older than its sources.
"""
from __future__ import annotations

import argparse
import json
import os
import struct
import zlib

import numpy as np

MAGIC = b"DBMMPACK"
VERSION = 1
HEADER = struct.Struct("<8sIIQIIQQQI4x")        # magic, version, dtype, N, D, reserved, x_bytes, names_bytes, source_bytes, crc
DTYPE_CODE = {np.dtype(np.float16): 1, np.dtype(np.float32): 2}
CODE_DTYPE = {v: k for k, v in DTYPE_CODE.items()}
SPLIT_ID = {"train": 0, "val": 1, "test": 2}


class PackError(RuntimeError):
    pass


def default_pack_path(embedding_json: str) -> str:
    return embedding_json + ".dbmm"


def _source_files(name: str, data_dir: str, embedding_json: str):
    files = [embedding_json]
    if name == "waterbirds":
        files.append(os.path.join(data_dir, "metadata.csv"))
    else:
        files += [os.path.join(data_dir, "list_attr_celeba.csv"), os.path.join(data_dir, "list_eval_partition.csv")]
    return files


def _stamp(path: str):
    st = os.stat(path)
    return [os.path.abspath(path), int(st.st_size), int(st.st_mtime_ns)]


def pack_arrays(out_path: str, x: np.ndarray, y, place, y_pred, split, filenames, source: dict | None = None,
                allow_fp16: bool = True) -> str:
    """Write one pack from in-memory arrays (rows in final order)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, d = x.shape
    labels = [np.asarray(a) for a in (y, place, y_pred, split)]
    for a, nm in zip(labels, ("y", "place", "y_pred", "split")):
        if a.shape != (n,) or a.min(initial=0) < -128 or a.max(initial=0) > 127:
            raise PackError(f"label column {nm}: expected {n} values in int8 range")
    if len(filenames) != n:
        raise PackError("one filename per row expected")
    x16 = x.astype(np.float16)
    lossless = allow_fp16 and bool(np.array_equal(x16.astype(np.float32), x))
    xs = x16 if lossless else x
    blob = [s.encode("utf-8") for s in filenames]
    offs = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum([len(b) for b in blob], out=offs[1:])
    names = offs.tobytes() + b"".join(blob)
    src = json.dumps(source or {}).encode("utf-8")
    body = [xs.tobytes()] + [a.astype(np.int8).tobytes() for a in labels] + [names, src]
    crc = 0
    for part in body:
        crc = zlib.crc32(part, crc)
    head = HEADER.pack(MAGIC, VERSION, DTYPE_CODE[xs.dtype], n, d, 0, xs.nbytes, len(names), len(src), crc & 0xFFFFFFFF)
    tmp = out_path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(head)
        for part in body:
            f.write(part)
    os.replace(tmp, out_path)
    return out_path


def convert(name: str, data_dir: str, embedding_json: str, out_path: str | None = None) -> str:
    """Reference files -> pack.  Row order: train, val, test; inside a split the metadata order (what the reference's
    Dataset indexes by), with the same per-row consistency assertion (data/waterbirds_embeddings.py:84-85)."""
    from . import data
    out_path = out_path or default_pack_path(embedding_json)
    xs, ys, ps, yps, sps, fns = [], [], [], [], [], []
    for split in ("train", "val", "test"):
        x, y, place, y_pred, files = data.read_split_arrays(name, data_dir, embedding_json, split)
        xs.append(x); ys.append(y); ps.append(place); yps.append(y_pred)
        sps.append(np.full(len(files), SPLIT_ID[split], dtype=np.int8)); fns += list(files)
    source = {"dataset": name, "files": [_stamp(p) for p in _source_files(name, data_dir, embedding_json)]}
    return pack_arrays(out_path, np.concatenate(xs), np.concatenate(ys), np.concatenate(ps), np.concatenate(yps),
                       np.concatenate(sps), fns, source)


class Pack:
    """Read side: memory-maps the file; `verify()` recomputes the checksum."""

    def __init__(self, path: str, verify: bool = True):
        self.path = path
        size = os.path.getsize(path)
        if size < HEADER.size:
            raise PackError(f"{path}: truncated header")
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        magic, ver, dt, n, d, _, x_bytes, names_bytes, src_bytes, crc = HEADER.unpack(bytes(self._mm[:HEADER.size]))
        if magic != MAGIC:
            raise PackError(f"{path}: not a dbmm pack")
        if ver != VERSION:
            raise PackError(f"{path}: pack version {ver}, this build reads {VERSION}")
        if dt not in CODE_DTYPE:
            raise PackError(f"{path}: unknown dtype code {dt}")
        self.n, self.d, self.dtype, self.crc = int(n), int(d), CODE_DTYPE[dt], int(crc)
        if x_bytes != self.n * self.d * self.dtype.itemsize:
            raise PackError(f"{path}: inconsistent header")
        want = HEADER.size + x_bytes + 4 * self.n + names_bytes + src_bytes
        if size != want:
            raise PackError(f"{path}: {size} bytes on disk, header describes {want}")
        o = HEADER.size
        self.x = self._mm[o:o + x_bytes].view(self.dtype).reshape(self.n, self.d); o += x_bytes
        cols = []
        for _ in range(4):
            cols.append(self._mm[o:o + self.n].view(np.int8)); o += self.n
        self.y, self.place, self.y_pred, self.split = cols
        offs = self._mm[o:o + 4 * (self.n + 1)].view(np.uint32)
        blob = bytes(self._mm[o + 4 * (self.n + 1):o + names_bytes]); o += names_bytes
        self.filenames = np.array([blob[offs[i]:offs[i + 1]].decode("utf-8") for i in range(self.n)], dtype=object) \
            if self.n else np.array([], dtype=object)
        self.source = json.loads(bytes(self._mm[o:o + src_bytes]).decode("utf-8")) if src_bytes else {}
        if verify:
            self.verify()

    def verify(self) -> None:
        crc, step = 0, 64 << 20
        for o in range(HEADER.size, self._mm.shape[0], step):
            crc = zlib.crc32(self._mm[o:o + step], crc)
        if (crc & 0xFFFFFFFF) != self.crc:
            raise PackError(f"{self.path}: checksum mismatch (file corrupted or truncated)")

    def is_fresh(self) -> bool:
        """True when every recorded source file still has the size and mtime it had at conversion."""
        try:
            return all(_stamp(p)[1:] == [sz, mt] for p, sz, mt in self.source.get("files", []))
        except OSError:
            return False

    def rows_of(self, split: str) -> np.ndarray:
        return np.nonzero(np.asarray(self.split) == SPLIT_ID[split])[0]

    def split_arrays(self, split: str, raw: bool = False):
        """(x, y, place, y_pred, filenames) of one split; raw=True keeps x in the stored dtype (fp16 stores are widened
        on the device by the dataset, dbmm_widen_f16)."""
        r = self.rows_of(split)
        lo, hi = (int(r[0]), int(r[-1]) + 1) if len(r) else (0, 0)
        if len(r) and hi - lo != len(r):
            raise PackError(f"{self.path}: rows of split {split} are not contiguous")
        sl = slice(lo, hi)
        return (np.asarray(self.x[sl]) if raw else np.asarray(self.x[sl], dtype=np.float32), np.asarray(self.y[sl]).astype(np.int64),
                np.asarray(self.place[sl]).astype(np.int64), np.asarray(self.y_pred[sl]).astype(np.int64),
                list(self.filenames[sl]))


_open_packs: dict = {}


def open_pack(path: str) -> Pack:
    key = (os.path.abspath(path), os.path.getmtime(path))
    if key not in _open_packs:
        _open_packs.clear()
        _open_packs[key] = Pack(path)
    return _open_packs[key]


def usable_pack(name: str, data_dir: str, embedding_json: str):
    """The pack next to `embedding_json` if it exists, verifies, was built for this dataset and is not stale; else None."""
    path = default_pack_path(embedding_json)
    if not os.path.exists(path):
        return None
    pk = open_pack(path)
    want = [os.path.abspath(p) for p in _source_files(name, data_dir, embedding_json)]
    have = [f[0] for f in pk.source.get("files", [])]
    if pk.source.get("dataset") != name or have != want or not pk.is_fresh():
        return None
    return pk


def load_split(path: str, split: str, device=None, data_dir=None, embedding_dir=None):
    from . import data
    x, y, place, y_pred, files = open_pack(path).split_arrays(split)
    return data.EmbeddingDataset(x, y, place, y_pred, files, split=split, device=device, data_dir=data_dir,
                                 embedding_dir=embedding_dir)


def main(argv=None):
    ap = argparse.ArgumentParser("dbmm.pack", description="convert the reference's embedding JSON + metadata CSVs to a packed store")
    ap.add_argument("--dataset", required=True, choices=["waterbirds", "celeba"])
    ap.add_argument("--data_dir", required=True)
    ap.add_argument("--image_embedding_dir", required=True, help="embedding JSON (clip_inference.py output)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args(argv)
    out = convert(a.dataset, a.data_dir, a.image_embedding_dir, a.out)
    pk = Pack(out)
    print(f"{out}: {pk.n} rows x {pk.d} ({pk.dtype}), {os.path.getsize(out) / 1e6:.1f} MB, crc32 {pk.crc:08x}, "
          f"splits {[int((np.asarray(pk.split) == s).sum()) for s in range(3)]}")
    return out


if __name__ == "__main__":
    main()
