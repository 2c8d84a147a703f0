"""Packed embedding store: conversion from the reference's JSON / CSV files is lossless and the
datasets built from a pack equal those built from the files; corruption, truncation and stale sources are detected."""
import os

import numpy as np
import pytest

import dbmm
from dbmm import data, pack, synth


@pytest.fixture(scope="module", params=["waterbirds", "celeba"])
def files(request, tmp_path_factory):
    sizes = ((40, 9, 5, 21), (12, 11, 6, 7), (13, 12, 5, 9))
    ds = synth.make_dataset(name=request.param, dim=64, group_sizes=sizes, seed=3)
    root = tmp_path_factory.mktemp(request.param)
    return request.param, ds, synth.write_reference_files(ds, str(root))


def test_pack_round_trip_equals_json_path(files):
    name, ds, paths = files
    ref = {s: data.read_split_arrays(name, paths["data_dir"], paths["image_embedding_dir"], s) for s in ("train", "val", "test")}
    out = pack.convert(name, paths["data_dir"], paths["image_embedding_dir"])
    assert out == paths["image_embedding_dir"] + ".dbmm"
    pk = pack.Pack(out)
    assert pk.dtype == np.float16            # synthetic embeddings are fp16-valued like CLIP's -> half the bytes, lossless
    assert pk.n == sum(len(r[4]) for r in ref.values()) and pk.d == 64
    for s, (x, y, place, y_pred, fns) in ref.items():
        px, py, pp, pyp, pf = pk.split_arrays(s)
        assert px.dtype == np.float32 and np.array_equal(px, x)
        assert np.array_equal(py, y) and np.array_equal(pp, place) and np.array_equal(pyp, y_pred) and pf == list(fns)
    # the loaders pick the pack up on their own and build identical datasets
    assert pack.usable_pack(name, paths["data_dir"], paths["image_embedding_dir"]) is not None
    a = data._build_split(name, paths["data_dir"], paths["image_embedding_dir"], "val")
    os.rename(out, out + ".off")
    try:
        b = data._build_split(name, paths["data_dir"], paths["image_embedding_dir"], "val")
    finally:
        os.rename(out + ".off", out)
    assert np.array_equal(a.x.cpu().numpy(), b.x.cpu().numpy()) and np.array_equal(a.group_array, b.group_array)
    assert np.array_equal(a.y_pred_array, b.y_pred_array) and list(a.filename_array) == list(b.filename_array)
    assert np.array_equal(a.group_ratio.numpy(), b.group_ratio.numpy())


def test_pack_keeps_fp32_when_fp16_would_lose_bits(tmp_path):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((7, 16)).astype(np.float32)              # not fp16-representable
    out = pack.pack_arrays(str(tmp_path / "p.dbmm"), x, [0] * 7, [1] * 7, [0] * 7, [0, 0, 0, 1, 1, 2, 2], [f"f{i}.jpg" for i in range(7)])
    pk = pack.Pack(out)
    assert pk.dtype == np.float32 and np.array_equal(pk.split_arrays("val")[0], x[3:5])
    assert pk.split_arrays("test")[4] == ["f5.jpg", "f6.jpg"]


def test_pack_detects_corruption_truncation_and_stale_sources(files, tmp_path):
    name, ds, paths = files
    out = pack.convert(name, paths["data_dir"], paths["image_embedding_dir"], str(tmp_path / "c.dbmm"))
    raw = bytearray(open(out, "rb").read())
    raw[pack.HEADER.size + 11] ^= 0x40
    bad = str(tmp_path / "bad.dbmm"); open(bad, "wb").write(raw)
    with pytest.raises(pack.PackError, match="checksum"):
        pack.Pack(bad)
    open(bad, "wb").write(raw[:-5])
    with pytest.raises(pack.PackError, match="bytes on disk"):
        pack.Pack(bad)
    open(bad, "wb").write(b"NOTAPACK" + bytes(raw[8:]))
    with pytest.raises(pack.PackError, match="not a dbmm pack"):
        pack.Pack(bad)
    pk = pack.Pack(out)
    assert pk.is_fresh()
    meta = pk.source["files"][1][0]
    st = os.stat(meta)
    os.utime(meta, ns=(st.st_atime_ns, st.st_mtime_ns + 10 ** 9))
    try:
        assert not pack.Pack(out).is_fresh()
    finally:
        os.utime(meta, ns=(st.st_atime_ns, st.st_mtime_ns))


def test_pack_empty_and_label_range(tmp_path):
    out = pack.pack_arrays(str(tmp_path / "e.dbmm"), np.zeros((0, 8), np.float32), [], [], [], [], [])
    pk = pack.Pack(out)
    assert pk.n == 0 and pk.split_arrays("train")[0].shape == (0, 8)
    with pytest.raises(pack.PackError):
        pack.pack_arrays(str(tmp_path / "r.dbmm"), np.zeros((1, 8), np.float32), [300], [0], [0], [0], ["a"])


def test_pack_cli(files, tmp_path, capsys):
    name, ds, paths = files
    out = pack.main(["--dataset", name, "--data_dir", paths["data_dir"], "--image_embedding_dir", paths["image_embedding_dir"],
                     "--out", str(tmp_path / "cli.dbmm")])
    assert os.path.exists(out) and "rows x 64" in capsys.readouterr().out
