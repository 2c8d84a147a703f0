"""ORACLE (test infrastructure / CPU baseline).  Recipe that stages the UNMODIFIED reference modules of the hot path
into `oracle/_ref/reference/` so that `bench.py --impl reference` and the `cpu_baseline` leg can run the reference's own
code on the GPU box's host cores (`/root/reference` does not exist there).

    python -m oracle.stage_reference          (also called by __graft_entry__.build() when /root/reference is present)

`oracle/_ref/` is git-ignored (nothing of the reference enters the history) but not gpurun-ignored, so the staged files
travel with the snapshot like the built `.so`.  Only the files the path imports are staged (final_main.py:1-30):
final_main.py, demo/util.py and the four embedding-dataset modules.  A MANIFEST with sha256 sums is written beside them.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref", "reference")
FILES = ("final_main.py", "demo/util.py", "data/waterbirds_embeddings.py", "data/waterbirds_embeddings_reg.py",
         "data/celeba_embeddings.py", "data/celeba_embeddings_reg.py")


def staged_root() -> str | None:
    """Where the reference can be imported from: the original tree in the build container, else the staged copy."""
    if os.path.exists(os.path.join(SRC, "final_main.py")):
        return SRC
    if os.path.exists(os.path.join(DST, "final_main.py")):
        return DST
    return None


def stage(verbose: bool = True) -> str | None:
    if not os.path.exists(os.path.join(SRC, "final_main.py")):
        return staged_root()
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=SRC, files=manifest), f, indent=1)
    if verbose:
        print(f"staged {len(FILES)} reference files into {DST}")
    return DST


if __name__ == "__main__":
    stage()
