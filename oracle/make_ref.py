"""oracle/make_ref.py -- TEST INFRASTRUCTURE ONLY (recipe; no reference source is committed).

Stages the UNMODIFIED Python sources of the reference that lie on or around the hot path into
``oracle/_ref/reference/`` so that they travel to the GPU box with the snapshot (``oracle/_ref/`` is
git-ignored, not gpurun-ignored).  The copy is byte-for-byte (``shutil.copy2``; a SHA-256 manifest is written
next to it) and is only ever produced from ``/root/reference`` in the build container by
``__graft_entry__.build()`` or ``python oracle/make_ref.py``.

Who may use the staged tree (the same rule as the rest of ``oracle/``):
  * ``tests/test_gpu_reference_callers.py`` / ``scripts/run_reference_callers.py`` -- run the reference's own
    ``get_basis.py`` and ``train.py`` as ``__main__`` (their own argparse, loaders, loops) against either the
    drop-in modules or the reference's modules;
  * ``bench.py --impl reference`` and the ``cpu_baseline`` leg -- time the reference's own ``model/`` classes on
    the host cores (``kind = "reference"``).
The product package never imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DEFAULT = "/root/reference"
DST = os.path.join(HERE, "_ref", "reference")

# files the callers import (train.py:22-29, get_basis.py:19-23) -- nothing from pretrain/ or the pre-processing
FILES = [
    "train.py", "get_basis.py",
    "model/dgcnn.py", "model/attention.py", "model/capl.py",
    "runs/__init__.py", "runs/eval.py",
    "util/util.py", "util/checkpoint_util.py", "util/logger.py",
    "dataloaders/__init__.py", "dataloaders/loader.py", "dataloaders/s3dis.py", "dataloaders/scannet.py",
]


def ref_dir() -> str | None:
    """the staged tree, or None when it has not been produced (then the tests that need it skip)"""
    return DST if os.path.exists(os.path.join(DST, "MANIFEST.json")) else None


def make(src: str = SRC_DEFAULT) -> str | None:
    if not os.path.isdir(src):
        return ref_dir()
    manifest = {}
    for rel in FILES:
        s = os.path.join(src, rel)
        if not os.path.exists(s):
            continue
        d = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copy2(s, d)
        manifest[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    out = make(sys.argv[1] if len(sys.argv) > 1 else SRC_DEFAULT)
    print(out if out else "reference not present; nothing staged")
