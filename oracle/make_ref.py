"""Recipe for oracle/_ref: the UNMODIFIED reference modules of the hot path, staged so that they can travel.

TEST / MEASUREMENT INFRASTRUCTURE.  The reference (Cyclones-Y/Deal-Yolo-Daya) is pure Python; nothing is compiled.
`/root/reference` exists only in the authoring container, so this script copies the two modules the path lives in
(`src/deal_yolo_data/core/processor.py`, `core/utils.py`) plus the three empty package `__init__.py` files, from
where they lie, into `oracle/_ref/` -- a directory that is git-ignored (reference sources never enter the
repository's history) but NOT gpurun-ignored, so the staged copy reaches the GPU box like a built `.so`.
`bench.py --impl reference` and the `cpu_baseline` leg import it through oracle/ref_loader.py and time the real
reference functions (`kind: "reference"`); without the staged copy they fall back to the port (`kind: "port"`).

    python -m oracle.make_ref            # run in the authoring container; __graft_entry__.build() does it too

A SHA-256 manifest of what was staged is written next to the copy, and oracle/ref_loader.py checks it on load.
"""
from __future__ import annotations

import hashlib
import json
import shutil
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_ROOT = Path("/root/reference")
DEST = HERE / "_ref"
FILES = [
    "src/__init__.py",
    "src/deal_yolo_data/__init__.py",
    "src/deal_yolo_data/core/__init__.py",
    "src/deal_yolo_data/core/processor.py",
    "src/deal_yolo_data/core/utils.py",
]


def stage(ref_root: Path = REF_ROOT, dest: Path = DEST):
    """Copy the path's modules byte for byte; returns the manifest (None when the reference is not present)."""
    if not (ref_root / FILES[-1]).exists():
        return None
    manifest = {}
    for rel in FILES:
        src, dst = ref_root / rel, dest / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(dst.read_bytes()).hexdigest()
    (dest / "MANIFEST.json").write_text(json.dumps({"source": str(ref_root), "sha256": manifest}, indent=1))
    return manifest


if __name__ == "__main__":
    m = stage()
    print("staged" if m else "reference not present; nothing staged", json.dumps(m, indent=1) if m else "")
