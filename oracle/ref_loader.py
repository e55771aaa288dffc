"""Import the staged, unmodified reference (oracle/_ref, made by oracle/make_ref.py) -- TEST / MEASUREMENT
INFRASTRUCTURE: only bench.py's CPU legs and tests/ use it.  Returns None when no staged copy exists (the
callers then time the port and say `kind: "port"`)."""
from __future__ import annotations

import hashlib
import importlib
import json
import sys
from pathlib import Path

DEST = Path(__file__).resolve().parent / "_ref"
_mod = None


def available() -> bool:
    return (DEST / "MANIFEST.json").exists() and (DEST / "src" / "deal_yolo_data" / "core" / "processor.py").exists()


def load():
    """The reference's `core.processor` module, verified against the staging manifest."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        return None
    manifest = json.loads((DEST / "MANIFEST.json").read_text())["sha256"]
    for rel, digest in manifest.items():
        if hashlib.sha256((DEST / rel).read_bytes()).hexdigest() != digest:
            raise RuntimeError(f"oracle/_ref/{rel} does not match its manifest: re-run python -m oracle.make_ref")
    if str(DEST) not in sys.path:
        sys.path.insert(0, str(DEST))
    for name in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        if not str(getattr(sys.modules[name], "__file__", "") or "").startswith(str(DEST)):
            del sys.modules[name]                        # another `src` package shadows the staged one
    _mod = importlib.import_module("src.deal_yolo_data.core.processor")
    return _mod


def run_hot_path_files(workdir, merged_csv, ref_csv=None, min_boxes=2, thr=0.7):
    """The bench chain through the reference's own step functions and CSV files: dedup (processor.py:111-164) ->
    reference filter (:166-219) -> ptList->bbox (:229-319) -> IoU filter (:321-407).  Same files and order as
    oracle/pipeline_port.run_hot_path_files."""
    ref = load()
    w = Path(workdir)
    cur = w / "deduplicate_result.csv"
    ref.deduplicate_csv_by_source(str(merged_csv), str(cur), verbose=False)
    if ref_csv is not None:
        ref.remove_duplicates_between_csv(str(cur), str(ref_csv), str(w / "filtered_main.csv"), verbose=False)
        cur = w / "filtered_main.csv"
    ref.process_csv_replace_ptlist(str(cur), str(w / "processed_replaced_ptlist.csv"), str(w / "processed_excluded.csv"))
    ref.filter_by_box_count_and_iou(str(w / "processed_replaced_ptlist.csv"), str(w / f"high_iou_{thr:.2f}.csv"),
                                    str(w / "other_data.csv"), min_boxes, thr)
