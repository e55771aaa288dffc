"""BASELINE.json config C1 as real CSV files: shared by the generator of the reference's SHA-256 pins
(tests/golden/make_c1_hashes.py) and by the tests that replay the same chain through the oracle port
(CPU) and through the CUDA drop-in (GPU)."""
from __future__ import annotations

import hashlib
from pathlib import Path

import numpy as np
import pandas as pd

from deal_yolo_daya_b200 import synth

ROWS = 10_000
SEED = 0
SRC = "source"
ANN = "结果字段-目标检测标签配置"


def sha256(path) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def write_inputs(td: Path) -> dict:
    """merged.csv = 10 k synthetic rows (5 % duplicate URLs, 10 % objects with a jittered twin);
    ref.csv = 3 k URLs of which a tenth also occur in merged.csv."""
    t = synth.make_table(SEED, 0, ROWS)
    rows = synth.table_to_rows(t)
    merged = td / "merged.csv"
    pd.DataFrame(rows, columns=[SRC, ANN]).to_csv(merged, index=False, encoding="utf-8-sig")
    ref_ids = synth.ref_ids_of(SEED, np.arange(3000), ROWS)
    ref = td / "ref.csv"
    pd.DataFrame({SRC: [synth.url_of(int(i)) for i in ref_ids]}).to_csv(ref, index=False, encoding="utf-8-sig")
    return {"merged": merged, "ref": ref}


def steps(mod, td: Path, paths: dict):
    """[(name, thunk, {key: produced file})] calling `mod`'s step functions positionally, the way
    ui/pages/processing.py does."""
    p = {k: td / f"{k}.csv" for k in ("dedup", "filtered", "rep", "exc", "hi70", "other70", "hi98", "other98")}
    return [
        ("dedup", lambda: mod.deduplicate_csv_by_source(str(paths["merged"]), str(p["dedup"])), {"dedup": p["dedup"]}),
        ("ref_filter", lambda: mod.remove_duplicates_between_csv(str(p["dedup"]), str(paths["ref"]), str(p["filtered"])),
         {"filtered": p["filtered"]}),
        ("replace_ptlist", lambda: mod.process_csv_replace_ptlist(str(p["filtered"]), str(p["rep"]), str(p["exc"])),
         {"rep": p["rep"]}),
        ("iou_0.70", lambda: mod.filter_by_box_count_and_iou(str(p["rep"]), str(p["hi70"]), str(p["other70"]), 2, 0.7),
         {"hi70": p["hi70"], "other70": p["other70"]}),
        ("iou_0.98", lambda: mod.filter_by_box_count_and_iou(str(p["rep"]), str(p["hi98"]), str(p["other98"]), 2, 0.98),
         {"hi98": p["hi98"], "other98": p["other98"]}),
    ]


# ---- label remap + split on the C1 chain's `other70.csv` (BASELINE config C5's rules at C1 size) ----
def mapping_frame() -> pd.DataFrame:
    """80 -> 20: cls{i} -> grp{i % 20}; a few labels left unmapped on purpose."""
    rows = [(synth.label_name(i), f"grp{i % 20:02d}") for i in range(synth.N_LABELS) if i % 17 != 5]
    return pd.DataFrame(rows, columns=["原标签", "新标签"])


def rules_frame() -> pd.DataFrame:
    """wide rules: 4 categories x 5 groups; grp19 undefined on purpose (unclassified rows)."""
    cats = {f"类别{c}": [f"grp{g:02d}" for g in range(c * 5, c * 5 + 5) if g != 19] for c in range(4)}
    n = max(len(v) for v in cats.values())
    return pd.DataFrame({k: v + [None] * (n - len(v)) for k, v in cats.items()})


def frame_digest(df: pd.DataFrame) -> str:
    return hashlib.sha256(df.to_csv(index=False).encode("utf-8")).hexdigest()
