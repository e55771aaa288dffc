"""Generate the golden fixtures by RUNNING THE UNMODIFIED REFERENCE.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

It writes small seeded inputs and the outputs the reference's own step
functions produce for them (processor.py:111-831) under tests/golden/.  The
fixtures pin both the CPU oracle (tests/test_oracle_golden.py) and the CUDA
drop-in (tests/test_gpu_dropin.py).  Nothing at test time reads /root/reference.

openpyxl is absent, so label-remap / split run through an in-memory shim for
pd.read_excel / pd.ExcelWriter / DataFrame.to_excel (reference code untouched).
"""
from __future__ import annotations

import contextlib
import gzip
import io
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import pandas as pd

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference")

from deal_yolo_daya_b200 import synth  # noqa: E402
from src.deal_yolo_data.core import processor as ref  # noqa: E402

ANN = "结果字段-目标检测标签配置"
NEW = "新_结果字段-目标检测标签配置"

# --------------------------------------------------------------------------
BOOK: dict[str, dict[str, pd.DataFrame]] = {}


class _Writer:
    def __init__(self, path, *a, **k):
        self.path = str(path); BOOK[self.path] = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _read_excel(path, sheet_name=None, **k):
    book = BOOK[str(path)]
    return book[sheet_name or next(iter(book))].copy()


def _to_excel(self, target, sheet_name="Sheet1", index=True, **k):
    if isinstance(target, _Writer):
        BOOK[target.path][sheet_name] = self.copy()
    else:
        BOOK.setdefault(str(target), {})[sheet_name] = self.copy()


pd.read_excel = _read_excel
pd.ExcelWriter = _Writer
pd.DataFrame.to_excel = _to_excel

# --------------------------------------------------------------------------

def box_obj(name, x1, y1, x2, y2, **extra):
    o = {"name": name, "polygon": {"ptList": [{"x": x1, "y": y1}, {"x": x2, "y": y1}, {"x": x2, "y": y2}, {"x": x1, "y": y2}]}}
    o.update(extra)
    return o


def edge_rows():
    """Hand-written known-answer rows (SURVEY.md §8a edge cases)."""
    J = lambda d: json.dumps(d, ensure_ascii=False)  # noqa: E731
    A = box_obj("cls01", 10, 10, 110, 110)
    rows = [
        # first-occurrence ties: 10.0 before 10 -> "10.0"; 5 before 5.0 -> "5"
        ("kat://tie", J({"width": 100, "height": 100, "objects": [
            {"name": "cls00", "polygon": {"ptList": [{"x": 10.0, "y": 5}, {"x": 10, "y": 5.0}, {"x": 10, "y": 7}]}}]})),
        # signed zeros
        ("kat://zeros", '{"width": 64, "height": 64, "objects": [{"name": "cls02", "polygon": {"ptList": '
                        '[{"x": 0.0, "y": -0.0}, {"x": -0.0, "y": 0.0}, {"x": 0.0, "y": -0.0}]}}, '
                        '{"name": "cls02", "polygon": {"ptList": [{"x": -0.0, "y": 0.0}, {"x": 0.0, "y": -0.0}]}}]}'),
        # polygons without a valid point -> null bbox
        ("kat://novalid", J({"width": 10, "height": 10, "objects": [
            {"name": "cls03", "polygon": {"ptList": []}},
            {"name": "cls03", "polygon": {"ptList": [{"x": 1}, "a", 3, {"y": 2}]}},
            {"name": "cls03"}, {"name": "cls03", "polygon": {}}]})),
        # null bbox truncates the IoU scan: [null, A, A] -> other ; [A, A, null] -> high
        ("kat://trunc_front", J({"width": 200, "height": 200, "objects": [{"name": "cls04", "polygon": {"ptList": []}}, A, A]})),
        ("kat://trunc_back", J({"width": 200, "height": 200, "objects": [A, A, {"name": "cls04", "polygon": {"ptList": []}}]})),
        # zero-area identical boxes never hit
        ("kat://zeroarea", J({"width": 50, "height": 50, "objects": [box_obj("cls05", 5, 5, 5, 20), box_obj("cls05", 5, 5, 5, 20)]})),
        # IoU exactly 70/100 -> high at thr 0.7 (>=), not at 0.98
        ("kat://boundary", J({"width": 20, "height": 20, "objects": [box_obj("cls06", 0, 0, 10, 10), box_obj("cls06", 0, 0, 10, 7)]})),
        ("kat://boundary_f", J({"width": 20, "height": 20, "objects": [box_obj("cls06", 0.0, 0.0, 10.0, 10.0), box_obj("cls06", 0.0, 0.0, 7.0, 10.0)]})),
        # one box / no boxes / no objects key
        ("kat://single", J({"width": 30, "height": 30, "objects": [box_obj("cls07", 1, 2, 3, 4)]})),
        ("kat://empty", J({"width": 30, "height": 30, "objects": []})),
        ("kat://noobjects", J({"width": 30, "height": 30})),
        # bad JSON -> None cell -> other
        ("kat://badjson", '{"width": 30, "objects": [oops'),
        # non-dict objects dropped; extra keys and key order preserved; unicode kept
        ("kat://mixed", J({"id": 7, "objects": [1, "x", box_obj("行人", 3.5, 4.25, 9.75, 8.5, score=0.5, attrs={"k": [1, 2]}), None,
                                                 box_obj("车,行人；cls09", 3.5, 4.25, 9.75, 8.5)], "height": 12, "width": 16, "tail": "末尾"})),
        # NaN / Infinity literals (json.loads accepts them); order-dependent min/max
        ("kat://nan_first", '{"width": 9, "height": 9, "objects": [{"name": "cls08", "polygon": {"ptList": '
                            '[{"x": NaN, "y": 1.0}, {"x": 2.0, "y": NaN}, {"x": 3.0, "y": 0.5}]}}]}'),
        ("kat://nan_mid", '{"width": 9, "height": 9, "objects": [{"name": "cls08", "polygon": {"ptList": '
                          '[{"x": 4.0, "y": 1.0}, {"x": NaN, "y": NaN}, {"x": 3.0, "y": 2.5}, {"x": Infinity, "y": -Infinity}]}},'
                          '{"name": "cls08", "polygon": {"ptList": [{"x": 4.0, "y": 1.0}, {"x": 3.0, "y": 2.5}, {"x": 1e999, "y": -1e999}]}}]}'),
        # width/height missing -> None -> float promotion of the whole column
        ("kat://nowh", J({"objects": [box_obj("cls10", 1, 1, 2, 2), box_obj("cls10", 1, 1, 2, 2)]})),
        # multi-token names for remap / split, object without name, empty name
        ("kat://labels", J({"width": 40, "height": 40, "objects": [box_obj("cls11, cls12|cls11", 1, 1, 9, 9), box_obj("未知标签", 2, 2, 8, 8),
                                                                   {"polygon": {"ptList": [{"x": 1, "y": 1}, {"x": 2, "y": 2}]}},
                                                                   box_obj("", 0, 0, 1, 1), box_obj("cls79;cls78", 0, 0, 30, 30)]})),
        # big integers that are still exact in fp64, negative coordinates
        ("kat://ints", J({"width": 4000, "height": 3000, "objects": [box_obj("cls13", -5, -7, 3999, 2999), box_obj("cls13", -5, -7, 3999, 2990)]})),
    ]
    return rows


def build_inputs():
    t = synth.make_table(seed=7, first_img=0, n_img=48)
    rows = synth.table_to_rows(t, decimals=1)
    main = pd.DataFrame(rows, columns=["source", ANN])
    main["source_file"] = "part_a.csv"
    edge = pd.DataFrame(edge_rows(), columns=["source", ANN])
    edge["source_file"] = "part_b.csv"
    # source-column cases: duplicates of earlier rows, empty (NaN) cells, the literal "nan",
    # an annotation-less row (excluded by step 4)
    extra = pd.DataFrame([
        (rows[3][0], rows[5][1], "part_c.csv"),
        ("kat://tie", edge_rows()[8][1], "part_c.csv"),
        ("", rows[6][1], "part_c.csv"),
        ("", rows[7][1], "part_c.csv"),
        ("nan", rows[8][1], "part_c.csv"),
        ("kat://noann", "", "part_c.csv"),
        ("kat://refhit", rows[9][1], "part_c.csv"),
        ("  kat://refhit", rows[9][1], "part_c.csv"),     # whitespace matters
        ("KAT://REFHIT", rows[9][1], "part_c.csv"),       # case matters
    ], columns=["source", ANN, "source_file"])
    full = pd.concat([main, edge, extra], ignore_index=True)
    refset = pd.DataFrame({"source": [rows[1][0], rows[20][0], "kat://refhit", "", "nan", "kat://unrelated", rows[1][0]],
                           "note": list("abcdefg")})
    mapping = pd.DataFrame({"old": [synth.label_name(i) for i in range(70)] + ["行人", " 车 ", "nan", ""],
                            "new": [f"grp{i % 20}" for i in range(70)] + ["person", "vehicle", "x", "y"]})
    rules = pd.DataFrame({
        "catA": ["grp0", "grp1", "grp2", "grp3,grp4", None],
        "catB": ["grp5", "grp6；grp7", "grp8", "grp9", "person"],
        "catC": ["grp10", "grp11", "grp12", "grp13", "grp14|vehicle"],
        "catD": ["grp15", "grp16", "grp17", "grp18", "cls79"],
    })
    return full, refset, mapping, rules


def gz_write(path: Path, text: str):
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(text.encode("utf-8"))


def jsonable(o):
    if isinstance(o, dict):
        return {str(k): jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, (Path,)):
        return o.name
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, float) and o != o:
        return None
    return o


def main():
    inp = HERE / "inputs"; exp = HERE / "expected"
    inp.mkdir(exist_ok=True); exp.mkdir(exist_ok=True)
    full, refset, mapping, rules = build_inputs()
    summary = {"versions": {"pandas": pd.__version__, "numpy": np.__version__, "python": sys.version.split()[0]}}
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        p_main = td / "merged.csv"; p_ref = td / "reference.csv"
        full.to_csv(p_main, index=False, encoding="utf-8-sig")
        refset.to_csv(p_ref, index=False, encoding="utf-8-sig")
        gz_write(inp / "merged.csv.gz", p_main.read_text(encoding="utf-8-sig"))
        gz_write(inp / "reference.csv.gz", p_ref.read_text(encoding="utf-8-sig"))
        (inp / "mapping.json").write_text(mapping.to_json(orient="records", force_ascii=False), encoding="utf-8")
        (inp / "rules.json").write_text(rules.to_json(orient="records", force_ascii=False), encoding="utf-8")

        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):
            for keep in ("first", "last"):
                out = td / f"dedup_{keep}.csv"
                d = ref.deduplicate_csv_by_source(str(p_main), str(out), "utf-8-sig", keep)
                gz_write(exp / f"dedup_{keep}.csv.gz", out.read_text(encoding="utf-8-sig"))
                summary[f"dedup_{keep}_rows"] = len(d)
            p_dedup = td / "dedup_first.csv"
            p_filt = td / "filtered_main.csv"
            f = ref.remove_duplicates_between_csv(str(p_dedup), str(p_ref), str(p_filt))
            summary["ref_filter_rows"] = len(f)
            gz_write(exp / "filtered_main.csv.gz", p_filt.read_text(encoding="utf-8-sig"))

            p_rep = td / "processed_replaced_ptlist.csv"; p_exc = td / "processed_excluded.csv"
            r = ref.process_csv_replace_ptlist(str(p_filt), str(p_rep), str(p_exc))
            summary["replace"] = jsonable(r)
            gz_write(exp / "processed_replaced_ptlist.csv.gz", p_rep.read_text(encoding="utf-8-sig"))
            gz_write(exp / "processed_excluded.csv.gz", p_exc.read_text(encoding="utf-8-sig"))

            for thr, mb in ((0.7, 2), (0.98, 2), (0.7, 3), (0.0, 1)):
                tag = f"{thr:.2f}_{mb}"
                p_hi = td / f"high_iou_{tag}.csv"; p_ot = td / f"other_{tag}.csv"
                ref.filter_by_box_count_and_iou(str(p_rep), str(p_hi), str(p_ot), mb, thr)
                gz_write(exp / f"high_iou_{tag}.csv.gz", p_hi.read_text(encoding="utf-8-sig"))
                gz_write(exp / f"other_{tag}.csv.gz", p_ot.read_text(encoding="utf-8-sig"))

            p_other = td / "other_0.70_2.csv"
            p_map = td / "label_mapping.xlsx"; BOOK[str(p_map)] = {"Sheet1": mapping}
            p_remap = td / "other_data_label_replaced.csv"
            p_diff = td / "diff.xlsx"; p_unm = td / "unmatched.xlsx"
            m = ref.replace_labels_by_mapping(str(p_other), str(p_map), str(p_remap), None, None, None, None,
                                              str(p_diff), str(p_unm))
            summary["remap"] = jsonable({"summary": m["summary"], "sample_diff": m["sample_diff"]})
            gz_write(exp / "other_data_label_replaced.csv.gz", p_remap.read_text(encoding="utf-8-sig"))
            gz_write(exp / "remap_diff.csv.gz", BOOK[str(p_diff)]["Sheet1"].to_csv(index=False))
            gz_write(exp / "remap_unmatched.csv.gz", BOOK[str(p_unm)]["Sheet1"].to_csv(index=False))

            p_rules = td / "rules.xlsx"; p_rules.touch(); BOOK[str(p_rules)] = {"Sheet1": rules}
            p_split = td / "split_by_category"
            s = ref.split_dataset_by_rules(str(p_remap), str(p_rules), str(p_split))
            summary["split"] = jsonable(s["summary"])
            summary["split_files"] = [Path(x).name for x in s["category_files"]]
            for path, sheets in list(BOOK.items()):
                if str(p_split) in path:
                    stem = Path(path).stem
                    for sh, df in sheets.items():
                        gz_write(exp / f"split__{stem}__{sh}.csv.gz", df.to_csv(index=False))
    (exp / "summary.json").write_text(json.dumps(summary, ensure_ascii=False, indent=1), encoding="utf-8")

    # the step that must crash: a None coordinate raises TypeError inside min() (processor.py:256)
    crash = pd.DataFrame([("kat://none", '{"objects": [{"polygon": {"ptList": [{"x": null, "y": 1}, {"x": 2, "y": 3}]}}]}')],
                         columns=["source", ANN])
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "crash.csv"; crash.to_csv(p, index=False, encoding="utf-8-sig")
        gz_write(inp / "crash_none.csv.gz", p.read_text(encoding="utf-8-sig"))
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                ref.process_csv_replace_ptlist(str(p), str(Path(td) / "o.csv"), None)
            raised = None
        except Exception as e:  # noqa: BLE001
            raised = type(e).__name__
    (exp / "crash_none.json").write_text(json.dumps({"raises": raised}), encoding="utf-8")
    total = sum(f.stat().st_size for f in HERE.rglob("*") if f.is_file())
    print("fixtures written,", total // 1024, "KiB; summary:", json.dumps(summary, ensure_ascii=False)[:600])


if __name__ == "__main__":
    main()
